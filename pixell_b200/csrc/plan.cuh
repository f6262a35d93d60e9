// plan.cuh -- the SHT plan object behind the C ABI (include/b200sht.h).
#pragma once
#include "legendre.cuh"
#include "ringfft.cuh"
#include "resample.cuh"
#include <map>
#include <vector>
#include <memory>
#include <string>

#define B2_MAX_GROUPS 8

struct b2_sht_plan {
	int lmax = 0, mmax = 0;
	int64_t lstride = 1;
	int64_t alm_span = 0;              // elements of one alm component touched by the layout
	bool alm_dense = false;            // the transform owns every element of the span
	DevBuf<int64_t> mstart;
	std::vector<int64_t> mstart_h;
	int nring = 0;
	int64_t nphi = 0, npix = 0;
	std::vector<int64_t> ringstart_h;
	int64_t map_lo = 0, map_hi = 0;    // element range of one map component touched by the rings
	int64_t row_pitch = 0;             // constant ring pitch (0: irregular)
	LegGeom geom;
	RingFft fft;
	// ring sets with per-ring nphi / phi0 (b2_sht_plan_rings_general): one RingFft per distinct nphi instead of `fft`
	std::vector<std::unique_ptr<RingFft>> groups;
	RingPack pack;                      // the groups' ring FFTs in a few launches (one per block size) instead of one per group
	std::vector<int64_t> npix_h;        // pixels of every ring (general plans)
	// the groups' launches are small (a cap ring pair each): they are spread over side streams so that they overlap
	std::vector<cudaStream_t> gstreams;
	std::vector<cudaEvent_t> gjoin;
	cudaEvent_t gfork = nullptr;
	bool dense_rings = true;           // the rings tile [map_lo, map_hi) without gaps
	std::map<int, std::unique_ptr<LegTables>> tables;   // by spin
	std::map<int, std::unique_ptr<LegStart>> starts;    // by spin: where every ring group's recurrence becomes live (see LegStart)
	DevBuf<double2> leg;               // [2][mmax+1][nring_pad]
	// Two-stage pipeline over the spin groups of one call: the FP64-bound Legendre kernels run on the caller's stream, the
	// HBM-bound stages (ring FFTs, theta weighting) on s_fft, so that group g+1's memory-bound work hides under group g's
	// Legendre kernel (and vice versa for alm -> map).  Groups alternate between two leg buffers ("lanes").
	DevBuf<double2> leg2;              // second lane, allocated by the first multi-group call
	DevBuf<double2> legb;              // four leg planes for batched synthesis (leg_alm2leg_batch), allocated on first use
	cudaStream_t s_fft = nullptr;      // high priority: its short kernels take the SM slots the long Legendre CTAs free
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_chunk[8] = {};
	// 2d plans
	bool is2d = false;
	std::string geometry;
	int ntheta = 0;
	std::unique_ptr<ThetaResampler> resamp;   // exact analysis on grids with ntheta < 2 lmax + 2
	DevBuf<double> w2d;                // direct quadrature weights / nphi (ring order of the plan)
	// staging for host-memory calls
	DevBuf<char> stage_alm, stage_map;
	cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	// host-memory calls are pipelined over three streams (copies in, kernels, copies out)
	cudaStream_t s_in = nullptr, s_out = nullptr, s_comp = nullptr;
	cudaEvent_t gev[2*B2_MAX_GROUPS] = {};      // per group: operands on the device, results ready
	double timing[4] = {0, 0, 0, 0};
	// the plan's scratch (leg, theta-stage buffers, staging) serves one call at a time: every call waits for the previous
	// one's last kernel, whatever streams the two run on (a device-memory call on the caller's stream followed by a
	// host-memory call on the plan's own streams would otherwise overlap in the scratch)
	cudaEvent_t ev_last = nullptr;
	// streamed host-memory transforms: the synthesis runs in chunks of ring pairs (pole -> equator) whose rows leave for the
	// host while the next chunk is computed; the adjoint Legendre stage runs in ranges of m whose alm leave likewise
	struct StreamChunk { int pair_lo, pair_hi, nrun, r0[2], nr[2]; };
	std::vector<StreamChunk> schunks;      // empty: rows of a chunk are not at most two dense row bands (no streaming)
	std::vector<int> mcuts;                // m range boundaries with about equal alm bytes (empty: alm layout not dense)
	cudaEvent_t sev[8] = {};
	// completion flags of the adjoint Legendre kernel (LegSignal): the alm of the last group of a host-memory call leave
	// range by range while the kernel is still running
	DevBuf<unsigned> sig_count;
	int *sig_flag_h = nullptr, *sig_flag_d = nullptr;      // mapped pinned host memory and its device address
	int sig_epoch = 0;
	// arrival flags for the synthesis kernels (the first group's alm of a host-memory call arrive range by range)
	DevBuf<int> gate_flag;
	int *gate_src_h = nullptr;         // pinned: the epoch value the copy stream writes into a range's flag
	int gate_epoch = 0;
	LegTables *get_tables(int spin);
	LegStart *get_start(int spin);      // nullptr when disabled (B2_NO_START_TABLE=1) or on allocation failure
	size_t bytes() const;
	~b2_sht_plan();
};

int gridweights_device(const char *geometry, int ntheta, double *out_host, double *out_dev);
int grid_theta_host(const char *geometry, int ntheta, std::vector<double> &theta);
