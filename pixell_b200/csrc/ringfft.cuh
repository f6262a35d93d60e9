// ringfft.cuh -- K3 (leg -> map, c2r ring FFT) and K4 (map -> leg, r2c ring FFT) of the SHT engine.
#pragma once
#include "fft_smem.cuh"
#include <vector>
#include <memory>

struct RingFft {
	int64_t nphi = 0;       // pixels per full circle
	int half = 0;           // 1: even nphi handled as a packed complex transform of length nphi/2
	int nfft = 0;           // complex transform length
	int mmax = 0;
	int xdir = 1;           // +1: phi grows with the pixel index, -1: decreases
	int64_t npix = 0;       // stored pixels per ring (<= nphi)
	int nring = 0;
	FftTables tab;
	DevBuf<double2> phase;  // exp(i m phi0), m = 0..mmax
	DevBuf<int64_t> ringstart;
	DevBuf<double> weight;  // empty: no weights
	// ring sets whose rings differ in nphi / phi0 (HEALPix, pixell/curvedsky.py:1192-1222) are served by one RingFft
	// per distinct nphi: block b of a launch works on ring ring_ids[b] (leg column, ringstart, weight are indexed by
	// the plan's ring number) with its own phi0s[b]; both empty for cylindrical maps
	DevBuf<int> ring_ids;
	DevBuf<double> phi0s;
	int threads = 256;
	size_t smem = 0;
	int twoff = 0;          // offset (elements) of the twiddle tables inside the dynamic shared memory
	int build(int64_t nphi, double phi0, int xdir, int64_t npix, int nring, const int64_t *ringstart,
	          const double *weight, int mmax);
	// host-only part of build / build_group (the FFT tables of this ring length): thread-safe, for parallel plan construction
	int prepare_tables(int64_t nphi);
	// group form: `ids` lists the plan rings with this nphi, phi0_of_id their phi0 (all rings' ringstart / weight arrays are passed whole)
	int build_group(int64_t nphi, int nids, const int *ids, const double *phi0_of_id, int nring_total, const int64_t *ringstart,
	                const double *weight, int mmax);
	size_t bytes() const { return tab.bytes() + phase.bytes() + ringstart.bytes() + weight.bytes() + ring_ids.bytes() + phi0s.bytes(); }
};

// leg: [ncomp][mmax+1][nring_pad] complex128 (device); map component c at map + c*map_cstride (elements of MapT)
// nrings > 0: only the rings [ring0, ring0 + nrings) (cylindrical plans)
int ring_leg2map(const RingFft &F, int ncomp, const double2 *leg, int64_t nring_pad,
                 void *map, int64_t map_cstride, int dtype, cudaStream_t st, int ring0 = 0, int nrings = 0);
int ring_map2leg(const RingFft &F, int ncomp, double2 *leg, int64_t nring_pad,
                 const void *map, int64_t map_cstride, int dtype, int use_weight, cudaStream_t st);

// All ring groups of a general plan (HEALPix) in a few launches: see k_leg2map_pack in ringfft.cu.  F0 is any group of the
// plan (mmax, phase table, ringstart and weight arrays are the plan's and common to all groups).
struct RingPack {
	struct Bucket { int threads, first, nblocks; size_t smem; };
	std::vector<Bucket> buckets;
	DevBuf<char> desc;          // GroupDesc per group
	DevBuf<int2> blocks;        // (group, position in the group) per block, bucket after bucket
	int build(const std::vector<std::unique_ptr<RingFft>> &groups);
	size_t bytes() const { return desc.bytes() + blocks.bytes(); }
};
int ring_leg2map_pack(const RingPack &P, const RingFft &F0, int ncomp, const double2 *leg, int64_t nring_pad,
                      void *map, int64_t map_cstride, int dtype, cudaStream_t st);
int ring_map2leg_pack(const RingPack &P, const RingFft &F0, int ncomp, double2 *leg, int64_t nring_pad,
                      const void *map, int64_t map_cstride, int dtype, int use_weight, cudaStream_t st);
