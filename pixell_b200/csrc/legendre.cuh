// legendre.cuh -- Legendre stage (K1 alm->leg, K2 leg->alm) of the B200 SHT engine.
// Replaces the theta-dependent half of ducc0's synthesis / adjoint_synthesis
// (call sites pixell/curvedsky.py:907-924, 936-960, 1068-1084).
#pragma once
#include "common.cuh"

// one ring pair: a primary ring and (optionally) its mirror image theta -> pi-theta
struct PairInfo {
	double x;        // cos(theta) of the primary ring
	double sh, ch;   // sin(theta/2), cos(theta/2)
	int rn, rs;      // ring indices of the primary and the mirror ring (-1: absent)
};

// alpha-normalised recurrence tables for one (lmax, mmax, spin):
//   G_{l+1} = (a_l x -+ b_l) G_l - G_{l-1},   n_l d^l_{m,+-s} = alpha_l G_l,   l = l0(m)..lmax, l0 = max(m,s)
struct LegTables {
	int lmax = -1, mmax = -1, spin = -1;
	DevBuf<int64_t> toff;     // [mmax+2] offset of row m
	DevBuf<double> a, b, alpha;
	DevBuf<double> pref;      // [mmax+1] magnitude of the l0 start value without the theta powers
	int build(int lmax, int mmax, int spin);
	size_t bytes() const { return toff.bytes() + a.bytes() + b.bytes() + alpha.bytes() + pref.bytes(); }
};

struct LegGeom {
	int nring = 0;            // rings (leg ring index = caller's ring index)
	int npair = 0;
	int64_t nring_pad = 0;    // leg row length (rings padded to a multiple of 32)
	DevBuf<PairInfo> pairs;   // sorted pole -> equator, padded to a multiple of 256 pairs with rn = -1
	int npair_pad = 0;
	std::vector<int> rn_h, rs_h;   // host copy of the pairs' ring indices (which rings a range of pairs covers)
	int build(int nring, const double *theta);
	size_t bytes() const { return pairs.bytes(); }
};

// Start table of one (plan geometry, spin): for every m and every group of 128 ring pairs (sorted order) the first
// window of 8 l in which a ring of the group is live, and the recurrence state of all its rings at the start of that
// window.  The Legendre kernels then begin there instead of running the scaled recurrence up from l = max(m, s):
// same numbers, without the pre-phase (about 12 % of the synthesis time at lmax 8000).
#define LEG_GROUP 128
#define LEG_NEVER 0x7fffffff
struct LegStart {
	int spin = -1, ngroup = 0;
	DevBuf<int> w;                    // [mmax+1][ngroup] first live window (LEG_NEVER: the group never contributes)
	DevBuf<double> p, pp, q, qp;      // [mmax+1][npair_pad] state at that window (spin 0: p, pp only)
	DevBuf<signed char> sp, sq;       // scale indices
	size_t bytes() const { return w.bytes() + p.bytes() + pp.bytes() + q.bytes() + qp.bytes() + sp.bytes() + sq.bytes(); }
};

// Completion signalling of the adjoint Legendre kernels: the orders are cut into ranges [cut[r], cut[r+1]), r < ncut - 1;
// count: device counters (zeroed by the caller before the launch), flag: mapped host memory, one int per range on its
// own 64-byte line, set to `epoch` by the CTA that completes the range.
#define LEG_MAXCUT 10
struct LegSignal {
	int ncut, epoch; int cut[LEG_MAXCUT]; unsigned *count; volatile int *flag;
	__host__ __device__ volatile int *flag_of(int r) const { return flag + 16*r; }
};

struct AlmLayout {
	int lmax, mmax;
	const int64_t *mstart_d;  // device [mmax+1]
	int64_t lstride;
};

// leg[ncomp_map][mmax+1][nring_pad] complex128; alm component c at alm + c*alm_cstride (complex elements)
// S (nullable): start table built by leg_build_start for this (T, G)
// pair_lo < pair_hi: only the ring pairs [pair_lo, pair_hi) of the sorted pair list (multiples of 256) are computed;
// gate (nullable): the kernel's CTAs wait for the arrival flag of their range of m (LegSignal, flags in device memory)
// m_lo < m_hi: only the orders m_lo <= m < m_hi (partial launches: results stream out while the rest is computed)
int leg_alm2leg(const LegTables &T, const LegGeom &G, const AlmLayout &L, int deriv1,
                const double2 *alm, int64_t alm_cstride, double2 *leg, cudaStream_t st, const LegStart *S = nullptr,
                int pair_lo = 0, int pair_hi = 0, const LegSignal *gate = nullptr);
int leg_leg2alm(const LegTables &T, const LegGeom &G, const AlmLayout &L, int deriv1,
                double2 *alm, int64_t alm_cstride, const double2 *leg, cudaStream_t st, const LegStart *S = nullptr,
                int m_lo = 0, int m_hi = 0, const LegSignal *sig = nullptr);
// batched synthesis: nb alm sets of one spin share the recurrence (nb = leg_batch_size(spin), or 2 for spin 0); member b reads
// alm + b*alm_bstride (+ c*alm_cstride for the second component) and writes leg + b*leg_bstride (second component of a
// spin > 0 member at + (mmax+1)*nring_pad, as in leg_alm2leg).  Bit-identical to nb calls of leg_alm2leg.
int leg_batch_size(int spin);
int leg_alm2leg_batch(const LegTables &T, const LegGeom &G, const AlmLayout &L, int nb,
                      const double2 *alm, int64_t alm_cstride, int64_t alm_bstride, double2 *leg, int64_t leg_bstride,
                      cudaStream_t st, const LegStart *S = nullptr);
int leg_build_start(LegStart &S, const LegTables &T, const LegGeom &G);
int dfma_peak_gflops(double *out);
int leg_set_variant(int which, int v);
