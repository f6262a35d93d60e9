// legendre.cu -- K1 (alm -> leg) and K2 (leg -> alm): the FP64-FMA-bound Legendre stage.
//
// Replaces the theta-dependent half of ducc0.sht.experimental.synthesis / adjoint_synthesis that
// pixell calls at pixell/curvedsky.py:907-924, 936-960, 1068-1084.  Not a port: the formulation is
// built for the B200 FP64 pipe (one DFMA per 2 clk per SM sub-partition, 64/clk/SM):
//   * one CTA per m; lanes = ring pairs (theta, pi-theta), R pairs per lane, so every recurrence
//     coefficient and alm value is a warp-uniform shared-memory broadcast that feeds 12R (spin>0)
//     or 4R (spin 0) DFMAs;
//   * alpha-normalised three-term recurrence in l for the Wigner functions n_l d^l_{m,-+s}
//     (2 DFMA per step and sequence), accumulation in the +- basis with north/south parity split;
//   * start values in extended-exponent form (binary powering), a rescaling pre-phase (window
//     variants A/B) and a check-free main phase (variant C); rings beyond the evanescent tail are
//     skipped (Airy-tail bound, see SeqConst::dead_sth);
//   * K2 reduces over rings with a halving warp-shuffle butterfly (about one 64-bit exchange per
//     output value and l), then across warps through shared memory, and accumulates into alm in
//     a fixed order: results are bit-reproducible run to run.
#include "legendre.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#define B2_DEF_SYNTH0 0
#define B2_DEF_ADJ0 9
#define B2_DEF_SYNTH2 5
#define B2_DEF_ADJ2 7

#define SMALLV 0x1p-512
// A sequence counts as "live" (enters the sums) once its magnitude has reached 2^-LIVE_EXP: what it misses before is
// below 2^-LIVE_EXP relative to O(1) harmonics (1e-27 for 90), far under double precision, and every ring joins
// the check-free main phase that much later in l (about 4 % fewer full-cost l steps at lmax 8000 than with the
// underflow-only threshold 2^-256).
#define LIVE_EXP 90

// ------------------------------------------------------------------------------------ tables

__global__ void k_build_tables(int lmax, int mmax, int s, const int64_t *toff, double *ta, double *tb, double *talpha)
{
	int m = blockIdx.x*blockDim.x + threadIdx.x;
	if (m > mmax) return;
	int l0 = m > s ? m : s;
	if (l0 > lmax) return;
	int n = lmax - l0 + 1;
	double *a = ta + toff[m], *b = tb ? tb + toff[m] : nullptr, *al = talpha + toff[m];
	double al_prev = 1.0, al_cur = 1.0;
	for (int i = 0; i < n; i++) {
		int l = l0 + i;
		double l1 = l + 1.0;
		double den = sqrt((l1 - m)*(l1 + m)*(l1 - s)*(l1 + s));
		double c0 = sqrt((2*l + 3.0)/(2*l + 1.0))*(2*l + 1.0)*l1/den;
		double c1 = (l == 0) ? 0.0 : c0*((double)m*s)/((double)l*l1);
		double c2 = (i == 0 || l == 0) ? 0.0 :
			sqrt((2*l + 3.0)/(2*l - 1.0))*l1*sqrt(((double)l - m)*((double)l + m)*((double)l - s)*((double)l + s))/(l*den);
		double al_next = (i == 0) ? 1.0 : c2*al_prev;
		al[i] = al_cur;
		a[i] = c0*al_cur/al_next;
		if (b) b[i] = c1*al_cur/al_next;
		al_prev = al_cur; al_cur = al_next;
	}
}

int LegTables::build(int lmax_, int mmax_, int spin_)
{
	lmax = lmax_; mmax = mmax_; spin = spin_;
	std::vector<int64_t> off(mmax + 2);
	int64_t tot = 0;
	for (int m = 0; m <= mmax; m++) { off[m] = tot; int l0 = std::max(m, spin); if (l0 <= lmax) tot += lmax - l0 + 1; }
	off[mmax + 1] = tot;
	// keep every row 16-byte friendly is not needed (rows are read element-wise)
	if (toff.upload(off)) return 1;
	if (a.alloc(tot) || alpha.alloc(tot)) return 1;
	if (spin > 0 && b.alloc(tot)) return 1;
	// start-value prefactors (host, long double): see legendre.cu header / DESIGN.md
	std::vector<double> pr(mmax + 1, 0.0);
	const long double pi = 3.141592653589793238462643383279502884L;
	{
		long double p2 = (2*spin + 1)/(4*pi);
		for (int m = spin; m <= mmax; m++) {
			pr[m] = (double)sqrtl(p2);
			p2 *= ((long double)(2*m + 3)*(2*m + 2))/(4.0L*(m + 1 + spin)*(m + 1 - spin));
		}
		for (int m = 0; m < spin && m <= mmax; m++) {
			long double f = (2*spin + 1)/(4*pi);     // n_s^2 (2s)!/((s+m)!(s-m)!)
			for (int k = 1; k <= 2*spin; k++) f *= k;
			for (int k = 1; k <= spin + m; k++) f /= k;
			for (int k = 1; k <= spin - m; k++) f /= k;
			pr[m] = (double)sqrtl(f);
		}
	}
	if (pref.upload(pr)) return 1;
	int nt = 128;
	k_build_tables<<<(mmax + nt)/nt, nt>>>(lmax, mmax, spin, toff.p, a.p, spin > 0 ? b.p : nullptr, alpha.p);
	B2_LAUNCH_CHECK();
	B2_CHECK(cudaDeviceSynchronize());
	return 0;
}

// ------------------------------------------------------------------------------------ ring pairs

int LegGeom::build(int nring_, const double *theta)
{
	nring = nring_;
	nring_pad = b2_round_up(nring, 32);
	std::vector<int> order(nring);
	for (int i = 0; i < nring; i++) order[i] = i;
	std::sort(order.begin(), order.end(), [&](int a, int b) { return theta[a] < theta[b]; });
	std::vector<char> used(nring, 0);
	std::vector<PairInfo> pr;
	int lo = 0, hi = nring - 1;
	const double tol = 1e-14*M_PI;
	// two-pointer match of theta_lo + theta_hi == pi on the sorted list
	while (lo <= hi) {
		int i = order[lo], j = order[hi];
		double s = theta[i] + theta[j] - M_PI;
		PairInfo p;
		if (lo < hi && fabs(s) < tol) { p.rn = i; p.rs = j; lo++; hi--; }
		else if (lo == hi || s < 0) { p.rn = i; p.rs = -1; lo++; }
		else { p.rn = j; p.rs = -1; hi--; }
		double t = theta[p.rn];
		p.x = cos(t);
		// half-angle functions accurate near both poles
		if (t <= M_PI_2) { p.sh = sin(0.5*t); p.ch = cos(0.5*t); }
		else { double u = M_PI - t; p.sh = cos(0.5*u); p.ch = sin(0.5*u); }
		pr.push_back(p);
	}
	// pole -> equator: chunks of neighbouring pairs become live at similar l
	std::stable_sort(pr.begin(), pr.end(), [](const PairInfo &a, const PairInfo &b) { return a.sh*a.ch < b.sh*b.ch; });
	npair = (int)pr.size();
	npair_pad = (int)b2_round_up(npair, 256);
	PairInfo dead; dead.x = 0; dead.sh = 0; dead.ch = 1; dead.rn = -1; dead.rs = -1;
	pr.resize(npair_pad, dead);
	rn_h.resize(npair_pad); rs_h.resize(npair_pad);
	for (int i = 0; i < npair_pad; i++) { rn_h[i] = pr[i].rn; rs_h[i] = pr[i].rs; }
	return pairs.upload(pr);
}

// ------------------------------------------------------------------------------------ device helpers

struct LegArgs {
	int lmax, mmax, spin, deriv1, npair;
	const int64_t *toff; const double *ta, *tb, *talpha, *pref;
	const PairInfo *pairs; int npair_pad;
	const int64_t *mstart; int64_t lstride;
	double2 *alm0, *alm1;
	double2 *leg0, *leg1;
	int64_t leg_mstride;
	// start table (st_w == nullptr: none; kernels then start every ring at l = max(m, s))
	const int *st_w; const double *st_p, *st_pp, *st_q, *st_qp; const signed char *st_sp, *st_sq; int ngroup;
	// partial launches (host-memory calls stream their results out while the rest is still being computed):
	// CTA b works on m = m0 + b; the synthesis kernels only touch the ring pairs [pair_lo, pair_hi) (multiples of 256)
	int m0, pair_lo, pair_hi;
	LegSignal sig;      // adjoint kernels: completion flags per range of m (sig.count == nullptr: none)
	int64_t alm_bstride, leg_bstride;      // batched synthesis (k_synth0b / k_synth2b): element strides between the batch members
};

// base^n = mant*2^ex with mant in [0.5,1) (or mant = 1, ex = 0 for n = 0; mant = 0 for base = 0)
__device__ __forceinline__ void scaled_pow(double base, int n, double &mant, int &ex)
{
	if (n == 0) { mant = 1.0; ex = 0; return; }
	if (base == 0.0) { mant = 0.0; ex = 0; return; }
	int be; double b = frexp(base, &be);
	double r = 1.0; int re = 0;
	while (true) {
		if (n & 1) { r *= b; re += be; int t; r = frexp(r, &t); re += t; }
		n >>= 1;
		if (!n) break;
		b *= b; be *= 2; { int t; b = frexp(b, &t); be += t; }
	}
	mant = r; ex = re;
}

__device__ __forceinline__ double ipow(double b, int n) { double r = 1.0; for (int i = 0; i < n; i++) r *= b; return r; }

// true value t*2^ex -> (v, sc) with value = v * 2^(512 sc), sc <= 0; sc == 0 ("live") iff |value| >= 2^-(LIVE_EXP+1)
__device__ __forceinline__ void init_scaled(double t, int ex, double &v, int &sc)
{
	if (t == 0.0) { v = 0.0; sc = 0; return; }
	int te; frexp(t, &te);
	int E = ex + te;
	if (E >= -LIVE_EXP) { v = ldexp(t, ex); sc = 0; }
	else { int k = (-LIVE_EXP - E + 511)/512; v = ldexp(t, ex + 512*k); sc = -k; }
}

// Once per window of <= 16 l: bring a scaled sequence back into range.  The test reads the exponent field
// with integer instructions (the FP64 pipe is the bottleneck of these kernels).  Within 16 steps a sequence
// grows by less than 2^100 (|a_l x| <= sqrt(2 lmax + 1)), so values stay far below overflow between tests,
// and a lane whose value passes 2^-LIVE_EXP inside a window joins the sums at the next window.
__device__ __forceinline__ void rescale(double &v, double &vp, int &sc)
{
	if (sc < 0 && (__double2hiint(v) & 0x7ff00000) >= ((1023 + 512 - LIVE_EXP) << 20)) { v *= SMALLV; vp *= SMALLV; sc++; }
}

// A ring contributes nothing for this m when m lies beyond the evanescent tail of the turning point
// m_t = (lmax+1/2) sin(theta): |n_l d^l_{m,s}| ~ exp(-0.94 delta^1.5/(sqrt(m) cos(theta))), delta = m - m_t;
// delta > 16 m^(1/3) puts the dropped values below exp(-60) ~ 1e-26 (checked against the oracle in
// tests/test_legendre_gpu.py at lmax up to 2000).  The threshold on sin(theta) is SeqConst::dead_sth.
struct SeqConst {      // per-m constants of the start values
	double pref; double sign_p, sign_q; int e, qc, qs, pc, ps;
	double dead_sth;    // rings with sin(theta) below this contribute nothing for this m
};

__device__ __forceinline__ SeqConst seq_const(int m, int s, int lmax, const double *pref)
{
	SeqConst c;
	c.pref = pref[m];
	c.sign_p = (m & 1) ? -1.0 : 1.0;
	c.sign_q = (m >= s && ((m - s) & 1)) ? -1.0 : 1.0;
	c.e = m > s ? m - s : 0;
	if (m >= s) { c.qc = 2*s; c.qs = 0; c.pc = 0; c.ps = 2*s; }
	else { c.qc = s + m; c.qs = s - m; c.pc = s - m; c.ps = s + m; }
	c.dead_sth = ((double)m - 16.0*cbrt((double)m) - s - 2)/(lmax + 0.5);
	return c;
}

// pairs are sorted by sin(theta), so the rings that are dead for this m form a prefix: index of the first live pair
__device__ __forceinline__ int first_live_pair(const PairInfo *pairs, int npair, double dead_sth)
{
	int lo = 0, hi = npair;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		PairInfo pi = pairs[mid];
		if (2.0*pi.sh*pi.ch < dead_sth) lo = mid + 1; else hi = mid;
	}
	return lo;
}

// start values of q = n d^{l0}_{m,+s}, p = (-1)^s n d^{l0}_{m,-s} for one ring
__device__ __forceinline__ bool init_pair(const PairInfo &pi, const SeqConst &c, int m, int s, int lmax,
	double &x, double &p, int &sp, double &q, int &sq)
{
	x = pi.x;
	double sth = 2.0*pi.sh*pi.ch;
	if (pi.rn < 0 || sth < c.dead_sth) { p = q = 0.0; sp = sq = 0; return false; }
	double mant; int ex;
	scaled_pow(sth, c.e, mant, ex);
	double base = c.pref*mant;
	init_scaled(c.sign_q*base*ipow(pi.ch, c.qc)*ipow(pi.sh, c.qs), ex, q, sq);
	init_scaled(c.sign_p*base*ipow(pi.ch, c.pc)*ipow(pi.sh, c.ps), ex, p, sp);
	return true;
}

// Warp reduction of the adjoint kernels.  Every lane holds NC*W partial sums v[slot*W + j] (NC component
// slots x W consecutive l); on return v[0] of lane L is the sum over all 32 lanes of one (component, j):
//   NC = 4: component (L >> 3) & 3, j = (L & 7) >> (3 - log2 W);   NC = 2: component (L >> 4) & 1, j = (L & 15) >> (4 - log2 W).
// Halving butterfly (each level sends one half of the values to the partner lane and keeps the other, about
// one 64-bit exchange per value in total).  The component levels need no register selects because the lanes
// store their slots permuted: slot s of lane L holds component s ^ ((L >> 3) & 3) (NC = 4) or s ^ (L >> 4)
// (NC = 2) -- the kernels arrange that by permuting each lane's ring inputs once per round.
template<int NC, int W> __device__ __forceinline__ void bfly_reduce(double (&v)[NC*W], int lane)
{
	static_assert((NC == 4 && W <= 8) || (NC == 2 && W <= 16), "window too long for the lane bits");
	if (NC == 4) {
		#pragma unroll
		for (int i = 0; i < 2*W; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i + 2*W], 16);
		#pragma unroll
		for (int i = 0; i < W; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i + W], 8);
	} else {
		#pragma unroll
		for (int i = 0; i < W; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i + W], 16);
	}
	int n = W;
	#pragma unroll
	for (int bit = (NC == 4 ? 4 : 8); bit >= 1; bit >>= 1) {
		if (n > 1) {
			n >>= 1;
			const bool up = (lane & bit) != 0;
			#pragma unroll
			for (int i = 0; i < W/2; i++) if (i < n) {
				double send = up ? v[i] : v[i + n];
				double keep = up ? v[i + n] : v[i];
				v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
			}
		} else {
			v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
		}
	}
}
// index (in the order [j][component]) of the element lane L holds after bfly_reduce
template<int NC, int W> __device__ __forceinline__ int bfly_element(int lane)
{
	constexpr int LW = (W == 1 ? 0 : W == 2 ? 1 : W == 4 ? 2 : W == 8 ? 3 : 4);
	constexpr int S4 = LW <= 3 ? 3 - LW : 0, S2 = 4 - LW;
	if (NC == 4) return NC*((lane & 7) >> S4) + ((lane >> 3) & 3);
	return NC*((lane & 15) >> S2) + ((lane >> 4) & 1);
}

// cp_async8 / cp_async16 (common.cuh): the next l tile travels global -> shared while the FP64 pipe works on the
// current one, without holding registers and without a load-to-use stall in the instruction stream

// CTAs of one warp (NW = 1) are fully warp-synchronous: no block barriers at all
template<int NW> __device__ __forceinline__ void cta_sync() { if (NW == 1) __syncwarp(); else __syncthreads(); }
template<int NW> __device__ __forceinline__ bool cta_or(bool v) { return NW == 1 ? __any_sync(0xffffffffu, v) : (__syncthreads_or(v) != 0); }

// smallest of the warps' (warp-uniform) values; slot: one shared int per CTA
template<int NW> __device__ __forceinline__ int cta_min(int v, int *slot)
{
	if (NW == 1) return v;
	if (threadIdx.x == 0) *slot = LEG_NEVER;
	__syncthreads();
	if ((threadIdx.x & 31) == 0) atomicMin(slot, v);
	__syncthreads();
	int r = *slot;
	__syncthreads();
	return r;
}

// The CTA of order m has written its alm row: count it in its range of m; whoever completes a range publishes the
// call's epoch in the range's flag (mapped host memory), which the host thread of a host-memory call polls to start that
// range's device -> host copy while the kernel is still working on the other ranges.
template<int NW> __device__ __forceinline__ void signal_done_body(const LegArgs &A)
{
	cta_sync<NW>();
	if (threadIdx.x == 0 && A.sig.count) {
		const int m = A.m0 + blockIdx.x;
		int r = 0;
		#pragma unroll 1
		while (r + 2 < A.sig.ncut && m >= A.sig.cut[r + 1]) r++;
		// release at GPU scope per CTA (a system-scope fence here costs the whole kernel 6 %: 107 -> 114 ms at C3); the one
		// CTA that completes a range acquires the others' rows through the counter and fences at system scope
		unsigned old;
		asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(&A.sig.count[r]), "r"(1u) : "memory");
		const unsigned done = old + 1u;
		if (done == (unsigned)(A.sig.cut[r + 1] - A.sig.cut[r])) {
			__threadfence_system();
			*A.sig.flag_of(r) = A.sig.epoch;
		}
	}
}

// SIG = 1: inlined, SIG = 2: behind a call (the kernels' register allocation, and with it the bank conflicts of their DFMA
// operands, differs between the two and from SIG = 0: measured on B200 at C3, 107 ms without, 114 ms with SIG = 1)
template<int NW> __device__ __noinline__ void signal_done_call(const LegArgs &A) { signal_done_body<NW>(A); }
template<int NW, int SIG> __device__ __forceinline__ void signal_done(const LegArgs &A)
{
	if (SIG == 1) signal_done_body<NW>(A); else signal_done_call<NW>(A);
}

// The mirror image for the synthesis kernels of a host-memory call: the alm of the first group arrive range by range
// (LegSignal::flag lives in device memory here, written by the copy stream after each range); a CTA waits for its range.
template<int NW> __device__ __forceinline__ void gate_wait(const LegArgs &A)
{
	if (threadIdx.x == 0) {
		const int m = A.m0 + blockIdx.x;
		int r = 0;
		#pragma unroll 1
		while (r + 2 < A.sig.ncut && m >= A.sig.cut[r + 1]) r++;
		const volatile int *f = A.sig.flag_of(r);
		while (*f != A.sig.epoch) __nanosleep(500);
		__threadfence();
	}
	cta_sync<NW>();
}

// ------------------------------------------------------------------------------------ spin 0

struct Tile0 { double ar, ai, a, pad; };            // alm*alpha (re, im), recurrence a_l

// one window of 8 l values, MODE 0: recurrence only (nobody live), 1: masked + rescale, 2: plain
template<int MODE, int R> __device__ __forceinline__ void synth0_window(const Tile0 *T,
	const double (&x)[R], double (&g)[R], double (&gp)[R], int (&sc)[R], double (&acc)[R][2][2])
{
	#pragma unroll
	for (int j = 0; j < 8; j++) {
		const Tile0 t = T[j];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			if (MODE != 0) {
				double gv = (MODE == 1) ? (sc[r] == 0 ? g[r] : 0.0) : g[r];
				acc[r][j & 1][0] = fma(gv, t.ar, acc[r][j & 1][0]);
				acc[r][j & 1][1] = fma(gv, t.ai, acc[r][j & 1][1]);
			}
			double ng = fma(t.a, x[r]*g[r], -gp[r]);
			gp[r] = g[r]; g[r] = ng;
		}
	}
	if (MODE != 2) {
		#pragma unroll
		for (int r = 0; r < R; r++) rescale(g[r], gp[r], sc[r]);
	}
}

template<int R, int NW, int MINB, int TL, int GATE> __global__ void __launch_bounds__(NW*32, MINB) k_synth0(LegArgs A)
{
	if (GATE) gate_wait<NW>(A);
	__shared__ __align__(16) Tile0 tiles[2][TL];
	__shared__ __align__(16) double2 raw_alm[TL];       // cp.async staging: alm, (alpha, a)
	__shared__ __align__(16) double raw_al[TL], raw_a[TL];
	__shared__ int wslot;
	const int m = A.m0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int lmax = A.lmax, l0 = m;
	const bool tab = A.st_w != nullptr && R*32 == LEG_GROUP;
	const int nl = lmax - l0 + 1, ntile = (nl + TL - 1)/TL;
	const double *ta = A.ta + A.toff[m], *tal = A.talpha + A.toff[m];
	const double2 *alm = A.alm0 + A.mstart[m];
	const SeqConst sc0 = seq_const(m, 0, lmax, A.pref);
	const int nchunk = A.npair_pad/(32*R);
	double2 *leg = A.leg0 + (int64_t)m*A.leg_mstride;
	// rounds made of dead rings only write zeros
	const int rlive = first_live_pair(A.pairs, A.npair, sc0.dead_sth)/(32*R*NW);
	const int round0 = max(rlive, A.pair_lo/(32*R*NW)), round1 = min(nchunk, A.pair_hi/(32*R));
	for (int i = A.pair_lo + tid; i < min(rlive*32*R*NW, A.pair_hi); i += NW*32) {
		PairInfo pi = A.pairs[i];
		if (pi.rn >= 0) leg[pi.rn] = make_double2(0, 0);
		if (pi.rs >= 0) leg[pi.rs] = make_double2(0, 0);
	}

	for (int round = round0; round*NW < round1; round++) {
		const int chunk = round*NW + warp;
		double x[R], g[R], gp[R], acc[R][2][2]; int sc[R], rn[R], rs[R];
		bool anyuse = false, use[R];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			PairInfo pi; pi.rn = -1; pi.rs = -1; pi.x = 0; pi.sh = 0; pi.ch = 1;
			if (chunk < nchunk) pi = A.pairs[(chunk*R + r)*32 + lane];
			double dummy; int dsc;
			if (!tab) { use[r] = init_pair(pi, sc0, m, 0, lmax, x[r], dummy, dsc, g[r], sc[r]); gp[r] = 0; }
			else {
				x[r] = pi.x;
				use[r] = pi.rn >= 0 && 2.0*pi.sh*pi.ch >= sc0.dead_sth;
				g[r] = gp[r] = 0; sc[r] = 0;
				if (chunk < nchunk) { int64_t k = (int64_t)m*A.npair_pad + (chunk*R + r)*32 + lane; g[r] = A.st_p[k]; gp[r] = A.st_pp[k]; sc[r] = A.st_sp[k]; }
			}
			rn[r] = pi.rn; rs[r] = pi.rs;
			acc[r][0][0] = acc[r][0][1] = acc[r][1][0] = acc[r][1][1] = 0;
			anyuse |= use[r];
		}
		// first window of 8 l in which this warp has work (start table), CTA-wide minimum for the shared tiles
		int wc = 0;
		if (tab) { wc = chunk < nchunk ? A.st_w[m*A.ngroup + chunk] : LEG_NEVER; if (wc == LEG_NEVER) anyuse = false; }
		const int wc_cta = tab ? cta_min<NW>(wc, &wslot) : 0;
		// a round in which no ring of the CTA can contribute only has to write zeros
		const bool wuse = __any_sync(0xffffffffu, anyuse);
		int phase = 0;
		if (cta_or<NW>(anyuse)) {
			// each of the first TL threads fetches one l of the next tile (issue), later scales it into the tile (finish)
			auto issue = [&](int tile) {
				#pragma unroll
				for (int t = tid; t < TL; t += NW*32) {      // one pass when TL <= NW*32
					int i = tile*TL + t;
					if (i < nl) {
						cp_async16(&raw_alm[t], &alm[(int64_t)(l0 + i)*A.lstride]);
						cp_async8(&raw_al[t], &tal[i]); cp_async8(&raw_a[t], &ta[i]);
					}
				}
				cp_async_commit();
			};
			auto finish = [&](int tile, int buf) {
				cp_async_wait_all();
				#pragma unroll
				for (int t = tid; t < TL; t += NW*32) {
					int i = tile*TL + t;
					Tile0 e; e.ar = e.ai = e.a = e.pad = 0;
					if (i < nl) { double al = raw_al[t]; double2 v = raw_alm[t]; e.ar = v.x*al; e.ai = v.y*al; e.a = raw_a[t]; }
					tiles[buf][t] = e;
				}
			};
			const int tile0 = min(wc_cta*8/TL, ntile - 1);
			issue(tile0); finish(tile0, tile0 & 1);
			cta_sync<NW>();
			for (int tile = tile0; tile < ntile; tile++) {
				const int buf = tile & 1;
				if (tile + 1 < ntile) issue(tile + 1);
				const int nwin = (min(TL, nl - tile*TL) + 7) >> 3;
				for (int w = 0; wuse && w < nwin; w++) {
					if (tile*(TL/8) + w < wc) continue;
					if (phase < 2) {      // liveness only grows: once every lane is live no more votes are needed
						bool mylive = true, anylive = false;
						#pragma unroll
						for (int r = 0; r < R; r++) { mylive &= (sc[r] == 0); anylive |= (sc[r] == 0 && use[r]); }
						phase = __all_sync(0xffffffffu, mylive) ? 2 : __any_sync(0xffffffffu, anylive) ? 1 : 0;
					}
					const Tile0 *T = &tiles[buf][w*8];
					if (phase == 2) synth0_window<2, R>(T, x, g, gp, sc, acc);
					else if (phase == 1) synth0_window<1, R>(T, x, g, gp, sc, acc);
					else synth0_window<0, R>(T, x, g, gp, sc, acc);
				}
				if (tile + 1 < ntile) finish(tile + 1, buf ^ 1);
				cta_sync<NW>();
			}
		}
		// set 0 holds the l = l0, l0+2, ... terms (parity sigma0 = (-1)^(l0+m) = +1 for spin 0)
		#pragma unroll
		for (int r = 0; r < R; r++) {
			if (rn[r] >= 0) leg[rn[r]] = make_double2(acc[r][0][0] + acc[r][1][0], acc[r][0][1] + acc[r][1][1]);
			if (rs[r] >= 0) leg[rs[r]] = make_double2(acc[r][0][0] - acc[r][1][0], acc[r][0][1] - acc[r][1][1]);
		}
	}
}

// adjoint, spin 0: a window of W l values -> NV = 2W partial sums v[slot*W + j] (slots: re, im) reduced over
// the warp by bfly_reduce<2, W>
template<int MODE, int R, int W> __device__ __forceinline__ void adj0_window(const double *Ta,
	const double (&x)[R], double (&g)[R], double (&gp)[R], int (&sc)[R],
	const double (&in)[R][2][2], double (&v)[2*W])
{
	#pragma unroll
	for (int j = 0; j < W; j++) {
		const double a = Ta[j];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			if (MODE != 0) {
				double gv = (MODE == 1) ? (sc[r] == 0 ? g[r] : 0.0) : g[r];
				v[j]     = fma(gv, in[r][j & 1][0], v[j]);
				v[W + j] = fma(gv, in[r][j & 1][1], v[W + j]);
			}
			double ng = fma(a, x[r]*g[r], -gp[r]);
			gp[r] = g[r]; g[r] = ng;
		}
	}
	if (MODE != 2) {
		#pragma unroll
		for (int r = 0; r < R; r++) rescale(g[r], gp[r], sc[r]);
	}
}

template<int R, int NW, int MINB, int TL, int W, int SIG> __global__ void __launch_bounds__(NW*32, MINB) k_adj0(LegArgs A)
{
	constexpr int NV = 2*W, NOUT = 2*TL;
	constexpr int NT = NW*32, NH = (NOUT + NT - 1)/NT;
	__shared__ double tiles[2][TL];
	__shared__ __align__(16) double red[2][NW][NOUT];
	__shared__ __align__(16) double olds[NOUT], alvs[NOUT];      // running sums and alpha_l of the tile's outputs (cp.async)
	const int m = A.m0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int lmax = A.lmax, l0 = m;
	const int nl = lmax - l0 + 1, ntile = (nl + TL - 1)/TL;
	const double *ta = A.ta + A.toff[m], *tal = A.talpha + A.toff[m];
	double *almr = (double*)(A.alm0 + A.mstart[m]);
	const SeqConst sc0 = seq_const(m, 0, lmax, A.pref);
	const int nchunk = A.npair_pad/(32*R);
	const double2 *leg = A.leg0 + (int64_t)m*A.leg_mstride;
	bool first = true;
	const int round0 = first_live_pair(A.pairs, A.npair, sc0.dead_sth)/(32*R*NW);
	// start table (single-warp CTAs): a chunk spans NG groups of LEG_GROUP pairs; each group's rings are injected
	// with their recorded state at the group's first live window, the row is zeroed once and every round adds to it
	constexpr int NG = (R*32)/LEG_GROUP;
	const bool tab = A.st_w != nullptr && NW == 1 && NG >= 1 && NG*LEG_GROUP == R*32;
	if (tab) {
		for (int i = tid; i < nl; i += NW*32) *(double2*)&almr[2*(int64_t)(l0 + i)*A.lstride] = make_double2(0, 0);
		__syncwarp();
		first = false;
	}

	for (int round = round0; round*NW < nchunk; round++) {
		const int chunk = round*NW + warp;
		double x[R], g[R], gp[R], in[R][2][2]; int sc[R];
		bool anyuse = false, use[R];
		int wg[NG > 0 ? NG : 1], wc = 0;
		if (tab) {
			wc = LEG_NEVER;
			#pragma unroll
			for (int k = 0; k < NG; k++) { wg[k] = A.st_w[m*A.ngroup + chunk*NG + k]; wc = min(wc, wg[k]); }
		}
		#pragma unroll
		for (int r = 0; r < R; r++) {
			PairInfo pi; pi.rn = -1; pi.rs = -1; pi.x = 0; pi.sh = 0; pi.ch = 1;
			if (chunk < nchunk) pi = A.pairs[(chunk*R + r)*32 + lane];
			double dummy; int dsc;
			if (!tab) { use[r] = init_pair(pi, sc0, m, 0, lmax, x[r], dummy, dsc, g[r], sc[r]); gp[r] = 0; }
			else {      // rings wait with zero state (contributing exact zeros) until their group is injected
				x[r] = pi.x; g[r] = gp[r] = 0; sc[r] = 0;
				use[r] = pi.rn >= 0 && 2.0*pi.sh*pi.ch >= sc0.dead_sth && wg[(r*32)/LEG_GROUP] != LEG_NEVER;
			}
			double2 gn = make_double2(0, 0), gs = make_double2(0, 0);
			if (use[r]) { gn = leg[pi.rn]; if (pi.rs >= 0) gs = leg[pi.rs]; }
			if (lane & 16) { gn = make_double2(gn.y, gn.x); gs = make_double2(gs.y, gs.x); }    // slot permutation of bfly_reduce
			in[r][0][0] = gn.x + gs.x; in[r][0][1] = gn.y + gs.y;    // l - l0 even
			in[r][1][0] = gn.x - gs.x; in[r][1][1] = gn.y - gs.y;    // l - l0 odd
			anyuse |= use[r];
		}
		if (tab && wc == LEG_NEVER) anyuse = false;
		const bool wuse = __any_sync(0xffffffffu, anyuse);
		int phase = 0;
		if (!cta_or<NW>(anyuse)) continue;
		// recurrence coefficients of a tile travel global -> shared with cp.async, one tile ahead
		auto issue = [&](int tile, int buf) {
			#pragma unroll
			for (int t = tid; t < TL; t += NT) {
				int i = tile*TL + t;
				if (i < nl) cp_async8(&tiles[buf][t], &ta[i]); else tiles[buf][t] = 0.0;
			}
			cp_async_commit();
		};
		const int tile0 = tab ? min(wc*8/TL, ntile - 1) : 0;
		issue(tile0, tile0 & 1); cp_async_wait_all();
		cta_sync<NW>();
		for (int tile = tile0; tile < ntile; tile++) {
			const int buf = tile & 1;
			if (tile + 1 < ntile) issue(tile + 1, buf ^ 1);
			const int nwin = (min(TL, nl - tile*TL) + W - 1)/W;
			// output threads fetch the running sum early (cp.async) so the read-modify-write latency hides behind the tile
			if constexpr (NW == 1) {
				// one l per lane: a_lm as one 16-byte access
				#pragma unroll
				for (int h = 0; h < TL/32; h++) {
					int o = lane + 32*h, i = tile*TL + o;
					if (i < nl) {
						cp_async8(&alvs[o], &tal[i]);
						if (!first) cp_async16(&olds[2*o], &almr[2*(int64_t)(l0 + i)*A.lstride]);
					}
				}
			} else {
				#pragma unroll
				for (int h = 0; h < NH; h++) {
					int e = tid + h*NT, i = tile*TL + (e >> 1);
					if (e < NOUT && i < nl) {
						cp_async8(&alvs[e], &tal[i]);
						if (!first) cp_async8(&olds[e], &almr[2*(int64_t)(l0 + i)*A.lstride + (e & 1)]);
					}
				}
			}
			cp_async_commit();
			// (two instances of the window loop: only a tile in which a ring group starts pays for the start-table check; left
			// in the common loop it compiles to 24 predicated loads per window, 8 % of the kernel's issue slots)
			auto windows = [&](auto CHECK) {
			#pragma unroll 1
			for (int w = 0; w < TL/W; w++) {
				double tot = 0;
				if (wuse && w < nwin && tile*TL + w*W >= wc*8) {
					if (decltype(CHECK)::value) {
						#pragma unroll
						for (int k = 0; k < NG; k++) if (tile*TL + w*W == wg[k]*8) {      // inject group k (warp-uniform)
							#pragma unroll
							for (int r = k*(LEG_GROUP/32); r < (k + 1)*(LEG_GROUP/32); r++) {
								int64_t kk = (int64_t)m*A.npair_pad + (chunk*R + r)*32 + lane;
								g[r] = A.st_p[kk]; gp[r] = A.st_pp[kk]; sc[r] = A.st_sp[kk];
							}
							phase = 0;
						}
					}
					double v[NV];
					#pragma unroll
					for (int i = 0; i < NV; i++) v[i] = 0;
					if (phase < 2) {
						bool mylive = true, anylive = false;
						#pragma unroll
						for (int r = 0; r < R; r++) { mylive &= (sc[r] == 0); anylive |= (sc[r] == 0 && use[r]); }
						phase = __all_sync(0xffffffffu, mylive) ? 2 : __any_sync(0xffffffffu, anylive) ? 1 : 0;
					}
					const double *T = &tiles[buf][w*W];
					if (phase == 2) adj0_window<2, R, W>(T, x, g, gp, sc, in, v);
					else if (phase == 1) adj0_window<1, R, W>(T, x, g, gp, sc, in, v);
					else adj0_window<0, R, W>(T, x, g, gp, sc, in, v);
					if (phase != 0) { bfly_reduce<2, W>(v, lane); tot = v[0]; }
				}
				// element 2 (l - l_window) + re/im; lanes holding the same element write the same value
				red[buf][warp][w*NV + bfly_element<2, W>(lane)] = tot;
			}
			};
			bool inj_tile = false;
			if (tab) {
				#pragma unroll
				for (int k = 0; k < NG; k++) inj_tile |= (wg[k] != LEG_NEVER && wg[k]*8 >= tile*TL && wg[k]*8 < (tile + 1)*TL);
			}
			if (inj_tile) windows(std::true_type()); else windows(std::false_type());
			cp_async_wait_all();
			cta_sync<NW>();
			if constexpr (NW == 1) {
				#pragma unroll
				for (int h = 0; h < TL/32; h++) {
					int o = lane + 32*h, i = tile*TL + o;
					if (i < nl) {
						double2 c = *(const double2*)&red[buf][0][2*o];
						double2 old = first ? make_double2(0, 0) : *(const double2*)&olds[2*o];
						double al = alvs[o];
						*(double2*)&almr[2*(int64_t)(l0 + i)*A.lstride] = make_double2(fma(c.x, al, old.x), fma(c.y, al, old.y));
					}
				}
			} else {
				#pragma unroll
				for (int h = 0; h < NH; h++) {
					int e = tid + h*NT, i = tile*TL + (e >> 1);
					if (e < NOUT && i < nl) {
						double s = 0;
						#pragma unroll
						for (int w = 0; w < NW; w++) s += red[buf][w][e];
						almr[2*(int64_t)(l0 + i)*A.lstride + (e & 1)] = (first ? 0.0 : olds[e]) + s*alvs[e];
					}
				}
			}
		}
		first = false;
	}
	if (first) for (int i = tid; i < nl; i += NW*32) { double *o = almr + 2*(int64_t)(l0 + i)*A.lstride; o[0] = 0; o[1] = 0; }
	if (SIG) signal_done<NW, SIG>(A);
}

// ------------------------------------------------------------------------------------ spin > 0

struct Tile2 { double a, b, apr, api, amr, ami; };    // recurrence a,b; A+ = -(E+iB)alpha/2, A- = -(E-iB)alpha/2

// north sums need sum_l p A+ and sum_l q A-, south sums sum_l sigma_l q A+ and sum_l sigma_l p A-: eight
// accumulators per ring pair, one DFMA each per l; sigma_l alternates and is folded into the operand sign
template<int MODE, int R> __device__ __forceinline__ void synth2_window(const Tile2 *T,
	const double (&x)[R], double (&p)[R], double (&pp)[R], double (&q)[R], double (&qp)[R],
	int (&sp)[R], int (&sq)[R], double (&acc)[R][8])
{
	#pragma unroll
	for (int j = 0; j < 8; j++) {
		const Tile2 t = T[j];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			{
				// masked windows (MODE 1) predicate the accumulations of a ring that is not live yet instead of selecting a zero
				// operand: no FSEL per operand (those windows are 6 % of the instructions and were 14 % of the samples)
				const double pv = p[r], qv = q[r];
				const double ps = (j & 1) ? -pv : pv, qs = (j & 1) ? -qv : qv;
				double (&c)[8] = acc[r];
				if (MODE == 2 || (MODE == 1 && sp[r] == 0)) {
					c[0] = fma(pv, t.apr, c[0]); c[1] = fma(pv, t.api, c[1]);
					c[2] = fma(ps, t.amr, c[2]); c[3] = fma(ps, t.ami, c[3]);
				}
				if (MODE == 2 || (MODE == 1 && sq[r] == 0)) {
					c[4] = fma(qs, t.apr, c[4]); c[5] = fma(qs, t.api, c[5]);
					c[6] = fma(qv, t.amr, c[6]); c[7] = fma(qv, t.ami, c[7]);
				}
			}
			double np = fma(fma(t.a, x[r],  t.b), p[r], -pp[r]);    // n = -s
			double nq = fma(fma(t.a, x[r], -t.b), q[r], -qp[r]);    // n = +s
			pp[r] = p[r]; p[r] = np; qp[r] = q[r]; q[r] = nq;
		}
	}
	if (MODE != 2) {
		#pragma unroll
		for (int r = 0; r < R; r++) { rescale(p[r], pp[r], sp[r]); rescale(q[r], qp[r], sq[r]); }
	}
}

template<int R, int NW, int MINB, int TL, int GATE> __global__ void __launch_bounds__(NW*32, MINB) k_synth2(LegArgs A)
{
	if (GATE) gate_wait<NW>(A);
	__shared__ __align__(16) Tile2 tiles[2][TL];
	__shared__ __align__(16) double2 raw_e[TL], raw_b[TL];       // cp.async staging: E, B, (a, b, alpha)
	__shared__ __align__(16) double raw_ta[TL], raw_tb[TL], raw_al[TL];
	__shared__ int wslot;
	const int m = A.m0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int lmax = A.lmax, s = A.spin, l0 = m > s ? m : s;
	const bool tab = A.st_w != nullptr && R*32 == LEG_GROUP;
	double2 *legq = A.leg0 + (int64_t)m*A.leg_mstride, *legu = A.leg1 + (int64_t)m*A.leg_mstride;
	const int nchunk = A.npair_pad/(32*R);
	if (l0 > lmax) {      // nothing to sum: zero this m column
		for (int i = A.pair_lo + tid; i < min(A.npair_pad, A.pair_hi); i += NW*32) {
			PairInfo pi = A.pairs[i];
			if (pi.rn >= 0) legq[pi.rn] = legu[pi.rn] = make_double2(0, 0);
			if (pi.rs >= 0) legq[pi.rs] = legu[pi.rs] = make_double2(0, 0);
		}
		return;
	}
	const int nl = lmax - l0 + 1, ntile = (nl + TL - 1)/TL;
	const double *ta = A.ta + A.toff[m], *tb = A.tb + A.toff[m], *tal = A.talpha + A.toff[m];
	const SeqConst sc0 = seq_const(m, s, lmax, A.pref);
	const double sigma0 = ((l0 + m + s) & 1) ? -1.0 : 1.0;
	const int64_t ms = A.mstart[m];
	const int rlive = first_live_pair(A.pairs, A.npair, sc0.dead_sth)/(32*R*NW);
	const int round0 = max(rlive, A.pair_lo/(32*R*NW)), round1 = min(nchunk, A.pair_hi/(32*R));
	for (int i = A.pair_lo + tid; i < min(rlive*32*R*NW, A.pair_hi); i += NW*32) {
		PairInfo pi = A.pairs[i];
		if (pi.rn >= 0) legq[pi.rn] = legu[pi.rn] = make_double2(0, 0);
		if (pi.rs >= 0) legq[pi.rs] = legu[pi.rs] = make_double2(0, 0);
	}

	for (int round = round0; round*NW < round1; round++) {
		const int chunk = round*NW + warp;
		double x[R], p[R], pp[R], q[R], qp[R], acc[R][8]; int sp[R], sq[R], rn[R], rs[R];
		bool anyuse = false, use[R];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			PairInfo pi; pi.rn = -1; pi.rs = -1; pi.x = 0; pi.sh = 0; pi.ch = 1;
			if (chunk < nchunk) pi = A.pairs[(chunk*R + r)*32 + lane];
			if (!tab) { use[r] = init_pair(pi, sc0, m, s, lmax, x[r], p[r], sp[r], q[r], sq[r]); pp[r] = qp[r] = 0; }
			else {
				x[r] = pi.x;
				use[r] = pi.rn >= 0 && 2.0*pi.sh*pi.ch >= sc0.dead_sth;
				p[r] = pp[r] = q[r] = qp[r] = 0; sp[r] = sq[r] = 0;
				if (chunk < nchunk) {
					int64_t k = (int64_t)m*A.npair_pad + (chunk*R + r)*32 + lane;
					p[r] = A.st_p[k]; pp[r] = A.st_pp[k]; q[r] = A.st_q[k]; qp[r] = A.st_qp[k]; sp[r] = A.st_sp[k]; sq[r] = A.st_sq[k];
				}
			}
			rn[r] = pi.rn; rs[r] = pi.rs;
			#pragma unroll
			for (int k = 0; k < 8; k++) acc[r][k] = 0;
			anyuse |= use[r];
		}
		int wc = 0;
		if (tab) { wc = chunk < nchunk ? A.st_w[m*A.ngroup + chunk] : LEG_NEVER; if (wc == LEG_NEVER) anyuse = false; }
		const int wc_cta = tab ? cta_min<NW>(wc, &wslot) : 0;
		const bool wuse = __any_sync(0xffffffffu, anyuse);
		int phase = 0;
		if (cta_or<NW>(anyuse)) {
			auto issue = [&](int tile) {
				#pragma unroll
				for (int t = tid; t < TL; t += NW*32) {      // one pass when TL <= NW*32
					int i = tile*TL + t;
					if (i < nl) {
						int64_t idx = ms + (int64_t)(l0 + i)*A.lstride;
						cp_async16(&raw_e[t], &A.alm0[idx]);
						if (!A.deriv1) cp_async16(&raw_b[t], &A.alm1[idx]);
						cp_async8(&raw_ta[t], &ta[i]); cp_async8(&raw_tb[t], &tb[i]); cp_async8(&raw_al[t], &tal[i]);
					}
				}
				cp_async_commit();
			};
			auto finish = [&](int tile, int buf) {
				cp_async_wait_all();
				#pragma unroll
				for (int t = tid; t < TL; t += NW*32) {
					int i = tile*TL + t;
					Tile2 e; e.a = e.b = e.apr = e.api = e.amr = e.ami = 0;
					if (i < nl) {
						int l = l0 + i;
						double2 E = raw_e[t], B = make_double2(0, 0);
						if (A.deriv1) { double f = sqrt((double)l*(l + 1.0)); E.x *= f; E.y *= f; }
						else B = raw_b[t];
						double h = -0.5*raw_al[t];
						e.a = raw_ta[t]; e.b = raw_tb[t];
						e.apr = h*(E.x - B.y); e.api = h*(E.y + B.x);
						e.amr = h*(E.x + B.y); e.ami = h*(E.y - B.x);
					}
					tiles[buf][t] = e;
				}
			};
			const int tile0 = min(wc_cta*8/TL, ntile - 1);
			issue(tile0); finish(tile0, tile0 & 1);
			cta_sync<NW>();
			for (int tile = tile0; tile < ntile; tile++) {
				const int buf = tile & 1;
				if (tile + 1 < ntile) issue(tile + 1);
				const int nwin = (min(TL, nl - tile*TL) + 7) >> 3;
				for (int w = 0; wuse && w < nwin; w++) {
					if (tile*(TL/8) + w < wc) continue;
					if (phase < 2) {
						bool mylive = true, anylive = false;
						#pragma unroll
						for (int r = 0; r < R; r++) {
							mylive &= (sp[r] == 0) & (sq[r] == 0);
							anylive |= use[r] & ((sp[r] == 0) | (sq[r] == 0));
						}
						phase = __all_sync(0xffffffffu, mylive) ? 2 : __any_sync(0xffffffffu, anylive) ? 1 : 0;
					}
					const Tile2 *T = &tiles[buf][w*8];
					if (phase == 2) synth2_window<2, R>(T, x, p, pp, q, qp, sp, sq, acc);
					else if (phase == 1) synth2_window<1, R>(T, x, p, pp, q, qp, sp, sq, acc);
					else synth2_window<0, R>(T, x, p, pp, q, qp, sp, sq, acc);
				}
				if (tile + 1 < ntile) finish(tile + 1, buf ^ 1);
				cta_sync<NW>();
			}
		}
		// Sp = sum p A+, Sq = sum q A-;  south: Sp' = sum sigma_l q A+, Sq' = sum sigma_l p A-
		// Q = Sp + Sq, U = i (Sq - Sp)
		#pragma unroll
		for (int r = 0; r < R; r++) {
			const double (&c)[8] = acc[r];
			if (rn[r] >= 0) {
				double spr = c[0], spi = c[1], sqr = c[6], sqi = c[7];
				legq[rn[r]] = make_double2(spr + sqr, spi + sqi);
				legu[rn[r]] = make_double2(spi - sqi, sqr - spr);
			}
			if (rs[r] >= 0) {
				double spr = sigma0*c[4], spi = sigma0*c[5], sqr = sigma0*c[2], sqi = sigma0*c[3];
				legq[rs[r]] = make_double2(spr + sqr, spi + sqi);
				legu[rs[r]] = make_double2(spi - sqi, sqr - spr);
			}
		}
	}
}

// adjoint, spin > 0: a window of W l values -> NV = 4W partial sums v[slot*W + j], slots {A+re, A+im, A-re, A-im}:
//   A+_l = sum_pairs p Z+N + sigma_l q Z+S,  A-_l = sum_pairs q Z-N + sigma_l p Z-S,  Z+- = Q +- iU
// zin[r][0..3] = Z+N (re,im), Z-N (re,im); zin[r][4..7] = sigma0 * (Z+S, Z-S).
// Slot permutation for bfly_reduce<4, W>: lanes with bit 3 set swap re <-> im of every Z; lanes with bit 4 set
// swap the roles of the two sequences (their "p" registers carry q and vice versa, Z+ <-> Z-), which only
// flips the sign of b_l in the recurrence: the tile stores (a, b, a, -b) and such lanes read the second pair.
struct TileAB { double a, b, a2, nb; };

template<int MODE, int R, int W> __device__ __forceinline__ void adj2_window(const double2 *Tab,
	const double (&x)[R], double (&p)[R], double (&pp)[R], double (&q)[R], double (&qp)[R],
	int (&sp)[R], int (&sq)[R], const double (&zin)[R][8], double (&v)[4*W])
{
	#pragma unroll
	for (int j = 0; j < W; j++) {
		const double2 ab = Tab[2*j];      // (a, b) or (a, -b), chosen by the caller's pointer offset
		#pragma unroll
		for (int r = 0; r < R; r++) {
			if (MODE != 0) {
				double pv = (MODE == 1) ? (sp[r] == 0 ? p[r] : 0.0) : p[r];
				double qv = (MODE == 1) ? (sq[r] == 0 ? q[r] : 0.0) : q[r];
				// sigma_l alternates: fold the sign into the south products
				double ps = (j & 1) ? -pv : pv, qs = (j & 1) ? -qv : qv;
				v[j]       = fma(pv, zin[r][0], fma(qs, zin[r][4], v[j]));
				v[W + j]   = fma(pv, zin[r][1], fma(qs, zin[r][5], v[W + j]));
				v[2*W + j] = fma(qv, zin[r][2], fma(ps, zin[r][6], v[2*W + j]));
				v[3*W + j] = fma(qv, zin[r][3], fma(ps, zin[r][7], v[3*W + j]));
			}
			double np = fma(fma(ab.x, x[r],  ab.y), p[r], -pp[r]);
			double nq = fma(fma(ab.x, x[r], -ab.y), q[r], -qp[r]);
			pp[r] = p[r]; p[r] = np; qp[r] = q[r]; q[r] = nq;
		}
	}
	if (MODE != 2) {
		#pragma unroll
		for (int r = 0; r < R; r++) { rescale(p[r], pp[r], sp[r]); rescale(q[r], qp[r], sq[r]); }
	}
}

template<int R, int NW, int MINB, int TL, int W, int SIG> __global__ void __launch_bounds__(NW*32, MINB) k_adj2(const __grid_constant__ LegArgs A)
{
	constexpr int NV = 4*W, NOUT = 4*TL;
	constexpr int NT = NW*32, NH = (NOUT + NT - 1)/NT;
	__shared__ __align__(16) TileAB tiles[2][TL];
	__shared__ __align__(16) double red[2][NW][NOUT];
	__shared__ __align__(16) double olds[NOUT], alvs[NOUT];      // running sums and alpha_l of the tile's outputs (cp.async)
	const int m = A.m0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int lmax = A.lmax, s = A.spin, l0 = m > s ? m : s;
	double *alme = (double*)A.alm0, *almb = (double*)A.alm1;
	const int64_t ms = A.mstart[m];
	// l < spin entries of the triangle are zero by definition
	for (int l = m + tid; l < l0 && l <= lmax; l += NW*32) {
		int64_t idx = 2*(ms + (int64_t)l*A.lstride);
		alme[idx] = alme[idx + 1] = 0;
		if (!A.deriv1) almb[idx] = almb[idx + 1] = 0;
	}
	if (l0 > lmax) { if (SIG) signal_done<NW, SIG>(A); return; }
	const int nl = lmax - l0 + 1, ntile = (nl + TL - 1)/TL;
	const double *ta = A.ta + A.toff[m], *tb = A.tb + A.toff[m], *tal = A.talpha + A.toff[m];
	const SeqConst sc0 = seq_const(m, s, lmax, A.pref);
	const double sigma0 = ((l0 + m + s) & 1) ? -1.0 : 1.0;
	const double2 *legq = A.leg0 + (int64_t)m*A.leg_mstride, *legu = A.leg1 + (int64_t)m*A.leg_mstride;
	const int nchunk = A.npair_pad/(32*R);
	bool first = true;
	const int round0 = first_live_pair(A.pairs, A.npair, sc0.dead_sth)/(32*R*NW);
	// with a start table the rounds begin at different l: the row is zeroed once and every round adds to it
	const bool tab = A.st_w != nullptr && R*32 == LEG_GROUP && NW == 1;
	if (tab) {
		for (int i = tid; i < nl; i += NW*32) {
			int64_t idx = ms + (int64_t)(l0 + i)*A.lstride;
			A.alm0[idx] = make_double2(0, 0);
			if (!A.deriv1) A.alm1[idx] = make_double2(0, 0);
		}
		__syncwarp();
		first = false;
	}

	for (int round = round0; round*NW < nchunk; round++) {
		const int chunk = round*NW + warp;
		double x[R], p[R], pp[R], q[R], qp[R], zin[R][8]; int sp[R], sq[R];
		bool anyuse = false, use[R];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			PairInfo pi; pi.rn = -1; pi.rs = -1; pi.x = 0; pi.sh = 0; pi.ch = 1;
			if (chunk < nchunk) pi = A.pairs[(chunk*R + r)*32 + lane];
			if (!tab) { use[r] = init_pair(pi, sc0, m, s, lmax, x[r], p[r], sp[r], q[r], sq[r]); pp[r] = qp[r] = 0; }
			else {
				x[r] = pi.x;
				use[r] = pi.rn >= 0 && 2.0*pi.sh*pi.ch >= sc0.dead_sth;
				int64_t k = (int64_t)m*A.npair_pad + (chunk*R + r)*32 + lane;
				p[r] = A.st_p[k]; pp[r] = A.st_pp[k]; q[r] = A.st_q[k]; qp[r] = A.st_qp[k]; sp[r] = A.st_sp[k]; sq[r] = A.st_sq[k];
			}
			double2 qn = make_double2(0, 0), un = qn, qs = qn, us = qn;
			if (use[r]) { qn = legq[pi.rn]; un = legu[pi.rn]; if (pi.rs >= 0) { qs = legq[pi.rs]; us = legu[pi.rs]; } }
			// Z+ = Q + iU = (Qr - Ui, Qi + Ur),  Z- = Q - iU = (Qr + Ui, Qi - Ur)
			zin[r][0] = qn.x - un.y; zin[r][1] = qn.y + un.x; zin[r][2] = qn.x + un.y; zin[r][3] = qn.y - un.x;
			zin[r][4] = sigma0*(qs.x - us.y); zin[r][5] = sigma0*(qs.y + us.x);
			zin[r][6] = sigma0*(qs.x + us.y); zin[r][7] = sigma0*(qs.y - us.x);
			// slot permutation of bfly_reduce (see adj2_window)
			if (lane & 8) {
				#pragma unroll
				for (int k = 0; k < 8; k += 2) { double t = zin[r][k]; zin[r][k] = zin[r][k + 1]; zin[r][k + 1] = t; }
			}
			if (lane & 16) {
				#pragma unroll
				for (int k = 0; k < 2; k++) {
					double t = zin[r][k]; zin[r][k] = zin[r][k + 2]; zin[r][k + 2] = t;
					t = zin[r][k + 4]; zin[r][k + 4] = zin[r][k + 6]; zin[r][k + 6] = t;
				}
				double t = p[r]; p[r] = q[r]; q[r] = t;
				t = pp[r]; pp[r] = qp[r]; qp[r] = t;
				int ti = sp[r]; sp[r] = sq[r]; sq[r] = ti;
			}
			anyuse |= use[r];
		}
		int wc = 0;
		if (tab) { wc = A.st_w[m*A.ngroup + chunk]; if (wc == LEG_NEVER) anyuse = false; }
		const bool wuse = __any_sync(0xffffffffu, anyuse);
		int phase = 0;
		if (!cta_or<NW>(anyuse)) continue;
		auto issue = [&](int tile, int buf) {
			#pragma unroll
			for (int t = tid; t < TL; t += NT) {
				int i = tile*TL + t;
				if (i < nl) { cp_async8(&tiles[buf][t].a, &ta[i]); cp_async8(&tiles[buf][t].b, &tb[i]); }
				else tiles[buf][t].a = tiles[buf][t].b = 0;
			}
			cp_async_commit();
		};
		auto finish = [&](int buf) {      // the issuing thread completes its entries: (a, b, a, -b)
			#pragma unroll
			for (int t = tid; t < TL; t += NT) { TileAB &e = tiles[buf][t]; e.a2 = e.a; e.nb = -e.b; }
		};
		const int tile0 = tab ? min(wc*8/TL, ntile - 1) : 0;
		issue(tile0, tile0 & 1); cp_async_wait_all(); finish(tile0 & 1);
		cta_sync<NW>();
		for (int tile = tile0; tile < ntile; tile++) {
			const int buf = tile & 1;
			if (tile + 1 < ntile) issue(tile + 1, buf ^ 1);
			const int nwin = (min(TL, nl - tile*TL) + W - 1)/W;
			// output threads: element e -> (l offset e >> 2, component e & 3); fetch the running sum early (cp.async)
			if constexpr (NW == 1) {
				// one l per lane: E_lm and B_lm as one 16-byte access each
				#pragma unroll
				for (int h = 0; h < TL/32; h++) {
					int o = lane + 32*h, i = tile*TL + o;
					if (i < nl) {
						cp_async8(&alvs[o], &tal[i]);
						if (!first) {
							int64_t idx = ms + (int64_t)(l0 + i)*A.lstride;
							cp_async16(&olds[4*o], &A.alm0[idx]);
							if (!A.deriv1) cp_async16(&olds[4*o + 2], &A.alm1[idx]);
						}
					}
				}
			} else {
				#pragma unroll
				for (int h = 0; h < NH; h++) {
					int e = tid + h*NT, i = tile*TL + (e >> 2), k = e & 3;
					if (e < NOUT && i < nl) {
						cp_async8(&alvs[e], &tal[i]);
						int64_t idx = 2*(ms + (int64_t)(l0 + i)*A.lstride) + (k & 1);
						if (!first && !(A.deriv1 && k >= 2)) cp_async8(&olds[e], &(k < 2 ? alme : almb)[idx]);
					}
				}
			}
			cp_async_commit();
			#pragma unroll 1
			for (int w = 0; w < TL/W; w++) {
				double tot = 0;
				if (wuse && w < nwin && tile*TL + w*W >= wc*8) {
					double v[NV];
					#pragma unroll
					for (int i = 0; i < NV; i++) v[i] = 0;
					if (phase < 2) {
						bool mylive = true, anylive = false;
						#pragma unroll
						for (int r = 0; r < R; r++) {
							mylive &= (sp[r] == 0) & (sq[r] == 0);
							anylive |= use[r] & ((sp[r] == 0) | (sq[r] == 0));
						}
						phase = __all_sync(0xffffffffu, mylive) ? 2 : __any_sync(0xffffffffu, anylive) ? 1 : 0;
					}
					const double2 *T = (const double2*)&tiles[buf][w*W] + ((lane >> 4) & 1);
					if (phase == 2) adj2_window<2, R, W>(T, x, p, pp, q, qp, sp, sq, zin, v);
					else if (phase == 1) adj2_window<1, R, W>(T, x, p, pp, q, qp, sp, sq, zin, v);
					else adj2_window<0, R, W>(T, x, p, pp, q, qp, sp, sq, zin, v);
					if (phase != 0) { bfly_reduce<4, W>(v, lane); tot = v[0]; }
				}
				red[buf][warp][w*NV + bfly_element<4, W>(lane)] = tot;      // element 4 (l - l_window) + component
			}
			cp_async_wait_all();
			if (tile + 1 < ntile) finish(buf ^ 1);
			cta_sync<NW>();
			if constexpr (NW == 1) {
				#pragma unroll
				for (int h = 0; h < TL/32; h++) {
					int o = lane + 32*h, i = tile*TL + o;
					if (i < nl) {
						const int l = l0 + i;
						const double2 c01 = *(const double2*)&red[buf][0][4*o], c23 = *(const double2*)&red[buf][0][4*o + 2];
						const double hh = 0.5*alvs[o];
						// E = -(A+ + A-)/2, B = (i/2)(A+ - A-)
						double2 E = make_double2(-hh*(c01.x + c23.x), -hh*(c01.y + c23.y));
						double2 Bv = make_double2(-hh*(c01.y - c23.y), hh*(c01.x - c23.x));
						int64_t idx = ms + (int64_t)l*A.lstride;
						if (A.deriv1) { double f = sqrt((double)l*(l + 1.0)); E.x *= f; E.y *= f; }
						if (!first) { double2 oe = *(const double2*)&olds[4*o]; E.x += oe.x; E.y += oe.y; }
						A.alm0[idx] = E;
						if (!A.deriv1) {
							if (!first) { double2 ob = *(const double2*)&olds[4*o + 2]; Bv.x += ob.x; Bv.y += ob.y; }
							A.alm1[idx] = Bv;
						}
					}
				}
			} else {
			#pragma unroll
			for (int h = 0; h < NH; h++) {
				int e = tid + h*NT, i = tile*TL + (e >> 2), k = e & 3;
				double c = 0;
				if (e < NOUT) {
					#pragma unroll
					for (int w = 0; w < NW; w++) c += red[buf][w][e];
				}
				// the 4 components of one l sit in 4 neighbouring lanes (NT and NOUT are multiples of 32)
				int base = lane & ~3;
				double c0 = __shfl_sync(0xffffffffu, c, base), c1 = __shfl_sync(0xffffffffu, c, base + 1);
				double c2 = __shfl_sync(0xffffffffu, c, base + 2), c3 = __shfl_sync(0xffffffffu, c, base + 3);
				if (e < NOUT && i < nl) {
					int l = l0 + i;
					double hh = 0.5*alvs[e], oldv = first ? 0.0 : olds[e];
					// E = -(A+ + A-)/2, B = (i/2)(A+ - A-)
					double val = k == 0 ? -hh*(c0 + c2) : k == 1 ? -hh*(c1 + c3) : k == 2 ? -hh*(c1 - c3) : hh*(c0 - c2);
					int64_t idx = 2*(ms + (int64_t)l*A.lstride) + (k & 1);
					if (A.deriv1) { if (k < 2) alme[idx] = oldv + val*sqrt((double)l*(l + 1.0)); }
					else (k < 2 ? alme : almb)[idx] = oldv + val;
				}
			}
			}
		}
		first = false;
	}
	if (first) for (int i = tid; i < nl; i += NW*32) {
		int64_t idx = 2*(ms + (int64_t)(l0 + i)*A.lstride);
		alme[idx] = alme[idx + 1] = 0;
		if (!A.deriv1) almb[idx] = almb[idx + 1] = 0;
	}
	if (SIG) signal_done<NW, SIG>(A);
}

// ------------------------------------------------------------------------------------ batched synthesis
// NB alm sets of one spin (Monte-Carlo realisations of one geometry, pixell/curvedsky.py:17-36 called in a loop) share
// the recurrence: per l and ring pair the 2 (spin 0) or 4 (spin s) recurrence DFMAs are paid once and only the 2 / 8
// accumulations per member, i.e. (2 + 2 NB)/NB instead of 4 and (4 + 8 NB)/NB instead of 12 FP64 instructions per member.
// The accumulations of one member are the very sequence of FMAs the single-map kernels execute, so the results are
// bit-identical to theirs.  Same structure as k_synth0 / k_synth2 (rounds of 32 R NW ring pairs, l tiles through cp.async,
// start table, live / masked / plain windows); whole launches only.

template<int NB> struct Tile0B { double a, pad; double ar[NB], ai[NB]; };

template<int MODE, int R, int NB> __device__ __forceinline__ void synth0b_window(const Tile0B<NB> *T,
	const double (&x)[R], double (&g)[R], double (&gp)[R], int (&sc)[R], double (&acc)[NB][R][2][2])
{
	#pragma unroll
	for (int j = 0; j < 8; j++) {
		const Tile0B<NB> t = T[j];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			if (MODE != 0) {
				double gv = (MODE == 1) ? (sc[r] == 0 ? g[r] : 0.0) : g[r];
				#pragma unroll
				for (int b = 0; b < NB; b++) {
					acc[b][r][j & 1][0] = fma(gv, t.ar[b], acc[b][r][j & 1][0]);
					acc[b][r][j & 1][1] = fma(gv, t.ai[b], acc[b][r][j & 1][1]);
				}
			}
			double ng = fma(t.a, x[r]*g[r], -gp[r]);
			gp[r] = g[r]; g[r] = ng;
		}
	}
	if (MODE != 2) {
		#pragma unroll
		for (int r = 0; r < R; r++) rescale(g[r], gp[r], sc[r]);
	}
}

template<int R, int NW, int MINB, int TL, int NB> __global__ void __launch_bounds__(NW*32, MINB) k_synth0b(LegArgs A)
{
	__shared__ __align__(16) Tile0B<NB> tiles[2][TL];
	__shared__ __align__(16) double2 raw_alm[NB][TL];
	__shared__ __align__(16) double raw_al[TL], raw_a[TL];
	__shared__ int wslot;
	const int m = A.m0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int lmax = A.lmax, l0 = m;
	const bool tab = A.st_w != nullptr && R*32 == LEG_GROUP;
	const int nl = lmax - l0 + 1, ntile = (nl + TL - 1)/TL;
	const double *ta = A.ta + A.toff[m], *tal = A.talpha + A.toff[m];
	const double2 *alm = A.alm0 + A.mstart[m];
	const SeqConst sc0 = seq_const(m, 0, lmax, A.pref);
	const int nchunk = A.npair_pad/(32*R);
	double2 *leg = A.leg0 + (int64_t)m*A.leg_mstride;
	const int rlive = first_live_pair(A.pairs, A.npair, sc0.dead_sth)/(32*R*NW);
	for (int i = tid; i < rlive*32*R*NW; i += NW*32) {
		PairInfo pi = A.pairs[i];
		#pragma unroll
		for (int b = 0; b < NB; b++) {
			if (pi.rn >= 0) leg[b*A.leg_bstride + pi.rn] = make_double2(0, 0);
			if (pi.rs >= 0) leg[b*A.leg_bstride + pi.rs] = make_double2(0, 0);
		}
	}
	for (int round = rlive; round*NW < nchunk; round++) {
		const int chunk = round*NW + warp;
		double x[R], g[R], gp[R], acc[NB][R][2][2]; int sc[R];
		bool anyuse = false, use[R];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			PairInfo pi; pi.rn = -1; pi.rs = -1; pi.x = 0; pi.sh = 0; pi.ch = 1;
			if (chunk < nchunk) pi = A.pairs[(chunk*R + r)*32 + lane];
			double dummy; int dsc;
			if (!tab) { use[r] = init_pair(pi, sc0, m, 0, lmax, x[r], dummy, dsc, g[r], sc[r]); gp[r] = 0; }
			else {
				x[r] = pi.x;
				use[r] = pi.rn >= 0 && 2.0*pi.sh*pi.ch >= sc0.dead_sth;
				g[r] = gp[r] = 0; sc[r] = 0;
				if (chunk < nchunk) { int64_t k = (int64_t)m*A.npair_pad + (chunk*R + r)*32 + lane; g[r] = A.st_p[k]; gp[r] = A.st_pp[k]; sc[r] = A.st_sp[k]; }
			}
			#pragma unroll
			for (int b = 0; b < NB; b++) acc[b][r][0][0] = acc[b][r][0][1] = acc[b][r][1][0] = acc[b][r][1][1] = 0;
			anyuse |= use[r];
		}
		int wc = 0;
		if (tab) { wc = chunk < nchunk ? A.st_w[m*A.ngroup + chunk] : LEG_NEVER; if (wc == LEG_NEVER) anyuse = false; }
		const int wc_cta = tab ? cta_min<NW>(wc, &wslot) : 0;
		const bool wuse = __any_sync(0xffffffffu, anyuse);
		int phase = 0;
		if (cta_or<NW>(anyuse)) {
			auto issue = [&](int tile) {
				int i = tile*TL + tid;
				if (tid < TL && i < nl) {
					#pragma unroll
					for (int b = 0; b < NB; b++) cp_async16(&raw_alm[b][tid], &alm[b*A.alm_bstride + (int64_t)(l0 + i)*A.lstride]);
					cp_async8(&raw_al[tid], &tal[i]); cp_async8(&raw_a[tid], &ta[i]);
				}
				cp_async_commit();
			};
			auto finish = [&](int tile, int buf) {
				cp_async_wait_all();
				int i = tile*TL + tid;
				if (tid < TL) {
					Tile0B<NB> t; t.a = t.pad = 0;
					#pragma unroll
					for (int b = 0; b < NB; b++) t.ar[b] = t.ai[b] = 0;
					if (i < nl) {
						double al = raw_al[tid]; t.a = raw_a[tid];
						#pragma unroll
						for (int b = 0; b < NB; b++) { double2 v = raw_alm[b][tid]; t.ar[b] = v.x*al; t.ai[b] = v.y*al; }
					}
					tiles[buf][tid] = t;
				}
			};
			const int tile0 = min(wc_cta*8/TL, ntile - 1);
			issue(tile0); finish(tile0, tile0 & 1);
			cta_sync<NW>();
			for (int tile = tile0; tile < ntile; tile++) {
				const int buf = tile & 1;
				if (tile + 1 < ntile) issue(tile + 1);
				const int nwin = (min(TL, nl - tile*TL) + 7) >> 3;
				for (int w = 0; wuse && w < nwin; w++) {
					if (tile*(TL/8) + w < wc) continue;
					if (phase < 2) {
						bool mylive = true, anylive = false;
						#pragma unroll
						for (int r = 0; r < R; r++) { mylive &= (sc[r] == 0); anylive |= (sc[r] == 0 && use[r]); }
						phase = __all_sync(0xffffffffu, mylive) ? 2 : __any_sync(0xffffffffu, anylive) ? 1 : 0;
					}
					const Tile0B<NB> *T = &tiles[buf][w*8];
					if (phase == 2) synth0b_window<2, R, NB>(T, x, g, gp, sc, acc);
					else if (phase == 1) synth0b_window<1, R, NB>(T, x, g, gp, sc, acc);
					else synth0b_window<0, R, NB>(T, x, g, gp, sc, acc);
				}
				if (tile + 1 < ntile) finish(tile + 1, buf ^ 1);
				cta_sync<NW>();
			}
		}
		#pragma unroll
		for (int r = 0; r < R; r++) {
			int rn = -1, rs = -1;
			if (chunk < nchunk) { const PairInfo &pi = A.pairs[(chunk*R + r)*32 + lane]; rn = pi.rn; rs = pi.rs; }
			#pragma unroll
			for (int b = 0; b < NB; b++) {
				if (rn >= 0) leg[b*A.leg_bstride + rn] = make_double2(acc[b][r][0][0] + acc[b][r][1][0], acc[b][r][0][1] + acc[b][r][1][1]);
				if (rs >= 0) leg[b*A.leg_bstride + rs] = make_double2(acc[b][r][0][0] - acc[b][r][1][0], acc[b][r][0][1] - acc[b][r][1][1]);
			}
		}
	}
}

template<int NB> struct Tile2B { double a, b; double apr[NB], api[NB], amr[NB], ami[NB]; };

template<int MODE, int R, int NB> __device__ __forceinline__ void synth2b_window(const Tile2B<NB> *T,
	const double (&x)[R], double (&p)[R], double (&pp)[R], double (&q)[R], double (&qp)[R],
	int (&sp)[R], int (&sq)[R], double (&acc)[NB][R][8])
{
	#pragma unroll
	for (int j = 0; j < 8; j++) {
		const Tile2B<NB> t = T[j];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			if (MODE != 0) {
				double pv = (MODE == 1) ? (sp[r] == 0 ? p[r] : 0.0) : p[r];
				double qv = (MODE == 1) ? (sq[r] == 0 ? q[r] : 0.0) : q[r];
				double ps = (j & 1) ? -pv : pv, qs = (j & 1) ? -qv : qv;
				#pragma unroll
				for (int b = 0; b < NB; b++) {
					double (&c)[8] = acc[b][r];
					c[0] = fma(pv, t.apr[b], c[0]); c[1] = fma(pv, t.api[b], c[1]);
					c[2] = fma(ps, t.amr[b], c[2]); c[3] = fma(ps, t.ami[b], c[3]);
					c[4] = fma(qs, t.apr[b], c[4]); c[5] = fma(qs, t.api[b], c[5]);
					c[6] = fma(qv, t.amr[b], c[6]); c[7] = fma(qv, t.ami[b], c[7]);
				}
			}
			double np = fma(fma(t.a, x[r],  t.b), p[r], -pp[r]);
			double nq = fma(fma(t.a, x[r], -t.b), q[r], -qp[r]);
			pp[r] = p[r]; p[r] = np; qp[r] = q[r]; q[r] = nq;
		}
	}
	if (MODE != 2) {
		#pragma unroll
		for (int r = 0; r < R; r++) { rescale(p[r], pp[r], sp[r]); rescale(q[r], qp[r], sq[r]); }
	}
}

// batch member b: alm (E, B) at alm0 / alm1 + b alm_bstride, leg (Q, U) at leg0 / leg1 + b leg_bstride
template<int R, int NW, int MINB, int TL, int NB> __global__ void __launch_bounds__(NW*32, MINB) k_synth2b(LegArgs A)
{
	__shared__ __align__(16) Tile2B<NB> tiles[2][TL];
	__shared__ __align__(16) double2 raw_e[NB][TL], raw_b[NB][TL];
	__shared__ __align__(16) double raw_ta[TL], raw_tb[TL], raw_al[TL];
	__shared__ int wslot;
	const int m = A.m0 + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int lmax = A.lmax, s = A.spin, l0 = m > s ? m : s;
	const bool tab = A.st_w != nullptr && R*32 == LEG_GROUP;
	double2 *legq = A.leg0 + (int64_t)m*A.leg_mstride, *legu = A.leg1 + (int64_t)m*A.leg_mstride;
	const int nchunk = A.npair_pad/(32*R);
	if (l0 > lmax) {
		for (int i = tid; i < A.npair_pad; i += NW*32) {
			PairInfo pi = A.pairs[i];
			#pragma unroll
			for (int b = 0; b < NB; b++) {
				if (pi.rn >= 0) legq[b*A.leg_bstride + pi.rn] = legu[b*A.leg_bstride + pi.rn] = make_double2(0, 0);
				if (pi.rs >= 0) legq[b*A.leg_bstride + pi.rs] = legu[b*A.leg_bstride + pi.rs] = make_double2(0, 0);
			}
		}
		return;
	}
	const int nl = lmax - l0 + 1, ntile = (nl + TL - 1)/TL;
	const double *ta = A.ta + A.toff[m], *tb = A.tb + A.toff[m], *tal = A.talpha + A.toff[m];
	const SeqConst sc0 = seq_const(m, s, lmax, A.pref);
	const double sigma0 = ((l0 + m + s) & 1) ? -1.0 : 1.0;
	const int64_t ms = A.mstart[m];
	const int rlive = first_live_pair(A.pairs, A.npair, sc0.dead_sth)/(32*R*NW);
	for (int i = tid; i < rlive*32*R*NW; i += NW*32) {
		PairInfo pi = A.pairs[i];
		#pragma unroll
		for (int b = 0; b < NB; b++) {
			if (pi.rn >= 0) legq[b*A.leg_bstride + pi.rn] = legu[b*A.leg_bstride + pi.rn] = make_double2(0, 0);
			if (pi.rs >= 0) legq[b*A.leg_bstride + pi.rs] = legu[b*A.leg_bstride + pi.rs] = make_double2(0, 0);
		}
	}
	for (int round = rlive; round*NW < nchunk; round++) {
		const int chunk = round*NW + warp;
		double x[R], p[R], pp[R], q[R], qp[R], acc[NB][R][8]; int sp[R], sq[R];
		bool anyuse = false, use[R];
		#pragma unroll
		for (int r = 0; r < R; r++) {
			PairInfo pi; pi.rn = -1; pi.rs = -1; pi.x = 0; pi.sh = 0; pi.ch = 1;
			if (chunk < nchunk) pi = A.pairs[(chunk*R + r)*32 + lane];
			if (!tab) { use[r] = init_pair(pi, sc0, m, s, lmax, x[r], p[r], sp[r], q[r], sq[r]); pp[r] = qp[r] = 0; }
			else {
				x[r] = pi.x;
				use[r] = pi.rn >= 0 && 2.0*pi.sh*pi.ch >= sc0.dead_sth;
				p[r] = pp[r] = q[r] = qp[r] = 0; sp[r] = sq[r] = 0;
				if (chunk < nchunk) {
					int64_t k = (int64_t)m*A.npair_pad + (chunk*R + r)*32 + lane;
					p[r] = A.st_p[k]; pp[r] = A.st_pp[k]; q[r] = A.st_q[k]; qp[r] = A.st_qp[k]; sp[r] = A.st_sp[k]; sq[r] = A.st_sq[k];
				}
			}
			#pragma unroll
			for (int b = 0; b < NB; b++) {
				#pragma unroll
				for (int k = 0; k < 8; k++) acc[b][r][k] = 0;
			}
			anyuse |= use[r];
		}
		int wc = 0;
		if (tab) { wc = chunk < nchunk ? A.st_w[m*A.ngroup + chunk] : LEG_NEVER; if (wc == LEG_NEVER) anyuse = false; }
		const int wc_cta = tab ? cta_min<NW>(wc, &wslot) : 0;
		const bool wuse = __any_sync(0xffffffffu, anyuse);
		int phase = 0;
		if (cta_or<NW>(anyuse)) {
			auto issue = [&](int tile) {
				int i = tile*TL + tid;
				if (tid < TL && i < nl) {
					int64_t idx = ms + (int64_t)(l0 + i)*A.lstride;
					#pragma unroll
					for (int b = 0; b < NB; b++) {
						cp_async16(&raw_e[b][tid], &A.alm0[b*A.alm_bstride + idx]);
						cp_async16(&raw_b[b][tid], &A.alm1[b*A.alm_bstride + idx]);
					}
					cp_async8(&raw_ta[tid], &ta[i]); cp_async8(&raw_tb[tid], &tb[i]); cp_async8(&raw_al[tid], &tal[i]);
				}
				cp_async_commit();
			};
			auto finish = [&](int tile, int buf) {
				cp_async_wait_all();
				int i = tile*TL + tid;
				if (tid < TL) {
					Tile2B<NB> t; t.a = t.b = 0;
					#pragma unroll
					for (int b = 0; b < NB; b++) t.apr[b] = t.api[b] = t.amr[b] = t.ami[b] = 0;
					if (i < nl) {
						double h = -0.5*raw_al[tid];
						t.a = raw_ta[tid]; t.b = raw_tb[tid];
						#pragma unroll
						for (int b = 0; b < NB; b++) {
							double2 E = raw_e[b][tid], B = raw_b[b][tid];
							t.apr[b] = h*(E.x - B.y); t.api[b] = h*(E.y + B.x);
							t.amr[b] = h*(E.x + B.y); t.ami[b] = h*(E.y - B.x);
						}
					}
					tiles[buf][tid] = t;
				}
			};
			const int tile0 = min(wc_cta*8/TL, ntile - 1);
			issue(tile0); finish(tile0, tile0 & 1);
			cta_sync<NW>();
			for (int tile = tile0; tile < ntile; tile++) {
				const int buf = tile & 1;
				if (tile + 1 < ntile) issue(tile + 1);
				const int nwin = (min(TL, nl - tile*TL) + 7) >> 3;
				for (int w = 0; wuse && w < nwin; w++) {
					if (tile*(TL/8) + w < wc) continue;
					if (phase < 2) {
						bool mylive = true, anylive = false;
						#pragma unroll
						for (int r = 0; r < R; r++) {
							mylive &= (sp[r] == 0) & (sq[r] == 0);
							anylive |= use[r] & ((sp[r] == 0) | (sq[r] == 0));
						}
						phase = __all_sync(0xffffffffu, mylive) ? 2 : __any_sync(0xffffffffu, anylive) ? 1 : 0;
					}
					const Tile2B<NB> *T = &tiles[buf][w*8];
					if (phase == 2) synth2b_window<2, R, NB>(T, x, p, pp, q, qp, sp, sq, acc);
					else if (phase == 1) synth2b_window<1, R, NB>(T, x, p, pp, q, qp, sp, sq, acc);
					else synth2b_window<0, R, NB>(T, x, p, pp, q, qp, sp, sq, acc);
				}
				if (tile + 1 < ntile) finish(tile + 1, buf ^ 1);
				cta_sync<NW>();
			}
		}
		// (the ring indices are read again here instead of being carried through the l loop: the accumulators need the registers)
		#pragma unroll
		for (int r = 0; r < R; r++) {
			int rn = -1, rs = -1;
			if (chunk < nchunk) { const PairInfo &pi = A.pairs[(chunk*R + r)*32 + lane]; rn = pi.rn; rs = pi.rs; }
			#pragma unroll
			for (int b = 0; b < NB; b++) {
				const double (&c)[8] = acc[b][r];
				if (rn >= 0) {
					double spr = c[0], spi = c[1], sqr = c[6], sqi = c[7];
					legq[b*A.leg_bstride + rn] = make_double2(spr + sqr, spi + sqi);
					legu[b*A.leg_bstride + rn] = make_double2(spi - sqi, sqr - spr);
				}
				if (rs >= 0) {
					double spr = sigma0*c[4], spi = sigma0*c[5], sqr = sigma0*c[2], sqi = sigma0*c[3];
					legq[b*A.leg_bstride + rs] = make_double2(spr + sqr, spi + sqi);
					legu[b*A.leg_bstride + rs] = make_double2(spi - sqi, sqr - spr);
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------ start table

// One warp per (group of LEG_GROUP ring pairs, m): runs the recurrence exactly as the kernels' pre-phase does (same
// window function, same rescaling cadence) until a ring of the group is live, and records window and state.
template<int SPIN0> __global__ void __launch_bounds__(32) k_start_build(LegArgs A, int *w_out, double *o_p, double *o_pp,
	double *o_q, double *o_qp, signed char *o_sp, signed char *o_sq)
{
	constexpr int R = LEG_GROUP/32;
	__shared__ __align__(16) Tile2 t2[8];
	__shared__ __align__(16) Tile0 t0[8];
	const int grp = blockIdx.x, m = blockIdx.y, lane = threadIdx.x;
	const int lmax = A.lmax, s = A.spin, l0 = m > s ? m : s;
	int *wdst = &w_out[m*A.ngroup + grp];
	const int64_t kbase = (int64_t)m*A.npair_pad + grp*LEG_GROUP + lane;
	double x[R], p[R], pp[R], q[R], qp[R]; int sp[R], sq[R]; bool use[R], anyuse = false;
	if (l0 <= lmax) {
		const SeqConst sc0 = seq_const(m, s, lmax, A.pref);
		#pragma unroll
		for (int r = 0; r < R; r++) {
			PairInfo pi = A.pairs[grp*LEG_GROUP + r*32 + lane];
			use[r] = init_pair(pi, sc0, m, s, lmax, x[r], p[r], sp[r], q[r], sq[r]);
			pp[r] = qp[r] = 0;
			anyuse |= use[r];
		}
	} else {
		#pragma unroll
		for (int r = 0; r < R; r++) { x[r] = p[r] = pp[r] = q[r] = qp[r] = 0; sp[r] = sq[r] = 0; use[r] = false; }
	}
	int wfound = LEG_NEVER;
	if (__any_sync(0xffffffffu, anyuse)) {
		const int nl = lmax - l0 + 1, nwin = (nl + 7) >> 3;
		const double *ta = A.ta + A.toff[m], *tb = SPIN0 ? nullptr : A.tb + A.toff[m];
		for (int w = 0; w < nwin; w++) {
			bool anylive = false;
			#pragma unroll
			for (int r = 0; r < R; r++) anylive |= SPIN0 ? (use[r] && sq[r] == 0) : (use[r] & ((sp[r] == 0) | (sq[r] == 0)));
			if (__any_sync(0xffffffffu, anylive)) { wfound = w; break; }
			if (lane < 8) {
				int i = w*8 + lane;
				if (SPIN0) { Tile0 t; t.ar = t.ai = t.pad = 0; t.a = i < nl ? ta[i] : 0.0; t0[lane] = t; }
				else { Tile2 t; t.apr = t.api = t.amr = t.ami = 0; t.a = i < nl ? ta[i] : 0.0; t.b = i < nl ? tb[i] : 0.0; t2[lane] = t; }
			}
			__syncwarp();
			if (SPIN0) { double acc[R][2][2]; synth0_window<0, R>(t0, x, q, qp, sq, acc); }
			else { double acc[R][8]; synth2_window<0, R>(t2, x, p, pp, q, qp, sp, sq, acc); }
			__syncwarp();
		}
	}
	if (lane == 0) *wdst = wfound;
	#pragma unroll
	for (int r = 0; r < R; r++) {
		const int64_t k = kbase + r*32;
		if (SPIN0) { o_p[k] = q[r]; o_pp[k] = qp[r]; o_sp[k] = (signed char)sq[r]; }      // spin 0 uses the q sequence (p = q)
		else { o_p[k] = p[r]; o_pp[k] = pp[r]; o_q[k] = q[r]; o_qp[k] = qp[r]; o_sp[k] = (signed char)sp[r]; o_sq[k] = (signed char)sq[r]; }
	}
}

// ------------------------------------------------------------------------------------ host entry points

// Kernel variants: R ring pairs per lane, NW warps per CTA, MINB resident CTAs per SM the register
// allocation is tuned for, TL l-values per shared-memory tile.  The default is the fastest measured
// on B200 (profiles/); B2_LEG_VARIANT=<synth0>,<adj0>,<synth2>,<adj2> selects others for tuning runs.
static int g_variant[4] = {-1, -1, -1, -1};
static void variant_init()
{
	if (g_variant[0] >= 0) return;
	int d[4] = {B2_DEF_SYNTH0, B2_DEF_ADJ0, B2_DEF_SYNTH2, B2_DEF_ADJ2};
	const char *e = getenv("B2_LEG_VARIANT");
	if (e) sscanf(e, "%d,%d,%d,%d", &d[0], &d[1], &d[2], &d[3]);
	for (int i = 0; i < 4; i++) g_variant[i] = d[i];
}
int leg_set_variant(int which, int v)
{
	B2_REQUIRE(which >= 0 && which < 4 && v >= 0, "leg_set_variant: argument out of range");
	variant_init();
	g_variant[which] = v;
	return 0;
}
static int sig_style() { static const int v = getenv("B2_SIG_STYLE") ? atoi(getenv("B2_SIG_STYLE")) : 1; return v; }
static int variant_of(int which) { variant_init(); return g_variant[which]; }

static LegArgs make_args(const LegTables &T, const LegGeom &G, const AlmLayout &L, int deriv1,
	double2 *alm, int64_t alm_cstride, double2 *leg, const LegStart *S = nullptr)
{
	LegArgs A;
	A.st_w = nullptr; A.st_p = A.st_pp = A.st_q = A.st_qp = nullptr; A.st_sp = A.st_sq = nullptr; A.ngroup = G.npair_pad/LEG_GROUP;
	if (S && S->spin == T.spin && S->w.n) {
		A.st_w = S->w.p; A.st_p = S->p.p; A.st_pp = S->pp.p; A.st_q = S->q.p; A.st_qp = S->qp.p; A.st_sp = S->sp.p; A.st_sq = S->sq.p;
	}
	A.lmax = L.lmax; A.mmax = L.mmax; A.spin = T.spin; A.deriv1 = deriv1;
	A.m0 = 0; A.pair_lo = 0; A.pair_hi = G.npair_pad;
	A.sig.count = nullptr; A.sig.flag = nullptr; A.sig.ncut = 0; A.sig.epoch = 0;
	A.alm_bstride = 0; A.leg_bstride = 0;
	A.toff = T.toff.p; A.ta = T.a.p; A.tb = T.b.p; A.talpha = T.alpha.p; A.pref = T.pref.p;
	A.pairs = G.pairs.p; A.npair_pad = G.npair_pad; A.npair = G.npair;
	A.mstart = L.mstart_d; A.lstride = L.lstride;
	A.alm0 = alm; A.alm1 = alm + alm_cstride;
	A.leg_mstride = G.nring_pad;
	A.leg0 = leg; A.leg1 = leg + (int64_t)(L.mmax + 1)*G.nring_pad;
	return A;
}

static int check_args(const LegTables &T, const AlmLayout &L, int deriv1)
{
	B2_REQUIRE(T.lmax == L.lmax && T.mmax >= L.mmax, "Legendre tables do not match the alm layout");
	B2_REQUIRE(!deriv1 || T.spin == 1, "DERIV1 mode needs spin-1 tables");
	return 0;
}

#define LAUNCH(K, ...) K<__VA_ARGS__><<<nm_launch, launch_threads<__VA_ARGS__>(), 0, st>>>(A)
template<int R, int NW, int... REST> constexpr int launch_threads() { return NW*32; }

int leg_build_start(LegStart &S, const LegTables &T, const LegGeom &G)
{
	// the table is built for mmax = T.mmax; mstart / alm are not touched by the build kernel
	S.spin = T.spin; S.ngroup = G.npair_pad/LEG_GROUP;
	const size_t nm = (size_t)T.mmax + 1, ne = nm*G.npair_pad;
	if (S.w.alloc(nm*S.ngroup) || S.p.alloc(ne) || S.pp.alloc(ne) || S.sp.alloc(ne)) return 1;
	if (T.spin > 0 && (S.q.alloc(ne) || S.qp.alloc(ne) || S.sq.alloc(ne))) return 1;
	AlmLayout L; L.lmax = T.lmax; L.mmax = T.mmax; L.mstart_d = nullptr; L.lstride = 1;
	LegArgs A = make_args(T, G, L, 0, nullptr, 0, nullptr);
	dim3 grid(S.ngroup, (unsigned)nm);
	if (T.spin == 0) k_start_build<1><<<grid, 32>>>(A, S.w.p, S.p.p, S.pp.p, nullptr, nullptr, S.sp.p, nullptr);
	else k_start_build<0><<<grid, 32>>>(A, S.w.p, S.p.p, S.pp.p, S.q.p, S.qp.p, S.sp.p, S.sq.p);
	B2_LAUNCH_CHECK();
	B2_CHECK(cudaDeviceSynchronize());
	return 0;
}

int leg_alm2leg(const LegTables &T, const LegGeom &G, const AlmLayout &L, int deriv1,
	const double2 *alm, int64_t alm_cstride, double2 *leg, cudaStream_t st, const LegStart *S, int pair_lo, int pair_hi, const LegSignal *gate)
{
	if (check_args(T, L, deriv1)) return 1;
	LegArgs A = make_args(T, G, L, deriv1, (double2*)alm, alm_cstride, leg, S);
	if (gate) {
		B2_REQUIRE(gate->ncut >= 2 && gate->ncut <= LEG_MAXCUT && gate->cut[0] == 0 && gate->cut[gate->ncut - 1] == L.mmax + 1 && gate->flag,
			"alm2leg: arrival ranges must cover all orders");
		A.sig = *gate;
	}
	const int nm_launch = L.mmax + 1;
	if (pair_hi > pair_lo) {
		B2_REQUIRE(pair_lo % 256 == 0 && (pair_hi % 256 == 0 || pair_hi >= G.npair_pad), "alm2leg: ring-pair ranges must be multiples of 256");
		A.pair_lo = pair_lo; A.pair_hi = std::min(pair_hi, G.npair_pad);
	}
	// template arguments: R, NW, MINB, TL (variant 0 = fastest measured on B200 at lmax 8000, profiles/r1*_tune_*)
	if (T.spin == 0) switch (variant_of(0)) {
		case 0: if (gate) LAUNCH(k_synth0, 4, 2, 8, 64, 1); else LAUNCH(k_synth0, 4, 2, 8, 64, 0); break;
		case 1: if (gate) LAUNCH(k_synth0, 4, 1, 16, 32, 1); else LAUNCH(k_synth0, 4, 1, 16, 32, 0); break;
		case 2: if (gate) LAUNCH(k_synth0, 2, 4, 6, 64, 1); else LAUNCH(k_synth0, 2, 4, 6, 64, 0); break;
		case 3: if (gate) LAUNCH(k_synth0, 4, 2, 8, 128, 1); else LAUNCH(k_synth0, 4, 2, 8, 128, 0); break;
		case 4: if (gate) LAUNCH(k_synth0, 4, 2, 8, 256, 1); else LAUNCH(k_synth0, 4, 2, 8, 256, 0); break;
		default: B2_REQUIRE(0, "unknown k_synth0 variant");
	} else switch (variant_of(2)) {
		case 0: if (gate) LAUNCH(k_synth2, 4, 2, 6, 64, 1); else LAUNCH(k_synth2, 4, 2, 6, 64, 0); break;
		case 1: if (gate) LAUNCH(k_synth2, 4, 4, 3, 64, 1); else LAUNCH(k_synth2, 4, 4, 3, 64, 0); break;
		case 2: if (gate) LAUNCH(k_synth2, 4, 1, 12, 32, 1); else LAUNCH(k_synth2, 4, 1, 12, 32, 0); break;
		case 3: if (gate) LAUNCH(k_synth2, 2, 4, 3, 64, 1); else LAUNCH(k_synth2, 2, 4, 3, 64, 0); break;
		case 4: if (gate) LAUNCH(k_synth2, 4, 2, 5, 64, 1); else LAUNCH(k_synth2, 4, 2, 5, 64, 0); break;
		case 5: if (gate) LAUNCH(k_synth2, 4, 2, 6, 128, 1); else LAUNCH(k_synth2, 4, 2, 6, 128, 0); break;
		case 6: if (gate) LAUNCH(k_synth2, 4, 2, 6, 256, 1); else LAUNCH(k_synth2, 4, 2, 6, 256, 0); break;
		default: B2_REQUIRE(0, "unknown k_synth2 variant");
	}
	B2_LAUNCH_CHECK();
	return 0;
}

int leg_leg2alm(const LegTables &T, const LegGeom &G, const AlmLayout &L, int deriv1,
	double2 *alm, int64_t alm_cstride, const double2 *leg, cudaStream_t st, const LegStart *S, int m_lo, int m_hi, const LegSignal *sig)
{
	if (check_args(T, L, deriv1)) return 1;
	LegArgs A = make_args(T, G, L, deriv1, alm, alm_cstride, (double2*)leg, S);
	if (sig) {
		B2_REQUIRE(sig->ncut >= 2 && sig->ncut <= LEG_MAXCUT && sig->cut[0] == 0 && sig->cut[sig->ncut - 1] == L.mmax + 1 && m_hi <= m_lo,
			"leg2alm: completion ranges must cover all orders of a whole launch");
		A.sig = *sig;
	}
	int nm_launch = L.mmax + 1;
	if (m_hi > m_lo) { A.m0 = m_lo; nm_launch = std::min(m_hi, L.mmax + 1) - m_lo; if (nm_launch <= 0) return 0; }
	// template arguments: R, NW, MINB, TL, W
	if (T.spin == 0) switch (variant_of(1)) {
		case 0: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 8, 1, 8, 32, 8, 2); else LAUNCH(k_adj0, 8, 1, 8, 32, 8, 1); } else LAUNCH(k_adj0, 8, 1, 8, 32, 8, 0); break;
		case 1: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 8, 1, 10, 32, 4, 2); else LAUNCH(k_adj0, 8, 1, 10, 32, 4, 1); } else LAUNCH(k_adj0, 8, 1, 10, 32, 4, 0); break;
		case 2: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 4, 4, 3, 64, 8, 2); else LAUNCH(k_adj0, 4, 4, 3, 64, 8, 1); } else LAUNCH(k_adj0, 4, 4, 3, 64, 8, 0); break;
		case 3: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 4, 1, 12, 32, 8, 2); else LAUNCH(k_adj0, 4, 1, 12, 32, 8, 1); } else LAUNCH(k_adj0, 4, 1, 12, 32, 8, 0); break;
		case 4: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 4, 1, 16, 32, 8, 2); else LAUNCH(k_adj0, 4, 1, 16, 32, 8, 1); } else LAUNCH(k_adj0, 4, 1, 16, 32, 8, 0); break;
		case 5: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 8, 1, 8, 32, 4, 2); else LAUNCH(k_adj0, 8, 1, 8, 32, 4, 1); } else LAUNCH(k_adj0, 8, 1, 8, 32, 4, 0); break;
		case 6: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 4, 1, 16, 32, 4, 2); else LAUNCH(k_adj0, 4, 1, 16, 32, 4, 1); } else LAUNCH(k_adj0, 4, 1, 16, 32, 4, 0); break;
		case 7: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 8, 1, 8, 128, 8, 2); else LAUNCH(k_adj0, 8, 1, 8, 128, 8, 1); } else LAUNCH(k_adj0, 8, 1, 8, 128, 8, 0); break;
		case 8: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 8, 1, 8, 256, 8, 2); else LAUNCH(k_adj0, 8, 1, 8, 256, 8, 1); } else LAUNCH(k_adj0, 8, 1, 8, 256, 8, 0); break;
		case 9: if (sig) { if (sig_style() == 2) LAUNCH(k_adj0, 8, 1, 8, 64, 8, 2); else LAUNCH(k_adj0, 8, 1, 8, 64, 8, 1); } else LAUNCH(k_adj0, 8, 1, 8, 64, 8, 0); break;
		default: B2_REQUIRE(0, "unknown k_adj0 variant");
	} else switch (variant_of(3)) {
		case 0: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 4, 1, 10, 32, 4, 2); else LAUNCH(k_adj2, 4, 1, 10, 32, 4, 1); } else LAUNCH(k_adj2, 4, 1, 10, 32, 4, 0); break;
		case 1: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 4, 1, 8, 32, 4, 2); else LAUNCH(k_adj2, 4, 1, 8, 32, 4, 1); } else LAUNCH(k_adj2, 4, 1, 8, 32, 4, 0); break;
		case 2: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 2, 4, 3, 32, 8, 2); else LAUNCH(k_adj2, 2, 4, 3, 32, 8, 1); } else LAUNCH(k_adj2, 2, 4, 3, 32, 8, 0); break;
		case 3: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 2, 1, 12, 32, 8, 2); else LAUNCH(k_adj2, 2, 1, 12, 32, 8, 1); } else LAUNCH(k_adj2, 2, 1, 12, 32, 8, 0); break;
		case 4: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 4, 1, 9, 32, 8, 2); else LAUNCH(k_adj2, 4, 1, 9, 32, 8, 1); } else LAUNCH(k_adj2, 4, 1, 9, 32, 8, 0); break;
		case 5: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 4, 1, 8, 32, 8, 2); else LAUNCH(k_adj2, 4, 1, 8, 32, 8, 1); } else LAUNCH(k_adj2, 4, 1, 8, 32, 8, 0); break;
		case 6: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 4, 1, 11, 32, 4, 2); else LAUNCH(k_adj2, 4, 1, 11, 32, 4, 1); } else LAUNCH(k_adj2, 4, 1, 11, 32, 4, 0); break;
		case 7: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 4, 1, 8, 64, 8, 2); else LAUNCH(k_adj2, 4, 1, 8, 64, 8, 1); } else LAUNCH(k_adj2, 4, 1, 8, 64, 8, 0); break;
		case 8: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 4, 1, 8, 128, 8, 2); else LAUNCH(k_adj2, 4, 1, 8, 128, 8, 1); } else LAUNCH(k_adj2, 4, 1, 8, 128, 8, 0); break;
		case 9: if (sig) { if (sig_style() == 2) LAUNCH(k_adj2, 4, 1, 10, 64, 4, 2); else LAUNCH(k_adj2, 4, 1, 10, 64, 4, 1); } else LAUNCH(k_adj2, 4, 1, 10, 64, 4, 0); break;
		default: B2_REQUIRE(0, "unknown k_adj2 variant");
	}
	B2_LAUNCH_CHECK();
	return 0;
}

int leg_batch_size(int spin) { return spin == 0 ? 4 : 2; }

int leg_alm2leg_batch(const LegTables &T, const LegGeom &G, const AlmLayout &L, int nb,
	const double2 *alm, int64_t alm_cstride, int64_t alm_bstride, double2 *leg, int64_t leg_bstride, cudaStream_t st, const LegStart *S)
{
	if (check_args(T, L, 0)) return 1;
	B2_REQUIRE(nb == leg_batch_size(T.spin) || (T.spin == 0 && nb == 2), "alm2leg_batch: unsupported batch size %d for spin %d", nb, T.spin);
	LegArgs A = make_args(T, G, L, 0, (double2*)alm, alm_cstride, leg, S);
	A.alm_bstride = alm_bstride; A.leg_bstride = leg_bstride;
	const int nm_launch = L.mmax + 1;
	// template arguments: R, NW, MINB, TL, NB (8 warps per SM: the accumulators of the batch take the registers)
	if (T.spin == 0) { if (nb == 4) LAUNCH(k_synth0b, 4, 2, 4, 64, 4); else LAUNCH(k_synth0b, 4, 2, 6, 64, 2); }
	else LAUNCH(k_synth2b, 4, 2, 4, 64, 2);
	B2_LAUNCH_CHECK();
	return 0;
}

// ------------------------------------------------------------------------------------ DFMA peak

// Sixteen independent chains d = fma(x, m, d) with m shared and x per chain: two operands to fetch per instruction, the
// most the register file delivers at full DFMA rate (d = fma(d, m, c) as used in round 1 measures 33-36 TFLOP/s instead of
// 36.9 because ptxas leaves some of its instructions with three fetched operands: scripts/ubench/ubench_dfma_operands.cu).
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double m_)
{
	double d[16], x[16];
	const double m = m_;
	#pragma unroll
	for (int i = 0; i < 16; i++) { d[i] = threadIdx.x*1e-9 + i; x[i] = 1e-9*(i + 1 + threadIdx.x); }
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int k = 0; k < 4; k++) {
			#pragma unroll
			for (int i = 0; i < 16; i++) d[i] = fma(x[i], m, d[i]);
		}
	}
	double sum = 0;
	#pragma unroll
	for (int i = 0; i < 16; i++) sum += d[i];
	out[blockIdx.x*blockDim.x + threadIdx.x] = sum;
}

int dfma_peak_gflops(double *out)
{
	int dev; B2_CHECK(cudaGetDevice(&dev));
	cudaDeviceProp pr; B2_CHECK(cudaGetDeviceProperties(&pr, dev));
	int nb = pr.multiProcessorCount*8, nt = 256, iters = 4096;
	DevBuf<double> buf; if (buf.alloc((size_t)nb*nt)) return 1;
	cudaEvent_t e0, e1; B2_CHECK(cudaEventCreate(&e0)); B2_CHECK(cudaEventCreate(&e1));
	double best = 0;
	for (int rep = 0; rep < 4; rep++) {
		B2_CHECK(cudaEventRecord(e0));
		k_dfma_peak<<<nb, nt>>>(buf.p, iters, 1.0000001);
		B2_CHECK(cudaEventRecord(e1));
		B2_CHECK(cudaEventSynchronize(e1));
		float ms; B2_CHECK(cudaEventElapsedTime(&ms, e0, e1));
		double gf = 2.0*64.0*iters*(double)nb*nt/(ms*1e-3)/1e9;
		if (rep > 0 && gf > best) best = gf;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	*out = best;
	return 0;
}
