// fft_smem.cuh -- in-place mixed-radix FFT on a shared-memory array.
// Building block of the ring FFTs (K3/K4), the theta weighting (K5) and the 2-D map FFT (K7);
// stands in for the pocketfft / ducc0.fft calls behind pixell/fft.py:33-60 and inside ducc's SHTs.
//
// Smooth lengths (prime factors <= FFT_MAX_RADIX): decimation-in-frequency passes; the transform
// leaves X[k] at position rev[k] (mixed-radix digit reversal, table built on the host) and consumers
// read through rev[] while they post-process or store, so there is no reordering pass.
// Radices 2, 3, 4, 5 are hard-wired; other prime factors go through a generic O(r^2) butterfly.
// Any other length: Bluestein's chirp-z algorithm inside the same shared-memory buffer -- multiply
// by the chirp, DIF transform of smooth length M >= 2n-1, multiply by the precomputed chirp
// spectrum (stored in digit-reversed order), DIT inverse transform (digit-reversed in, natural
// out), multiply by the chirp; rev[] is the identity for such plans.
#pragma once
#include "common.cuh"

#define FFT_MAX_FAC 20
#define FFT_MAX_RADIX 64
#define FFT_TWLO 128       // fine twiddle table: tw[0..128); coarse table: tw[128 i]

struct FftDesc {
	int n;                 // transform length seen by the caller
	int nsmem;             // complex elements of shared memory the transform needs (n, or M for Bluestein)
	int nfac;
	int fac[FFT_MAX_FAC];  // factors of the in-memory transform (of n, or of M for Bluestein)
	int nt;                // length of the in-memory transform (n or M)
	int twmul;             // stride of w_nt in the twiddle table tw
	int ntab;
	const double2 *tw;     // tw[k] = exp(-2 pi i k / ntab), ntab a multiple of n (caller-visible table)
	const int *rev;        // position of X[k] after fft_smem (identity for Bluestein); apply fft_pad() to it
	// when every factor is a power of two the digit reversal is a handful of shifts: bits per factor, one nibble each
	// (first factor lowest), 0 = use the table; rev_total = log2(nt).  Consumers call fft_rev(): no dependent global load
	// in front of their shared-memory reads.
	unsigned rev_bits; int rev_total;
	// fast path (all factors in {2,3,4,5,8,16}): register butterflies, padded shared memory, two-level
	// twiddle tables in shared memory (no global loads inside the passes)
	int fast;
	int pad_shift;         // element i lives at s[i + (i >> pad_shift)] (31: no padding)
	int ntw_hi;            // entries of the coarse twiddle table (ceil(ntab / FFT_TWLO))
	// Bluestein only
	int bluestein;
	const double2 *btw;    // exp(-2 pi i k / M), k < M
	const double2 *chirp;  // exp(-i pi k^2 / n), k < n
	const double2 *bhat;   // FFT_M(conj chirp, wrapped)/M, in digit-reversed order
};

struct FftTables {
	FftDesc d;
	DevBuf<double2> tw, btw, chirp, bhat;
	DevBuf<int> rev;
	// n: complex transform length, ntab: twiddle table length (multiple of n; 2n for packed real transforms)
	int build(int n, int ntab);
	// build() in two steps: prepare() computes the tables on the host (pure CPU work, safe to run for many plans in
	// parallel threads), commit() uploads them; build() skips the first step when prepare() already ran for (n, ntab)
	struct HostTables { std::vector<double2> tw, btw, chirp, bhat; std::vector<int> rev; int n = 0, ntab = 0; bool ready = false; } host;
	int prepare(int n, int ntab);
	int commit();
	static bool smooth(int64_t n);
	static int bluestein_len(int n);
	// shared-memory elements a transform of length n needs (padding included; twiddle tables not included)
	static int64_t smem_len(int64_t n);
	static bool fast_ok(int64_t n);
	static int pad_shift_of(int64_t n);
	// shared-memory elements of the twiddle tables of this plan (0 on the slow path)
	int twsm_len() const { return d.fast ? d.ntw_hi + FFT_TWLO : 0; }
	size_t bytes() const { return tw.bytes() + rev.bytes() + btw.bytes() + chirp.bytes() + bhat.bytes(); }
};

__device__ __forceinline__ double2 cj(double2 w, bool c) { if (c) w.y = -w.y; return w; }

// multiply by -i (forward) or +i (inverse)
template<bool INV> __device__ __forceinline__ double2 mul_mi(double2 a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

template<bool INV, bool DIT> __device__ __noinline__ void fft_generic_bfly(double2 *p, int r, int m, int j, int tws,
	const double2 *tw, int ntab)
{
	double2 u[FFT_MAX_RADIX];
	for (int q = 0; q < r; q++) { u[q] = p[q*m]; if (DIT && j && q) u[q] = cmul(u[q], cj(tw[tws*j*q], INV)); }
	int wr = ntab/r;
	for (int k = 0; k < r; k++) {
		double2 y = u[0];
		int t = 0;
		for (int q = 1; q < r; q++) {
			t += k; if (t >= r) t -= r;
			double2 w = cj(tw[wr*t], INV);
			y.x += u[q].x*w.x - u[q].y*w.y; y.y += u[q].x*w.y + u[q].y*w.x;
		}
		if (!DIT && j && k) y = cmul(y, cj(tw[tws*j*k], INV));
		p[k*m] = y;
	}
}

// One set of passes over s (length nt = prod fac) by all threads of the CTA; ends with __syncthreads().
// DIT = false: natural order in, digit-reversed out.  DIT = true: digit-reversed in, natural out.
// `nbatch` independent transforms may sit back to back in s.
template<bool INV, bool DIT> __device__ void fft_passes(double2 *s, int nt, int nfac, const int *fac,
	const double2 *tw, int twmul, int ntab, int tid, int nthreads, int nbatch)
{
	int Ls = DIT ? 1 : nt;
	for (int ff = 0; ff < nfac; ff++) {
		const int f = DIT ? nfac - 1 - ff : ff;
		const int r = fac[f];
		if (DIT) Ls *= r;
		const int m = Ls/r;
		const int nb = (nt/r)*nbatch;
		const int tws = twmul*(nt/Ls);
		for (int b = tid; b < nb; b += nthreads) {
			int blk = b/m, j = b - blk*m;
			double2 *p = s + (int64_t)blk*Ls + j;
			if (r == 4) {
				double2 u0 = p[0], u1 = p[m], u2 = p[2*m], u3 = p[3*m];
				if (DIT && j) { u1 = cmul(u1, cj(tw[tws*j], INV)); u2 = cmul(u2, cj(tw[2*tws*j], INV)); u3 = cmul(u3, cj(tw[3*tws*j], INV)); }
				double2 t0 = cadd(u0, u2), t1 = csub(u0, u2), t2 = cadd(u1, u3), t3 = mul_mi<INV>(csub(u1, u3));
				double2 y0 = cadd(t0, t2), y1 = cadd(t1, t3), y2 = csub(t0, t2), y3 = csub(t1, t3);
				if (!DIT && j) { y1 = cmul(y1, cj(tw[tws*j], INV)); y2 = cmul(y2, cj(tw[2*tws*j], INV)); y3 = cmul(y3, cj(tw[3*tws*j], INV)); }
				p[0] = y0; p[m] = y1; p[2*m] = y2; p[3*m] = y3;
			} else if (r == 2) {
				double2 u0 = p[0], u1 = p[m];
				if (DIT && j) u1 = cmul(u1, cj(tw[tws*j], INV));
				double2 y0 = cadd(u0, u1), y1 = csub(u0, u1);
				if (!DIT && j) y1 = cmul(y1, cj(tw[tws*j], INV));
				p[0] = y0; p[m] = y1;
			} else if (r == 3) {
				const double s3 = 0.86602540378443864676;
				double2 u0 = p[0], u1 = p[m], u2 = p[2*m];
				if (DIT && j) { u1 = cmul(u1, cj(tw[tws*j], INV)); u2 = cmul(u2, cj(tw[2*tws*j], INV)); }
				double2 t = cadd(u1, u2), dd = csub(u1, u2);
				double2 y0 = cadd(u0, t);
				double2 a = make_double2(u0.x - 0.5*t.x, u0.y - 0.5*t.y);
				double2 bb = cscale(mul_mi<INV>(dd), s3);
				double2 y1 = cadd(a, bb), y2 = csub(a, bb);
				if (!DIT && j) { y1 = cmul(y1, cj(tw[tws*j], INV)); y2 = cmul(y2, cj(tw[2*tws*j], INV)); }
				p[0] = y0; p[m] = y1; p[2*m] = y2;
			} else if (r == 5) {
				const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
				const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
				double2 u0 = p[0], u1 = p[m], u2 = p[2*m], u3 = p[3*m], u4 = p[4*m];
				if (DIT && j) {
					u1 = cmul(u1, cj(tw[tws*j], INV)); u2 = cmul(u2, cj(tw[2*tws*j], INV));
					u3 = cmul(u3, cj(tw[3*tws*j], INV)); u4 = cmul(u4, cj(tw[4*tws*j], INV));
				}
				double2 a1 = cadd(u1, u4), b1 = csub(u1, u4), a2 = cadd(u2, u3), b2 = csub(u2, u3);
				double2 y0 = make_double2(u0.x + a1.x + a2.x, u0.y + a1.y + a2.y);
				double2 e1 = make_double2(u0.x + c1*a1.x + c2*a2.x, u0.y + c1*a1.y + c2*a2.y);
				double2 e2 = make_double2(u0.x + c2*a1.x + c1*a2.x, u0.y + c2*a1.y + c1*a2.y);
				double2 o1 = mul_mi<INV>(make_double2(s1*b1.x + s2*b2.x, s1*b1.y + s2*b2.y));
				double2 o2 = mul_mi<INV>(make_double2(s2*b1.x - s1*b2.x, s2*b1.y - s1*b2.y));
				double2 y1 = cadd(e1, o1), y4 = csub(e1, o1), y2 = cadd(e2, o2), y3 = csub(e2, o2);
				if (!DIT && j) {
					y1 = cmul(y1, cj(tw[tws*j], INV)); y2 = cmul(y2, cj(tw[2*tws*j], INV));
					y3 = cmul(y3, cj(tw[3*tws*j], INV)); y4 = cmul(y4, cj(tw[4*tws*j], INV));
				}
				p[0] = y0; p[m] = y1; p[2*m] = y2; p[3*m] = y3; p[4*m] = y4;
			} else {
				fft_generic_bfly<INV, DIT>(p, r, m, j, tws, tw, ntab);
			}
		}
		__syncthreads();
		if (!DIT) Ls = m;
	}
}


// ------------------------------------------------------------------------------------ fast path

__device__ __forceinline__ int fft_pad(const FftDesc &d, int i) { return i + (i >> d.pad_shift); }
__device__ __forceinline__ int fft_rev(const FftDesc &d, int k)
{
	if (d.rev_bits == 0) return __ldg(&d.rev[k]);
	int pos = 0, sh = d.rev_total;
	unsigned bits = d.rev_bits;
	#pragma unroll
	for (int f = 0; f < 8; f++) {
		const int b = bits & 15;
		if (b) { sh -= b; pos |= (k & ((1 << b) - 1)) << sh; k >>= b; }
		bits >>= 4;
	}
	return pos;
}

// copy the two-level twiddle tables into shared memory (twsm: d.ntw_hi + FFT_TWLO elements); the caller
// synchronises before the first fft_smem call
__device__ __forceinline__ void fft_load_tw(double2 *twsm, const FftDesc &d, int tid, int nthreads)
{
	if (!d.fast) return;
	for (int i = tid; i < d.ntw_hi; i += nthreads) twsm[i] = d.tw[i*FFT_TWLO];
	for (int i = tid; i < FFT_TWLO; i += nthreads) twsm[d.ntw_hi + i] = d.tw[i < d.ntab ? i : 0];
}

// exp(-2 pi i e / ntab) (conjugated for INV) from the tables: one complex product
template<bool INV> __device__ __forceinline__ double2 fft_tw(const double2 *twsm, int nhi, int e)
{
	double2 a = twsm[e >> 7], b = twsm[nhi + (e & (FFT_TWLO - 1))];
	double2 w = make_double2(fma(a.x, b.x, -a.y*b.y), fma(a.x, b.y, a.y*b.x));
	if (INV) w.y = -w.y;
	return w;
}

template<bool INV> __device__ __forceinline__ void r4(double2 &a, double2 &b, double2 &c, double2 &d)
{
	double2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = mul_mi<INV>(csub(b, d));
	a = cadd(t0, t2); b = cadd(t1, t3); c = csub(t0, t2); d = csub(t1, t3);
}
// multiply by exp(-+ 2 pi i k / 16), k = 1, 2, 3 (forward sign for INV = false)
template<bool INV, int K> __device__ __forceinline__ double2 mul_w16(double2 a)
{
	const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
	double wr = K == 1 ? c1 : K == 2 ? h : s1, wi = K == 1 ? s1 : K == 2 ? h : c1;      // w = wr - i wi (forward)
	if (INV) return make_double2(fma(a.x, wr, -a.y*wi), fma(a.y, wr, a.x*wi));
	return make_double2(fma(a.x, wr, a.y*wi), fma(a.y, wr, -a.x*wi));
}

// in-register DFTs: natural order in, natural order out
template<bool INV> __device__ __forceinline__ void dft8(double2 (&u)[8])
{
	r4<INV>(u[0], u[2], u[4], u[6]);      // even samples -> E[k], k = 0..3 in u[0], u[2], u[4], u[6]
	r4<INV>(u[1], u[3], u[5], u[7]);      // odd samples  -> O[k]
	u[3] = mul_w16<INV, 2>(u[3]);         // O[1] w8
	u[5] = mul_mi<INV>(u[5]);             // O[2] w8^2 = -+i
	u[7] = mul_mi<INV>(mul_w16<INV, 2>(u[7]));      // O[3] w8^3
	double2 y[8];
	#pragma unroll
	for (int k = 0; k < 4; k++) { y[k] = cadd(u[2*k], u[2*k + 1]); y[k + 4] = csub(u[2*k], u[2*k + 1]); }
	#pragma unroll
	for (int k = 0; k < 8; k++) u[k] = y[k];
}

template<bool INV> __device__ __forceinline__ void dft16(double2 (&u)[16])
{
	// x[4 a + b]: DFT over a for each b, twiddle w16^(b k1), DFT over b; X[k1 + 4 k2]
	#pragma unroll
	for (int b = 0; b < 4; b++) r4<INV>(u[b], u[4 + b], u[8 + b], u[12 + b]);      // u[4 k1 + b] = T[b][k1]
	u[5]  = mul_w16<INV, 1>(u[5]);  u[6]  = mul_w16<INV, 2>(u[6]);  u[7]  = mul_w16<INV, 3>(u[7]);
	u[9]  = mul_w16<INV, 2>(u[9]);  u[10] = mul_mi<INV>(u[10]);     u[11] = mul_mi<INV>(mul_w16<INV, 2>(u[11]));
	u[13] = mul_w16<INV, 3>(u[13]); u[14] = mul_mi<INV>(mul_w16<INV, 2>(u[14]));
	{ double2 t = mul_w16<INV, 1>(u[15]); u[15] = make_double2(-t.x, -t.y); }      // w16^9 = -w16
	double2 y[16];
	#pragma unroll
	for (int k1 = 0; k1 < 4; k1++) {
		double2 a = u[4*k1], b = u[4*k1 + 1], c = u[4*k1 + 2], d = u[4*k1 + 3];
		r4<INV>(a, b, c, d);
		y[k1] = a; y[k1 + 4] = b; y[k1 + 8] = c; y[k1 + 12] = d;
	}
	#pragma unroll
	for (int k = 0; k < 16; k++) u[k] = y[k];
}

template<bool INV> __device__ __forceinline__ void dft3(double2 (&u)[3])
{
	const double s3 = 0.86602540378443864676;
	double2 t = cadd(u[1], u[2]), dd = csub(u[1], u[2]);
	double2 a = make_double2(u[0].x - 0.5*t.x, u[0].y - 0.5*t.y), bb = cscale(mul_mi<INV>(dd), s3);
	u[0] = cadd(u[0], t); u[1] = cadd(a, bb); u[2] = csub(a, bb);
}

template<bool INV> __device__ __forceinline__ void dft5(double2 (&u)[5])
{
	const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
	const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
	double2 a1 = cadd(u[1], u[4]), b1 = csub(u[1], u[4]), a2 = cadd(u[2], u[3]), b2 = csub(u[2], u[3]);
	double2 e1 = make_double2(u[0].x + c1*a1.x + c2*a2.x, u[0].y + c1*a1.y + c2*a2.y);
	double2 e2 = make_double2(u[0].x + c2*a1.x + c1*a2.x, u[0].y + c2*a1.y + c1*a2.y);
	double2 o1 = mul_mi<INV>(make_double2(s1*b1.x + s2*b2.x, s1*b1.y + s2*b2.y));
	double2 o2 = mul_mi<INV>(make_double2(s2*b1.x - s1*b2.x, s2*b1.y - s1*b2.y));
	u[0] = make_double2(u[0].x + a1.x + a2.x, u[0].y + a1.y + a2.y);
	u[1] = cadd(e1, o1); u[4] = csub(e1, o1); u[2] = cadd(e2, o2); u[3] = csub(e2, o2);
}

template<int R, bool INV> __device__ __forceinline__ void dft_small(double2 (&u)[R])
{
	if constexpr (R == 2) { double2 a = u[0]; u[0] = cadd(a, u[1]); u[1] = csub(a, u[1]); }
	else if constexpr (R == 3) dft3<INV>(u);
	else if constexpr (R == 4) r4<INV>(u[0], u[1], u[2], u[3]);
	else if constexpr (R == 5) dft5<INV>(u);
	else if constexpr (R == 8) dft8<INV>(u);
	else dft16<INV>(u);
}

// u[k] *= w1^k, k = 1..R-1; the powers are built by products of depth <= 4
template<int R> __device__ __forceinline__ void mul_powers(double2 (&u)[R], double2 w1)
{
	u[1] = cmul(u[1], w1);
	if constexpr (R > 2) {
		double2 w2 = cmul(w1, w1);
		u[2] = cmul(u[2], w2);
		if constexpr (R > 3) {
			double2 w3 = cmul(w2, w1);
			u[3] = cmul(u[3], w3);
			if constexpr (R > 4) {
				double2 w4 = cmul(w2, w2);
				u[4] = cmul(u[4], w4);
				if constexpr (R > 5) {
					double2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
					u[5] = cmul(u[5], w5); u[6] = cmul(u[6], w6); u[7] = cmul(u[7], w7);
					if constexpr (R > 8) {
						double2 w8 = cmul(w4, w4);
						u[8] = cmul(u[8], w8);
						u[9] = cmul(u[9], cmul(w8, w1)); u[10] = cmul(u[10], cmul(w8, w2)); u[11] = cmul(u[11], cmul(w8, w3));
						u[12] = cmul(u[12], cmul(w8, w4)); u[13] = cmul(u[13], cmul(w8, w5)); u[14] = cmul(u[14], cmul(w8, w6));
						u[15] = cmul(u[15], cmul(w8, w7));
					}
				}
			}
		}
	}
}

// one decimation-in-frequency pass of radix R over blocks of length Ls (all threads; ends with __syncthreads())
template<int R, bool INV> __device__ __forceinline__ void fft_pass_fast(double2 *s, const FftDesc &d, int Ls,
	int tid, int nthreads, int nbatch, const double2 *twsm)
{
	const int m = Ls/R, nbf = d.nt/R, total = nbf*nbatch, tws = d.twmul*(d.nt/Ls), nhi = d.ntw_hi;
	for (int b = tid; b < total; b += nthreads) {
		const int line = b/nbf, bb = b - line*nbf;
		const int blk = bb/m, j = bb - blk*m;
		double2 *S = s + (int64_t)line*d.nsmem;
		const int base = blk*Ls + j;
		double2 u[R];
		#pragma unroll
		for (int q = 0; q < R; q++) u[q] = S[fft_pad(d, base + q*m)];
		dft_small<R, INV>(u);
		if (j) mul_powers<R>(u, fft_tw<INV>(twsm, nhi, tws*j));      // u[k] *= w^k, w = exp(-+ 2 pi i j / Ls)
		#pragma unroll
		for (int q = 0; q < R; q++) S[fft_pad(d, base + q*m)] = u[q];
	}
	__syncthreads();
}

template<bool INV> __device__ void fft_fast(double2 *s, const FftDesc &d, int tid, int nthreads, int nbatch, const double2 *twsm)
{
	int Ls = d.nt;
	for (int f = 0; f < d.nfac; f++) {
		const int r = d.fac[f];
		switch (r) {
			case 2:  fft_pass_fast<2, INV>(s, d, Ls, tid, nthreads, nbatch, twsm); break;
			case 3:  fft_pass_fast<3, INV>(s, d, Ls, tid, nthreads, nbatch, twsm); break;
			case 4:  fft_pass_fast<4, INV>(s, d, Ls, tid, nthreads, nbatch, twsm); break;
			case 5:  fft_pass_fast<5, INV>(s, d, Ls, tid, nthreads, nbatch, twsm); break;
			case 8:  fft_pass_fast<8, INV>(s, d, Ls, tid, nthreads, nbatch, twsm); break;
			default: fft_pass_fast<16, INV>(s, d, Ls, tid, nthreads, nbatch, twsm); break;
		}
		Ls /= r;
	}
}

// In-place FFT of s[0..n) by all `nthreads` threads of the CTA (ends with __syncthreads()); the caller
// must have synchronised after filling s and must provide d.nsmem elements per transform.
// Result X[k] is at s[d.rev[k]].  nbatch > 1 (smooth lengths only): transforms back to back in s.
// Fast plans (d.fast): element i of transform b lives at s[b*d.nsmem + fft_pad(d, i)] and twsm must point at the
// twiddle tables filled by fft_load_tw; other plans ignore twsm and use no padding.
template<bool INV> __device__ void fft_smem(double2 *s, const FftDesc &d, int tid, int nthreads, int nbatch = 1, const double2 *twsm = nullptr)
{
	if (d.fast) { fft_fast<INV>(s, d, tid, nthreads, nbatch, twsm); return; }
	if (!d.bluestein) {
		fft_passes<INV, false>(s, d.nt, d.nfac, d.fac, d.tw, d.twmul, d.ntab, tid, nthreads, nbatch);
		return;
	}
	const int n = d.n, M = d.nt;
	// inverse transform = conj(forward(conj x))
	for (int j = tid; j < M; j += nthreads) {
		double2 v = make_double2(0, 0);
		if (j < n) { v = s[j]; if (INV) v.y = -v.y; v = cmul(v, d.chirp[j]); }
		s[j] = v;
	}
	__syncthreads();
	fft_passes<false, false>(s, M, d.nfac, d.fac, d.btw, 1, M, tid, nthreads, 1);
	for (int j = tid; j < M; j += nthreads) s[j] = cmul(s[j], d.bhat[j]);
	__syncthreads();
	fft_passes<true, true>(s, M, d.nfac, d.fac, d.btw, 1, M, tid, nthreads, 1);
	for (int k = tid; k < n; k += nthreads) {
		double2 v = cmul(s[k], d.chirp[k]);
		if (INV) v.y = -v.y;
		s[k] = v;
	}
	__syncthreads();
}
