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

struct FftDesc {
	int n;                 // transform length seen by the caller
	int nsmem;             // complex elements of shared memory the transform needs (n, or M for Bluestein)
	int nfac;
	int fac[FFT_MAX_FAC];  // factors of the in-memory transform (of n, or of M for Bluestein)
	int nt;                // length of the in-memory transform (n or M)
	int twmul;             // stride of w_nt in the twiddle table tw
	int ntab;
	const double2 *tw;     // tw[k] = exp(-2 pi i k / ntab), ntab a multiple of n (caller-visible table)
	const int *rev;        // position of X[k] after fft_smem (identity for Bluestein)
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
	static bool smooth(int64_t n);
	static int bluestein_len(int n);
	// shared-memory elements a transform of length n needs
	static int64_t smem_len(int64_t n) { return smooth(n) ? n : bluestein_len((int)n); }
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

// In-place FFT of s[0..n) by all `nthreads` threads of the CTA (ends with __syncthreads()); the caller
// must have synchronised after filling s and must provide d.nsmem elements per transform.
// Result X[k] is at s[d.rev[k]].  nbatch > 1 (smooth lengths only): transforms back to back in s.
template<bool INV> __device__ void fft_smem(double2 *s, const FftDesc &d, int tid, int nthreads, int nbatch = 1)
{
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
