// fft_smem.cuh -- in-place mixed-radix decimation-in-frequency FFT on a shared-memory array.
// Building block of the ring FFTs (K3/K4), the theta resampling (K5) and the 2-D map FFT (K7);
// stands in for the pocketfft / ducc0.fft calls behind pixell/fft.py:33-60 and inside ducc's SHTs.
//
// The transform leaves X[k] at position rev[k] (mixed-radix digit reversal, table built on the
// host); consumers read through rev[] while they post-process or store, so no reordering pass.
// Radices 2, 3, 4, 5 are hard-wired; any other prime factor <= FFT_MAX_RADIX goes through a generic
// O(r^2) butterfly (rare: nphi = 61 in the reference's round-trip test).
#pragma once
#include "common.cuh"

#define FFT_MAX_FAC 16
#define FFT_MAX_RADIX 64

struct FftDesc {
	int n;                 // transform length
	int nfac;
	int fac[FFT_MAX_FAC];
	int twmul;             // ntab / n: stride of w_n in the twiddle table
	int ntab;
	const double2 *tw;     // tw[k] = exp(-2 pi i k / ntab)
	const int *rev;        // position of X[k] after the DIF passes
};

struct FftTables {
	FftDesc d;
	DevBuf<double2> tw;
	DevBuf<int> rev;
	// n: complex transform length, ntab: twiddle table length (multiple of n; 2n for packed real transforms)
	int build(int n, int ntab);
	static bool supported(int64_t n);
	size_t bytes() const { return tw.bytes() + rev.bytes(); }
};

template<bool INV> __device__ __forceinline__ double2 twid(const FftDesc &d, int idx)
{
	double2 w = d.tw[idx];
	if (INV) w.y = -w.y;
	return w;
}

// multiply by -i (forward) or +i (inverse)
template<bool INV> __device__ __forceinline__ double2 mul_mi(double2 a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

template<bool INV> __device__ __noinline__ void fft_generic_bfly(double2 *p, int r, int m, int j, int tws, const FftDesc &d)
{
	double2 u[FFT_MAX_RADIX];
	for (int q = 0; q < r; q++) u[q] = p[q*m];
	int wr = d.ntab/r;
	for (int k = 0; k < r; k++) {
		double2 y = u[0];
		int t = 0;
		for (int q = 1; q < r; q++) {
			t += k; if (t >= r) t -= r;
			double2 w = twid<INV>(d, wr*t);
			y.x += u[q].x*w.x - u[q].y*w.y; y.y += u[q].x*w.y + u[q].y*w.x;
		}
		if (j && k) y = cmul(y, twid<INV>(d, tws*j*k));
		p[k*m] = y;
	}
}

// In-place FFT of s[0..n) by all `nthreads` threads of the CTA (ends with __syncthreads()).
// `nfft` independent transforms of length d.n may be laid out back to back in s (batch in one CTA).
template<bool INV> __device__ void fft_smem(double2 *s, const FftDesc &d, int tid, int nthreads, int nbatch = 1)
{
	int Ls = d.n;
	for (int f = 0; f < d.nfac; f++) {
		const int r = d.fac[f], m = Ls/r;
		const int nb = (d.n/r)*nbatch;
		const int tws = d.twmul*(d.n/Ls);
		for (int b = tid; b < nb; b += nthreads) {
			int blk = b/m, j = b - blk*m;
			double2 *p = s + (int64_t)blk*Ls + j;
			if (r == 4) {
				double2 u0 = p[0], u1 = p[m], u2 = p[2*m], u3 = p[3*m];
				double2 t0 = cadd(u0, u2), t1 = csub(u0, u2), t2 = cadd(u1, u3), t3 = mul_mi<INV>(csub(u1, u3));
				double2 y0 = cadd(t0, t2), y1 = cadd(t1, t3), y2 = csub(t0, t2), y3 = csub(t1, t3);
				if (j) {
					double2 w1 = twid<INV>(d, tws*j), w2 = twid<INV>(d, 2*tws*j), w3 = twid<INV>(d, 3*tws*j);
					y1 = cmul(y1, w1); y2 = cmul(y2, w2); y3 = cmul(y3, w3);
				}
				p[0] = y0; p[m] = y1; p[2*m] = y2; p[3*m] = y3;
			} else if (r == 2) {
				double2 u0 = p[0], u1 = p[m];
				double2 y0 = cadd(u0, u1), y1 = csub(u0, u1);
				if (j) y1 = cmul(y1, twid<INV>(d, tws*j));
				p[0] = y0; p[m] = y1;
			} else if (r == 3) {
				const double s3 = 0.86602540378443864676;
				double2 u0 = p[0], u1 = p[m], u2 = p[2*m];
				double2 t = cadd(u1, u2), dd = csub(u1, u2);
				double2 y0 = cadd(u0, t);
				double2 a = make_double2(u0.x - 0.5*t.x, u0.y - 0.5*t.y);
				double2 bb = cscale(mul_mi<INV>(dd), s3);
				double2 y1 = cadd(a, bb), y2 = csub(a, bb);
				if (j) { y1 = cmul(y1, twid<INV>(d, tws*j)); y2 = cmul(y2, twid<INV>(d, 2*tws*j)); }
				p[0] = y0; p[m] = y1; p[2*m] = y2;
			} else if (r == 5) {
				const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
				const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
				double2 u0 = p[0], u1 = p[m], u2 = p[2*m], u3 = p[3*m], u4 = p[4*m];
				double2 a1 = cadd(u1, u4), b1 = csub(u1, u4), a2 = cadd(u2, u3), b2 = csub(u2, u3);
				double2 y0 = make_double2(u0.x + a1.x + a2.x, u0.y + a1.y + a2.y);
				double2 e1 = make_double2(u0.x + c1*a1.x + c2*a2.x, u0.y + c1*a1.y + c2*a2.y);
				double2 e2 = make_double2(u0.x + c2*a1.x + c1*a2.x, u0.y + c2*a1.y + c1*a2.y);
				double2 o1 = mul_mi<INV>(make_double2(s1*b1.x + s2*b2.x, s1*b1.y + s2*b2.y));
				double2 o2 = mul_mi<INV>(make_double2(s2*b1.x - s1*b2.x, s2*b1.y - s1*b2.y));
				double2 y1 = cadd(e1, o1), y4 = csub(e1, o1), y2 = cadd(e2, o2), y3 = csub(e2, o2);
				if (j) {
					y1 = cmul(y1, twid<INV>(d, tws*j)); y2 = cmul(y2, twid<INV>(d, 2*tws*j));
					y3 = cmul(y3, twid<INV>(d, 3*tws*j)); y4 = cmul(y4, twid<INV>(d, 4*tws*j));
				}
				p[0] = y0; p[m] = y1; p[2*m] = y2; p[3*m] = y3; p[4*m] = y4;
			} else {
				fft_generic_bfly<INV>(p, r, m, j, tws, d);
			}
		}
		__syncthreads();
		Ls = m;
	}
}
