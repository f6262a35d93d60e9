// fft2d.cu -- K7: batched 1-D / 2-D DFTs over strided arrays: the engine behind the pixell.fft plug-in
// (reference pixell/fft.py:8-113 engines, :133-209 fft/ifft/rfft/irfft; pixell/enmap.py:1307-1337 enmap.fft/ifft).
//
// A transform is one or two "axis passes".  A pass transforms every line along one axis of a strided
// array: a CTA loads a tile of nb lines into shared memory (coalesced along whichever of the transform
// axis / the innermost other axis is contiguous), runs fft_smem on them and stores them with the output
// strides.  Lines longer than the shared-memory capacity are split decimation-in-frequency over P CTAs
// (each forms y_p[j] = w_n^{jp} sum_q x[j + q n/P] w_P^{qp} while loading and owns the outputs k = p mod P).
// Real transforms ride on the same kernel through the element loaders / storers: r2c loads reals and
// stores k <= n/2, c2r loads the Hermitian extension of the half spectrum and stores real parts.
// Normalisation (and any caller-supplied factor) is a multiplication in the store of the last pass.
#include "../../include/b200sht.h"
#include "fft_smem.cuh"
#include <algorithm>
#include <stdlib.h>
#include <memory>

enum { LK_C128, LK_C64, LK_F64, LK_F32, LK_H128, LK_H64 };
enum { SK_C128, SK_C64, SK_HC128, SK_HC64, SK_F64, SK_F32 };

struct AxisArgs {
	FftDesc d;
	int n, P, nl, nb, jfast, lk, sk, lstride, twoff;   // lstride: shared-memory elements per line; twoff: twiddle tables
	int64_t n_in, n_o1, n_o2;                    // lines are enumerated by (inner, outer1, outer2) indices
	int64_t is_t, is_in, is_o1, is_o2;           // input strides in elements of the input type
	int64_t os_t, os_in, os_o1, os_o2;
	const void *in; void *out; double scale;
};

__device__ __forceinline__ double2 fft_ld(const AxisArgs &A, int64_t base, int j)
{
	switch (A.lk) {
		case LK_C128: return ((const double2*)A.in)[base + j*A.is_t];
		case LK_C64: { float2 v = ((const float2*)A.in)[base + j*A.is_t]; return make_double2(v.x, v.y); }
		case LK_F64: return make_double2(((const double*)A.in)[base + j*A.is_t], 0.0);
		case LK_F32: return make_double2((double)((const float*)A.in)[base + j*A.is_t], 0.0);
		case LK_H128: {
			bool up = 2*j > A.n; int jj = up ? A.n - j : j;
			double2 v = ((const double2*)A.in)[base + jj*A.is_t]; if (up) v.y = -v.y; return v;
		}
		default: {
			bool up = 2*j > A.n; int jj = up ? A.n - j : j;
			float2 v = ((const float2*)A.in)[base + jj*A.is_t]; return make_double2(v.x, up ? -(double)v.y : (double)v.y);
		}
	}
}

__device__ __forceinline__ void fft_st(const AxisArgs &A, int64_t base, int k, double2 v)
{
	v.x *= A.scale; v.y *= A.scale;
	switch (A.sk) {
		case SK_C128: ((double2*)A.out)[base + k*A.os_t] = v; break;
		case SK_C64: ((float2*)A.out)[base + k*A.os_t] = make_float2((float)v.x, (float)v.y); break;
		case SK_HC128: if (2*k <= A.n) ((double2*)A.out)[base + k*A.os_t] = v; break;
		case SK_HC64: if (2*k <= A.n) ((float2*)A.out)[base + k*A.os_t] = make_float2((float)v.x, (float)v.y); break;
		case SK_F64: ((double*)A.out)[base + k*A.os_t] = v.x; break;
		default: ((float*)A.out)[base + k*A.os_t] = (float)v.x; break;
	}
}

template<bool INV> __global__ void k_fft_axis(AxisArgs A)
{
	extern __shared__ __align__(16) double2 s[];
	const int tid = threadIdx.x, T = blockDim.x, p = blockIdx.y;
	const int nb = A.nb, nl = A.nl, P = A.P, ls = A.lstride;
	const int64_t ntile = (A.n_in + nb - 1)/nb;
	const int64_t outer = blockIdx.x/ntile, i0 = (blockIdx.x % ntile)*nb;
	const int64_t o1 = outer/A.n_o2, o2 = outer % A.n_o2;
	const int nbv = (int)min((int64_t)nb, A.n_in - i0);
	const int64_t bin = o1*A.is_o1 + o2*A.is_o2 + i0*A.is_in, bout = o1*A.os_o1 + o2*A.os_o2 + i0*A.os_in;
	const int tot = nb*nl;
	const double2 *twsm = s + A.twoff;
	fft_load_tw(s + A.twoff, A.d, tid, T);
	for (int idx = tid; idx < tot; idx += T) {
		int line, j;
		if (A.jfast) { line = idx/nl; j = idx - line*nl; } else { j = idx/nb; line = idx - j*nb; }
		double2 acc = make_double2(0, 0);
		if (line < nbv) {
			const int64_t b = bin + line*A.is_in;
			if (P == 1) acc = fft_ld(A, b, j);
			else {
				for (int q = 0; q < P; q++) {
					double2 v = fft_ld(A, b, j + q*nl);
					int e = (q*p) % P;
					if (e) v = cmul(v, cj(A.d.tw[(A.n/P)*e], INV));
					acc = cadd(acc, v);
				}
				if (p) acc = cmul(acc, cj(A.d.tw[j*p], INV));
			}
		}
		s[line*ls + fft_pad(A.d, j)] = acc;
	}
	__syncthreads();
	fft_smem<INV>(s, A.d, tid, T, nb, twsm);
	for (int idx = tid; idx < tot; idx += T) {
		int line, kk;
		if (A.jfast) { line = idx/nl; kk = idx - line*nl; } else { kk = idx/nb; line = idx - kk*nb; }
		if (line < nbv) fft_st(A, bout + line*A.os_in, p + P*kk, s[line*ls + fft_pad(A.d, __ldg(&A.d.rev[kk]))]);
	}
}

// Specialised passes (P in {1, 2, 4, 8} known at compile time, loads of a thread issued together):
//   MODE 0  complex lines (any loader / storer of the generic kernel)
//   MODE 1  packed real-to-complex, forward: a real line of even length n is read as nc = n/2 complex numbers
//           z_j = x_2j + i x_2j+1 (one 16-byte load each), transformed with length nc, and untangled on the way out:
//           X_k = [(Z_k + conj Z_{nc-k}) - i w_n^k (Z_k - conj Z_{nc-k})]/2, k = 0..nc (needs P <= 2: both partners
//           of a pair then live in the same CTA)
//   MODE 2  packed complex-to-real, backward: Z_k = (X_k + conj X_{nc-k}) + i conj(w_n^k) (X_k - conj X_{nc-k}) is
//           formed while loading, the inverse transform of length nc yields (x_2k, x_2k+1) pairs
enum { FM_C2C, FM_R2C_PACKED, FM_C2R_PACKED };

template<bool INV, int MODE, int P> __global__ void __launch_bounds__(512) k_fft_axis2(AxisArgs A)
{
	extern __shared__ __align__(16) double2 s[];
	// the P CTAs of one tile are neighbours in the grid, so they run together and all but the first read the lines from L2
	const int tid = threadIdx.x, T = blockDim.x, p = blockIdx.x % P;
	const int64_t tile = blockIdx.x/P;
	const int nb = A.nb, nl = A.nl, ls = A.lstride;
	const int nc = nl*P;                                   // complex transform length (n, or n/2 for the packed modes)
	const int64_t ntile = (A.n_in + nb - 1)/nb;
	const int64_t outer = tile/ntile, i0 = (tile % ntile)*nb;
	const int64_t o1 = outer/A.n_o2, o2 = outer % A.n_o2;
	const int nbv = (int)min((int64_t)nb, A.n_in - i0);
	const int64_t bin = o1*A.is_o1 + o2*A.is_o2 + i0*A.is_in, bout = o1*A.os_o1 + o2*A.os_o2 + i0*A.os_in;
	const int tot = nb*nl;
	const int twq = A.d.twmul/P;                            // table step of w_nc
	const double2 *twsm = s + A.twoff;
	fft_load_tw(s + A.twoff, A.d, tid, T);
	double2 wq[P];
	#pragma unroll
	for (int q = 0; q < P; q++) wq[q] = cj(A.d.tw[(A.d.ntab/P)*((q*p) % P)], INV);
	#pragma unroll 4
	for (int idx = tid; idx < tot; idx += T) {
		int line, j;
		if (A.jfast) { line = idx/nl; j = idx - line*nl; } else { j = idx/nb; line = idx - j*nb; }
		double2 v[P];
		const int64_t b = bin + line*A.is_in;
		#pragma unroll
		for (int q = 0; q < P; q++) {
			const int jj = j + q*nl;
			if (line >= nbv) v[q] = make_double2(0, 0);
			else if (MODE == FM_C2C) v[q] = ((const double2*)A.in)[b + jj*A.is_t];      // complex128 only (see run_pass)
			else if (MODE == FM_R2C_PACKED) {
				if (A.lk == LK_F64) v[q] = ((const double2*)((const double*)A.in + b))[jj];
				else { float2 f = ((const float2*)((const float*)A.in + b))[jj]; v[q] = make_double2(f.x, f.y); }
			} else {
				const int jm = nc - jj;
				double2 xa, xb;
				if (A.lk == LK_C128) { xa = ((const double2*)A.in)[b + jj*A.is_t]; xb = ((const double2*)A.in)[b + jm*A.is_t]; }
				else { float2 fa = ((const float2*)A.in)[b + jj*A.is_t], fb = ((const float2*)A.in)[b + jm*A.is_t]; xa = make_double2(fa.x, fa.y); xb = make_double2(fb.x, fb.y); }
				double2 sm = make_double2(xa.x + xb.x, xa.y - xb.y), df = make_double2(xa.x - xb.x, xa.y + xb.y);
				double2 w = __ldg(&A.d.tw[jj]); w.y = -w.y;                // e^{+2 pi i jj/n}
				double2 u = cmul(df, w);
				v[q] = make_double2(sm.x - u.y, sm.y + u.x);
			}
		}
		double2 acc = v[0];
		#pragma unroll
		for (int q = 1; q < P; q++) acc = cadd(acc, p ? cmul(v[q], wq[q]) : v[q]);
		if (P > 1 && p) acc = cmul(acc, cj(__ldg(&A.d.tw[twq*j*p]), INV));
		s[line*ls + fft_pad(A.d, j)] = acc;
	}
	__syncthreads();
	fft_smem<INV>(s, A.d, tid, T, nb, twsm);
	#pragma unroll 4
	for (int idx = tid; idx < tot; idx += T) {
		int line, kk;
		if (A.jfast) { line = idx/nl; kk = idx - line*nl; } else { kk = idx/nb; line = idx - kk*nb; }
		if (line >= nbv) continue;
		const int k = p + P*kk;
		const double2 x = s[line*ls + fft_pad(A.d, __ldg(&A.d.rev[kk]))];
		const int64_t bo = bout + line*A.os_in;
		if (MODE == FM_C2C) ((double2*)A.out)[bo + k*A.os_t] = make_double2(x.x*A.scale, x.y*A.scale);
		else if (MODE == FM_C2R_PACKED) {
			if (A.sk == SK_F64) ((double2*)((double*)A.out + bo))[k] = make_double2(x.x*A.scale, x.y*A.scale);
			else ((float2*)((float*)A.out + bo))[k] = make_float2((float)(x.x*A.scale), (float)(x.y*A.scale));
		} else {
			// partner nc - k has the same residue mod P (P <= 2)
			const int kp = k ? nc - k : 0;
			const double2 y = s[line*ls + fft_pad(A.d, __ldg(&A.d.rev[(kp - p)/P]))];
			double2 sm = make_double2(x.x + y.x, x.y - y.y), df = make_double2(x.x - y.x, x.y + y.y);
			double2 u = cmul(df, __ldg(&A.d.tw[k]));                    // w_n^k (Z_k - conj Z_{nc-k})
			double2 X = make_double2(0.5*(sm.x + u.y), 0.5*(sm.y - u.x));
			const int savek = A.sk;      // complex storer without the k <= n/2 test
			if (savek == SK_HC128 || savek == SK_C128) ((double2*)A.out)[bo + k*A.os_t] = make_double2(X.x*A.scale, X.y*A.scale);
			else ((float2*)A.out)[bo + k*A.os_t] = make_float2((float)(X.x*A.scale), (float)(X.y*A.scale));
			if (k == 0) {
				double2 Xn = make_double2((x.x - x.y)*A.scale, 0.0);
				if (savek == SK_HC128 || savek == SK_C128) ((double2*)A.out)[bo + nc*A.os_t] = Xn;
				else ((float2*)A.out)[bo + nc*A.os_t] = make_float2((float)Xn.x, 0.f);
			}
		}
	}
}

// ------------------------------------------------------------------------------------ plan

struct ArrayDesc { int64_t stride[4]; int kind; };      // element strides; kind = LK_* (as input) / SK_* (as output)

struct AxisPass {
	int axis = 0, n = 0, P = 1, nl = 0, nb = 1, threads = 256;
	size_t smem = 0;
	FftTables tab;
};

struct b2_fft_plan {
	int ndim = 0, naxes = 0, axes[2] = {0, 0}, kind = 0, dtype = 0;
	int64_t shape[4] = {1, 1, 1, 1}, cshape[4] = {1, 1, 1, 1};      // full (real) shape, half-spectrum shape
	int64_t istride[4] = {0, 0, 0, 0}, ostride[4] = {0, 0, 0, 0};
	int64_t in_span = 0, out_span = 0;                              // elements touched in the caller's arrays
	AxisPass pass[2];
	AxisPass packed;       // real axis as a half-length complex transform (even length, line fits two CTAs); n = 0: unavailable
	DevBuf<char> work, stage_in, stage_out;
};

static const size_t FFT_TILE_ELEMS = 6144;
// shared memory one CTA may use (B2_FFT_SMEM_KB overrides, for tuning)
static size_t fft_smem_max() { const char *e = getenv("B2_FFT_SMEM_KB"); return (size_t)(e ? atoi(e) : 200)*1024; }
#define FFT_SMEM_MAX fft_smem_max()

// strided: the axis is not the contiguous one, so a CTA should hold at least two neighbouring lines (full 32-byte sectors)
static int setup_pass(AxisPass &ps, int axis, int n, bool can_batch, bool strided)
{
	ps.axis = axis; ps.n = n;
	int P = 1;
	const int minb = (strided && can_batch) ? 2 : 1;
	while ((size_t)(minb*FftTables::smem_len(n/P) + n/FFT_TWLO + FFT_TWLO + 1)*sizeof(double2) > FFT_SMEM_MAX) {
		int np = P + 1;
		while (np <= 64 && n % np) np++;
		B2_REQUIRE(np <= 64, "fft: a transform of length %d does not fit in shared memory (no usable split)", n);
		P = np;
	}
	ps.P = P; ps.nl = n/P;
	if (ps.tab.build(ps.nl, n)) return 1;
	ps.nb = 1;
	if (can_batch && FftTables::smooth(ps.nl)) ps.nb = (int)std::max<size_t>(minb, FFT_TILE_ELEMS/FftTables::smem_len(ps.nl));
	return 0;
}

// the real axis of even length n as a complex transform of length n/2 over at most two CTAs
static int setup_packed(AxisPass &ps, int axis, int n)
{
	ps.n = 0;
	if (n % 2 || n < 4) return 0;
	const int nc = n/2;
	for (int P = 1; P <= 2; P++) {
		if (nc % P) continue;
		if ((size_t)(FftTables::smem_len(nc/P) + n/FFT_TWLO + FFT_TWLO + 1)*sizeof(double2) > FFT_SMEM_MAX) continue;
		if (!FftTables::fast_ok(nc/P)) return 0;
		ps.axis = axis; ps.n = n; ps.P = P; ps.nl = nc/P;
		if (ps.tab.build(ps.nl, n)) return 1;
		ps.nb = (int)std::max<size_t>(1, FFT_TILE_ELEMS/FftTables::smem_len(ps.nl));
		return 0;
	}
	return 0;
}

extern "C" int b2_fft_plan_create(b2_fft_plan **out, int ndim, const int64_t *shape, const int64_t *istride,
	const int64_t *ostride, int naxes, const int *axes, int kind, int dtype)
{
	B2_REQUIRE(out && shape && istride && ostride && axes, "fft plan: null argument");
	B2_REQUIRE(ndim >= 1 && ndim <= 4, "fft plan: 1 to 4 dimensions are supported (got %d)", ndim);
	B2_REQUIRE(naxes == 1 || naxes == 2, "fft plan: 1 or 2 transform axes are supported (got %d)", naxes);
	B2_REQUIRE(kind == B2_FFT_C2C || kind == B2_FFT_R2C || kind == B2_FFT_C2R, "fft plan: bad kind");
	B2_REQUIRE(dtype == B2_F64 || dtype == B2_F32, "fft plan: bad dtype");
	std::unique_ptr<b2_fft_plan> p(new b2_fft_plan());
	p->ndim = ndim; p->naxes = naxes; p->kind = kind; p->dtype = dtype;
	for (int i = 0; i < naxes; i++) {
		int a = axes[i] < 0 ? axes[i] + ndim : axes[i];
		B2_REQUIRE(a >= 0 && a < ndim, "fft plan: axis %d out of range", axes[i]);
		B2_REQUIRE(i == 0 || a != p->axes[0], "fft plan: repeated axis");
		p->axes[i] = a;
	}
	const int last = p->axes[naxes - 1];
	for (int d = 0; d < ndim; d++) {
		B2_REQUIRE(shape[d] >= 1 && shape[d] < (1LL << 31), "fft plan: bad extent %lld", (long long)shape[d]);
		B2_REQUIRE(istride[d] >= 0 && ostride[d] >= 0, "fft plan: negative strides are not supported");
		p->shape[d] = shape[d]; p->cshape[d] = shape[d];
		p->istride[d] = istride[d]; p->ostride[d] = ostride[d];
	}
	if (kind != B2_FFT_C2C) p->cshape[last] = shape[last]/2 + 1;
	const int64_t *ish = kind == B2_FFT_C2R ? p->cshape : p->shape, *osh = kind == B2_FFT_R2C ? p->cshape : p->shape;
	p->in_span = 1; p->out_span = 1;
	for (int d = 0; d < ndim; d++) { p->in_span += (ish[d] - 1)*istride[d]; p->out_span += (osh[d] - 1)*ostride[d]; }
	// pass 0 = the real (last listed) axis for r2c, the other axis for c2r, so that real data is touched once
	int order[2] = {p->axes[naxes - 1], p->axes[0]};
	if (kind == B2_FFT_C2R && naxes == 2) std::swap(order[0], order[1]);
	for (int i = 0; i < naxes; i++)
		if (setup_pass(p->pass[i], order[i], (int)p->shape[order[i]], true, istride[order[i]] != 1 || ostride[order[i]] != 1)) return 1;
	if (kind != B2_FFT_C2C && setup_packed(p->packed, last, (int)p->shape[last])) return 1;
	*out = p.release();
	return 0;
}

extern "C" void b2_fft_plan_destroy(b2_fft_plan *plan) { delete plan; }

// ------------------------------------------------------------------------------------ execution

// run one pass over `axis` of an array with extents dims[], reading src (strides ss, loader lk) and writing dst
template<bool INV, int MODE, int P> static int launch_axis2(const AxisArgs &A, dim3 grid, int threads, size_t smem, cudaStream_t st)
{
	if (smem > 48*1024) B2_CHECK(cudaFuncSetAttribute(k_fft_axis2<INV, MODE, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_fft_axis2<INV, MODE, P><<<grid, threads, smem, st>>>(A);
	B2_LAUNCH_CHECK();
	return 0;
}

static int run_pass(b2_fft_plan *p, AxisPass &ps, const int64_t *dims, const void *src, const int64_t *ss, int lk,
	void *dst, const int64_t *ds, int sk, bool inverse, double scale, cudaStream_t st, int mode = FM_C2C)
{
	AxisArgs A;
	A.d = ps.tab.d; A.n = ps.n; A.P = ps.P; A.nl = ps.nl; A.lk = lk; A.sk = sk; A.scale = scale;
	A.in = src; A.out = dst;
	// the other dimensions: the one with the smallest input stride becomes the tile ("inner") dimension
	int others[3], no = 0;
	for (int d = 0; d < p->ndim; d++) if (d != ps.axis) others[no++] = d;
	std::sort(others, others + no, [&](int a, int b) {
		bool ta = dims[a] > 1, tb = dims[b] > 1;
		if (ta != tb) return ta;
		return ss[a] < ss[b];
	});
	int64_t ext[3] = {1, 1, 1}, sin[3] = {0, 0, 0}, sout[3] = {0, 0, 0};
	for (int i = 0; i < no; i++) { ext[i] = dims[others[i]]; sin[i] = ss[others[i]]; sout[i] = ds[others[i]]; }
	A.n_in = ext[0]; A.n_o1 = ext[1]; A.n_o2 = ext[2];
	A.is_t = ss[ps.axis]; A.is_in = sin[0]; A.is_o1 = sin[1]; A.is_o2 = sin[2];
	A.os_t = ds[ps.axis]; A.os_in = sout[0]; A.os_o1 = sout[1]; A.os_o2 = sout[2];
	A.jfast = (A.n_in == 1 || A.is_t <= A.is_in) ? 1 : 0;
	A.nb = (int)std::min<int64_t>(ps.nb, A.n_in);
	if (mode != FM_C2C) A.jfast = 1;
	A.lstride = ps.tab.d.nsmem;
	A.twoff = A.nb*ps.tab.d.nsmem;
	size_t smem = sizeof(double2)*(size_t)(A.twoff + ps.tab.twsm_len());
	int threads = (int)std::min<int64_t>(512, std::max<int64_t>(64, b2_round_up((int64_t)A.nb*ps.nl/4, 32)));
	int64_t ntile = (A.n_in + A.nb - 1)/A.nb;
	int64_t nblk = ntile*A.n_o1*A.n_o2;
	B2_REQUIRE(nblk < (1LL << 31), "fft: too many lines for one launch");
	dim3 grid((unsigned)nblk, ps.P);
	const bool plain128 = (mode != FM_C2C) || (lk == LK_C128 && sk == SK_C128);
	if (ps.tab.d.fast && plain128 && (ps.P == 1 || ps.P == 2 || ps.P == 4 || ps.P == 8) && nblk*ps.P < (1LL << 31)) {
		grid = dim3((unsigned)(nblk*ps.P), 1);
		threads = (int)std::min<int64_t>(512, std::max<int64_t>(64, b2_round_up((int64_t)A.nb*ps.nl/16, 32)));
		#define AX2(INV, MODE) (ps.P == 1 ? launch_axis2<INV, MODE, 1>(A, grid, threads, smem, st) : ps.P == 2 ? launch_axis2<INV, MODE, 2>(A, grid, threads, smem, st) : \
			ps.P == 4 ? launch_axis2<INV, MODE, 4>(A, grid, threads, smem, st) : launch_axis2<INV, MODE, 8>(A, grid, threads, smem, st))
		#define AX2P(INV, MODE) (ps.P == 1 ? launch_axis2<INV, MODE, 1>(A, grid, threads, smem, st) : launch_axis2<INV, MODE, 2>(A, grid, threads, smem, st))
		if (mode == FM_R2C_PACKED) return AX2P(false, FM_R2C_PACKED);
		if (mode == FM_C2R_PACKED) return AX2P(true, FM_C2R_PACKED);
		return inverse ? AX2(true, FM_C2C) : AX2(false, FM_C2C);
		#undef AX2
		#undef AX2P
	}
	B2_REQUIRE(mode == FM_C2C, "fft: internal error (packed pass without a fast plan)");
	if (inverse) {
		if (smem > 48*1024) B2_CHECK(cudaFuncSetAttribute(k_fft_axis<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_fft_axis<true><<<grid, threads, smem, st>>>(A);
	} else {
		if (smem > 48*1024) B2_CHECK(cudaFuncSetAttribute(k_fft_axis<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_fft_axis<false><<<grid, threads, smem, st>>>(A);
	}
	B2_LAUNCH_CHECK();
	return 0;
}

extern "C" int b2_fft_execute(b2_fft_plan *p, const void *in, void *out, int forward, double scale, int mem, void *stream)
{
	B2_REQUIRE(p && in && out, "fft: null argument");
	B2_REQUIRE(mem == B2_MEM_HOST || mem == B2_MEM_DEVICE, "fft: bad memory kind");
	B2_REQUIRE(p->kind != B2_FFT_R2C || forward, "fft: a real-to-complex plan only runs forward");
	B2_REQUIRE(p->kind != B2_FFT_C2R || !forward, "fft: a complex-to-real plan only runs backward");
	cudaStream_t st = (cudaStream_t)stream;
	const bool f32 = p->dtype == B2_F32, inverse = !forward;
	const size_t csz = f32 ? 8 : 16, rsz = f32 ? 4 : 8;
	const size_t isz = p->kind == B2_FFT_R2C ? rsz : csz, osz = p->kind == B2_FFT_C2R ? rsz : csz;
	// ---- host arrays are staged through the device (whole span of the strided view)
	const void *din = in; void *dout = out;
	if (mem == B2_MEM_HOST) {
		size_t nin = (size_t)p->in_span*isz, nout = (size_t)p->out_span*osz;
		if (p->stage_in.n < nin && p->stage_in.alloc(nin)) return 1;
		B2_CHECK(cudaMemcpyAsync(p->stage_in.p, in, nin, cudaMemcpyHostToDevice, st));
		din = p->stage_in.p;
		if (in == out) dout = p->stage_in.p;
		else {
			if (p->stage_out.n < nout && p->stage_out.alloc(nout)) return 1;
			// a strided output view keeps the caller's values between its elements
			bool dense = true; { int64_t n = 1; const int64_t *osh = p->kind == B2_FFT_R2C ? p->cshape : p->shape; for (int d = 0; d < p->ndim; d++) n *= osh[d]; dense = (n == p->out_span); }
			if (!dense) B2_CHECK(cudaMemcpyAsync(p->stage_out.p, out, nout, cudaMemcpyHostToDevice, st));
			dout = p->stage_out.p;
		}
	}
	const int c_lk = f32 ? LK_C64 : LK_C128, c_sk = f32 ? SK_C64 : SK_C128;
	const int r_lk = f32 ? LK_F32 : LK_F64, r_sk = f32 ? SK_F32 : SK_F64;
	const int h_lk = f32 ? LK_H64 : LK_H128, h_sk = f32 ? SK_HC64 : SK_HC128;
	// compact complex128 work array in the half-spectrum (or full complex) shape
	int64_t wstride[4] = {0, 0, 0, 0};
	const int64_t *wshape = p->kind == B2_FFT_C2C ? p->shape : p->cshape;
	{ int64_t acc = 1; for (int d = p->ndim - 1; d >= 0; d--) { wstride[d] = acc; acc *= wshape[d]; } }
	auto need_work = [&]() -> int {
		int64_t n = 1; for (int d = 0; d < p->ndim; d++) n *= wshape[d];
		if (p->work.n < (size_t)n*16) return p->work.alloc((size_t)n*16);
		return 0;
	};
	// the packed real passes need unit stride along the real axis and 16-byte (8 for float32) aligned line starts
	const int last = p->axes[p->naxes - 1];
	auto can_pack = [&](const void *ptr, const int64_t *str) -> bool {
		if (p->packed.n == 0 || str[last] != 1) return false;
		if (((uintptr_t)ptr) % (2*rsz)) return false;
		for (int d = 0; d < p->ndim; d++) if (d != last && p->shape[d] > 1 && (str[d] % 2)) return false;
		return true;
	};
	int rc = 0;
	if (p->naxes == 1) {
		AxisPass &a = p->pass[0];
		const bool alias = (din == dout) && a.P > 1;
		B2_REQUIRE(!alias, "fft: in-place transforms of lines longer than %d elements are not supported", (int)(FFT_SMEM_MAX/16));
		if (p->kind == B2_FFT_C2C) rc = run_pass(p, a, p->shape, din, p->istride, c_lk, dout, p->ostride, c_sk, inverse, scale, st);
		else if (p->kind == B2_FFT_R2C) {
			if (can_pack(din, p->istride)) rc = run_pass(p, p->packed, p->shape, din, p->istride, r_lk, dout, p->ostride, h_sk, false, scale, st, FM_R2C_PACKED);
			else rc = run_pass(p, a, p->shape, din, p->istride, r_lk, dout, p->ostride, h_sk, false, scale, st);
		} else {
			if (can_pack(dout, p->ostride)) rc = run_pass(p, p->packed, p->shape, din, p->istride, c_lk, dout, p->ostride, r_sk, true, scale, st, FM_C2R_PACKED);
			else rc = run_pass(p, a, p->shape, din, p->istride, h_lk, dout, p->ostride, r_sk, true, scale, st);
		}
	} else {
		AxisPass &a = p->pass[0], &b = p->pass[1];
		if (p->kind == B2_FFT_C2R) {
			// c2c along the first listed axis on the half spectrum (into the work array), then c2r along the last
			if (need_work()) return 1;
			rc = run_pass(p, a, p->cshape, din, p->istride, c_lk, p->work.p, wstride, SK_C128, true, 1.0, st);
			if (!rc) {
				if (can_pack(dout, p->ostride)) rc = run_pass(p, p->packed, p->shape, p->work.p, wstride, LK_C128, dout, p->ostride, r_sk, true, scale, st, FM_C2R_PACKED);
				else rc = run_pass(p, b, p->shape, p->work.p, wstride, LK_H128, dout, p->ostride, r_sk, true, scale, st);
			}
		} else {
			// first pass along the last listed axis; the second runs in place on the output when its lines fit one CTA
			const bool r2c = p->kind == B2_FFT_R2C;
			const bool inplace2 = (b.P == 1) && !(a.P > 1 && din == dout);
			const int64_t *dims1 = p->shape, *dims2 = r2c ? p->cshape : p->shape;
			const int lk1 = r2c ? r_lk : c_lk;
			const bool pack = r2c && can_pack(din, p->istride);
			if (inplace2) {
				if (pack) rc = run_pass(p, p->packed, dims1, din, p->istride, lk1, dout, p->ostride, h_sk, false, 1.0, st, FM_R2C_PACKED);
				else rc = run_pass(p, a, dims1, din, p->istride, lk1, dout, p->ostride, r2c ? h_sk : c_sk, inverse, 1.0, st);
				if (!rc) rc = run_pass(p, b, dims2, dout, p->ostride, c_lk, dout, p->ostride, c_sk, inverse, scale, st);
			} else {
				if (need_work()) return 1;
				if (pack) rc = run_pass(p, p->packed, dims1, din, p->istride, lk1, p->work.p, wstride, SK_HC128, false, 1.0, st, FM_R2C_PACKED);
				else rc = run_pass(p, a, dims1, din, p->istride, lk1, p->work.p, wstride, r2c ? SK_HC128 : SK_C128, inverse, 1.0, st);
				if (!rc) rc = run_pass(p, b, dims2, p->work.p, wstride, LK_C128, dout, p->ostride, c_sk, inverse, scale, st);
			}
		}
	}
	if (rc) return 1;
	if (mem == B2_MEM_HOST) {
		B2_CHECK(cudaMemcpyAsync(out, dout, (size_t)p->out_span*osz, cudaMemcpyDeviceToHost, st));
		B2_CHECK(cudaStreamSynchronize(st));
	}
	return 0;
}
