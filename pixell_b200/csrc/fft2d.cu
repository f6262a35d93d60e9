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
#include "tfft.cuh"
#include <algorithm>
#include <stdlib.h>
#include <memory>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

enum { LK_C128, LK_C64, LK_F64, LK_F32, LK_H128, LK_H64 };
enum { SK_C128, SK_C64, SK_HC128, SK_HC64, SK_F64, SK_F32 };

struct AxisArgs {
	FftDesc d;
	int n, P, nl, nb, jfast, lk, sk, lstride, twoff;   // lstride: shared-memory elements per line; twoff: twiddle tables
	int64_t n_in, n_o1, n_o2;                    // lines are enumerated by (inner, outer1, outer2) indices
	int64_t is_t, is_in, is_o1, is_o2;           // input strides in elements of the input type
	int64_t os_t, os_in, os_o1, os_o2;
	const void *in; void *out; double scale;
	int async;                                   // 16-byte aligned complex128 / float64-pair input: tiles are fetched with cp.async
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


// (line, j) of the elements idx = tid, tid + T, ... of a tile of nb lines x nl elements without a division per element:
// jfast: idx = line*nl + j (threads run along the line), else idx = j*nb + line (threads run across neighbouring lines)
struct TileIter {
	int line, j, dl, dj, lim, jfast;
	__device__ __forceinline__ TileIter(int tid, int T, int nl, int nb, int jfast_) : jfast(jfast_) {
		if (jfast) { lim = nl; line = tid/nl; j = tid - line*nl; dl = T/nl; dj = T - dl*nl; }
		else { lim = nb; j = tid/nb; line = tid - j*nb; dj = T/nb; dl = T - dj*nb; }
	}
	__device__ __forceinline__ void next() {
		line += dl; j += dj;
		if (jfast) { if (j >= lim) { j -= lim; line++; } }
		else if (line >= lim) { line -= lim; j++; }
	}
};

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
		if (line < nbv) fft_st(A, bout + line*A.os_in, p + P*kk, s[line*ls + fft_pad(A.d, fft_rev(A.d, kk))]);
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
	TileIter it(tid, T, nl, nb, A.jfast);
	#pragma unroll 4
	for (int idx = tid; idx < tot; idx += T, it.next()) {
		const int line = it.line, j = it.j;
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
	TileIter ot(tid, T, nl, nb, A.jfast);
	#pragma unroll 4
	for (int idx = tid; idx < tot; idx += T, ot.next()) {
		const int line = ot.line, kk = ot.j;
		if (line >= nbv) continue;
		const int k = p + P*kk;
		const double2 x = s[line*ls + fft_pad(A.d, fft_rev(A.d, kk))];
		const int64_t bo = bout + line*A.os_in;
		if (MODE == FM_C2C) ((double2*)A.out)[bo + k*A.os_t] = make_double2(x.x*A.scale, x.y*A.scale);
		else if (MODE == FM_C2R_PACKED) {
			if (A.sk == SK_F64) ((double2*)((double*)A.out + bo))[k] = make_double2(x.x*A.scale, x.y*A.scale);
			else ((float2*)((float*)A.out + bo))[k] = make_float2((float)(x.x*A.scale), (float)(x.y*A.scale));
		} else {
			// partner nc - k has the same residue mod P (P <= 2)
			const int kp = k ? nc - k : 0;
			const double2 y = s[line*ls + fft_pad(A.d, fft_rev(A.d, (kp - p)/P))];
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


// Lines that do not fit one CTA (or whose tile of neighbouring lines does not): a thread-block cluster of P CTAs
// holds the tile in its distributed shared memory.  CTA r loads the r-th contiguous block of every line (each
// element comes from DRAM once), the cluster does the radix-P decimation-in-frequency step in place across the
// P shared memories (y_p[j] = w_n^{jp} sum_q x[j + q n/P] w_P^{qp}, one CTA per quarter of the j range), every CTA
// transforms its y_p (length n/P) locally, and the outputs k = p + P kk are either stored by their owner (strided
// axis: whole 64-byte row segments) or gathered through distributed shared memory so that CTA r writes the r-th
// contiguous block of every line (contiguous axis: no 16-byte interleaving between CTAs).
template<bool INV, int MODE, int P> __global__ void __launch_bounds__(512) k_fft_cl(AxisArgs A)
{
	extern __shared__ __align__(16) double2 s[];
	cg::cluster_group cl = cg::this_cluster();
	const int tid = threadIdx.x, T = blockDim.x, r = P > 1 ? (int)cl.block_rank() : 0;
	const int64_t tile = blockIdx.x/P;
	const int nb = A.nb, nl = A.nl, ls = A.lstride;
	const int nc = nl*P;                                   // complex transform length (n, or n/2 for the packed modes)
	const int64_t ntile = (A.n_in + nb - 1)/nb;
	const int64_t outer = tile/ntile, i0 = (tile % ntile)*nb;
	const int64_t o1 = outer/A.n_o2, o2 = outer % A.n_o2;
	const int nbv = (int)min((int64_t)nb, A.n_in - i0);
	const int64_t bin = o1*A.is_o1 + o2*A.is_o2 + i0*A.is_in, bout = o1*A.os_o1 + o2*A.os_o2 + i0*A.os_in;
	const int tot = nb*nl;
	const int twq = A.d.twmul/P, nhi = A.d.ntw_hi;          // table step of w_nc
	const double2 *twsm = s + A.twoff;
	auto sync_all = [&]() { if (P > 1) cl.sync(); else __syncthreads(); };
	// ---- block r of every line
	if (P == 1 && MODE == FM_C2R_PACKED && A.async) {
		// raw half spectrum X_0 .. X_nc by cp.async (X_nc of line b in the spare slot xn[b]), then the pairs (k, nc - k)
		// become Z_k = (X_k + conj X_{nc-k}) + i conj(w_n^k) (X_k - conj X_{nc-k}) and Z_{nc-k} in place
		double2 *xn = s + A.twoff + nhi + FFT_TWLO;
		TileIter it(tid, T, nl, nb, 1);
		for (int idx = tid; idx < tot; idx += T, it.next()) {
			const int line = it.line, jl = it.j;
			double2 *dst = &s[line*ls + fft_pad(A.d, jl)];
			if (line >= nbv) *dst = make_double2(0, 0);
			else cp_async16_cg(dst, (const double2*)A.in + bin + line*A.is_in + jl*A.is_t);
		}
		for (int b = tid; b < nb; b += T) { if (b < nbv) cp_async16_cg(&xn[b], (const double2*)A.in + bin + b*A.is_in + nc*A.is_t); else xn[b] = make_double2(0, 0); }
		cp_async_commit();
		fft_load_tw(s + A.twoff, A.d, tid, T);
		cp_async_wait_all();
		__syncthreads();
		const int nh = nc/2 + 1;
		TileIter pt(tid, T, nh, nb, 1);
		#pragma unroll 2
		for (int idx = tid; idx < nb*nh; idx += T, pt.next()) {
			const int line = pt.line, k = pt.j, kp = nc - k;
			double2 *pa = &s[line*ls + fft_pad(A.d, k)], *pb = k ? &s[line*ls + fft_pad(A.d, kp)] : &xn[line];
			const double2 xa = *pa, xb = *pb;
			const double2 sm = make_double2(xa.x + xb.x, xa.y - xb.y), df = make_double2(xa.x - xb.x, xa.y + xb.y);
			const double2 u = cmul(df, fft_tw<true>(twsm, nhi, k));
			*pa = make_double2(sm.x - u.y, sm.y + u.x);
			if (k && kp != k) *pb = make_double2(sm.x + u.y, u.x - sm.y);      // Z_{nc-k} = conj(sm) + i conj(u)
		}
	} else if (MODE != FM_C2R_PACKED && A.async) {
		// raw elements: the whole tile is put in flight at once (cp.async), nothing waits on a register
		TileIter it(tid, T, nl, nb, A.jfast);
		for (int idx = tid; idx < tot; idx += T, it.next()) {
			const int line = it.line, jl = it.j;
			const int jj = r*nl + jl;
			const int64_t b = bin + line*A.is_in;
			double2 *dst = &s[line*ls + fft_pad(A.d, jl)];
			if (line >= nbv) *dst = make_double2(0, 0);
			else if (MODE == FM_C2C) cp_async16_cg(dst, (const double2*)A.in + b + jj*A.is_t);
			else cp_async16_cg(dst, (const double2*)((const double*)A.in + b) + jj);
		}
		cp_async_commit();
		fft_load_tw(s + A.twoff, A.d, tid, T);
		cp_async_wait_all();
	} else {
		fft_load_tw(s + A.twoff, A.d, tid, T);
		if (MODE == FM_C2R_PACKED) __syncthreads();
		TileIter it(tid, T, nl, nb, A.jfast);
		#pragma unroll 4
		for (int idx = tid; idx < tot; idx += T, it.next()) {
			const int line = it.line, jl = it.j;
			const int jj = r*nl + jl;
			const int64_t b = bin + line*A.is_in;
			double2 v = make_double2(0, 0);
			if (line < nbv) {
				if (MODE == FM_C2C) v = ((const double2*)A.in)[b + jj*A.is_t];
				else if (MODE == FM_R2C_PACKED) {
					if (A.lk == LK_F64) v = ((const double2*)((const double*)A.in + b))[jj];
					else { float2 f = ((const float2*)((const float*)A.in + b))[jj]; v = make_double2(f.x, f.y); }
				} else {
					// the mirrored element belongs to another CTA of this cluster, which loads it at about the same time (L2)
					const int jm = nc - jj;
					double2 xa, xb;
					if (A.lk == LK_C128) { xa = ((const double2*)A.in)[b + jj*A.is_t]; xb = ((const double2*)A.in)[b + jm*A.is_t]; }
					else { float2 fa = ((const float2*)A.in)[b + jj*A.is_t], fb = ((const float2*)A.in)[b + jm*A.is_t]; xa = make_double2(fa.x, fa.y); xb = make_double2(fb.x, fb.y); }
					double2 sm = make_double2(xa.x + xb.x, xa.y - xb.y), df = make_double2(xa.x - xb.x, xa.y + xb.y);
					double2 u = cmul(df, fft_tw<true>(twsm, nhi, jj));          // e^{+2 pi i jj/n}
					v = make_double2(sm.x - u.y, sm.y + u.x);
				}
			}
			s[line*ls + fft_pad(A.d, jl)] = v;
		}
	}
	sync_all();
	// ---- radix-P step across the cluster, in place: this CTA combines its share of the j range
	if constexpr (P > 1) {
		double2 *S[P];
		#pragma unroll
		for (int q = 0; q < P; q++) S[q] = cl.map_shared_rank(s, q);
		const int nq = (nl + P - 1)/P, j0 = r*nq, nmine = max(0, min(nq, nl - j0));
		for (int idx = tid; idx < nb*nmine; idx += T) {
			int line, jl;
			if (A.jfast) { line = idx/nmine; jl = j0 + idx - line*nmine; } else { jl = j0 + idx/nb; line = idx % nb; }
			const int off = line*ls + fft_pad(A.d, jl);
			double2 u[P];
			#pragma unroll
			for (int q = 0; q < P; q++) u[q] = S[q][off];
			dft_small<P, INV>(u);
			if (jl) mul_powers<P>(u, fft_tw<INV>(twsm, nhi, twq*jl));
			#pragma unroll
			for (int q = 0; q < P; q++) S[q][off] = u[q];
		}
		cl.sync();
	}
	fft_smem<INV>(s, A.d, tid, T, nb, twsm);
	if (P == 1 ? MODE == FM_C2C : (A.os_t != 1 && MODE == FM_C2C)) {
		// ---- strided axis: the owner stores its outputs k = r + P kk
		TileIter ot(tid, T, nl, nb, A.jfast);
		#pragma unroll 4
		for (int idx = tid; idx < tot; idx += T, ot.next()) {
			const int line = ot.line, kk = ot.j;
			if (line >= nbv) continue;
			const double2 x = s[line*ls + fft_pad(A.d, fft_rev(A.d, kk))];
			((double2*)A.out)[bout + line*A.os_in + (int64_t)(r + P*kk)*A.os_t] = make_double2(x.x*A.scale, x.y*A.scale);
		}
		return;         // the last remote access (radix-P step) lies before the previous cluster barrier
	}
	// ---- contiguous axis: CTA r writes block r of every line, gathered from the owners' shared memories
	if (P > 1) cl.sync();
	auto owner = [&](int k) -> const double2* { return P > 1 ? cl.map_shared_rank(s, k % P) : s; };
	if (P == 1 && MODE == FM_R2C_PACKED) {
		// one CTA owns the whole line: the pair (k, nc - k) shares its two inputs and its twiddle
		const int nh = nc/2 + 1;                  // k = 0 .. nc/2
		const bool c128 = (A.sk == SK_HC128 || A.sk == SK_C128);
		TileIter pt(tid, T, nh, nb, 1);
		#pragma unroll 2
		for (int idx = tid; idx < nb*nh; idx += T, pt.next()) {
			const int line = pt.line, k = pt.j, kp = nc - k;
			if (line >= nbv) continue;
			const double2 x = s[line*ls + fft_pad(A.d, fft_rev(A.d, k))];
			const double2 y = k ? s[line*ls + fft_pad(A.d, fft_rev(A.d, kp))] : x;
			const double2 sm = make_double2(x.x + y.x, x.y - y.y), df = make_double2(x.x - y.x, x.y + y.y);
			const double2 u = cmul(df, fft_tw<false>(twsm, nhi, k));       // w_n^k (Z_k - conj Z_{nc-k})
			const double h = 0.5*A.scale;
			const double2 Xk = make_double2(h*(sm.x + u.y), h*(sm.y - u.x));
			// X_{nc-k} = [conj(sm) - i conj(u)]/2;  for k = 0 this is the Nyquist term Re Z_0 - Im Z_0
			const double2 Xp = make_double2(h*(sm.x - u.y), -h*(sm.y + u.x));
			const int64_t bo = bout + line*A.os_in;
			if (c128) { ((double2*)A.out)[bo + k*A.os_t] = Xk; if (kp != k) ((double2*)A.out)[bo + kp*A.os_t] = Xp; }
			else {
				((float2*)A.out)[bo + k*A.os_t] = make_float2((float)Xk.x, (float)Xk.y);
				if (kp != k) ((float2*)A.out)[bo + kp*A.os_t] = make_float2((float)Xp.x, (float)Xp.y);
			}
		}
		return;
	}
	TileIter gt(tid, T, nl, nb, A.jfast);
	#pragma unroll 2
	for (int idx = tid; idx < tot; idx += T, gt.next()) {
		const int line = gt.line, kl = gt.j;
		if (line >= nbv) continue;
		const int k = r*nl + kl;
		const double2 x = owner(k)[line*ls + fft_pad(A.d, fft_rev(A.d, k/P))];
		const int64_t bo = bout + line*A.os_in;
		if (MODE == FM_C2C) ((double2*)A.out)[bo + k*A.os_t] = make_double2(x.x*A.scale, x.y*A.scale);
		else if (MODE == FM_C2R_PACKED) {
			if (A.sk == SK_F64) ((double2*)((double*)A.out + bo))[k] = make_double2(x.x*A.scale, x.y*A.scale);
			else ((float2*)((float*)A.out + bo))[k] = make_float2((float)(x.x*A.scale), (float)(x.y*A.scale));
		} else {
			const int kp = k ? nc - k : 0;
			const double2 y = owner(kp)[line*ls + fft_pad(A.d, fft_rev(A.d, kp/P))];
			double2 sm = make_double2(x.x + y.x, x.y - y.y), df = make_double2(x.x - y.x, x.y + y.y);
			double2 u = cmul(df, fft_tw<false>(twsm, nhi, k));             // w_n^k (Z_k - conj Z_{nc-k})
			double2 X = make_double2(0.5*(sm.x + u.y), 0.5*(sm.y - u.x));
			const bool c128 = (A.sk == SK_HC128 || A.sk == SK_C128);
			if (c128) ((double2*)A.out)[bo + k*A.os_t] = make_double2(X.x*A.scale, X.y*A.scale);
			else ((float2*)A.out)[bo + k*A.os_t] = make_float2((float)(X.x*A.scale), (float)(X.y*A.scale));
			if (k == 0) {
				double2 Xn = make_double2((x.x - x.y)*A.scale, 0.0);
				if (c128) ((double2*)A.out)[bo + nc*A.os_t] = Xn;
				else ((float2*)A.out)[bo + nc*A.os_t] = make_float2((float)Xn.x, 0.f);
			}
		}
	}
	if (P > 1) cl.sync();          // remote shared memory stays alive until every reader is done
}

// ------------------------------------------------------------------------------------ plan

struct ArrayDesc { int64_t stride[4]; int kind; };      // element strides; kind = LK_* (as input) / SK_* (as output)

struct AxisPass {
	int axis = 0, n = 0, P = 1, nl = 0, nb = 1, threads = 256;
	bool cluster = false;      // P > 1 as a thread-block cluster (k_fft_cl) instead of P CTAs that each read whole lines
	bool strided = false;
	std::unique_ptr<AxisPass> alt;      // the same pass without a cluster (element kinds k_fft_cl does not handle), built on first use
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
	TfPlan *tf = nullptr;  // TMA-pipelined path (tfft.cu) for float64 transforms over the last two axes; null: not applicable
	~b2_fft_plan() { if (tf) tfft_plan_destroy(tf); }
};

static const size_t FFT_TILE_ELEMS = 6144;
// shared memory one CTA may use (B2_FFT_SMEM_KB overrides, for tuning)
static size_t fft_smem_max() { const char *e = getenv("B2_FFT_SMEM_KB"); return (size_t)(e ? atoi(e) : 200)*1024; }
#define FFT_SMEM_MAX fft_smem_max()

// shared memory that lets three CTAs share an SM (their load / transform / store phases then overlap)
static const size_t FFT_SMEM_SMALL = 73*1024;
static size_t tile_bytes(int nb, int nl, int ntab) { return (size_t)(nb*FftTables::smem_len(nl) + ntab/FFT_TWLO + FFT_TWLO + 1)*sizeof(double2); }
// elements per tile of short lines: three CTAs per SM on the cp.async path
static size_t tile_elems();
// B2_FFT_LEGACY=1: the register-staged kernels only (comparison runs).  B2_FFT_CLUSTER=1: split long lines over
// thread-block clusters instead of P independent CTAs (measured slower on B200 so far, profiles/r1n_*: off by default).
static bool no_cluster() { const char *e = getenv("B2_FFT_LEGACY"); return e && atoi(e); }
static bool use_cluster() { const char *e = getenv("B2_FFT_CLUSTER"); return e && atoi(e) && !no_cluster(); }

static size_t tile_elems() { return no_cluster() ? FFT_TILE_ELEMS : 4096; }

// cluster split of a line of complex length nc (tables of length ntab) held nb lines at a time: smallest P in {2, 4, 8}
// whose tile fits FFT_SMEM_SMALL, else the largest that fits FFT_SMEM_MAX; 0: none
static int cluster_split(int nc, int nb, int ntab)
{
	if (!use_cluster()) return 0;
	int best = 0;
	for (int P = 2; P <= 8; P *= 2) {
		if (nc % P || !FftTables::fast_ok(nc/P) || nc/P < P) continue;
		size_t b = tile_bytes(nb, nc/P, ntab);
		if (b <= FFT_SMEM_SMALL) return P;
		if (b <= FFT_SMEM_MAX) best = P;
	}
	return best;
}

// strided: the axis is not the contiguous one, so a CTA should hold neighbouring lines (64-byte row segments)
static int setup_pass(AxisPass &ps, int axis, int n, bool can_batch, bool strided, bool allow_cluster = true)
{
	ps.axis = axis; ps.n = n; ps.cluster = false; ps.strided = strided;
	const int wantb = (strided && can_batch) ? 4 : 1;
	if (allow_cluster && FftTables::fast_ok(n) && tile_bytes(wantb, n, n) > FFT_SMEM_SMALL) {
		int P = cluster_split(n, wantb, n);
		if (P) {
			ps.P = P; ps.nl = n/P; ps.cluster = true;
			if (ps.tab.build(ps.nl, n)) return 1;
			ps.nb = wantb;
			while (can_batch && tile_bytes(2*ps.nb, ps.nl, n) <= FFT_SMEM_SMALL && (size_t)2*ps.nb*ps.nl <= FFT_TILE_ELEMS) ps.nb *= 2;
			return 0;
		}
	}
	int P = 1;
	const bool four = allow_cluster && !no_cluster() && FftTables::fast_ok(n) && tile_bytes(4, n, n) <= FFT_SMEM_SMALL;
	const int minb = (strided && can_batch) ? (four ? 4 : 2) : 1;
	while ((size_t)(minb*FftTables::smem_len(n/P) + n/FFT_TWLO + FFT_TWLO + 1)*sizeof(double2) > FFT_SMEM_MAX) {
		int np = P + 1;
		while (np <= 64 && n % np) np++;
		B2_REQUIRE(np <= 64, "fft: a transform of length %d does not fit in shared memory (no usable split)", n);
		P = np;
	}
	ps.P = P; ps.nl = n/P;
	if (ps.tab.build(ps.nl, n)) return 1;
	ps.nb = 1;
	if (can_batch && FftTables::smooth(ps.nl)) ps.nb = (int)std::max<size_t>(minb, tile_elems()/FftTables::smem_len(ps.nl));
	return 0;
}

// the real axis of even length n as a complex transform of length n/2 (one CTA, a cluster, or two CTAs)
static int setup_packed(AxisPass &ps, int axis, int n)
{
	ps.n = 0; ps.cluster = false;
	if (n % 2 || n < 4) return 0;
	const int nc = n/2;
	if (FftTables::fast_ok(nc) && tile_bytes(1, nc, n) > FFT_SMEM_SMALL) {
		int P = cluster_split(nc, 1, n);
		if (P) {
			ps.axis = axis; ps.n = n; ps.P = P; ps.nl = nc/P; ps.cluster = true; ps.nb = 1;
			return ps.tab.build(ps.nl, n);
		}
	}
	for (int P = 1; P <= 2; P++) {
		if (nc % P) continue;
		if ((size_t)(FftTables::smem_len(nc/P) + n/FFT_TWLO + FFT_TWLO + 1)*sizeof(double2) > FFT_SMEM_MAX) continue;
		if (!FftTables::fast_ok(nc/P)) return 0;
		ps.axis = axis; ps.n = n; ps.P = P; ps.nl = nc/P;
		if (ps.tab.build(ps.nl, n)) return 1;
		ps.nb = (int)std::max<size_t>(1, tile_elems()/FftTables::smem_len(ps.nl));
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
	if (tfft_eligible(kind, dtype, ndim, p->shape, p->istride, p->ostride, naxes, p->axes) &&
		tfft_plan_create(&p->tf, kind, ndim, p->shape, p->istride, p->ostride)) p->tf = nullptr;      // fall back to the axis passes
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

template<bool INV, int MODE, int P> static int launch_cl(const AxisArgs &A, unsigned nblk, int threads, size_t smem, cudaStream_t st)
{
	if (smem > 48*1024) B2_CHECK(cudaFuncSetAttribute(k_fft_cl<INV, MODE, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(nblk*P, 1, 1); cfg.blockDim = dim3(threads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = P; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
	cfg.attrs = at; cfg.numAttrs = 1;
	B2_CHECK(cudaLaunchKernelEx(&cfg, k_fft_cl<INV, MODE, P>, A));
	g_b2_launches++;
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
	size_t smem = sizeof(double2)*(size_t)(A.twoff + ps.tab.twsm_len() + A.nb);      // + one spare element per line (packed c2r: X_nc)
	int threads = (int)std::min<int64_t>(512, std::max<int64_t>(64, b2_round_up((int64_t)A.nb*ps.nl/4, 32)));
	int64_t ntile = (A.n_in + A.nb - 1)/A.nb;
	int64_t nblk = ntile*A.n_o1*A.n_o2;
	B2_REQUIRE(nblk < (1LL << 31), "fft: too many lines for one launch");
	dim3 grid((unsigned)nblk, ps.P);
	const bool plain128 = (mode != FM_C2C) || (lk == LK_C128 && sk == SK_C128);
	if (ps.cluster && !plain128) {
		if (!ps.alt) { ps.alt.reset(new AxisPass()); if (setup_pass(*ps.alt, ps.axis, ps.n, true, ps.strided, false)) return 1; }
		return run_pass(p, *ps.alt, dims, src, ss, lk, dst, ds, sk, inverse, scale, st, mode);
	}
	A.async = ((uintptr_t)src % 16 == 0) && (mode == FM_C2C ? lk == LK_C128 : mode == FM_R2C_PACKED ? lk == LK_F64 : lk == LK_C128);
	if ((ps.cluster || (ps.P == 1 && plain128 && !no_cluster())) && ps.tab.d.fast) {
		B2_REQUIRE(nblk*ps.P < (1LL << 31), "fft: too many lines for one launch");
		threads = (int)std::min<int64_t>(512, std::max<int64_t>(64, b2_round_up((int64_t)A.nb*ps.nl/16, 32)));
		#define CL(INV, MODE) (ps.P == 1 ? launch_cl<INV, MODE, 1>(A, (unsigned)nblk, threads, smem, st) : ps.P == 2 ? launch_cl<INV, MODE, 2>(A, (unsigned)nblk, threads, smem, st) : ps.P == 4 ? launch_cl<INV, MODE, 4>(A, (unsigned)nblk, threads, smem, st) : \
			launch_cl<INV, MODE, 8>(A, (unsigned)nblk, threads, smem, st))
		if (mode == FM_R2C_PACKED) return CL(false, FM_R2C_PACKED);
		if (mode == FM_C2R_PACKED) return CL(true, FM_C2R_PACKED);
		return inverse ? CL(true, FM_C2C) : CL(false, FM_C2C);
		#undef CL
	}
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
	if (p->tf && ((uintptr_t)din % 16 == 0) && ((uintptr_t)dout % 16 == 0)) rc = tfft_execute(p->tf, din, dout, forward, scale, st);
	else if (p->naxes == 1) {
		AxisPass &a = p->pass[0];
		const bool alias = (din == dout) && a.P > 1 && !(a.cluster && p->dtype == B2_F64 && p->kind == B2_FFT_C2C);
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
			const bool cl64 = p->dtype == B2_F64;      // cluster passes load a whole tile before they store: in place is safe
			const bool a_inplace_ok = (a.P == 1) || (a.cluster && cl64 && !r2c);
			const bool inplace2 = (b.P == 1 || (b.cluster && cl64)) && (a_inplace_ok || din != dout);
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

// ------------------------------------------------------------------------------------ QU <-> EB rotation

// enmap.queb_rotmat + map_mul (pixell/enmap.py:1391-1400, 1418-1427) applied in place to a pair of Fourier-space
// components: (a, b) <- (c a - s b, s a + c b), c + i s = exp(i spin atan2(sign lx, ly)).  Spin 2 uses
// cos 2phi = (ly^2 - lx^2)/l^2, sin 2phi = 2 sign lx ly/l^2 (no transcendental functions); l = 0 gives the identity.
template<typename C> __global__ void k_queb_rotate(C *a, C *b, int64_t batch_stride, int ny, int nx,
	const double *ly, const double *lx, int spin, int sign)
{
	const int x = blockIdx.x*blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x >= nx) return;
	const double vy = ly[y], vx = sign*lx[x];
	double c, sn;
	if (spin == 2) {
		const double l2 = vy*vy + vx*vx;
		if (l2 > 0) { const double inv = 1.0/l2; c = (vy*vy - vx*vx)*inv; sn = 2.0*vx*vy*inv; } else { c = 1.0; sn = 0.0; }
	} else sincos(spin*atan2(vx, vy), &sn, &c);
	const int64_t i = (int64_t)blockIdx.z*batch_stride + (int64_t)y*nx + x;
	const C va = a[i], vb = b[i];
	C ra, rb;
	ra.x = c*va.x - sn*vb.x; ra.y = c*va.y - sn*vb.y;
	rb.x = sn*va.x + c*vb.x; rb.y = sn*va.y + c*vb.y;
	a[i] = ra; b[i] = rb;
}

extern "C" int b2_queb_rotate(void *data, int64_t comp_stride, int64_t nbatch, int64_t batch_stride, int ny, int nx,
	const double *ly, const double *lx, int spin, int sign, int dtype, int mem, void *stream)
{
	B2_REQUIRE(data && ly && lx, "queb_rotate: null argument");
	B2_REQUIRE(ny >= 1 && nx >= 1 && nbatch >= 1 && nbatch < 65536, "queb_rotate: bad extents");
	B2_REQUIRE(dtype == B2_F64 || dtype == B2_F32, "queb_rotate: bad dtype");
	B2_REQUIRE(mem == B2_MEM_HOST || mem == B2_MEM_DEVICE, "queb_rotate: bad memory kind");
	B2_REQUIRE(sign == 1 || sign == -1, "queb_rotate: sign must be +1 or -1");
	if (spin == 0) return 0;
	cudaStream_t st = (cudaStream_t)stream;
	const size_t csz = dtype == B2_F32 ? 8 : 16;
	DevBuf<double> l; if (l.alloc((size_t)ny + nx)) return 1;
	B2_CHECK(cudaMemcpyAsync(l.p, ly, sizeof(double)*ny, cudaMemcpyHostToDevice, st));
	B2_CHECK(cudaMemcpyAsync(l.p + ny, lx, sizeof(double)*nx, cudaMemcpyHostToDevice, st));
	// the pair (a, b) of batch k starts at data + k*batch_stride and data + k*batch_stride + comp_stride
	char *d = (char*)data;
	DevBuf<char> stage;
	const size_t span = (size_t)((nbatch - 1)*batch_stride + comp_stride + (int64_t)ny*nx)*csz;
	if (mem == B2_MEM_HOST) {
		B2_REQUIRE(comp_stride > 0 && batch_stride >= 0, "queb_rotate: host arrays need non-negative strides");
		if (stage.alloc(span)) return 1;
		B2_CHECK(cudaMemcpyAsync(stage.p, data, span, cudaMemcpyHostToDevice, st));
		d = stage.p;
	}
	dim3 grid((nx + 255)/256, ny, (unsigned)nbatch);
	B2_REQUIRE(ny <= 65535, "queb_rotate: more than 65535 rows are not supported");
	if (dtype == B2_F64) k_queb_rotate<double2><<<grid, 256, 0, st>>>((double2*)d, (double2*)d + comp_stride, batch_stride, ny, nx, l.p, l.p + ny, spin, sign);
	else k_queb_rotate<float2><<<grid, 256, 0, st>>>((float2*)d, (float2*)d + comp_stride, batch_stride, ny, nx, l.p, l.p + ny, spin, sign);
	B2_LAUNCH_CHECK();
	if (mem == B2_MEM_HOST) B2_CHECK(cudaMemcpyAsync(data, stage.p, span, cudaMemcpyDeviceToHost, st));
	B2_CHECK(cudaStreamSynchronize(st));      // the l table (and the staging buffer) are released on return
	return 0;
}

// ------------------------------------------------------------------------------------ Fourier-space filter
// data[b][y][x] *= fy[y]*fx[x] (separable, e.g. a Gaussian beam exp(-l^2 sigma^2/2) = gy(ly) gx(lx)) or *= f2[y][x]: the
// harmonic filter between enmap.fft and enmap.ifft (pixell/enmap.py:1429-1439 smooth_gauss, apply_window :1441-1460) in one
// pass over the array: two 16-byte elements per thread, rows by blockIdx.y, batch members by blockIdx.z.
template<typename C, typename Rr> __global__ void __launch_bounds__(256) k_fourier_filter(C *data, int64_t batch_stride, int64_t row_stride,
	int ny, int nx, const Rr *fy, const Rr *fx, const Rr *f2)
{
	const int y = blockIdx.y;
	C *row = data + (int64_t)blockIdx.z*batch_stride + (int64_t)y*row_stride;
	const Rr gy = fy ? fy[y] : (Rr)1;
	for (int x = blockIdx.x*blockDim.x + threadIdx.x; x < nx; x += gridDim.x*blockDim.x) {
		const Rr g = f2 ? f2[(int64_t)y*nx + x] : gy*fx[x];
		C v = row[x]; v.x *= g; v.y *= g; row[x] = v;
	}
}

extern "C" int b2_fourier_filter(void *data, int64_t nbatch, int64_t batch_stride, int64_t row_stride, int ny, int nx,
	const void *fy, const void *fx, const void *f2, int dtype, void *stream)
{
	B2_REQUIRE(data && ((fy && fx) || f2), "fourier_filter: need fy and fx, or f2");
	B2_REQUIRE(ny >= 1 && nx >= 1 && nbatch >= 1 && nbatch < 65536 && ny <= 65535, "fourier_filter: bad extents");
	B2_REQUIRE(dtype == B2_F64 || dtype == B2_F32, "fourier_filter: bad dtype");
	cudaStream_t st = (cudaStream_t)stream;
	dim3 grid((unsigned)std::min<int64_t>((nx + 511)/512, 64), ny, (unsigned)nbatch);
	if (dtype == B2_F64) k_fourier_filter<double2, double><<<grid, 256, 0, st>>>((double2*)data, batch_stride, row_stride, ny, nx, (const double*)fy, (const double*)fx, (const double*)f2);
	else k_fourier_filter<float2, float><<<grid, 256, 0, st>>>((float2*)data, batch_stride, row_stride, ny, nx, (const float*)fy, (const float*)fx, (const float*)f2);
	B2_LAUNCH_CHECK();
	return 0;
}
