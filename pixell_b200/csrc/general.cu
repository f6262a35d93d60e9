// general.cu -- K8: synthesis at arbitrary positions (ducc0.sht.experimental.synthesis_general as called by
// pixell/curvedsky.py:993-1016, alm2map_pos :174-207, alm2map_general :796-820; consumer lensing.py:468-492).
//
// For every component and m, leg_m(theta) (K1, on a Clenshaw-Curtis ring set with nt >= lmax + 2 rings) is a
// trigonometric polynomial of degree <= lmax on the circle theta in [0, 2 pi) with the symmetry
// leg_m(2 pi - theta) = (-1)^(m+s) leg_m(theta).  Its Fourier coefficients c_{k,m} turn the field into a 2-D
// Fourier series  f(theta, phi) = sum_{k,m} c_{k,m} e^{i(k theta + m phi)}, which a type-2 non-uniform FFT evaluates
// at the requested points: divide by the transform of the interpolation kernel, zero-pad to an M x M grid
// (M >= 2 (2 lmax + 1)), inverse FFT (K7, complex-to-real along phi), and interpolate with the "exponential of
// semicircle" kernel exp(beta (sqrt(1 - z^2) - 1)) over W x W grid points (W = 13, beta = 2.3 W: ~1e-12).
// The kernels here are the three steps around the FFTs; the host side (pixell_b200/sht.py) strings them together.
#include "../../include/b200sht.h"
#include "common.cuh"

// ext[c][m][N] <- leg[c][m][nring_pad] on the CC rings j = 0..nt-1: ext[j] = leg[j], ext[N - j] = sig leg[j], N = 2 (nt - 1)
__global__ void k_gen_extend(const double2 *leg, double2 *ext, int nt, int64_t nring_pad, int N, int spin, int nm)
{
	const int64_t col = blockIdx.y;              // c*nm + m
	const int m = (int)(col % nm);
	const int j = blockIdx.x*blockDim.x + threadIdx.x;
	if (j >= N) return;
	const double sig = ((m + spin) & 1) ? -1.0 : 1.0;
	double2 v;
	if (j < nt) v = leg[col*nring_pad + j];
	else { v = leg[col*nring_pad + (N - j)]; v.x *= sig; v.y *= sig; }
	ext[col*N + j] = v;
}

// G[c][k mod M][m] <- C[c][m][k mod N] / (P_k P_m), |k| <= lmax, m <= mmax; the m = 0 column is made Hermitian in k.
// C holds the theta-FFT of ext (scaled by 1/N by the caller's FFT); corr[k] = 1/P_k, k = 0..lmax.
// 32 x 32 tiles through shared memory: reads run along k, writes along m.
__global__ void k_gen_scatter(const double2 *C, double2 *G, int lmax, int nm, int N, int M, int64_t mrow, const double *corr)
{
	__shared__ double2 tile[32][33];
	const int c = blockIdx.z;
	const int k0 = blockIdx.x*32 - lmax, m0 = blockIdx.y*32;         // signed k of this tile's first row
	const double2 *Cc = C + (int64_t)c*nm*N;
	double2 *Gc = G + (int64_t)c*M*mrow;
	for (int r = threadIdx.y; r < 32; r += blockDim.y) {
		const int m = m0 + r, k = k0 + threadIdx.x;
		double2 v = make_double2(0, 0);
		if (m < nm && k <= lmax) {
			v = Cc[(int64_t)m*N + (k < 0 ? k + N : k)];
			if (m == 0) { double2 w = Cc[(k > 0 ? N - k : -k)]; v = make_double2(0.5*(v.x + w.x), 0.5*(v.y - w.y)); }
			const double f = corr[k < 0 ? -k : k]*corr[m];
			v.x *= f; v.y *= f;
		}
		tile[r][threadIdx.x] = v;
	}
	__syncthreads();
	for (int r = threadIdx.y; r < 32; r += blockDim.y) {
		const int k = k0 + r, m = m0 + threadIdx.x;
		if (m < nm && k <= lmax) Gc[(int64_t)(k < 0 ? k + M : k)*mrow + m] = tile[threadIdx.x][r];
	}
}

// out[c][i] = sum_{a,b < W} g[c][(ia + a) mod M][(ib + b) mod M] psi(theta_i; a) psi(phi_i; b)
template<int W> __global__ void k_gen_interp(const double *g, int ncomp, int M, const double *loc, int64_t npos,
	double beta, double *out, int64_t out_cstride)
{
	const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= npos) return;
	const double scale = M/6.283185307179586476925286766559;
	double wt[2][W]; int i0[2];
	#pragma unroll
	for (int d = 0; d < 2; d++) {
		double u = loc[2*i + d]*scale;
		u -= floor(u/M)*M;                        // [0, M)
		const int first = (int)ceil(u - 0.5*W);
		i0[d] = first;
		#pragma unroll
		for (int a = 0; a < W; a++) {
			const double z = (first + a - u)*(2.0/W);
			const double t = 1.0 - z*z;
			wt[d][a] = t > 0 ? exp(beta*(sqrt(t) - 1.0)) : 0.0;
		}
	}
	int cols[W];
	#pragma unroll
	for (int b = 0; b < W; b++) { int j = i0[1] + b; j %= M; if (j < 0) j += M; cols[b] = j; }
	for (int c = 0; c < ncomp; c++) {
		const double *gc = g + (int64_t)c*M*M;
		double acc = 0;
		#pragma unroll 1
		for (int a = 0; a < W; a++) {
			int r = i0[0] + a; r %= M; if (r < 0) r += M;
			const double *row = gc + (int64_t)r*M;
			double s = 0;
			#pragma unroll
			for (int b = 0; b < W; b++) s = fma(row[cols[b]], wt[1][b], s);
			acc = fma(s, wt[0][a], acc);
		}
		out[c*out_cstride + i] = acc;
	}
}

extern "C" int b2_general_extend(const void *leg_dev, void *ext_dev, int ncomp, int nm, int nt, int64_t nring_pad, int spin, void *stream)
{
	B2_REQUIRE(leg_dev && ext_dev && ncomp >= 1 && nm >= 1 && nt >= 2 && nring_pad >= nt, "general_extend: bad arguments");
	const int N = 2*(nt - 1);
	B2_REQUIRE((int64_t)ncomp*nm <= 65535, "general_extend: too many columns for one launch");
	dim3 grid((N + 255)/256, (unsigned)(ncomp*nm));
	k_gen_extend<<<grid, 256, 0, (cudaStream_t)stream>>>((const double2*)leg_dev, (double2*)ext_dev, nt, nring_pad, N, spin, nm);
	B2_LAUNCH_CHECK();
	return 0;
}

extern "C" int b2_general_scatter(const void *coef_dev, void *grid_dev, int ncomp, int lmax, int nm, int N, int M,
	const double *corr_dev, void *stream)
{
	B2_REQUIRE(coef_dev && grid_dev && corr_dev, "general_scatter: null argument");
	B2_REQUIRE(N >= 2*lmax + 1 && M >= 2*(2*lmax + 1) && M % 2 == 0 && nm <= lmax + 1, "general_scatter: grid too small for lmax %d", lmax);
	const int64_t mrow = M/2 + 1;
	cudaStream_t st = (cudaStream_t)stream;
	B2_CHECK(cudaMemsetAsync(grid_dev, 0, sizeof(double2)*(size_t)ncomp*M*mrow, st));
	dim3 grid((2*lmax + 1 + 31)/32, (nm + 31)/32, ncomp), block(32, 8);
	k_gen_scatter<<<grid, block, 0, st>>>((const double2*)coef_dev, (double2*)grid_dev, lmax, nm, N, M, mrow, corr_dev);
	B2_LAUNCH_CHECK();
	return 0;
}

extern "C" int b2_general_interp(const void *fine_dev, int ncomp, int M, const double *loc_dev, int64_t npos, int W, double beta,
	void *out_dev, int64_t out_comp_stride, void *stream)
{
	B2_REQUIRE(fine_dev && loc_dev && out_dev && ncomp >= 1 && npos >= 0, "general_interp: bad arguments");
	B2_REQUIRE(W == 13 || W == 8, "general_interp: kernel width %d is not built (8 or 13)", W);
	B2_REQUIRE(M >= 2*W, "general_interp: grid smaller than the kernel");
	if (npos == 0) return 0;
	cudaStream_t st = (cudaStream_t)stream;
	const int64_t nblk = (npos + 127)/128;
	B2_REQUIRE(nblk < (1LL << 31), "general_interp: too many positions for one launch");
	if (W == 13) k_gen_interp<13><<<(unsigned)nblk, 128, 0, st>>>((const double*)fine_dev, ncomp, M, loc_dev, npos, beta, (double*)out_dev, out_comp_stride);
	else k_gen_interp<8><<<(unsigned)nblk, 128, 0, st>>>((const double*)fine_dev, ncomp, M, loc_dev, npos, beta, (double*)out_dev, out_comp_stride);
	B2_LAUNCH_CHECK();
	return 0;
}

// ------------------------------------------------------------------------------------ adjoint (type-1 direction)

// fine[c][r][j] += val[c][i] psi(theta_i; a) psi(phi_i; b): transpose of k_gen_interp (native FP64 atomics; points that
// are close in memory are close on the sky for map-shaped inputs, so most collisions stay inside a warp's L2 lines)
template<int W> __global__ void k_gen_spread(double *g, int ncomp, int M, const double *loc, int64_t npos,
	double beta, const double *val, int64_t val_cstride)
{
	const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= npos) return;
	const double scale = M/6.283185307179586476925286766559;
	double wt[2][W]; int i0[2];
	#pragma unroll
	for (int d = 0; d < 2; d++) {
		double u = loc[2*i + d]*scale;
		u -= floor(u/M)*M;
		const int first = (int)ceil(u - 0.5*W);
		i0[d] = first;
		#pragma unroll
		for (int a = 0; a < W; a++) {
			const double z = (first + a - u)*(2.0/W);
			const double t = 1.0 - z*z;
			wt[d][a] = t > 0 ? exp(beta*(sqrt(t) - 1.0)) : 0.0;
		}
	}
	int cols[W];
	#pragma unroll
	for (int b = 0; b < W; b++) { int j = i0[1] + b; j %= M; if (j < 0) j += M; cols[b] = j; }
	for (int c = 0; c < ncomp; c++) {
		double *gc = g + (int64_t)c*M*M;
		const double v = val[c*val_cstride + i];
		#pragma unroll 1
		for (int a = 0; a < W; a++) {
			int r = i0[0] + a; r %= M; if (r < 0) r += M;
			double *row = gc + (int64_t)r*M;
			const double va = v*wt[0][a];
			#pragma unroll
			for (int b = 0; b < W; b++) atomicAdd(&row[cols[b]], va*wt[1][b]);
		}
	}
}

// C[c][m][k mod N] <- G[c][k mod M][m] corr[|k|] corr[m] for |k| <= lmax, 0 for the other k: transpose of k_gen_scatter
// (without the m = 0 symmetrisation, which the real part taken by the Legendre adjoint supplies)
__global__ void k_gen_gather(double2 *C, const double2 *G, int lmax, int nm, int N, int M, int64_t mrow, const double *corr)
{
	__shared__ double2 tile[32][33];
	const int c = blockIdx.z;
	const int k0 = blockIdx.x*32 - lmax, m0 = blockIdx.y*32;
	double2 *Cc = C + (int64_t)c*nm*N;
	const double2 *Gc = G + (int64_t)c*M*mrow;
	for (int r = threadIdx.y; r < 32; r += blockDim.y) {
		const int k = k0 + r, m = m0 + threadIdx.x;
		double2 v = make_double2(0, 0);
		if (m < nm && k <= lmax) {
			v = Gc[(int64_t)(k < 0 ? k + M : k)*mrow + m];
			const double f = corr[k < 0 ? -k : k]*corr[m];
			v.x *= f; v.y *= f;
		}
		tile[r][threadIdx.x] = v;
	}
	__syncthreads();
	for (int r = threadIdx.y; r < 32; r += blockDim.y) {
		const int m = m0 + r, k = k0 + threadIdx.x;
		if (m < nm && k <= lmax) Cc[(int64_t)m*N + (k < 0 ? k + N : k)] = tile[threadIdx.x][r];
	}
}

// leg[c][m][j] <- ext[j] + sig ext[N - j] (1 <= j <= nt - 2), ext[j] at the two poles: transpose of k_gen_extend
__global__ void k_gen_fold(double2 *leg, const double2 *ext, int nt, int64_t nring_pad, int N, int spin, int nm)
{
	const int64_t col = blockIdx.y;
	const int m = (int)(col % nm);
	const int j = blockIdx.x*blockDim.x + threadIdx.x;
	if (j >= nring_pad) return;
	double2 v = make_double2(0, 0);
	if (j < nt) {
		v = ext[col*N + j];
		if (j >= 1 && j <= nt - 2) {
			const double sig = ((m + spin) & 1) ? -1.0 : 1.0;
			const double2 w = ext[col*N + (N - j)];
			v.x += sig*w.x; v.y += sig*w.y;
		}
	}
	leg[col*nring_pad + j] = v;
}

extern "C" int b2_general_spread(void *fine_dev, int ncomp, int M, const double *loc_dev, int64_t npos, int W, double beta,
	const void *val_dev, int64_t val_comp_stride, void *stream)
{
	B2_REQUIRE(fine_dev && loc_dev && val_dev && ncomp >= 1 && npos >= 0, "general_spread: bad arguments");
	B2_REQUIRE(W == 13 || W == 8, "general_spread: kernel width %d is not built (8 or 13)", W);
	B2_REQUIRE(M >= 2*W, "general_spread: grid smaller than the kernel");
	cudaStream_t st = (cudaStream_t)stream;
	B2_CHECK(cudaMemsetAsync(fine_dev, 0, sizeof(double)*(size_t)ncomp*M*M, st));
	if (npos == 0) return 0;
	const int64_t nblk = (npos + 127)/128;
	B2_REQUIRE(nblk < (1LL << 31), "general_spread: too many positions for one launch");
	if (W == 13) k_gen_spread<13><<<(unsigned)nblk, 128, 0, st>>>((double*)fine_dev, ncomp, M, loc_dev, npos, beta, (const double*)val_dev, val_comp_stride);
	else k_gen_spread<8><<<(unsigned)nblk, 128, 0, st>>>((double*)fine_dev, ncomp, M, loc_dev, npos, beta, (const double*)val_dev, val_comp_stride);
	B2_LAUNCH_CHECK();
	return 0;
}

extern "C" int b2_general_gather(void *coef_dev, const void *grid_dev, int ncomp, int lmax, int nm, int N, int M,
	const double *corr_dev, void *stream)
{
	B2_REQUIRE(coef_dev && grid_dev && corr_dev, "general_gather: null argument");
	B2_REQUIRE(N >= 2*lmax + 1 && M >= 2*(2*lmax + 1) && M % 2 == 0 && nm <= lmax + 1, "general_gather: grid too small for lmax %d", lmax);
	cudaStream_t st = (cudaStream_t)stream;
	B2_CHECK(cudaMemsetAsync(coef_dev, 0, sizeof(double2)*(size_t)ncomp*nm*N, st));
	dim3 grid((2*lmax + 1 + 31)/32, (nm + 31)/32, ncomp), block(32, 8);
	k_gen_gather<<<grid, block, 0, st>>>((double2*)coef_dev, (const double2*)grid_dev, lmax, nm, N, M, (int64_t)(M/2 + 1), corr_dev);
	B2_LAUNCH_CHECK();
	return 0;
}

extern "C" int b2_general_fold(void *leg_dev, const void *ext_dev, int ncomp, int nm, int nt, int64_t nring_pad, int spin, void *stream)
{
	B2_REQUIRE(leg_dev && ext_dev && ncomp >= 1 && nm >= 1 && nt >= 2 && nring_pad >= nt, "general_fold: bad arguments");
	B2_REQUIRE((int64_t)ncomp*nm <= 65535, "general_fold: too many columns for one launch");
	dim3 grid((unsigned)((nring_pad + 255)/256), (unsigned)(ncomp*nm));
	k_gen_fold<<<grid, 256, 0, (cudaStream_t)stream>>>((double2*)leg_dev, (const double2*)ext_dev, nt, nring_pad, 2*(nt - 1), spin, nm);
	B2_LAUNCH_CHECK();
	return 0;
}
