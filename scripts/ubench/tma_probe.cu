// tma_probe.cu -- checks the TMA behaviours the tile-DFT kernels rely on: tensor maps whose strides are not
// monotonic (a row-permuting view), zero fill of out-of-range box parts on load, clipping on store.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu
#include "../../pixell_b200/csrc/tma.cuh"
#include <stdio.h>
#include <vector>
#include <complex>

struct Maps { CUtensorMap ld, st; };

__global__ void k_probe(const __grid_constant__ Maps M, int c0, int c2, int box_bytes)
{
	extern __shared__ __align__(128) unsigned char sm[];
	__shared__ uint64_t bar;
	if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
	__syncthreads();
	if (threadIdx.x == 0) {
		mbar_expect_tx(&bar, box_bytes);
		tma_load_4d(sm, &M.ld, &bar, c0, 0, c2, 0);
	}
	mbar_wait(&bar, 0);
	// touch the data with the generic proxy: negate imaginary parts (conjugate)
	double *d = (double*)sm;
	for (int i = threadIdx.x; i < box_bytes/16; i += blockDim.x) d[2*i + 1] = -d[2*i + 1];
	fence_proxy_async();
	__syncthreads();
	if (threadIdx.x == 0) { tma_store_4d(&M.st, sm, c0, 0, c2, 0); bulk_commit(); bulk_wait_all<0>(); }
}

int main()
{
	const int N1 = 8, N2 = 4, ncol = 40, pitch = 41;      // rows = N1*N2 = 32; row = j1*N2 + j2
	const int rows = N1*N2, w = 16;
	std::vector<std::complex<double>> A((size_t)rows*pitch), B((size_t)rows*pitch, {-7, -7});
	for (int r = 0; r < rows; r++) for (int c = 0; c < pitch; c++) A[(size_t)r*pitch + c] = {(double)r, (double)c};
	double *dA, *dB;
	cudaMalloc(&dA, A.size()*16); cudaMalloc(&dB, B.size()*16);
	cudaMemcpy(dA, A.data(), A.size()*16, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size()*16, cudaMemcpyHostToDevice);
	Maps M;
	// load view: d0 = doubles in a row (ncol columns valid), d1 = j1 (stride N2 rows), d2 = j2 (stride 1 row), d3 = 1
	uint64_t dims[4] = {2*(uint64_t)ncol, N1, N2, 1};
	uint64_t st_ld[3] = {(uint64_t)N2*pitch*16, (uint64_t)pitch*16, (uint64_t)rows*pitch*16};
	uint32_t box[4] = {2*w, N1, 1, 1};
	int rc = b2_make_map_f64(&M.ld, dA, dims, st_ld, box);
	printf("encode load map (non-monotonic strides): rc=%d\n", rc);
	// store view: row' = j2*N1 + k1: d1 = k1 (stride 1 row), d2 = j2 (stride N1 rows)
	uint64_t st_st[3] = {(uint64_t)pitch*16, (uint64_t)N1*pitch*16, (uint64_t)rows*pitch*16};
	rc |= b2_make_map_f64(&M.st, dB, dims, st_st, box);
	printf("encode store map: rc=%d\n", rc);
	if (rc) return 1;
	int bad = 0;
	for (int cb = 0; cb < 3; cb++) for (int j2 = 0; j2 < N2; j2++) {
		k_probe<<<1, 128, w*N1*16>>>(M, 2*w*cb + (cb == 2 ? -2*3 : 0), j2, w*N1*16);      // third block starts at column 29: unaligned + 5 columns out of range
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
	}
	cudaMemcpy(B.data(), dB, B.size()*16, cudaMemcpyDeviceToHost);
	for (int j1 = 0; j1 < N1; j1++) for (int j2 = 0; j2 < N2; j2++) for (int c = 0; c < pitch; c++) {
		std::complex<double> got = B[(size_t)(j2*N1 + j1)*pitch + c];
		std::complex<double> want = c < ncol ? std::conj(A[(size_t)(j1*N2 + j2)*pitch + c]) : std::complex<double>(-7, -7);
		if (got != want) { if (bad < 10) printf("mismatch j1=%d j2=%d c=%d got (%g,%g) want (%g,%g)\n", j1, j2, c, got.real(), got.imag(), want.real(), want.imag()); bad++; }
	}
	printf(bad ? "TMA PROBE FAILED (%d mismatches)\n" : "TMA PROBE OK: permuted strides, unaligned box start, OOB clip on store%.0d\n", bad);
	return bad != 0;
}
