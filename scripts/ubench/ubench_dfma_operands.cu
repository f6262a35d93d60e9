// DFMA issue rate against the number of distinct register operands per instruction (register-file read bandwidth):
//   A  d = fma(d, m, c)        one varying operand (m, c shared by all chains)
//   B  d = fma(x, m, d)        two varying operands (x per chain, loop invariant)
//   C  d = fma(x, y, d)        three varying operands (x, y per chain, loop invariant)
//   D  n = fma(c, g, -p); p = g; g = n      the Legendre recurrence step: three varying operands, two of them fresh results
//   E  Legendre window: per l one shared tile value t, per ring r: acc = fma(g_r, t, acc_r); g_r' = fma(ax_r, g_r, -gp_r)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dfma_operands ubench_dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
#define NCH 16
template<int MODE> __global__ void __launch_bounds__(128) k(double *out, int iters, double m_, double c_)
{
	double d[NCH], x[NCH], y[NCH], p[NCH];
	const double m = m_, c = c_;
	for (int i = 0; i < NCH; i++) { d[i] = threadIdx.x*1e-9 + i; x[i] = 1.0 + 1e-9*(i + threadIdx.x); y[i] = 1.0 - 1e-9*(i + 2*threadIdx.x); p[i] = 0.5*i; }
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int u = 0; u < 4; u++) {
			#pragma unroll
			for (int i = 0; i < NCH; i++) {
				if (MODE == 0) d[i] = fma(d[i], m, c);
				else if (MODE == 1) d[i] = fma(x[i], m, d[i]);
				else if (MODE == 2) d[i] = fma(x[i], y[i], d[i]);
				else if (MODE == 3) { double n = fma(x[i], d[i], -p[i]); p[i] = d[i]; d[i] = n; }
			}
		}
	}
	double s = 0;
	for (int i = 0; i < NCH; i++) s += d[i] + p[i];
	out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
// MODE E: R rings, per l: t (from shared memory), a (shared memory); per ring: 1 coefficient fma, 1 recurrence fma, 2 accumulations
template<int R> __global__ void __launch_bounds__(128) kE(double *out, int iters, const double *tab)
{
	__shared__ double2 T[64];
	if (threadIdx.x < 64) T[threadIdx.x] = make_double2(tab[2*threadIdx.x], tab[2*threadIdx.x + 1]);
	__syncthreads();
	double g[R], gp[R], xr[R], a0[R], a1[R];
	for (int r = 0; r < R; r++) { g[r] = 1e-3*(r + 1 + threadIdx.x); gp[r] = 0; xr[r] = 0.3 + 1e-4*(r + threadIdx.x); a0[r] = a1[r] = 0; }
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int j = 0; j < 16; j++) {
			const double2 t = T[(j + it) & 63];
			#pragma unroll
			for (int r = 0; r < R; r++) {
				a0[r] = fma(g[r], t.x, a0[r]);
				a1[r] = fma(g[r], t.y, a1[r]);
				double ng = fma(t.x*xr[r], g[r], -gp[r]);
				gp[r] = g[r]; g[r] = ng;
			}
		}
	}
	double s = 0;
	for (int r = 0; r < R; r++) s += a0[r] + a1[r] + g[r];
	out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
template<typename F> static double run(F launch, double flop)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	launch(); cudaDeviceSynchronize();
	cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	return flop/(ms*1e-3)/1e12;
}
int main()
{
	int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
	const int blocks = nsm*8, threads = 128, iters = 4000;
	double *out; cudaMalloc(&out, sizeof(double)*blocks*threads);
	double *tab; cudaMalloc(&tab, 128*sizeof(double)); cudaMemset(tab, 0, 128*sizeof(double));
	const double n = (double)blocks*threads*iters;
	printf("A one varying operand     %.2f TFLOP/s\n", run([&]{ k<0><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, n*4*NCH*2));
	printf("B two varying operands    %.2f TFLOP/s\n", run([&]{ k<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, n*4*NCH*2));
	printf("C three varying operands  %.2f TFLOP/s\n", run([&]{ k<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, n*4*NCH*2));
	printf("D recurrence step         %.2f TFLOP/s\n", run([&]{ k<3><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, n*4*NCH*2));
	printf("E Legendre window R=4     %.2f TFLOP/s (4 FP64 instructions per ring and l)\n", run([&]{ kE<4><<<blocks, threads>>>(out, iters, tab); }, n*16*4*4*2));
	printf("E Legendre window R=8     %.2f TFLOP/s\n", run([&]{ kE<8><<<blocks, threads>>>(out, iters, tab); }, n*16*8*4*2));
	return 0;
}
