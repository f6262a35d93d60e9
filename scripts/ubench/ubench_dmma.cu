// FP64 throughput of the vector pipe (DFMA), of the tensor pipe (DMMA m8n8k4) and of both issued together, per SM and chip.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// MODE 0: DFMA only (NF chains), 1: DMMA only (NM accumulator pairs), 2: both
template<int MODE, int NF, int NM> __global__ void __launch_bounds__(256) k(double *out, int iters)
{
	double f[NF > 0 ? NF : 1], c[NM > 0 ? 2*NM : 2];
	const double m = 1.0000001, add = 1e-9;
	for (int i = 0; i < NF; i++) f[i] = threadIdx.x*1e-9 + i;
	for (int i = 0; i < 2*NM; i++) c[i] = 0;
	double a = 1.0 + threadIdx.x*1e-12, b = 1e-9*threadIdx.x;
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int rep = 0; rep < 4; rep++) {
			if (MODE != 1) {
				#pragma unroll
				for (int i = 0; i < NF; i++) f[i] = fma(f[i], m, add);
			}
			if (MODE != 0) {
				#pragma unroll
				for (int i = 0; i < NM; i++) dmma(c[2*i], c[2*i + 1], a, b);
			}
		}
	}
	double s = 0;
	for (int i = 0; i < NF; i++) s += f[i];
	for (int i = 0; i < 2*NM; i++) s += c[i];
	out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

template<int MODE, int NF, int NM> static void run(const char *name, int nsm, int warps_per_sm)
{
	int nb = nsm*(warps_per_sm*32/256 > 0 ? warps_per_sm*32/256 : 1), nt = warps_per_sm*32 >= 256 ? 256 : warps_per_sm*32;
	double *buf; cudaMalloc(&buf, sizeof(double)*nb*nt);
	int iters = 4096;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int r = 0; r < 4; r++) {
		cudaEventRecord(e0); k<MODE, NF, NM><<<nb, nt>>>(buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
	}
	double nwarp = (double)nb*nt/32;
	double fma_v = (MODE != 1) ? nwarp*32.0*NF*4*iters : 0, fma_t = (MODE != 0) ? nwarp*256.0*NM*4*iters : 0;
	printf("%-34s warps/SM %2d  %8.3f ms  vector %6.2f TF  tensor %6.2f TF  total %6.2f TFLOP/s\n", name, warps_per_sm, best,
		2*fma_v/best/1e9, 2*fma_t/best/1e9, 2*(fma_v + fma_t)/best/1e9);
	cudaFree(buf);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int nsm = p.multiProcessorCount;
	printf("%s, %d SMs\n", p.name, nsm);
	for (int w : {8, 16, 32}) {
		run<0, 8, 0>("DFMA only (8 chains)", nsm, w);
		run<1, 0, 4>("DMMA only (4 accumulator pairs)", nsm, w);
		run<1, 0, 8>("DMMA only (8 accumulator pairs)", nsm, w);
		run<2, 8, 1>("DFMA x8 + DMMA x1 per step", nsm, w);
		run<2, 8, 2>("DFMA x8 + DMMA x2 per step", nsm, w);
		run<2, 4, 2>("DFMA x4 + DMMA x2 per step", nsm, w);
		run<2, 4, 4>("DFMA x4 + DMMA x4 per step", nsm, w);
	}
	return 0;
}
