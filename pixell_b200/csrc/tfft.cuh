// tfft.cuh -- the TMA-pipelined 2-D FFT path (tfft.cu) as seen by the FFT plan object (fft2d.cu)
#pragma once
#include "common.cuh"
struct TfPlan;
// float64 transforms over the last two axes of a 2- or 3-dimensional array with unit stride along the last axis
bool tfft_eligible(int kind, int dtype, int ndim, const int64_t *shape, const int64_t *istride, const int64_t *ostride, int naxes, const int *axes);
int  tfft_plan_create(TfPlan **out, int kind, int ndim, const int64_t *shape, const int64_t *istride, const int64_t *ostride);
int  tfft_execute(TfPlan *p, const void *in, void *out, int forward, double scale, cudaStream_t st);
void tfft_plan_destroy(TfPlan *p);
