// fft2d.cu -- K7: batched 1-D / 2-D DFT over strided arrays (the pixell.fft engine plug-in).
#include "../../include/b200sht.h"
#include "fft_smem.cuh"

struct b2_fft_plan { int dummy; };

extern "C" int b2_fft_plan_create(b2_fft_plan **out, int ndim, const int64_t *shape, const int64_t *istride,
	const int64_t *ostride, int naxes, const int *axes, int kind, int dtype)
{
	b2_set_error("b2_fft_plan_create: the 2-D FFT engine is not built yet");
	return 1;
}
extern "C" int b2_fft_execute(b2_fft_plan *plan, const void *in, void *out, int forward, double scale, int mem, void *stream)
{
	b2_set_error("b2_fft_execute: the 2-D FFT engine is not built yet");
	return 1;
}
extern "C" void b2_fft_plan_destroy(b2_fft_plan *plan) { delete plan; }
