// common.cuh -- shared declarations of libb200sht (error handling, device helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <vector>
#include <string>

void b2_set_error(const char *fmt, ...);

#define B2_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	b2_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return 1; } } while (0)
#define B2_REQUIRE(cond, ...) do { if (!(cond)) { b2_set_error(__VA_ARGS__); return 1; } } while (0)
extern long long g_b2_launches;
#define B2_LAUNCH_CHECK() do { g_b2_launches++; B2_CHECK(cudaGetLastError()); } while (0)

static inline int64_t b2_round_up(int64_t a, int64_t b) { return (a + b - 1)/b*b; }

// owned device buffer (freed with the plan)
template<typename T> struct DevBuf {
	T *p = nullptr; size_t n = 0;
	int alloc(size_t count) {
		release(); n = count;
		if (count == 0) return 0;
		cudaError_t e = cudaMalloc((void**)&p, count*sizeof(T));
		if (e != cudaSuccess) { b2_set_error("cudaMalloc(%zu bytes) failed: %s", count*sizeof(T), cudaGetErrorString(e)); p = nullptr; n = 0; return 1; }
		return 0;
	}
	int upload(const std::vector<T> &h) {
		if (alloc(h.size())) return 1;
		if (h.empty()) return 0;
		cudaError_t e = cudaMemcpy(p, h.data(), h.size()*sizeof(T), cudaMemcpyHostToDevice);
		if (e != cudaSuccess) { b2_set_error("cudaMemcpy H2D failed: %s", cudaGetErrorString(e)); return 1; }
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
	size_t bytes() const { return n*sizeof(T); }
	~DevBuf() { release(); }
	DevBuf() {}
	DevBuf(const DevBuf&) = delete; DevBuf &operator=(const DevBuf&) = delete;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { /* a*conj(b) */ return make_double2(a.x*b.x + a.y*b.y, a.y*b.x - a.x*b.y); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x+b.x, a.y+b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x-b.x, a.y-b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x*s, a.y*s); }

// cp.async (LDGSTS): global -> shared copies that hold no registers and do not stall the issuing thread;
// .ca keeps the line in L1 (small tables that are re-read), .cg goes through L2 only (streamed map data)
__device__ __forceinline__ void cp_async8(void *sm, const void *g)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(sm)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(void *sm, const void *g)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(sm)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void *sm, const void *g)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(sm)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
