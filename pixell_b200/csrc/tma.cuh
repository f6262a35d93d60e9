// tma.cuh -- Blackwell/Hopper bulk-copy primitives used by the FFT family: TMA tensor loads and stores
// (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG), 1-D bulk stores (cp.async.bulk, SASS UBLKCP), mbarrier
// transaction barriers and the proxy fences between generic and asynchronous shared-memory accesses.
// Host side: tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// ------------------------------------------------------------------------------------ device

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the asynchronous proxy (TMA completes transactions on them)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// generic-proxy writes to shared memory -> visible to the asynchronous proxy (before a TMA store reads them, and
// before a TMA load overwrites a buffer the generic proxy has used)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m)
{
	asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 4-D tiled tensor load: global (through the tensor map) -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
		:: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 4-D tiled tensor store: shared -> global; completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *m, const void *smem_src, int c0, int c1, int c2, int c3)
{
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
		:: "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 1-D bulk store of `bytes` (multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_store_1d(void *gdst, const void *smem_src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
		:: "l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
// 1-D bulk load
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until all but the N most recent bulk groups of this thread have finished READING their shared-memory source
template<int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template<int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }

// named barrier among `nthreads` threads (multiple of 32); id 0 is __syncthreads
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

// ------------------------------------------------------------------------------------ host

typedef CUresult (*b2_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline b2_encode_tiled_fn b2_get_encode_tiled()
{
	static b2_encode_tiled_fn fn = nullptr;
	if (!fn) {
		void *p = nullptr; cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
			fn = (b2_encode_tiled_fn)p;
	}
	return fn;
}

// float64 tensor of rank 4: dims[0] is the unit-stride dimension (in doubles); strides_bytes[i] is the byte stride of
// dims[i+1] (multiples of 16); box[] the tile extents.  Returns 0 on success.
static inline int b2_make_map_f64(CUtensorMap *m, const void *base, const uint64_t dims[4], const uint64_t strides_bytes[3], const uint32_t box[4])
{
	b2_encode_tiled_fn enc = b2_get_encode_tiled();
	if (!enc) return -1;
	cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
	cuuint64_t gs[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
	cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
	cuuint32_t es[4] = {1, 1, 1, 1};
	CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(base), gd, gs, bx, es,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return r == CUDA_SUCCESS ? 0 : (int)r;
}
