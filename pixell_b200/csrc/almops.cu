// almops.cu -- K6: the alm helpers of pixell.cmisc (reference cython/cmisc_core.c:16-304,
// cython/cmisc.pyx:8-191) as HBM-bound kernels.  All kernels walk the alm m-row by m-row with l as
// the fast (coalesced) index; reductions are done in a fixed order (bit-reproducible).
#include "common.cuh"
#include <algorithm>

#define B2_F64 0
#define B2_F32 1
#define MB 32        // m values per CTA in alm2cl

template<typename T> struct cplx_of;
template<> struct cplx_of<double> { typedef double2 type; };
template<> struct cplx_of<float>  { typedef float2 type; };

// ---- alm2cl: cmisc_core.c:16-110.  partial[mb][l] = sum over the m block, then a fixed-order sum.
template<typename T> __global__ void k_alm2cl_partial(int lmax, int mmax, const int64_t *mstart,
	const typename cplx_of<T>::type *a1, const typename cplx_of<T>::type *a2, double *partial)
{
	int l = blockIdx.x*blockDim.x + threadIdx.x;
	int mb = blockIdx.y;
	if (l > lmax) return;
	double acc = 0;
	const int m0 = mb*MB, m1 = min(min(m0 + MB - 1, mmax), l);
	// fixed trip count: the loads of one thread are issued together
	#pragma unroll 8
	for (int k = 0; k < MB; k++) {
		const int m = m0 + k;
		if (m <= m1) {
			int64_t i = mstart[m] + l;
			typename cplx_of<T>::type u = a1[i], v = a2[i];
			// m = 0 uses the real parts only (cmisc_core.c:26, :58, :90)
			if (m == 0) acc += (double)u.x*(double)v.x*0.5;
			else acc += (double)u.x*(double)v.x + (double)u.y*(double)v.y;
		}
	}
	partial[(int64_t)mb*(lmax + 1) + l] = acc;
}

template<typename CT> __global__ void k_alm2cl_final(int lmax, int nmb, const double *partial, CT *cl)
{
	int l = blockIdx.x*blockDim.x + threadIdx.x;
	if (l > lmax) return;
	double s = 0;
	for (int b = 0; b < nmb; b++) s += partial[(int64_t)b*(lmax + 1) + l];
	cl[l] = (CT)(s*(2.0/(2*l + 1)));
}

// ---- lmul: cmisc_core.c:159-182
template<typename T> __global__ void k_lmul(int lmax, int mmax, const int64_t *mstart,
	typename cplx_of<T>::type *alm, int lfmax, const T *lfun)
{
	int m = blockIdx.y;
	int l = m + blockIdx.x*blockDim.x + threadIdx.x;
	if (l > lmax) return;
	int64_t i = mstart[m] + l;
	T v = l <= lfmax ? lfun[l] : (T)0;
	typename cplx_of<T>::type a = alm[i];
	a.x *= v; a.y *= v;
	alm[i] = a;
}

// ---- lmatmul: cmisc_core.c:185-274.  In-place safe: a thread reads all M inputs before it writes.
#define LMAT_MAX 8
template<typename T> __global__ void k_lmatmul(int N, int M, int lmax, int mmax, const int64_t *mstart,
	const typename cplx_of<T>::type *alm, int64_t acs, int lfmax, const T *lmat,
	typename cplx_of<T>::type *oalm, int64_t ocs)
{
	int m = blockIdx.y;
	int l = m + blockIdx.x*blockDim.x + threadIdx.x;
	if (l > lmax) return;
	int64_t i = mstart[m] + l;
	typename cplx_of<T>::type in[LMAT_MAX];
	for (int c = 0; c < M; c++) in[c] = alm[c*acs + i];
	for (int r = 0; r < N; r++) {
		T vr = 0, vi = 0;
		if (l <= lfmax) for (int c = 0; c < M; c++) {
			T f = lmat[((int64_t)r*M + c)*(lfmax + 1) + l];
			vr += f*in[c].x; vi += f*in[c].y;
		}
		typename cplx_of<T>::type o; o.x = vr; o.y = vi;
		oalm[r*ocs + i] = o;
	}
}

// ---- transpose_alm: cmisc_core.c:116-156.  The k-th element in m-major address order holds the
// k-th (l,m) pair of the l-major enumeration and moves to that pair's address.
template<typename C> __global__ void k_transpose_alm(int lmax, int mmax, const int64_t *mstart, const C *ialm, C *oalm)
{
	int m = blockIdx.y;
	int l = m + blockIdx.x*blockDim.x + threadIdx.x;
	if (l > lmax) return;
	// rank of (m,l) in m-major order
	int64_t k = (int64_t)m*(lmax + 1) - (int64_t)m*(m - 1)/2 + (l - m);
	// k-th pair of the l-major enumeration: rows l' <= mmax hold l'+1 entries, later rows mmax+1
	int64_t ntri = (int64_t)(mmax + 1)*(mmax + 2)/2;
	int64_t lp, mp;
	if (k < ntri) {
		lp = (int64_t)((sqrt(8.0*(double)k + 1.0) - 1.0)*0.5);
		while (lp*(lp + 1)/2 > k) lp--;
		while ((lp + 1)*(lp + 2)/2 <= k) lp++;
		mp = k - lp*(lp + 1)/2;
	} else {
		int64_t r = k - ntri;
		lp = mmax + 1 + r/(mmax + 1);
		mp = r % (mmax + 1);
	}
	oalm[mstart[mp] + lp] = ialm[mstart[m] + l];
}

// ---- transfer_alm: cmisc.pyx:131-150
template<typename C> __global__ void k_transfer_alm(int lmax, int mmax, const int64_t *ms1, int64_t ls1, const C *a1,
	const int64_t *ms2, int64_t ls2, C *a2)
{
	int m = blockIdx.y;
	int l = m + blockIdx.x*blockDim.x + threadIdx.x;
	if (l > lmax) return;
	a2[ms2[m] + l*ls2] = a1[ms1[m] + l*ls1];
}

// ------------------------------------------------------------------------------------ host side

struct Staged {       // host <-> device staging of one buffer
	void *dev = nullptr; void *host = nullptr; size_t bytes = 0; bool owned = false, out = false;
	int in(const void *p, size_t nbytes, int mem, bool copy_in, bool copy_out, cudaStream_t st) {
		bytes = nbytes; out = copy_out;
		if (mem == 1 || p == nullptr) { dev = (void*)p; return 0; }
		host = (void*)p; owned = true;
		B2_CHECK(cudaMalloc(&dev, std::max<size_t>(nbytes, 16)));
		if (copy_in) B2_CHECK(cudaMemcpyAsync(dev, p, nbytes, cudaMemcpyHostToDevice, st));
		return 0;
	}
	int finish(cudaStream_t st) {
		if (owned && out) B2_CHECK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
		return 0;
	}
	~Staged() { if (owned && dev) cudaFree(dev); }
};

// device copies of mstart arrays and the alm2cl scratch are kept between calls (a cudaMalloc / cudaFree pair per
// call would cost more than the kernels); one small cache per process, guarded by a mutex
#include <mutex>
#include <list>
struct MsEntry { int dev; std::vector<int64_t> key; DevBuf<int64_t> buf; };
static std::mutex g_ms_mutex;
static std::list<MsEntry> g_ms_cache;

static int mstart_dev(const int64_t *mstart, int n, const int64_t **out)
{
	int dev = 0; B2_CHECK(cudaGetDevice(&dev));
	std::lock_guard<std::mutex> lock(g_ms_mutex);
	for (auto it = g_ms_cache.begin(); it != g_ms_cache.end(); ++it)
		if (it->dev == dev && (int)it->key.size() == n && std::equal(mstart, mstart + n, it->key.begin())) {
			g_ms_cache.splice(g_ms_cache.begin(), g_ms_cache, it);
			*out = g_ms_cache.front().buf.p; return 0;
		}
	if (g_ms_cache.size() >= 16) g_ms_cache.pop_back();
	g_ms_cache.emplace_front();
	MsEntry &e = g_ms_cache.front();
	e.dev = dev; e.key.assign(mstart, mstart + n);
	if (e.buf.upload(e.key)) { g_ms_cache.pop_front(); return 1; }
	*out = e.buf.p;
	return 0;
}

static int64_t span_of(int lmax, int mmax, const int64_t *mstart, int64_t lstride)
{
	int64_t hi = 0;
	for (int m = 0; m <= mmax; m++) hi = std::max(hi, mstart[m] + (int64_t)lmax*lstride + 1);
	return hi;
}

static int check_mstart(int lmax, int mmax, const int64_t *mstart)
{
	B2_REQUIRE(lmax >= 0 && mmax >= 0 && mmax <= lmax && mstart, "bad alm layout (lmax=%d mmax=%d)", lmax, mmax);
	for (int m = 0; m <= mmax; m++) B2_REQUIRE(mstart[m] + m >= 0, "negative alm index for m=%d", m);
	return 0;
}

extern "C" int b2_alm2cl(int lmax, int mmax, const int64_t *mstart, int dtype, const void *alm1, const void *alm2,
	int cl_dtype, void *cl, int mem, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (check_mstart(lmax, mmax, mstart)) return 1;
	B2_REQUIRE(alm1 && cl, "alm2cl: null buffer");
	if (!alm2) alm2 = alm1;
	size_t esz = dtype == B2_F64 ? 16 : 8, csz = cl_dtype == B2_F64 ? 8 : 4;
	int64_t span = span_of(lmax, mmax, mstart, 1);
	const int64_t *msp; if (mstart_dev(mstart, mmax + 1, &msp)) return 1;
	Staged a, b, c;
	if (a.in(alm1, span*esz, mem, true, false, st)) return 1;
	if (alm2 == alm1) b.dev = a.dev; else if (b.in(alm2, span*esz, mem, true, false, st)) return 1;
	if (c.in(cl, (lmax + 1)*csz, mem, false, true, st)) return 1;
	int nmb = mmax/MB + 1;
	// scratch from the stream-ordered pool: per call, per device, safe with concurrent streams / threads
	struct { double *p; } partial = {nullptr};
	B2_CHECK(cudaMallocAsync((void**)&partial.p, sizeof(double)*(size_t)nmb*(lmax + 1), st));
	dim3 grid((lmax + 256)/256, nmb);
	if (dtype == B2_F64) k_alm2cl_partial<double><<<grid, 256, 0, st>>>(lmax, mmax, msp, (const double2*)a.dev, (const double2*)b.dev, partial.p);
	else                 k_alm2cl_partial<float><<<grid, 256, 0, st>>>(lmax, mmax, msp, (const float2*)a.dev, (const float2*)b.dev, partial.p);
	B2_LAUNCH_CHECK();
	if (cl_dtype == B2_F64) k_alm2cl_final<double><<<(lmax + 256)/256, 256, 0, st>>>(lmax, nmb, partial.p, (double*)c.dev);
	else                    k_alm2cl_final<float><<<(lmax + 256)/256, 256, 0, st>>>(lmax, nmb, partial.p, (float*)c.dev);
	B2_LAUNCH_CHECK();
	B2_CHECK(cudaFreeAsync(partial.p, st));
	if (c.finish(st)) return 1;
	if (mem == 0) B2_CHECK(cudaStreamSynchronize(st));      // staging buffers die here
	return 0;
}

extern "C" int b2_lmul(int lmax, int mmax, const int64_t *mstart, int dtype, void *alm, int lfmax, const void *lfun,
	int mem, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (check_mstart(lmax, mmax, mstart)) return 1;
	B2_REQUIRE(alm && lfun && lfmax >= 0, "lmul: bad arguments");
	size_t esz = dtype == B2_F64 ? 16 : 8;
	int64_t span = span_of(lmax, mmax, mstart, 1);
	const int64_t *msp; if (mstart_dev(mstart, mmax + 1, &msp)) return 1;
	Staged a, f;
	if (a.in(alm, span*esz, mem, true, true, st)) return 1;
	if (f.in(lfun, (size_t)(lfmax + 1)*esz/2, mem, true, false, st)) return 1;
	dim3 grid((lmax + 256)/256, mmax + 1);
	if (dtype == B2_F64) k_lmul<double><<<grid, 256, 0, st>>>(lmax, mmax, msp, (double2*)a.dev, lfmax, (const double*)f.dev);
	else                 k_lmul<float><<<grid, 256, 0, st>>>(lmax, mmax, msp, (float2*)a.dev, lfmax, (const float*)f.dev);
	B2_LAUNCH_CHECK();
	if (a.finish(st)) return 1;
	if (mem == 0) B2_CHECK(cudaStreamSynchronize(st));
	return 0;
}

extern "C" int b2_lmatmul(int N, int M, int lmax, int mmax, const int64_t *mstart, int dtype,
	const void *alm, int64_t acs, int lfmax, const void *lmat, void *oalm, int64_t ocs, int mem, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (check_mstart(lmax, mmax, mstart)) return 1;
	B2_REQUIRE(alm && lmat && oalm && lfmax >= 0, "lmatmul: bad arguments");
	B2_REQUIRE(N >= 1 && M >= 1 && M <= LMAT_MAX && N <= LMAT_MAX, "lmatmul supports up to %d components", LMAT_MAX);
	size_t esz = dtype == B2_F64 ? 16 : 8;
	int64_t span = span_of(lmax, mmax, mstart, 1);
	const int64_t *msp; if (mstart_dev(mstart, mmax + 1, &msp)) return 1;
	Staged a, o, f;
	bool inplace = (alm == oalm);
	if (mem == 0) {
		B2_REQUIRE(acs >= span && ocs >= span, "lmatmul: component stride smaller than the alm span");
		if (a.in(alm, ((M - 1)*acs + span)*esz, mem, true, false, st)) return 1;
		if (inplace) { o.dev = a.dev; }
		else if (o.in(oalm, ((N - 1)*ocs + span)*esz, mem, true, true, st)) return 1;
	} else { a.dev = (void*)alm; o.dev = oalm; }
	if (f.in(lmat, (size_t)N*M*(lfmax + 1)*esz/2, mem, true, false, st)) return 1;
	dim3 grid((lmax + 256)/256, mmax + 1);
	if (dtype == B2_F64) k_lmatmul<double><<<grid, 256, 0, st>>>(N, M, lmax, mmax, msp, (const double2*)a.dev, acs, lfmax, (const double*)f.dev, (double2*)o.dev, ocs);
	else                 k_lmatmul<float><<<grid, 256, 0, st>>>(N, M, lmax, mmax, msp, (const float2*)a.dev, acs, lfmax, (const float*)f.dev, (float2*)o.dev, ocs);
	B2_LAUNCH_CHECK();
	if (mem == 0) {
		if (inplace) B2_CHECK(cudaMemcpyAsync(oalm, a.dev, ((N - 1)*ocs + span)*esz, cudaMemcpyDeviceToHost, st));
		else if (o.finish(st)) return 1;
	}
	if (mem == 0) B2_CHECK(cudaStreamSynchronize(st));
	return 0;
}

extern "C" int b2_transpose_alm(int lmax, int mmax, const int64_t *mstart, int dtype, const void *ialm, void *oalm,
	int mem, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (check_mstart(lmax, mmax, mstart)) return 1;
	B2_REQUIRE(ialm && oalm && ialm != oalm, "transpose_alm: needs distinct input and output");
	size_t esz = dtype == B2_F64 ? 16 : 8;
	int64_t span = span_of(lmax, mmax, mstart, 1);
	const int64_t *msp; if (mstart_dev(mstart, mmax + 1, &msp)) return 1;
	Staged a, o;
	if (a.in(ialm, span*esz, mem, true, false, st)) return 1;
	if (o.in(oalm, span*esz, mem, true, true, st)) return 1;
	dim3 grid((lmax + 256)/256, mmax + 1);
	if (dtype == B2_F64) k_transpose_alm<double2><<<grid, 256, 0, st>>>(lmax, mmax, msp, (const double2*)a.dev, (double2*)o.dev);
	else                 k_transpose_alm<float2><<<grid, 256, 0, st>>>(lmax, mmax, msp, (const float2*)a.dev, (float2*)o.dev);
	B2_LAUNCH_CHECK();
	if (o.finish(st)) return 1;
	if (mem == 0) B2_CHECK(cudaStreamSynchronize(st));
	return 0;
}

extern "C" int b2_transfer_alm(int lmax1, int mmax1, const int64_t *mstart1, int64_t ls1, const void *alm1,
	int lmax2, int mmax2, const int64_t *mstart2, int64_t ls2, void *alm2, int dtype, int mem, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (check_mstart(lmax1, mmax1, mstart1) || check_mstart(lmax2, mmax2, mstart2)) return 1;
	B2_REQUIRE(alm1 && alm2 && ls1 >= 1 && ls2 >= 1, "transfer_alm: bad arguments");
	size_t esz = dtype == B2_F64 ? 16 : 8;
	int lmax = std::min(lmax1, lmax2), mmax = std::min(mmax1, mmax2);
	const int64_t *m1p, *m2p;
	if (mstart_dev(mstart1, mmax1 + 1, &m1p) || mstart_dev(mstart2, mmax2 + 1, &m2p)) return 1;
	Staged a, o;
	if (a.in(alm1, span_of(lmax1, mmax1, mstart1, ls1)*esz, mem, true, false, st)) return 1;
	if (o.in(alm2, span_of(lmax2, mmax2, mstart2, ls2)*esz, mem, true, true, st)) return 1;
	dim3 grid((lmax + 256)/256, mmax + 1);
	if (dtype == B2_F64) k_transfer_alm<double2><<<grid, 256, 0, st>>>(lmax, mmax, m1p, ls1, (const double2*)a.dev, m2p, ls2, (double2*)o.dev);
	else                 k_transfer_alm<float2><<<grid, 256, 0, st>>>(lmax, mmax, m1p, ls1, (const float2*)a.dev, m2p, ls2, (float2*)o.dev);
	B2_LAUNCH_CHECK();
	if (o.finish(st)) return 1;
	if (mem == 0) B2_CHECK(cudaStreamSynchronize(st));
	return 0;
}

// ------------------------------------------------------------------------------------ K8: random alm on the device
// curvedsky.rand_alm (pixell/curvedsky.py:61-77) = rand_alm_white (:620-628: unit normals filled in memory order of the
// l-major array, fill_gauss :602-606, then transposed to m-major) + colouring with ps^(1/2)/sqrt(2) (:70-72, lmul) + the
// m = 0 fix (:73-74), fused: one thread per (l, m) draws the white pairs of every component, mixes them and writes the
// m-major element.  The stream is counter based (Philox4x32-10, key = seed): the pair of element j of the l-major array of
// component c, j = l (l + 1)/2 + m (mmax = lmax; rectangular part after the triangle otherwise), comes from counter
// c nlm + j -- the reference's fill order, so realisations at two lmax share their large scales in component 0, as the
// reference's do.  Box-Muller on two 53-bit uniforms.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
	#pragma unroll
	for (int r = 0; r < 10; r++) {
		const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u*c0;
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u*c2;
		const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double2 philox_normal_pair(uint64_t counter, uint64_t seed)
{
	uint32_t x[4];
	philox4x32_10((uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
	const uint64_t a = ((uint64_t)x[1] << 32) | x[0], b = ((uint64_t)x[3] << 32) | x[2];
	const double u1 = ((double)(a >> 11) + 0.5)*0x1p-53, u2 = ((double)(b >> 11) + 0.5)*0x1p-53;      // (0, 1)
	const double r = sqrt(-2.0*log(u1));
	double sn, cs; sincospi(2.0*u2, &sn, &cs);
	return make_double2(r*cs, r*sn);
}

#define RA_MAXC 4
template<typename T> __global__ void k_rand_alm(int lmax, int mmax, const int64_t *mstart, int ncomp, uint64_t seed, int64_t nlm,
	const double *ps12 /* [ncomp][ncomp][lmax+1] or null: white */, typename cplx_of<T>::type *alm, int64_t cstride)
{
	const int m = blockIdx.y;
	const int l = m + blockIdx.x*blockDim.x + threadIdx.x;
	if (l > lmax) return;
	// position in the l-major array (cmisc_core.c:116-156: the triangle l <= mmax first, then rows of mmax + 1)
	const int64_t j = l <= mmax ? (int64_t)l*(l + 1)/2 + m : (int64_t)(mmax + 1)*(mmax + 2)/2 + (int64_t)(l - mmax - 1)*(mmax + 1) + m;
	double2 w[RA_MAXC];
	for (int c = 0; c < ncomp; c++) w[c] = philox_normal_pair((uint64_t)(c*nlm + j), seed);
	const int64_t i = mstart[m] + l;
	for (int r = 0; r < ncomp; r++) {
		double2 v;
		if (ps12) {
			v = make_double2(0, 0);
			for (int c = 0; c < ncomp; c++) { const double f = ps12[((int64_t)r*ncomp + c)*(lmax + 1) + l]*0.70710678118654752440; v.x += f*w[c].x; v.y += f*w[c].y; }
			if (m == 0) { v.x *= 1.41421356237309504880; v.y = 0; }
		} else v = w[r];
		typename cplx_of<T>::type o; o.x = (T)v.x; o.y = (T)v.y;
		alm[r*cstride + i] = o;
	}
}

extern "C" int b2_rand_alm(int lmax, int mmax, const int64_t *mstart, int ncomp, uint64_t seed, const double *ps12,
	int dtype, void *alm, int64_t alm_cstride, int mem, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (check_mstart(lmax, mmax, mstart)) return 1;
	B2_REQUIRE(alm && ncomp >= 1 && ncomp <= RA_MAXC, "rand_alm: 1 to %d components are supported", RA_MAXC);
	B2_REQUIRE(dtype == B2_F64 || dtype == B2_F32, "rand_alm: bad dtype");
	const size_t esz = dtype == B2_F64 ? 16 : 8;
	const int64_t span = span_of(lmax, mmax, mstart, 1);
	B2_REQUIRE(ncomp == 1 || alm_cstride >= span, "rand_alm: component stride shorter than one alm");
	int64_t nlm = 0; for (int m = 0; m <= mmax; m++) nlm += lmax - m + 1;
	const int64_t *msp; if (mstart_dev(mstart, mmax + 1, &msp)) return 1;
	Staged a, f;
	const size_t abytes = ((size_t)(ncomp - 1)*alm_cstride + span)*esz;
	if (a.in(alm, abytes, mem, true, true, st)) return 1;      // entries outside the layout keep the caller's values
	if (ps12 && f.in(ps12, sizeof(double)*(size_t)ncomp*ncomp*(lmax + 1), mem, true, false, st)) return 1;
	dim3 grid((lmax + 256)/256, mmax + 1);
	if (dtype == B2_F64) k_rand_alm<double><<<grid, 256, 0, st>>>(lmax, mmax, msp, ncomp, seed, nlm, (const double*)f.dev, (double2*)a.dev, alm_cstride);
	else                 k_rand_alm<float><<<grid, 256, 0, st>>>(lmax, mmax, msp, ncomp, seed, nlm, (const double*)f.dev, (float2*)a.dev, alm_cstride);
	B2_LAUNCH_CHECK();
	if (a.finish(st)) return 1;
	if (mem == 0) B2_CHECK(cudaStreamSynchronize(st));
	return 0;
}
