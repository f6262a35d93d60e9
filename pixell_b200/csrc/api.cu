// api.cu -- the C ABI of libb200sht.so (include/b200sht.h): plans, transforms, grid weights.
#include "../../include/b200sht.h"
#include "plan.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <thread>
#include <atomic>

// ------------------------------------------------------------------------------------ errors

static thread_local char g_err[1024] = "";
void b2_set_error(const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char *b2_last_error(void) { return g_err; }
extern "C" int b2_version(void) { return 100; }
long long g_b2_launches = 0;
extern "C" int64_t b2_launch_count(void) { return g_b2_launches; }

extern "C" int b2_init(int device)
{
	int n = 0;
	B2_CHECK(cudaGetDeviceCount(&n));
	B2_REQUIRE(device >= 0 && device < n, "device %d out of range (%d visible)", device, n);
	B2_CHECK(cudaSetDevice(device));
	B2_CHECK(cudaFree(0));
	return 0;
}
extern "C" int b2_device_synchronize(void) { B2_CHECK(cudaDeviceSynchronize()); return 0; }
extern "C" int b2_dfma_peak_gflops(double *out) { return dfma_peak_gflops(out); }
extern "C" int b2_set_leg_variant(int which, int v) { return leg_set_variant(which, v); }

// ------------------------------------------------------------------------------------ grids

int grid_theta_host(const char *g, int n, std::vector<double> &theta)
{
	std::string s(g);
	theta.resize(n);
	for (int k = 0; k < n; k++) {
		// ring colatitudes of the named equiangular grids (pole offsets: pixell/curvedsky.py:1334-1342)
		if      (s == "CC")     theta[k] = n > 1 ? k*M_PI/(n - 1) : 0.0;
		else if (s == "F1")     theta[k] = (k + 0.5)*M_PI/n;
		else if (s == "MW")     theta[k] = (2*k + 1)*M_PI/(2*n - 1);
		else if (s == "MWflip") theta[k] = 2*k*M_PI/(2*n - 1);
		else if (s == "DH")     theta[k] = k*M_PI/n;
		else if (s == "F2")     theta[k] = (k + 1)*M_PI/(n + 1);
		else { b2_set_error("unknown geometry '%s'", g); return 1; }
		if (theta[k] > M_PI) theta[k] = M_PI;      // k pi/(n-1) can round one ulp above pi for the last Clenshaw-Curtis ring
	}
	return 0;
}

// interpolatory ring weights (sum = 4 pi).  kind 0: Fourier-series rule on a circle of N points with
// node offset o2/2 (CC, F1, MW, MWflip); kind 1: Fejer-2 nodes (k+1) pi/(nn+1), shifted by `shift` rings (DH)
__global__ void k_gridweights(double *w, int n, int kind, int N, int o2, int K, int cc, int nn, int shift)
{
	int k = blockIdx.x*blockDim.x + threadIdx.x;
	if (k >= n) return;
	if (kind == 0) {
		long long t = 2LL*k + o2;                      // theta_k = t pi / N
		double sum = 0;
		for (int j = K/2; j >= 1; j--) {
			double c = (cc && 2*j == n - 1) ? 1.0 : 2.0;
			long long r = (2LL*j*t) % (2LL*N);
			sum += c*cospi((double)r/(double)N)/(4.0*j*j - 1.0);
		}
		bool single = (t % N) == 0;                    // pole ring: appears once on the circle
		w[k] = 4.0*M_PI/N*(single ? 1.0 : 2.0)*(1.0 - sum);
	} else {
		int kk = k - shift;
		if (kk < 0) { w[k] = 0; return; }
		long long t = kk + 1;                          // theta = t pi/(nn+1)
		double sum = 0;
		for (int j = (nn + 1)/2; j >= 1; j--) {
			long long r = ((2LL*j - 1)*t) % (2LL*(nn + 1));
			sum += sinpi((double)r/(double)(nn + 1))/(2.0*j - 1.0);
		}
		w[k] = 2.0*M_PI*(4.0/(nn + 1))*sinpi((double)t/(double)(nn + 1))*sum;
	}
}

int gridweights_device(const char *g, int n, double *out_host, double *out_dev)
{
	std::string s(g);
	B2_REQUIRE(n >= 1, "get_gridweights: ntheta must be positive");
	DevBuf<double> tmp;
	double *d = out_dev;
	if (!d) { if (tmp.alloc(n)) return 1; d = tmp.p; }
	int kind = 0, N = 0, o2 = 0, cc = 0, nn = 0, shift = 0;
	if      (s == "CC")     { N = 2*(n - 1); o2 = 0; cc = 1; }
	else if (s == "F1")     { N = 2*n;       o2 = 1; }
	else if (s == "MW")     { N = 2*n - 1;   o2 = 1; }
	else if (s == "MWflip") { N = 2*n - 1;   o2 = 0; }
	else if (s == "F2")     { kind = 1; nn = n; shift = 0; }
	else if (s == "DH")     { kind = 1; nn = n - 1; shift = 1; }
	else { b2_set_error("unknown geometry '%s'", g); return 1; }
	if (kind == 0 && N == 0) {        // CC with a single ring
		double v = 4.0*M_PI;
		B2_CHECK(cudaMemcpy(d, &v, sizeof(double), cudaMemcpyHostToDevice));
	} else {
		k_gridweights<<<(n + 127)/128, 128>>>(d, n, kind, N, o2, n - 1, cc, nn, shift);
		B2_LAUNCH_CHECK();
	}
	if (out_host) B2_CHECK(cudaMemcpy(out_host, d, n*sizeof(double), cudaMemcpyDeviceToHost));
	else B2_CHECK(cudaDeviceSynchronize());
	return 0;
}

extern "C" int b2_gridweights(const char *geometry, int ntheta, double *out)
{
	B2_REQUIRE(geometry && out, "get_gridweights: null argument");
	return gridweights_device(geometry, ntheta, out, nullptr);
}

// ------------------------------------------------------------------------------------ plans

b2_sht_plan::~b2_sht_plan()
{
	for (auto &e : ev) if (e) cudaEventDestroy(e);
	for (auto &e : gev) if (e) cudaEventDestroy(e);
	for (auto &g : gstreams) if (g) cudaStreamDestroy(g);
	for (auto &e : gjoin) if (e) cudaEventDestroy(e);
	for (auto &e : sev) if (e) cudaEventDestroy(e);
	if (gfork) cudaEventDestroy(gfork);
	if (s_in) cudaStreamDestroy(s_in);
	if (s_out) cudaStreamDestroy(s_out);
	if (s_comp) cudaStreamDestroy(s_comp);
	if (s_fft) cudaStreamDestroy(s_fft);
	if (sig_flag_h) cudaFreeHost(sig_flag_h);
	if (gate_src_h) cudaFreeHost(gate_src_h);
	if (ev_last) cudaEventDestroy(ev_last);
	if (ev_fork) cudaEventDestroy(ev_fork);
	if (ev_join) cudaEventDestroy(ev_join);
	for (auto &e : ev_ready) if (e) cudaEventDestroy(e);
	for (auto &e : ev_free) if (e) cudaEventDestroy(e);
	for (auto &e : ev_chunk) if (e) cudaEventDestroy(e);
}

size_t b2_sht_plan::bytes() const
{
	size_t b = mstart.bytes() + geom.bytes() + fft.bytes() + leg.bytes() + leg2.bytes() + legb.bytes() + w2d.bytes() + stage_alm.bytes() + stage_map.bytes();
	for (auto &g : groups) b += g->bytes();
	b += pack.bytes();
	for (auto &t : tables) b += t.second->bytes();
	for (auto &t : starts) if (t.second) b += t.second->bytes();
	if (resamp) b += resamp->bytes();
	return b;
}

LegTables *b2_sht_plan::get_tables(int spin)
{
	auto it = tables.find(spin);
	if (it != tables.end()) return it->second.get();
	std::unique_ptr<LegTables> t(new LegTables());
	if (t->build(lmax, mmax, spin)) return nullptr;
	LegTables *p = t.get();
	tables[spin] = std::move(t);
	return p;
}

LegStart *b2_sht_plan::get_start(int spin)
{
	static const bool off = getenv("B2_NO_START_TABLE") && atoi(getenv("B2_NO_START_TABLE")) != 0;
	if (off) return nullptr;
	auto it = starts.find(spin);
	if (it != starts.end()) return it->second.get();
	LegTables *T = get_tables(spin);
	if (!T) return nullptr;
	std::unique_ptr<LegStart> s(new LegStart());
	if (leg_build_start(*s, *T, geom)) { starts[spin] = nullptr; return nullptr; }      // e.g. out of memory: run without
	LegStart *p = s.get();
	starts[spin] = std::move(s);
	return p;
}

// chunks of the streamed host-memory path (see b2_sht_plan::schunks / mcuts).  Both lists shrink towards their end: what
// stays exposed is the copy of the last chunk, and the later chunks are the expensive ones to compute (rings near the
// equator, resp. nothing left to hide behind), so their copies still finish under the next chunk's kernels.
static void plan_stream_setup(b2_sht_plan *p)
{
	static const int enable = getenv("B2_STREAM_CHUNKS") ? atoi(getenv("B2_STREAM_CHUNKS")) : 1;
	p->schunks.clear(); p->mcuts.clear();
	if (enable < 1) return;
	const int np = p->geom.npair_pad;
	const bool dense_rows = p->nphi > 0 && p->nring > 1 && (p->row_pitch == p->npix || p->row_pitch == -p->npix);
	if (dense_rows && np >= 2048) {
		// units of 256 ring pairs, shares 6 : 4 : 3 : 2 : 1
		const int unit = 256, U = np/unit;
		static const int share[5] = {6, 4, 3, 2, 1};
		std::vector<int> bounds(1, 0);
		for (int k = 0, acc = 0; k < 5; k++) {
			acc += share[k];
			int b = std::max(bounds.back() + 1, (U*acc + 8)/16);
			if (k == 4) b = U;
			if (b > U) b = U;
			if (b > bounds.back()) bounds.push_back(b);
		}
		std::vector<b2_sht_plan::StreamChunk> ch;
		bool ok = true;
		for (size_t k = 0; k + 1 < bounds.size() && ok; k++) {
			b2_sht_plan::StreamChunk c; c.pair_lo = bounds[k]*unit; c.pair_hi = (k + 2 == bounds.size()) ? np : bounds[k + 1]*unit; c.nrun = 0;
			std::vector<int> rings;
			for (int i = c.pair_lo; i < c.pair_hi; i++) { if (p->geom.rn_h[i] >= 0) rings.push_back(p->geom.rn_h[i]); if (p->geom.rs_h[i] >= 0) rings.push_back(p->geom.rs_h[i]); }
			std::sort(rings.begin(), rings.end());
			for (size_t i = 0; i < rings.size() && ok; ) {
				size_t j = i + 1;
				while (j < rings.size() && rings[j] == rings[j - 1] + 1) j++;
				if (c.nrun == 2) { ok = false; break; }
				c.r0[c.nrun] = rings[i]; c.nr[c.nrun] = (int)(j - i); c.nrun++;
				i = j;
			}
			if (c.nrun > 0) ch.push_back(c);
		}
		if (ok && ch.size() >= 2 && ch.size() <= 8) p->schunks = ch;
	}
	if (p->alm_dense && p->lstride == 1 && p->mmax >= 256) {
		bool inc = true;
		for (int m = 1; m <= p->mmax && inc; m++) inc = p->mstart_h[m] + m == p->mstart_h[m - 1] + p->lmax + 1;
		if (inc) {
			// ranges of m holding 5 : 5 : 5 : 4 : 3 : 2 of the coefficients
			static const int share[6] = {5, 5, 5, 4, 3, 2};
			int64_t total = 0, acc = 0; for (int m = 0; m <= p->mmax; m++) total += p->lmax - m + 1;
			p->mcuts.push_back(0);
			int64_t want = share[0];
			for (int m = 0, k = 0; m <= p->mmax && k < 5; m++) {
				acc += p->lmax - m + 1;
				if (acc*24 >= total*want) { if (m + 1 <= p->mmax) p->mcuts.push_back(m + 1); k++; want += share[k]; }
			}
			if (p->mcuts.back() != p->mmax + 1) p->mcuts.push_back(p->mmax + 1);
		}
	}
}

static int plan_common(b2_sht_plan *p, int nring, const double *theta, int64_t nphi, double phi0, int xdir,
	int64_t npix, const int64_t *ringstart, const double *weight, int lmax, int mmax, const int64_t *mstart, int64_t lstride)
{
	B2_REQUIRE(nring >= 1 && theta && ringstart, "plan: need at least one ring");
	B2_REQUIRE(lmax >= 0 && mmax >= 0 && mmax <= lmax, "plan: need 0 <= mmax <= lmax (got lmax=%d mmax=%d)", lmax, mmax);
	B2_REQUIRE(mstart && lstride >= 1, "plan: bad alm layout");
	for (int r = 0; r < nring; r++) {
		B2_REQUIRE(theta[r] >= 0 && theta[r] <= M_PI, "plan: theta[%d]=%g outside [0,pi]", r, theta[r]);
		B2_REQUIRE(ringstart[r] >= 0, "plan: negative ringstart");
	}
	p->lmax = lmax; p->mmax = mmax; p->lstride = lstride;
	p->mstart_h.assign(mstart, mstart + mmax + 1);
	p->alm_span = 0;
	for (int m = 0; m <= mmax; m++) {
		B2_REQUIRE(mstart[m] + (int64_t)m*lstride >= 0, "plan: negative alm index for m=%d", m);
		p->alm_span = std::max(p->alm_span, mstart[m] + (int64_t)lmax*lstride + 1);
	}
	{
		int64_t owned = 0;
		for (int m = 0; m <= mmax; m++) owned += lmax - m + 1;
		p->alm_dense = (lstride == 1 && owned == p->alm_span);
	}
	if (p->mstart.upload(p->mstart_h)) return 1;
	p->nring = nring; p->nphi = nphi; p->npix = npix;
	p->ringstart_h.assign(ringstart, ringstart + nring);
	p->map_lo = *std::min_element(ringstart, ringstart + nring);
	p->map_hi = *std::max_element(ringstart, ringstart + nring) + npix;
	p->row_pitch = 0;
	if (nring > 1) {
		int64_t d = ringstart[1] - ringstart[0];
		bool ok = d != 0;
		for (int r = 1; r < nring && ok; r++) ok = (ringstart[r] - ringstart[r - 1] == d);
		if (ok) p->row_pitch = d;
	} else p->row_pitch = npix;
	if (p->geom.build(nring, theta)) return 1;
	if (nphi > 0 && p->fft.build(nphi, phi0, xdir, npix, nring, ringstart, weight, mmax)) return 1;      // nphi <= 0: ring groups follow
	if (p->leg.alloc((size_t)2*(mmax + 1)*p->geom.nring_pad)) return 1;
	B2_CHECK(cudaMemset(p->leg.p, 0, p->leg.bytes()));
	for (auto &e : p->ev) B2_CHECK(cudaEventCreate(&e));
	plan_stream_setup(p);
	return 0;
}

extern "C" int b2_sht_plan_rings(b2_sht_plan **out, int nring, const double *theta, int64_t nphi, double phi0,
	int xdir, int64_t npix_ring, const int64_t *ringstart, const double *weight,
	int lmax, int mmax, const int64_t *mstart, int64_t lstride)
{
	B2_REQUIRE(out, "plan: null output pointer");
	std::unique_ptr<b2_sht_plan> p(new b2_sht_plan());
	if (plan_common(p.get(), nring, theta, nphi, phi0, xdir, npix_ring, ringstart, weight, lmax, mmax, mstart, lstride)) return 1;
	*out = p.release();
	return 0;
}

extern "C" int b2_sht_plan_rings_general(b2_sht_plan **out, int nring, const double *theta, const int64_t *nphi, const double *phi0,
	const int64_t *ringstart, const double *weight, int lmax, int mmax, const int64_t *mstart, int64_t lstride)
{
	B2_REQUIRE(out && nphi && phi0, "plan: null argument");
	std::unique_ptr<b2_sht_plan> p(new b2_sht_plan());
	if (plan_common(p.get(), nring, theta, 0, 0.0, 1, 0, ringstart, weight, lmax, mmax, mstart, lstride)) return 1;
	std::map<int64_t, std::vector<int>> by_nphi;
	int64_t total = 0;
	p->npix_h.assign(nphi, nphi + nring);
	p->map_hi = 0;
	for (int r = 0; r < nring; r++) {
		B2_REQUIRE(nphi[r] >= 1, "plan: ring %d has nphi=%lld", r, (long long)nphi[r]);
		by_nphi[nphi[r]].push_back(r);
		total += nphi[r];
		p->map_hi = std::max(p->map_hi, ringstart[r] + nphi[r]);
	}
	p->dense_rings = (total == p->map_hi - p->map_lo);
	p->row_pitch = 0; p->nphi = 0; p->npix = 0;
	// the FFT tables of the distinct ring lengths (HEALPix nside 2048: 2048 lengths, about half of them Bluestein) are
	// host work in long double: computed by all host cores, uploaded afterwards
	{
		std::vector<std::unique_ptr<RingFft>> pre;
		for (auto &g : by_nphi) { pre.emplace_back(new RingFft()); (void)g; }
		std::vector<int64_t> lens; for (auto &g : by_nphi) lens.push_back(g.first);
		unsigned nth = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
		std::atomic<size_t> next(0); std::atomic<int> failed(0);
		std::vector<std::thread> pool;
		for (unsigned t = 0; t < nth && t < lens.size(); t++) pool.emplace_back([&]() {
			for (size_t i = next++; i < lens.size(); i = next++) if (pre[i]->prepare_tables(lens[i])) failed = 1;
		});
		for (auto &th : pool) th.join();
		B2_REQUIRE(!failed, "plan: FFT table construction failed");
		size_t i = 0;
		for (auto &g : by_nphi) {
			std::vector<double> ph(g.second.size());
			for (size_t k = 0; k < ph.size(); k++) ph[k] = phi0[g.second[k]];
			if (pre[i]->build_group(g.first, (int)g.second.size(), g.second.data(), ph.data(), nring, ringstart, weight, mmax)) return 1;
			p->groups.push_back(std::move(pre[i])); i++;
		}
	}
	if (p->pack.build(p->groups)) return 1;
	const int ns = (int)std::min<size_t>(16, p->groups.size());
	p->gstreams.assign(ns, nullptr); p->gjoin.assign(ns, nullptr);
	for (int i = 0; i < ns; i++) {
		B2_CHECK(cudaStreamCreateWithFlags(&p->gstreams[i], cudaStreamNonBlocking));
		B2_CHECK(cudaEventCreateWithFlags(&p->gjoin[i], cudaEventDisableTiming));
	}
	B2_CHECK(cudaEventCreateWithFlags(&p->gfork, cudaEventDisableTiming));
	*out = p.release();
	return 0;
}

// packed launches (RingPack) unless B2_NO_PACK=1 asks for one launch per group on the side streams
static bool use_pack(const b2_sht_plan *p)
{
	static const bool off = getenv("B2_NO_PACK") && atoi(getenv("B2_NO_PACK"));
	return !off && !p->pack.buckets.empty();
}

// ring FFTs of a general plan: fork the groups over the side streams, join back into st
template<typename F> static int run_groups(b2_sht_plan *p, cudaStream_t st, F launch)
{
	const int ns = (int)p->gstreams.size();
	B2_CHECK(cudaEventRecord(p->gfork, st));
	for (int i = 0; i < ns; i++) B2_CHECK(cudaStreamWaitEvent(p->gstreams[i], p->gfork, 0));
	for (size_t g = 0; g < p->groups.size(); g++) if (launch(*p->groups[g], p->gstreams[g % ns])) return 1;
	for (int i = 0; i < ns; i++) {
		B2_CHECK(cudaEventRecord(p->gjoin[i], p->gstreams[i]));
		B2_CHECK(cudaStreamWaitEvent(st, p->gjoin[i], 0));
	}
	return 0;
}

static int maxlmax_2d(const std::string &g, int ny)
{
	// pixell/curvedsky.py:1349-1353
	if (g == "CC") return ny - 2;
	if (g == "DH") return (ny - 2)/2;
	if (g == "F2") return (ny - 1)/2;
	return ny - 1;
}

extern "C" int b2_sht_plan_2d(b2_sht_plan **out, const char *geometry, int ntheta, int64_t nphi, double phi0,
	int flip_y, int flip_x, int lmax, int mmax, const int64_t *mstart, int64_t lstride)
{
	B2_REQUIRE(out && geometry, "plan: null argument");
	std::unique_ptr<b2_sht_plan> p(new b2_sht_plan());
	std::vector<double> theta;
	if (grid_theta_host(geometry, ntheta, theta)) return 1;
	std::vector<int64_t> rs(ntheta);
	for (int k = 0; k < ntheta; k++) rs[k] = (int64_t)(flip_y ? ntheta - 1 - k : k)*nphi;
	if (plan_common(p.get(), ntheta, theta.data(), nphi, phi0, flip_x ? -1 : 1, nphi, rs.data(), nullptr, lmax, mmax, mstart, lstride)) return 1;
	p->is2d = true; p->geometry = geometry; p->ntheta = ntheta;
	// exact analysis: either direct quadrature weights or the folded theta weighting (K5)
	if (lmax <= maxlmax_2d(p->geometry, ntheta)) {
		if (ThetaResampler::needed(p->geometry, ntheta, lmax)) {
			p->resamp.reset(new ThetaResampler());
			if (p->resamp->build(p->geometry, ntheta, nphi, lmax, mmax, p->geom.nring_pad)) return 1;
		} else {
			std::vector<double> w(ntheta);
			if (gridweights_device(geometry, ntheta, w.data(), nullptr)) return 1;
			for (auto &v : w) v /= (double)nphi;
			if (p->w2d.upload(w)) return 1;
		}
	}
	*out = p.release();
	return 0;
}

extern "C" void b2_sht_plan_destroy(b2_sht_plan *plan) { delete plan; }
extern "C" int64_t b2_sht_plan_bytes(const b2_sht_plan *plan) { return plan ? (int64_t)plan->bytes() : 0; }

// ------------------------------------------------------------------------------------ conversions

__global__ void k_c64_to_c128(const float2 *in, double2 *out, int64_t n)
{
	int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x;
	if (i < n) { float2 v = in[i]; out[i] = make_double2(v.x, v.y); }
}
__global__ void k_c128_to_c64(const double2 *in, float2 *out, int64_t n)
{
	int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x;
	if (i < n) { double2 v = in[i]; out[i] = make_float2((float)v.x, (float)v.y); }
}
__global__ void k_scale_rows(double2 *leg, const double *w, int nring, int64_t nring_pad, int64_t nrow)
{
	int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x;
	if (i >= nrow*nring_pad) return;
	int r = (int)(i % nring_pad);
	if (r < nring) { double2 v = leg[i]; double s = w[r]; leg[i] = make_double2(v.x*s, v.y*s); }
}

// ------------------------------------------------------------------------------------ execution

enum { OP_SYNTH, OP_ADJ_SYNTH, OP_ANALYSIS, OP_ADJ_ANALYSIS };

// B2_TRACE=1: stage timeline of every host-memory call on stderr (debugging aid; events are created per mark)
struct Trace {
	bool on = false; std::vector<std::pair<std::string, cudaEvent_t>> ev;
	Trace() { on = getenv("B2_TRACE") && atoi(getenv("B2_TRACE")); }
	void mark(const char *name, int g, cudaStream_t s) {
		if (!on) return;
		cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return;
		cudaEventRecord(e, s);
		char b[96]; snprintf(b, sizeof b, "g%d %s", g, name); ev.emplace_back(b, e);
	}
	void dump(const char *title) {
		if (!on || ev.empty()) return;
		cudaDeviceSynchronize();
		std::vector<std::pair<float, std::string>> rows;
		for (auto &x : ev) { float t = 0; cudaEventElapsedTime(&t, ev[0].second, x.second); rows.emplace_back(t, x.first); }
		for (auto &x : ev) cudaEventDestroy(x.second);
		std::sort(rows.begin(), rows.end());
		fprintf(stderr, "[b2 trace] %s\n", title);
		for (auto &r : rows) fprintf(stderr, "  %9.3f ms  %s\n", r.first, r.second.c_str());
		ev.clear();
	}
};
static Trace g_trace;
#define TR(name, s) g_trace.mark(name, G.gi, s)

struct Exec {
	b2_sht_plan *p; int op, spin, mode, dtype, mem; cudaStream_t st;      // st: compute stream (Legendre kernels)
	cudaStream_t sf;                                                      // stream of the ring FFT / theta stages (== st: no overlap)
	cudaStream_t s_in, s_out;                                             // copy streams (host-memory calls)
	int nca, ncm;              // alm / map components
	size_t asz, msz;           // bytes per complex alm element / real map element
};

// one spin group in flight: where its operands live on the device
struct GroupCtx {
	void *alm; int64_t alm_cs; void *map; int64_t map_cs;      // caller's operands
	double2 *dalm; int64_t dalm_cs; float2 *tmp32;             // complex128 alm on the device (+ complex64 scratch)
	void *dmap; int64_t dmap_cs; char *dmap_base;              // map on the device (element offsets as in the caller's array)
	bool alm_direct, streamed;      // streamed: the results already left for the host chunk by chunk
	int gi; int lane; double2 *leg; bool lane_reused; bool last, poll, gated; LegSignal gate;      // poll: alm ranges leave as the kernel's flags come in      // which of the plan's two leg buffers this group works in
	cudaEvent_t ev_in, ev_done;
};

static int copy_map(Exec &E, void *host, void *dev, bool to_dev, cudaStream_t st)
{
	b2_sht_plan *p = E.p;
	// host component pointer `host` addresses element 0; the rings occupy [map_lo, map_hi)
	char *h = (char*)host + p->map_lo*E.msz; char *d = (char*)dev;
	cudaMemcpyKind kind = to_dev ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
	if (p->groups.empty() ? (p->row_pitch == p->npix || p->row_pitch == -p->npix || p->nring == 1) : p->dense_rings) {
		if (to_dev) B2_CHECK(cudaMemcpyAsync(d, h, (p->map_hi - p->map_lo)*E.msz, kind, st));
		else        B2_CHECK(cudaMemcpyAsync(h, d, (p->map_hi - p->map_lo)*E.msz, kind, st));
	} else {
		// only the ring pixels move (rows with gaps between them)
		for (int r = 0; r < p->nring; r++) {
			size_t off = (p->ringstart_h[r] - p->map_lo)*E.msz;
			const int64_t np = p->groups.empty() ? p->npix : p->npix_h[r];
			if (to_dev) B2_CHECK(cudaMemcpyAsync(d + off, h + off, np*E.msz, kind, st));
			else        B2_CHECK(cudaMemcpyAsync(h + off, d + off, np*E.msz, kind, st));
		}
	}
	return 0;
}

static inline bool to_map_op(int op) { return op == OP_SYNTH || op == OP_ADJ_ANALYSIS; }
static size_t group_alm_bytes(const Exec &E) { return (size_t)E.nca*E.p->alm_span*16 + (E.dtype == B2_F32 ? (size_t)E.nca*E.p->alm_span*8 : 0); }
static size_t group_map_bytes(const Exec &E) { return (size_t)E.ncm*(size_t)(E.p->map_hi - E.p->map_lo)*E.msz; }

// phase 1: operands to the device (copy stream); salm / smap: this group's staging areas
static int group_stage_in(Exec &E, GroupCtx &G, char *salm, char *smap)
{
	b2_sht_plan *p = E.p;
	const bool to_map = to_map_op(E.op);
	G.alm_direct = (E.mem == B2_MEM_DEVICE && E.dtype == B2_F64);
	G.tmp32 = nullptr; G.streamed = false;
	if (G.alm_direct) { G.dalm = (double2*)G.alm; G.dalm_cs = G.alm_cs; }
	else {
		G.dalm = (double2*)salm; G.dalm_cs = p->alm_span;
		G.tmp32 = (float2*)(salm + (size_t)E.nca*p->alm_span*16);
		// the input direction needs the values; the output direction needs them too so that entries the
		// transform does not own survive the round trip
		// first group of a host-memory synthesis: nothing hides this copy, so it goes range by range of m, each followed by
		// its arrival flag, and the Legendre kernel's CTAs start as soon as their own range is there (B2_GATE_ALM=0: off)
		static const bool gate_alm = !(getenv("B2_GATE_ALM") && !atoi(getenv("B2_GATE_ALM")));
		if (gate_alm && to_map && G.gi == 0 && E.mem == B2_MEM_HOST && E.dtype == B2_F64 && E.s_in != E.st
			&& p->mcuts.size() >= 3 && (int)p->mcuts.size() <= LEG_MAXCUT) {
			if (!p->gate_src_h) {
				B2_CHECK(cudaHostAlloc((void**)&p->gate_src_h, LEG_MAXCUT*sizeof(int), cudaHostAllocDefault));
				if (p->gate_flag.alloc(16*LEG_MAXCUT)) return 1;
				B2_CHECK(cudaMemset(p->gate_flag.p, 0, p->gate_flag.bytes()));
			}
			LegSignal &sg = G.gate; sg.ncut = (int)p->mcuts.size(); sg.epoch = ++p->gate_epoch; sg.count = nullptr; sg.flag = p->gate_flag.p;
			for (int i = 0; i < sg.ncut; i++) sg.cut[i] = p->mcuts[i];
			for (int r = 0; r + 1 < sg.ncut; r++) {
				const int m_lo = p->mcuts[r], m_hi = p->mcuts[r + 1];
				const int64_t lo = p->mstart_h[m_lo] + m_lo, hi = p->mstart_h[m_hi - 1] + p->lmax + 1;
				for (int c = 0; c < E.nca; c++)
					B2_CHECK(cudaMemcpyAsync(G.dalm + (size_t)c*G.dalm_cs + lo, (char*)G.alm + ((size_t)c*G.alm_cs + lo)*16, (size_t)(hi - lo)*16, cudaMemcpyHostToDevice, E.s_in));
				p->gate_src_h[r] = sg.epoch;
				B2_CHECK(cudaMemcpyAsync((void*)sg.flag_of(r), &p->gate_src_h[r], sizeof(int), cudaMemcpyHostToDevice, E.s_in));
			}
			G.gated = true;
		}
		for (int c = 0; c < E.nca && !G.gated; c++) {
			if (!to_map && p->alm_dense) break;      // every entry of the span is overwritten
			char *src = (char*)G.alm + (size_t)c*G.alm_cs*E.asz;
			if (E.dtype == B2_F64) B2_CHECK(cudaMemcpyAsync(G.dalm + c*G.dalm_cs, src, p->alm_span*16, cudaMemcpyDefault, E.s_in));
			else if (E.mem == B2_MEM_HOST) B2_CHECK(cudaMemcpyAsync(G.tmp32 + c*p->alm_span, src, p->alm_span*8, cudaMemcpyHostToDevice, E.s_in));
		}
	}
	G.dmap = G.map; G.dmap_cs = G.map_cs; G.dmap_base = nullptr;
	if (E.mem == B2_MEM_HOST) {
		size_t span = (size_t)(p->map_hi - p->map_lo);
		G.dmap_cs = (int64_t)span; G.dmap_base = smap;
		G.dmap = smap - p->map_lo*E.msz;      // so that element offsets keep their meaning
		if (!to_map) for (int c = 0; c < E.ncm; c++)
			if (copy_map(E, (char*)G.map + (size_t)c*G.map_cs*E.msz, smap + (size_t)c*span*E.msz, true, E.s_in)) return 1;
	}
	if (E.s_in != E.st) B2_CHECK(cudaEventRecord(G.ev_in, E.s_in));
	TR("h2d done", E.s_in);
	return 0;
}

// phase 2: the transform.  Legendre kernels (and the alm precision conversions) on E.st, the HBM-bound stages (ring
// FFTs, theta weighting) on E.sf; E.sf == E.st runs everything in order on one stream.
static int hop(cudaStream_t from, cudaStream_t to, cudaEvent_t ev)
{
	if (from == to) return 0;
	B2_CHECK(cudaEventRecord(ev, from));
	B2_CHECK(cudaStreamWaitEvent(to, ev, 0));
	return 0;
}

static int group_compute(Exec &E, GroupCtx &G)
{
	b2_sht_plan *p = E.p;
	const bool to_map = to_map_op(E.op);
	const bool two = E.sf != E.st;
	LegTables *T = p->get_tables(E.spin);
	if (!T) return 1;
	AlmLayout L; L.lmax = p->lmax; L.mmax = p->mmax; L.mstart_d = p->mstart.p; L.lstride = p->lstride;
	const int deriv1 = E.mode == B2_MODE_DERIV1;
	double2 *leg = G.leg;
	B2_CHECK(cudaEventRecord(p->ev[0], E.st));
	TR("compute enqueued (stream position)", E.st);
	if (E.s_in != E.st) {
		if (!G.gated) B2_CHECK(cudaStreamWaitEvent(E.st, G.ev_in, 0));      // gated: the kernel's CTAs wait for their own range of alm
		if (two) B2_CHECK(cudaStreamWaitEvent(E.sf, G.ev_in, 0));
	}
	const LegSignal *gate = G.gated ? &G.gate : nullptr;
	// the group that used this leg buffer before must be through with it: its last reader ran on the other stream
	if (two && G.lane_reused) B2_CHECK(cudaStreamWaitEvent(to_map ? E.st : E.sf, p->ev_free[G.lane], 0));
	if (!G.alm_direct && E.dtype == B2_F32 && !(!to_map && p->alm_dense)) {
		for (int c = 0; c < E.nca; c++) {
			const float2 *s32 = E.mem == B2_MEM_HOST ? G.tmp32 + c*p->alm_span : (const float2*)((char*)G.alm + (size_t)c*G.alm_cs*E.asz);
			k_c64_to_c128<<<(unsigned)((p->alm_span + 255)/256), 256, 0, E.st>>>(s32, G.dalm + c*G.dalm_cs, p->alm_span);
			B2_LAUNCH_CHECK();
		}
	}
	B2_CHECK(cudaEventRecord(p->ev[1], to_map ? E.st : E.sf));
	const bool host64 = E.mem == B2_MEM_HOST && E.dtype == B2_F64 && E.s_out != E.st;
	if (to_map && E.op == OP_SYNTH && host64 && !p->schunks.empty() && p->groups.empty()) {
		// chunks of ring pairs, pole -> equator: Legendre synthesis and ring FFTs of a chunk, then its rows go to the host
		// on the copy stream while the next chunk is computed; only the last chunk's copy is exposed
		const size_t span = (size_t)(p->map_hi - p->map_lo);
		B2_CHECK(cudaEventRecord(p->ev[5], E.st));
		for (size_t c = 0; c < p->schunks.size(); c++) {
			const b2_sht_plan::StreamChunk &C = p->schunks[c];
			if (leg_alm2leg(*T, p->geom, L, deriv1, G.dalm, G.dalm_cs, leg, E.st, p->get_start(E.spin), C.pair_lo, C.pair_hi, gate)) return 1;
			if (c + 1 == p->schunks.size()) B2_CHECK(cudaEventRecord(p->ev[2], E.st));
			TR("K1 chunk done", E.st);
			if (two) {
				if (!p->ev_chunk[c]) B2_CHECK(cudaEventCreateWithFlags(&p->ev_chunk[c], cudaEventDisableTiming));
				if (hop(E.st, E.sf, p->ev_chunk[c])) return 1;
			}
			for (int r = 0; r < C.nrun; r++)
				if (ring_leg2map(p->fft, E.ncm, leg, p->geom.nring_pad, G.dmap, G.dmap_cs, E.dtype, E.sf, C.r0[r], C.nr[r])) return 1;
			if (!p->sev[c]) B2_CHECK(cudaEventCreateWithFlags(&p->sev[c], cudaEventDisableTiming));
			B2_CHECK(cudaEventRecord(p->sev[c], E.sf));
			TR("K3 chunk done", E.sf);
			B2_CHECK(cudaStreamWaitEvent(E.s_out, p->sev[c], 0));
			for (int r = 0; r < C.nrun; r++) {
				const int64_t a = p->ringstart_h[C.r0[r]], b = p->ringstart_h[C.r0[r] + C.nr[r] - 1];
				const int64_t lo = std::min(a, b), nel = std::max(a, b) + p->npix - lo;
				for (int k = 0; k < E.ncm; k++)
					B2_CHECK(cudaMemcpyAsync((char*)G.map + ((size_t)k*G.map_cs + lo)*E.msz, G.dmap_base + ((size_t)k*span + (lo - p->map_lo))*E.msz,
						(size_t)nel*E.msz, cudaMemcpyDeviceToHost, E.s_out));
			}
			TR("d2h chunk done", E.s_out);
		}
		B2_CHECK(cudaEventRecord(p->ev[3], E.sf));
		B2_CHECK(cudaEventRecord(p->ev[4], E.sf));
		if (two) B2_CHECK(cudaEventRecord(p->ev_free[G.lane], E.sf));
		G.streamed = true;
	} else if (to_map) {
		B2_CHECK(cudaEventRecord(p->ev[5], E.st));
		if (leg_alm2leg(*T, p->geom, L, deriv1, G.dalm, G.dalm_cs, leg, E.st, p->get_start(E.spin), 0, 0, gate)) return 1;
		B2_CHECK(cudaEventRecord(p->ev[2], E.st));
		TR("K1 done", E.st);
		if (hop(E.st, E.sf, p->ev_ready[G.lane])) return 1;
		if (E.op == OP_ADJ_ANALYSIS) {
			if (p->resamp) { if (p->resamp->apply(leg, E.ncm, E.spin, E.sf, true)) return 1; }
			else {
				int64_t nrow = (int64_t)E.ncm*(p->mmax + 1);
				k_scale_rows<<<(unsigned)((nrow*p->geom.nring_pad + 255)/256), 256, 0, E.sf>>>(leg, p->w2d.p, p->nring, p->geom.nring_pad, nrow);
				B2_LAUNCH_CHECK();
			}
		}
		B2_CHECK(cudaEventRecord(p->ev[3], E.sf));
		if (p->groups.empty()) { if (ring_leg2map(p->fft, E.ncm, leg, p->geom.nring_pad, G.dmap, G.dmap_cs, E.dtype, E.sf)) return 1; }
		else if (use_pack(p)) { if (ring_leg2map_pack(p->pack, *p->groups[0], E.ncm, leg, p->geom.nring_pad, G.dmap, G.dmap_cs, E.dtype, E.sf)) return 1; }
		else if (run_groups(p, E.sf, [&](const RingFft &g, cudaStream_t s) { return ring_leg2map(g, E.ncm, leg, p->geom.nring_pad, G.dmap, G.dmap_cs, E.dtype, s); })) return 1;
		B2_CHECK(cudaEventRecord(p->ev[4], E.sf));
		TR("K3 done", E.sf);
		if (two) B2_CHECK(cudaEventRecord(p->ev_free[G.lane], E.sf));
	} else {
		if (p->groups.empty()) { if (ring_map2leg(p->fft, E.ncm, leg, p->geom.nring_pad, G.dmap, G.dmap_cs, E.dtype, E.op == OP_ADJ_SYNTH, E.sf)) return 1; }
		else if (use_pack(p)) { if (ring_map2leg_pack(p->pack, *p->groups[0], E.ncm, leg, p->geom.nring_pad, G.dmap, G.dmap_cs, E.dtype, E.op == OP_ADJ_SYNTH, E.sf)) return 1; }
		else if (run_groups(p, E.sf, [&](const RingFft &g, cudaStream_t s) { return ring_map2leg(g, E.ncm, leg, p->geom.nring_pad, G.dmap, G.dmap_cs, E.dtype, E.op == OP_ADJ_SYNTH, s); })) return 1;
		B2_CHECK(cudaEventRecord(p->ev[2], E.sf));
		TR("K4 done", E.sf);
		if (E.op == OP_ANALYSIS) {
			if (p->resamp) { if (p->resamp->apply(leg, E.ncm, E.spin, E.sf)) return 1; }
			else {
				int64_t nrow = (int64_t)E.ncm*(p->mmax + 1);
				k_scale_rows<<<(unsigned)((nrow*p->geom.nring_pad + 255)/256), 256, 0, E.sf>>>(leg, p->w2d.p, p->nring, p->geom.nring_pad, nrow);
				B2_LAUNCH_CHECK();
			}
		}
		B2_CHECK(cudaEventRecord(p->ev[3], E.sf));
		TR("K5 done", E.sf);
		if (hop(E.sf, E.st, p->ev_ready[G.lane])) return 1;
		B2_CHECK(cudaEventRecord(p->ev[5], E.st));
		// Last group of a host-memory call: nothing is left to hide its alm copy behind, so the kernel reports every range
		// of m it completes (LegSignal) and execute_groups sends that range to the host while the kernel works on the rest.
		// (Separate launches per range were measured too: every launch ends in a tail of single-warp CTAs that costs more
		// than the copy it hides, C3 map2alm 212 -> 245 ms.)  B2_STREAM_ALM=0 disables it.
		static const bool stream_alm = !(getenv("B2_STREAM_ALM") && !atoi(getenv("B2_STREAM_ALM")));
		if (stream_alm && host64 && !G.alm_direct && G.last && p->mcuts.size() >= 3 && (int)p->mcuts.size() <= LEG_MAXCUT) {
			if (!p->sig_flag_h) {
				B2_CHECK(cudaHostAlloc((void**)&p->sig_flag_h, 16*LEG_MAXCUT*sizeof(int), cudaHostAllocMapped));
				memset(p->sig_flag_h, 0, 16*LEG_MAXCUT*sizeof(int));
				B2_CHECK(cudaHostGetDevicePointer((void**)&p->sig_flag_d, p->sig_flag_h, 0));
				if (p->sig_count.alloc(LEG_MAXCUT)) return 1;
			}
			LegSignal sg; sg.ncut = (int)p->mcuts.size(); sg.epoch = ++p->sig_epoch;
			for (int i = 0; i < sg.ncut; i++) sg.cut[i] = p->mcuts[i];
			sg.count = p->sig_count.p; sg.flag = p->sig_flag_d;
			B2_CHECK(cudaMemsetAsync(p->sig_count.p, 0, LEG_MAXCUT*sizeof(unsigned), E.st));
			if (leg_leg2alm(*T, p->geom, L, deriv1, G.dalm, G.dalm_cs, leg, E.st, p->get_start(E.spin), 0, 0, &sg)) return 1;
			G.streamed = true; G.poll = true;
		} else if (leg_leg2alm(*T, p->geom, L, deriv1, G.dalm, G.dalm_cs, leg, E.st, p->get_start(E.spin))) return 1;
		B2_CHECK(cudaEventRecord(p->ev[4], E.st));
		TR("K2 done", E.st);
		if (two) B2_CHECK(cudaEventRecord(p->ev_free[G.lane], E.st));
		if (!G.alm_direct && E.dtype == B2_F32) {
			for (int c = 0; c < E.nca; c++) {
				float2 *d32 = E.mem == B2_MEM_HOST ? G.tmp32 + c*p->alm_span : (float2*)((char*)G.alm + (size_t)c*G.alm_cs*E.asz);
				k_c128_to_c64<<<(unsigned)((p->alm_span + 255)/256), 256, 0, E.st>>>(G.dalm + c*G.dalm_cs, d32, p->alm_span);
				B2_LAUNCH_CHECK();
			}
		}
	}
	p->timing[0] = to_map ? 1 : -1;
	if (E.s_out != E.st) B2_CHECK(cudaEventRecord(G.ev_done, to_map ? E.sf : E.st));
	return 0;
}

// phase 3: results back to the caller (copy stream)
static int group_stage_out(Exec &E, GroupCtx &G)
{
	b2_sht_plan *p = E.p;
	const bool to_map = to_map_op(E.op);
	if (G.streamed) return 0;
	if (E.s_out != E.st) B2_CHECK(cudaStreamWaitEvent(E.s_out, G.ev_done, 0));
	struct Done { GroupCtx &G; cudaStream_t s; ~Done() { TR("d2h done", s); } } done_mark{G, E.s_out};
	if (to_map) {
		if (E.mem == B2_MEM_HOST) {
			size_t span = (size_t)(p->map_hi - p->map_lo);
			for (int c = 0; c < E.ncm; c++)
				if (copy_map(E, (char*)G.map + (size_t)c*G.map_cs*E.msz, G.dmap_base + (size_t)c*span*E.msz, false, E.s_out)) return 1;
		}
	} else if (!G.alm_direct) {
		for (int c = 0; c < E.nca; c++) {
			char *dst = (char*)G.alm + (size_t)c*G.alm_cs*E.asz;
			if (E.dtype == B2_F64) B2_CHECK(cudaMemcpyAsync(dst, G.dalm + c*G.dalm_cs, p->alm_span*16, cudaMemcpyDefault, E.s_out));
			else if (E.mem == B2_MEM_HOST) B2_CHECK(cudaMemcpyAsync(dst, G.tmp32 + c*p->alm_span, p->alm_span*8, cudaMemcpyDeviceToHost, E.s_out));
		}
	}
	return 0;
}

static int check_exec_args(b2_sht_plan *plan, int op, int spin, int mode, int dtype, int mem)
{
	B2_REQUIRE(spin >= 0 && spin <= 32, "transform: spin %d out of range", spin);
	B2_REQUIRE(dtype == B2_F64 || dtype == B2_F32, "transform: bad dtype");
	B2_REQUIRE(mem == B2_MEM_HOST || mem == B2_MEM_DEVICE, "transform: bad memory kind");
	B2_REQUIRE(mode == B2_MODE_STANDARD || (mode == B2_MODE_DERIV1 && spin == 1), "transform: DERIV1 needs spin=1");
	if (op == OP_ANALYSIS || op == OP_ADJ_ANALYSIS) {
		B2_REQUIRE(plan->is2d, "analysis_2d needs a plan made by b2_sht_plan_2d");
		B2_REQUIRE(plan->resamp || plan->w2d.n, "lmax=%d too large for geometry %s with %d rings", plan->lmax, plan->geometry.c_str(), plan->ntheta);
	}
	return 0;
}

// Host side of LegSignal: as soon as the kernel has published a range of m, its coefficients go to the host.
static int poll_and_copy(Exec &E, GroupCtx &G)
{
	b2_sht_plan *p = E.p;
	const int nr = (int)p->mcuts.size() - 1;
	bool drained = false;
	static const int dbg = getenv("B2_SIG_DEBUG") ? atoi(getenv("B2_SIG_DEBUG")) : 0;      // 1: no polling, 2: poll, copy at the end
	if (dbg == 1) { B2_CHECK(cudaStreamSynchronize(E.st)); drained = true; }
	if (dbg == 2) { for (int r = 0; r < nr; r++) { volatile int *f = p->sig_flag_h + 16*r; while (*f != p->sig_epoch) {} } drained = true; }
	for (int r = 0; r < nr; r++) {
		volatile int *f = p->sig_flag_h + 16*r;
		for (unsigned spin = 0; !drained && *f != p->sig_epoch; spin++) {
			// the stream running dry (kernel finished, or failed) also ends the wait: stream order then covers the copy
			if ((spin & 255) == 255 && cudaStreamQuery(E.st) != cudaErrorNotReady) { drained = true; B2_CHECK(cudaStreamSynchronize(E.st)); }
		}
		const int m_lo = p->mcuts[r], m_hi = p->mcuts[r + 1];
		const int64_t lo = p->mstart_h[m_lo] + m_lo, hi = p->mstart_h[m_hi - 1] + p->lmax + 1;
		for (int k = 0; k < E.nca; k++)
			B2_CHECK(cudaMemcpyAsync((char*)G.alm + ((size_t)k*G.alm_cs + lo)*16, G.dalm + (size_t)k*G.dalm_cs + lo, (size_t)(hi - lo)*16, cudaMemcpyDeviceToHost, E.s_out));
		g_trace.mark("alm range copy queued (s_out position)", G.gi, E.s_out);
	}
	return 0;
}

// Runs a list of spin groups.  Device memory: everything on the caller's stream, in order.  Host memory: three
// streams -- the H2D copies of group g+1 and the D2H copies of group g-1 overlap the kernels of group g.
static int execute_groups(b2_sht_plan *plan, int op, int ngroups, const int *spins, int mode, int dtype,
	void *const *alms, const int64_t *alm_cs, void *const *maps, const int64_t *map_cs, int mem, void *stream)
{
	B2_REQUIRE(plan && ngroups >= 1 && ngroups <= B2_MAX_GROUPS, "transform: bad group count");
	Exec E; E.p = plan; E.op = op; E.mode = mode; E.dtype = dtype; E.mem = mem; E.st = (cudaStream_t)stream;
	E.asz = dtype == B2_F64 ? 16 : 8; E.msz = dtype == B2_F64 ? 8 : 4;
	E.s_in = E.s_out = E.st;
	const bool pipelined = (mem == B2_MEM_HOST);
	if (pipelined) {
		if (!plan->s_in) { B2_CHECK(cudaStreamCreateWithFlags(&plan->s_in, cudaStreamNonBlocking)); B2_CHECK(cudaStreamCreateWithFlags(&plan->s_out, cudaStreamNonBlocking)); }
		if (!plan->s_comp) B2_CHECK(cudaStreamCreateWithFlags(&plan->s_comp, cudaStreamNonBlocking));
		E.s_in = plan->s_in; E.s_out = plan->s_out;
		if (!E.st) E.st = plan->s_comp;      // never the legacy stream: it would serialise with the copy streams' neighbours
	}
	// one call at a time on this plan's scratch (see b2_sht_plan::ev_last)
	if (!plan->ev_last) B2_CHECK(cudaEventCreateWithFlags(&plan->ev_last, cudaEventDisableTiming));
	else B2_CHECK(cudaStreamWaitEvent(E.st, plan->ev_last, 0));
	// several groups: two-stage pipeline (see b2_sht_plan::s_fft)
	// Off by default (B2_OVERLAP=1 enables it): measured on B200 at C3 it gains nothing -- the Legendre grids keep every SM's
	// register file full, a 512-thread FFT CTA only fits once six of their CTAs have left one SM, and what does get in
	// displaces FP64 work one for one (map2alm 174.6 ms in order, 177.4 ms pipelined).
	static const bool overlap = getenv("B2_OVERLAP") && atoi(getenv("B2_OVERLAP"));
	E.sf = E.st;
	if (ngroups > 1 && overlap) {
		if (!plan->s_fft) {
			int lo = 0, hi = 0;
			B2_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
			B2_CHECK(cudaStreamCreateWithPriority(&plan->s_fft, cudaStreamNonBlocking, hi));
			B2_CHECK(cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming));
			B2_CHECK(cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming));
			for (int i = 0; i < 2; i++) {
				B2_CHECK(cudaEventCreateWithFlags(&plan->ev_ready[i], cudaEventDisableTiming));
				B2_CHECK(cudaEventCreateWithFlags(&plan->ev_free[i], cudaEventDisableTiming));
			}
		}
		if (!plan->leg2.n) {
			if (plan->leg2.alloc(plan->leg.n)) return 1;
			B2_CHECK(cudaMemsetAsync(plan->leg2.p, 0, plan->leg2.bytes(), E.st));
		}
		E.sf = plan->s_fft;
		B2_CHECK(cudaEventRecord(plan->ev_fork, E.st));
		B2_CHECK(cudaStreamWaitEvent(E.sf, plan->ev_fork, 0));
	}
	GroupCtx G[B2_MAX_GROUPS];
	size_t off_a[B2_MAX_GROUPS + 1] = {0}, off_m[B2_MAX_GROUPS + 1] = {0};
	for (int g = 0; g < ngroups; g++) {
		B2_REQUIRE(alms[g] && maps[g], "transform: null argument");
		if (check_exec_args(plan, op, spins[g], mode, dtype, mem)) return 1;
		E.spin = spins[g]; E.ncm = E.spin == 0 ? 1 : 2; E.nca = (E.spin == 0 || mode == B2_MODE_DERIV1) ? 1 : 2;
		const bool direct = (mem == B2_MEM_DEVICE && dtype == B2_F64);
		off_a[g + 1] = off_a[g] + (direct ? 0 : b2_round_up((int64_t)group_alm_bytes(E), 256));
		off_m[g + 1] = off_m[g] + (mem == B2_MEM_HOST ? b2_round_up((int64_t)group_map_bytes(E), 256) : 0);
	}
	if (plan->stage_alm.n < off_a[ngroups] && plan->stage_alm.alloc(off_a[ngroups])) return 1;
	if (plan->stage_map.n < off_m[ngroups] && plan->stage_map.alloc(off_m[ngroups])) return 1;
	if (pipelined) for (int g = 0; g < ngroups; g++) {
		if (!plan->gev[2*g]) { B2_CHECK(cudaEventCreateWithFlags(&plan->gev[2*g], cudaEventDisableTiming)); B2_CHECK(cudaEventCreateWithFlags(&plan->gev[2*g + 1], cudaEventDisableTiming)); }
	}
	auto setup = [&](int g) {
		E.spin = spins[g]; E.ncm = E.spin == 0 ? 1 : 2; E.nca = (E.spin == 0 || mode == B2_MODE_DERIV1) ? 1 : 2;
		G[g].alm = alms[g]; G[g].alm_cs = alm_cs[g]; G[g].map = maps[g]; G[g].map_cs = map_cs[g];
		G[g].ev_in = plan->gev[2*g]; G[g].ev_done = plan->gev[2*g + 1];
		G[g].gi = g; G[g].last = (g == ngroups - 1); G[g].lane = (E.sf != E.st) ? (g & 1) : 0; G[g].leg = G[g].lane ? plan->leg2.p : plan->leg.p; G[g].lane_reused = g >= 2;
	};
	for (int g = 0; g < ngroups; g++) { G[g].poll = false; G[g].gated = false; }
	g_trace.mark("start", -1, E.s_in);
	// all copies in are queued first (they run back to back on the copy stream), then each group's kernels and copies out
	for (int g = 0; g < ngroups; g++) { setup(g); if (group_stage_in(E, G[g], plan->stage_alm.p + off_a[g], plan->stage_map.p + off_m[g])) return 1; }
	for (int g = 0; g < ngroups; g++) {
		setup(g);
		if (group_compute(E, G[g])) return 1;
		if (group_stage_out(E, G[g])) return 1;
	}
	for (int g = 0; g < ngroups; g++) if (G[g].poll) {
		E.spin = spins[g]; E.ncm = E.spin == 0 ? 1 : 2; E.nca = (E.spin == 0 || mode == B2_MODE_DERIV1) ? 1 : 2;
		if (poll_and_copy(E, G[g])) return 1;
	}
	if (E.sf != E.st) {
		B2_CHECK(cudaEventRecord(plan->ev_join, E.sf));
		B2_CHECK(cudaStreamWaitEvent(E.st, plan->ev_join, 0));
	}
	B2_CHECK(cudaEventRecord(plan->ev_last, E.st));
	if (mem == B2_MEM_HOST) { B2_CHECK(cudaStreamSynchronize(E.s_out)); B2_CHECK(cudaStreamSynchronize(E.st)); }
	if (mem == B2_MEM_HOST) g_trace.dump(to_map_op(op) ? "alm -> map" : "map -> alm");
	return 0;
}

// Device-resident float64 synthesis of a batch of alm sets of one spin (Monte-Carlo realisations): blocks of
// leg_batch_size(spin) members go through the batched Legendre kernels, which pay the recurrence once per block; bit-identical
// to member-by-member calls.  done = the number of members finished (the caller runs the rest one by one).
static int execute_synth_batch(b2_sht_plan *p, int spin, int nbatch, const double2 *alm, int64_t acs, int64_t abs_,
	double *map, int64_t mcs, int64_t mbs, cudaStream_t st, int &done)
{
	static const bool off = getenv("B2_BATCH") && !atoi(getenv("B2_BATCH"));
	done = 0;
	if (off || nbatch < 2 || !p->groups.empty() || p->nphi <= 0) return 0;
	LegTables *T = p->get_tables(spin);
	if (!T) return 1;
	const int ncm = spin == 0 ? 1 : 2;
	const int64_t plane = (int64_t)(p->mmax + 1)*p->geom.nring_pad;
	if (!p->legb.n && p->legb.alloc((size_t)4*plane)) return 1;
	if (!p->ev_last) B2_CHECK(cudaEventCreateWithFlags(&p->ev_last, cudaEventDisableTiming));
	else B2_CHECK(cudaStreamWaitEvent(st, p->ev_last, 0));
	AlmLayout L; L.lmax = p->lmax; L.mmax = p->mmax; L.mstart_d = p->mstart.p; L.lstride = p->lstride;
	while (nbatch - done >= 2) {
		const int nb = (spin == 0 && nbatch - done >= 4) ? 4 : 2;
		if (leg_alm2leg_batch(*T, p->geom, L, nb, alm + (int64_t)done*abs_, acs, abs_, p->legb.p, ncm*plane, st, p->get_start(spin))) return 1;
		for (int b = 0; b < nb; b++)
			if (ring_leg2map(p->fft, ncm, p->legb.p + (int64_t)b*ncm*plane, p->geom.nring_pad, map + (int64_t)(done + b)*mbs, mcs, B2_F64, st)) return 1;
		done += nb;
	}
	B2_CHECK(cudaEventRecord(p->ev_last, st));
	return 0;
}

static int execute(b2_sht_plan *plan, int op, int spin, int mode, int dtype, int nbatch,
	void *alm, int64_t alm_cstride, int64_t alm_bstride, void *map, int64_t map_cstride, int64_t map_bstride,
	int mem, void *stream)
{
	B2_REQUIRE(plan && alm && map, "transform: null argument");
	B2_REQUIRE(nbatch >= 1, "transform: nbatch must be >= 1");
	if (check_exec_args(plan, op, spin, mode, dtype, mem)) return 1;
	const size_t asz = dtype == B2_F64 ? 16 : 8, msz = dtype == B2_F64 ? 8 : 4;
	if (op == OP_SYNTH && mem == B2_MEM_DEVICE && dtype == B2_F64 && mode == B2_MODE_STANDARD && nbatch >= 2) {
		int done = 0;
		if (execute_synth_batch(plan, spin, nbatch, (const double2*)alm, alm_cstride, alm_bstride, (double*)map, map_cstride, map_bstride, (cudaStream_t)stream, done)) return 1;
		alm = (char*)alm + (size_t)done*alm_bstride*asz; map = (char*)map + (size_t)done*map_bstride*msz; nbatch -= done;
		if (nbatch == 0) return 0;
	}
	// batches run as groups of the same spin, B2_MAX_GROUPS at a time
	for (int b0 = 0; b0 < nbatch; b0 += B2_MAX_GROUPS) {
		int ng = std::min(B2_MAX_GROUPS, nbatch - b0);
		int spins[B2_MAX_GROUPS]; void *alms[B2_MAX_GROUPS], *maps[B2_MAX_GROUPS]; int64_t acs[B2_MAX_GROUPS], mcs[B2_MAX_GROUPS];
		for (int g = 0; g < ng; g++) {
			spins[g] = spin; acs[g] = alm_cstride; mcs[g] = map_cstride;
			alms[g] = (char*)alm + (size_t)(b0 + g)*alm_bstride*asz; maps[g] = (char*)map + (size_t)(b0 + g)*map_bstride*msz;
		}
		if (execute_groups(plan, op, ng, spins, mode, dtype, alms, acs, maps, mcs, mem, stream)) return 1;
	}
	return 0;
}

extern "C" int b2_sht_execute_groups(b2_sht_plan *plan, int op, int ngroups, const int *spins, int mode, int dtype,
	void *const *alm, const int64_t *alm_cstride, void *const *map, const int64_t *map_cstride, int mem, void *stream)
{
	B2_REQUIRE(plan && spins && alm && alm_cstride && map && map_cstride, "execute_groups: null argument");
	B2_REQUIRE(op >= OP_SYNTH && op <= OP_ADJ_ANALYSIS, "execute_groups: bad operation");
	return execute_groups(plan, op, ngroups, spins, mode, dtype, alm, alm_cstride, map, map_cstride, mem, stream);
}

extern "C" int b2_synthesis(b2_sht_plan *plan, int spin, int mode, int dtype, int nbatch,
	const void *alm, int64_t acs, int64_t abs_, void *map, int64_t mcs, int64_t mbs, int mem, void *stream)
{ return execute(plan, OP_SYNTH, spin, mode, dtype, nbatch, (void*)alm, acs, abs_, map, mcs, mbs, mem, stream); }

extern "C" int b2_adjoint_synthesis(b2_sht_plan *plan, int spin, int mode, int dtype, int nbatch,
	void *alm, int64_t acs, int64_t abs_, const void *map, int64_t mcs, int64_t mbs, int mem, void *stream)
{ return execute(plan, OP_ADJ_SYNTH, spin, mode, dtype, nbatch, alm, acs, abs_, (void*)map, mcs, mbs, mem, stream); }

extern "C" int b2_analysis_2d(b2_sht_plan *plan, int spin, int dtype, int nbatch,
	void *alm, int64_t acs, int64_t abs_, const void *map, int64_t mcs, int64_t mbs, int mem, void *stream)
{ return execute(plan, OP_ANALYSIS, spin, B2_MODE_STANDARD, dtype, nbatch, alm, acs, abs_, (void*)map, mcs, mbs, mem, stream); }

extern "C" int b2_adjoint_analysis_2d(b2_sht_plan *plan, int spin, int dtype, int nbatch,
	const void *alm, int64_t acs, int64_t abs_, void *map, int64_t mcs, int64_t mbs, int mem, void *stream)
{ return execute(plan, OP_ADJ_ANALYSIS, spin, B2_MODE_STANDARD, dtype, nbatch, (void*)alm, acs, abs_, map, mcs, mbs, mem, stream); }

extern "C" int b2_sht_last_timing(b2_sht_plan *p, double out[4])
{
	B2_REQUIRE(p && out, "timing: null argument");
	B2_CHECK(cudaEventSynchronize(p->ev[4]));
	B2_CHECK(cudaEventSynchronize(p->ev[2])); B2_CHECK(cudaEventSynchronize(p->ev[3]));
	float t01, tleg, tfft, t23;
	const bool to_map = p->timing[0] >= 0;
	B2_CHECK(cudaEventElapsedTime(&t01, p->ev[0], p->ev[1]));
	// ev5 sits right before the Legendre kernel on its stream: ev5..ev2 (alm -> map) or ev5..ev4 (map -> alm); the ring
	// FFTs are ev3..ev4 resp. ev1..ev2, the theta stage ev2..ev3 (with several groups in flight that one includes waiting)
	B2_CHECK(cudaEventElapsedTime(&tleg, p->ev[5], to_map ? p->ev[2] : p->ev[4]));
	B2_CHECK(cudaEventElapsedTime(&tfft, to_map ? p->ev[3] : p->ev[1], to_map ? p->ev[4] : p->ev[2]));
	B2_CHECK(cudaEventElapsedTime(&t23, p->ev[2], p->ev[3]));
	out[0] = tleg; out[1] = tfft; out[2] = t23; out[3] = t01;
	return 0;
}

extern "C" int b2_alm2leg(b2_sht_plan *p, int spin, int mode, const void *alm_dev, int64_t acs, void *leg_dev, void *stream)
{
	B2_REQUIRE(p && alm_dev && leg_dev, "alm2leg: null argument");
	LegTables *T = p->get_tables(spin); if (!T) return 1;
	AlmLayout L; L.lmax = p->lmax; L.mmax = p->mmax; L.mstart_d = p->mstart.p; L.lstride = p->lstride;
	return leg_alm2leg(*T, p->geom, L, mode == B2_MODE_DERIV1, (const double2*)alm_dev, acs, (double2*)leg_dev, (cudaStream_t)stream, p->get_start(spin));
}

extern "C" int b2_theta_weighting(b2_sht_plan *p, int spin, int ncomp, int adjoint, void *leg_dev, void *stream)
{
	B2_REQUIRE(p && leg_dev, "theta_weighting: null argument");
	B2_REQUIRE(p->resamp, "theta_weighting: this plan integrates with plain ring weights");
	return p->resamp->apply((double2*)leg_dev, ncomp, spin, (cudaStream_t)stream, adjoint != 0);
}

extern "C" int b2_leg2alm(b2_sht_plan *p, int spin, int mode, void *alm_dev, int64_t acs, const void *leg_dev, void *stream)
{
	B2_REQUIRE(p && alm_dev && leg_dev, "leg2alm: null argument");
	LegTables *T = p->get_tables(spin); if (!T) return 1;
	AlmLayout L; L.lmax = p->lmax; L.mmax = p->mmax; L.mstart_d = p->mstart.p; L.lstride = p->lstride;
	return leg_leg2alm(*T, p->geom, L, mode == B2_MODE_DERIV1, (double2*)alm_dev, acs, (const double2*)leg_dev, (cudaStream_t)stream, p->get_start(spin));
}
