// ringfft.cu -- K3 / K4: the per-ring FFT stage between leg[comp][m][ring] and the map rows.
// Replaces the ring-FFT half of ducc0's synthesis / adjoint_synthesis (pixell/curvedsky.py:907-960)
// including what pixell does around it on the host: the phi0 phase, x flips (as a conjugation),
// cut rows (npix < nphi) and quadrature weights (curvedsky.py:852-868) are all fused here, so the
// map is touched exactly once and no flipped/padded copy (map2buffer, :1384-1411) is ever made.
//
// One CTA per (ring, component); the ring lives in shared memory as a packed half-length complex
// sequence (even nphi) or a full complex one (odd nphi); fft_smem.cuh does the passes.
#include "ringfft.cuh"
#include <map>
#include <algorithm>
#include <complex>
#include <type_traits>

// ------------------------------------------------------------------------------------ tables

bool FftTables::smooth(int64_t n)
{
	if (n < 1) return false;
	for (int p = 2; p <= FFT_MAX_RADIX && n > 1; p++) while (n % p == 0) n /= p;
	return n == 1;
}

int FftTables::bluestein_len(int n)
{
	for (int M = 2*n - 1; ; M++) {
		int r = M;
		while (r % 2 == 0) r /= 2;
		while (r % 3 == 0) r /= 3;
		while (r % 5 == 0) r /= 5;
		if (r == 1) return M;
	}
}

typedef std::complex<long double> cld;
static const long double TAU = 6.283185307179586476925286766559005768L;

// plain recursive mixed-radix FFT in long double (host, plan time only); roots[k] = exp(-2 pi i k/N) for the top length N
static void host_fft(std::vector<cld> &x, const std::vector<cld> &roots)
{
	size_t n = x.size();
	if (n <= 1) return;
	size_t r = n;
	for (size_t p = 2; p*p <= n; p++) if (n % p == 0) { r = p; break; }
	size_t m = n/r, step = roots.size()/n;
	std::vector<std::vector<cld>> sub(r, std::vector<cld>(m));
	for (size_t q = 0; q < r; q++) { for (size_t j = 0; j < m; j++) sub[q][j] = x[j*r + q]; host_fft(sub[q], roots); }
	for (size_t k = 0; k < n; k++) {
		cld acc = 0;
		for (size_t q = 0; q < r; q++) acc += sub[q][k % m]*roots[((q*k) % n)*step];
		x[k] = acc;
	}
}

static void factorize(int n, int *fac, int &nfac)
{
	nfac = 0;
	int rem = n;
	while (rem % 4 == 0) { fac[nfac++] = 4; rem /= 4; }
	for (int p = 2; p <= FFT_MAX_RADIX; p++) while (rem % p == 0) { fac[nfac++] = p; rem /= p; }
}

bool FftTables::fast_ok(int64_t n)
{
	if (n < 2) return false;
	while (n % 2 == 0) n /= 2;
	while (n % 3 == 0) n /= 3;
	while (n % 5 == 0) n /= 5;
	return n == 1;
}

// fast path: odd radices first (long strides), then a leftover 2 or 4, then 16s, then 8s; the last radix sets the padding
static void factorize_fast(int n, int *fac, int &nfac)
{
	nfac = 0;
	int rem = n, a = 0;
	while (rem % 5 == 0) { fac[nfac++] = 5; rem /= 5; }
	while (rem % 3 == 0) { fac[nfac++] = 3; rem /= 3; }
	while (rem % 2 == 0) { a++; rem /= 2; }
	if (a == 1) fac[nfac++] = 2;
	else if (a == 2) fac[nfac++] = 4;
	else if (a == 5) { fac[nfac++] = 4; fac[nfac++] = 8; }
	else if (a > 0) {
		int y = 0;
		while ((a - 3*y) % 4 != 0) y++;
		for (int i = 0; i < (a - 3*y)/4; i++) fac[nfac++] = 16;
		for (int i = 0; i < y; i++) fac[nfac++] = 8;
	}
}

int FftTables::pad_shift_of(int64_t n)
{
	if (!fast_ok(n)) return 31;
	int fac[FFT_MAX_FAC], nfac;
	factorize_fast((int)n, fac, nfac);
	int last = fac[nfac - 1];
	return last == 16 ? 4 : last == 8 ? 3 : last == 4 ? 2 : 31;
}

int64_t FftTables::smem_len(int64_t n)
{
	if (fast_ok(n)) return n + (n >> pad_shift_of(n));
	return smooth(n) ? n : bluestein_len((int)n);
}

static std::vector<int> digit_reversal(int n, const int *fac, int nfac)
{
	std::vector<int> rv(n);
	for (int k = 0; k < n; k++) {
		int kk = k, pos = 0, len = n;
		for (int f = 0; f < nfac; f++) { int r = fac[f]; len /= r; pos += (kk % r)*len; kk /= r; }
		rv[k] = pos;
	}
	return rv;
}

static std::vector<double2> twiddles(int n)
{
	std::vector<double2> t(n);
	for (int k = 0; k < n; k++) {
		long double a = TAU*(long double)k/(long double)n;
		t[k].x = (double)cosl(a); t[k].y = (double)-sinl(a);
	}
	return t;
}

int FftTables::prepare(int n, int ntab)
{
	B2_REQUIRE(n >= 1 && ntab % n == 0, "bad FFT table request n=%d ntab=%d", n, ntab);
	host = HostTables(); host.n = n; host.ntab = ntab;
	d.n = n; d.ntab = ntab; d.twmul = ntab/n;
	host.tw = twiddles(ntab);
	d.fast = 0; d.pad_shift = 31; d.ntw_hi = 0; d.bluestein = 0;
	d.tw = nullptr; d.rev = nullptr; d.btw = nullptr; d.chirp = nullptr; d.bhat = nullptr; d.rev_bits = 0; d.rev_total = 0;
	if (fast_ok(n)) {
		d.nt = n; d.fast = 1;
		factorize_fast(n, d.fac, d.nfac);
		d.pad_shift = pad_shift_of(n); d.nsmem = (int)smem_len(n); d.ntw_hi = (ntab + FFT_TWLO - 1)/FFT_TWLO;
		host.rev = digit_reversal(n, d.fac, d.nfac);
	} else if (smooth(n)) {
		d.nt = n; d.nsmem = n;
		factorize(n, d.fac, d.nfac);
		B2_REQUIRE(d.nfac <= FFT_MAX_FAC, "too many FFT factors");
		host.rev = digit_reversal(n, d.fac, d.nfac);
	} else {
		// Bluestein: chirp c_k = exp(-i pi k^2/n); X = c .* IFFT_M(FFT_M(x .* c) .* FFT_M(conj c wrapped))
		const int M = bluestein_len(n);
		d.bluestein = 1; d.nt = M; d.nsmem = M;
		factorize(M, d.fac, d.nfac);
		std::vector<int> rv = digit_reversal(M, d.fac, d.nfac);
		host.chirp.resize(n);
		std::vector<cld> b(M, cld(0, 0)), roots(M);
		for (int k = 0; k < M; k++) { long double a = TAU*(long double)k/(long double)M; roots[k] = cld(cosl(a), -sinl(a)); }
		for (int k = 0; k < n; k++) {
			long long k2 = ((long long)k*k) % (2LL*n);
			long double a = TAU*(long double)k2/(2.0L*n);
			host.chirp[k].x = (double)cosl(a); host.chirp[k].y = (double)-sinl(a);
			cld bc(cosl(a), sinl(a));
			b[k] = bc; if (k) b[M - k] = bc;
		}
		host_fft(b, roots);
		host.bhat.resize(M);
		for (int k = 0; k < M; k++) { host.bhat[rv[k]].x = (double)(b[k].real()/M); host.bhat[rv[k]].y = (double)(b[k].imag()/M); }
		host.rev.resize(n);
		for (int k = 0; k < n; k++) host.rev[k] = k;
		host.btw = twiddles(M);
	}
	// arithmetic digit reversal for power-of-two factorisations (checked against the table)
	if (!d.bluestein && d.nfac <= 8) {
		unsigned bits = 0; int total = 0; bool ok = true;
		for (int f = 0; f < d.nfac && ok; f++) {
			int r = d.fac[f], b = 0;
			while ((1 << b) < r) b++;
			ok = ((1 << b) == r) && b >= 1 && b <= 15;
			bits |= (unsigned)b << (4*f); total += b;
		}
		if (ok && (1 << total) == d.nt) {
			for (int k = 0; k < d.nt && ok; k++) {
				int pos = 0, sh = total, kk = k;
				for (int f = 0; f < d.nfac; f++) { int b = (bits >> (4*f)) & 15; sh -= b; pos |= (kk & ((1 << b) - 1)) << sh; kk >>= b; }
				ok = (pos == host.rev[k]);
			}
			if (ok) { d.rev_bits = bits; d.rev_total = total; }
		}
	}
	host.ready = true;
	return 0;
}

int FftTables::commit()
{
	B2_REQUIRE(host.ready, "FFT tables were not prepared");
	if (tw.upload(host.tw) || rev.upload(host.rev)) return 1;
	d.tw = tw.p; d.rev = rev.p;
	if (d.bluestein) {
		if (btw.upload(host.btw) || chirp.upload(host.chirp) || bhat.upload(host.bhat)) return 1;
		d.btw = btw.p; d.chirp = chirp.p; d.bhat = bhat.p;
	}
	host = HostTables();
	return 0;
}

int FftTables::build(int n, int ntab)
{
	if (!(host.ready && host.n == n && host.ntab == ntab) && prepare(n, ntab)) return 1;
	return commit();
}

__global__ void k_phase(double2 *ph, int mmax, double phi0)
{
	int m = blockIdx.x*blockDim.x + threadIdx.x;
	if (m > mmax) return;
	double s, c; sincos((double)m*phi0, &s, &c);
	ph[m] = make_double2(c, s);
}

int RingFft::prepare_tables(int64_t nphi_)
{
	const int hf = (nphi_ % 2 == 0) ? 1 : 0;
	return tab.prepare((int)(hf ? nphi_/2 : nphi_), (int)nphi_);
}

int RingFft::build(int64_t nphi_, double phi0, int xdir_, int64_t npix_, int nring_, const int64_t *rs,
	const double *w, int mmax_)
{
	nphi = nphi_; xdir = xdir_ < 0 ? -1 : 1; npix = npix_; nring = nring_; mmax = mmax_;
	B2_REQUIRE(nphi >= 1 && npix >= 1 && npix <= nphi, "bad ring description: nphi=%lld npix=%lld", (long long)nphi, (long long)npix);
	half = (nphi % 2 == 0) ? 1 : 0;
	nfft = (int)(half ? nphi/2 : nphi);
	if (tab.build(nfft, (int)nphi)) return 1;
	twoff = (int)FftTables::smem_len(nfft) + 1;      // data (one extra element for the packed real transform), then twiddle tables
	smem = sizeof(double2)*(size_t)(twoff + tab.twsm_len());
	B2_REQUIRE(smem <= 227*1024, "nphi=%lld needs %zu bytes of shared memory per ring (limit 227 KB)", (long long)nphi, smem);
	if (phase.alloc(mmax + 1)) return 1;
	k_phase<<<(mmax + 128)/128, 128>>>(phase.p, mmax, phi0);
	B2_LAUNCH_CHECK();
	std::vector<int64_t> r(rs, rs + nring);
	if (ringstart.upload(r)) return 1;
	if (w) { std::vector<double> ww(w, w + nring); if (weight.upload(ww)) return 1; }
	threads = (int)std::min<int64_t>(512, std::max<int64_t>(64, b2_round_up(nfft/4, 32)));
	B2_CHECK(cudaDeviceSynchronize());
	return 0;
}

int RingFft::build_group(int64_t nphi_, int nids, const int *ids, const double *phi0_of_id, int nring_total, const int64_t *rs,
	const double *w, int mmax_)
{
	B2_REQUIRE(nids >= 1 && ids && phi0_of_id, "bad ring group");
	if (build(nphi_, 0.0, 1, nphi_, nring_total, rs, w, mmax_)) return 1;
	nring = nids;                          // blocks per launch; ringstart / weight stay indexed by the plan's ring number
	std::vector<int> iv(ids, ids + nids);
	std::vector<double> pv(phi0_of_id, phi0_of_id + nids);
	if (ring_ids.upload(iv) || phi0s.upload(pv)) return 1;
	return 0;
}

// ------------------------------------------------------------------------------------ kernels

struct RingArgs {
	FftDesc d;
	int half, nfft, mmax, xdir, nring, twoff;
	int64_t nphi, npix, nring_pad;
	const double2 *phase; const int64_t *ringstart; const double *weight;
	const int *ring_ids; const double *phi0s;      // ring groups (null: block b = ring ring0 + b, phases from the table)
	int ring0;
	double2 *leg; void *map; int64_t map_cstride;
};

// e^{i m phi0} of this block's ring
__device__ __forceinline__ double2 ring_phase(const RingArgs &A, int m, int bx)
{
	if (!A.phi0s) return __ldg(&A.phase[m]);
	double sn, c; sincos((double)m*A.phi0s[bx], &sn, &c);
	return make_double2(c, sn);
}

// TAB = true: the plan's table (cylindrical maps).  The choice is made OUTSIDE the hot loops: a branch inside them keeps the
// loads of an unrolled loop from being issued together (one outstanding load per thread, profiles/r3m_fftk_stalls.txt).
template<bool TAB> __device__ __forceinline__ double2 ring_phase_t(const RingArgs &A, int m, int bx)
{
	if (TAB) return __ldg(&A.phase[m]);
	double sn, c; sincos((double)m*A.phi0s[bx], &sn, &c);
	return make_double2(c, sn);
}

__device__ __forceinline__ double2 leg_phase(const RingArgs &A, const double2 *legc, int m, int bx)
{
	double2 g = cmul(legc[(int64_t)m*A.nring_pad], ring_phase(A, m, bx));
	if (A.xdir < 0) g.y = -g.y;
	return g;
}

// bx: index of this block's ring in the launch (the block index, or the position inside its group for the packed launches)
template<typename MapT> __device__ __forceinline__ void leg2map_body(const RingArgs &A, const int bx)
{
	extern __shared__ __align__(16) double2 s[];
	const int ring = A.ring_ids ? A.ring_ids[bx] : A.ring0 + bx, comp = blockIdx.y, tid = threadIdx.x, T = blockDim.x;
	const double2 *legc = A.leg + ((int64_t)comp*(A.mmax + 1))*A.nring_pad + ring;
	MapT *row = (MapT*)A.map + (int64_t)comp*A.map_cstride + A.ringstart[ring];
	const int n = (int)A.nphi, nf = A.nfft, mmax = A.mmax;
	#define SI(i) fft_pad(A.d, (i))
	const double2 *twsm = s + A.twoff;
	fft_load_tw(s + A.twoff, A.d, tid, T);
	if (A.half) {
		if (mmax < nf) {
			// no aliasing: X[k] = leg_k e^{i k phi0} for k <= mmax, 0 above; branch-free so that the loads of a thread overlap
			const bool flip = A.xdir < 0;
			auto fill = [&](auto TAB) {
				#pragma unroll 8
				for (int k = tid; k <= nf; k += T) {
					const int kc = min(k, mmax);
					double2 g = cmul(__ldg(&legc[(int64_t)kc*A.nring_pad]), ring_phase_t<decltype(TAB)::value>(A, kc, bx));
					if (flip) g.y = -g.y;
					if (k == 0) g = make_double2(g.x, 0.0);
					if (k > mmax) g = make_double2(0, 0);
					s[SI(k)] = g;
				}
			};
			if (A.phi0s) fill(std::false_type()); else fill(std::true_type());
		} else {
			// half spectrum X[0..nf] of the real ring, |m| aliased mod nphi
			for (int k = tid; k <= nf; k += T) {
				double2 acc = make_double2(0, 0);
				if (k == 0 || k == nf) {
					for (int m = k; m <= mmax; m += n) { double2 g = leg_phase(A, legc, m, bx); acc.x += (m == 0 ? 1.0 : 2.0)*g.x; }
				} else {
					for (int m = k; m <= mmax; m += n) acc = cadd(acc, leg_phase(A, legc, m, bx));
					for (int m = n - k; m <= mmax; m += n) acc = cadd(acc, cconj(leg_phase(A, legc, m, bx)));
				}
				s[SI(k)] = acc;
			}
		}
		__syncthreads();
		// Z[k] = (X[k] + conj X[nf-k]) + i e^{+2 pi i k/n} (X[k] - conj X[nf-k]); z = IFFT(Z) packs (x_2j, x_2j+1)
		#pragma unroll 4
		for (int k = tid; 2*k <= nf; k += T) {
			if (k == 0) { double x0 = s[0].x, xn = s[SI(nf)].x; s[0] = make_double2(x0 + xn, x0 - xn); }
			else {
				int kk = nf - k;
				double2 a = s[SI(k)], b = s[SI(kk)];
				double2 s1 = make_double2(a.x + b.x, a.y - b.y), d1 = make_double2(a.x - b.x, a.y + b.y);
				double2 w = __ldg(&A.d.tw[k]); w.y = -w.y;
				double2 wd = cmul(w, d1);
				s[SI(k)] = make_double2(s1.x - wd.y, s1.y + wd.x);
				if (kk != k) s[SI(kk)] = make_double2(s1.x + wd.y, -s1.y + wd.x);
			}
		}
		__syncthreads();
		fft_smem<true>(s, A.d, tid, T, 1, twsm);
		#pragma unroll 8
		for (int64_t i = tid; i < A.npix; i += T) {
			double2 z = s[SI(fft_rev(A.d, (int)(i >> 1)))];
			row[i] = (MapT)((i & 1) ? z.y : z.x);
		}
	} else {
		for (int k = tid; k < n; k += T) {
			double2 acc = make_double2(0, 0);
			for (int m = k; m <= mmax; m += n) { double2 g = leg_phase(A, legc, m, bx); if (m == 0) g.y = 0; acc = cadd(acc, g); }
			for (int m = (k == 0 ? n : n - k); m <= mmax; m += n) acc = cadd(acc, cconj(leg_phase(A, legc, m, bx)));
			s[SI(k)] = acc;
		}
		__syncthreads();
		fft_smem<true>(s, A.d, tid, T, 1, twsm);
		for (int64_t i = tid; i < A.npix; i += T) row[i] = (MapT)s[SI(A.d.rev[i])].x;
	}
}

template<typename MapT> __device__ __forceinline__ void map2leg_body(const RingArgs &A, const int bx)
{
	extern __shared__ __align__(16) double2 s[];
	const int ring = A.ring_ids ? A.ring_ids[bx] : A.ring0 + bx, comp = blockIdx.y, tid = threadIdx.x, T = blockDim.x;
	double2 *legc = A.leg + ((int64_t)comp*(A.mmax + 1))*A.nring_pad + ring;
	const MapT *row = (const MapT*)A.map + (int64_t)comp*A.map_cstride + A.ringstart[ring];
	const int n = (int)A.nphi, nf = A.nfft, mmax = A.mmax;
	const double wgt = A.weight ? A.weight[ring] : 1.0;
	const double2 *twsm = s + A.twoff;
	fft_load_tw(s + A.twoff, A.d, tid, T);
	if (A.half) {
		#pragma unroll 8
		for (int j = tid; j < nf; j += T) {
			int64_t i = 2*(int64_t)j;
			double a = i < A.npix ? (double)row[i] : 0.0, b = i + 1 < A.npix ? (double)row[i + 1] : 0.0;
			s[SI(j)] = make_double2(a, b);
		}
		__syncthreads();
		fft_smem<false>(s, A.d, tid, T, 1, twsm);
		if (mmax < nf) {
			// no aliasing (k = m): branch-free so that the table loads and the scattered stores of a thread overlap
			const bool flip = A.xdir < 0;
			#pragma unroll 4
			for (int m = tid; m <= mmax; m += T) {
				const int k2 = m == 0 ? 0 : nf - m;
				double2 zk = s[SI(fft_rev(A.d, m))], zc = cconj(s[SI(fft_rev(A.d, k2))]);
				double2 e = make_double2(0.5*(zk.x + zc.x), 0.5*(zk.y + zc.y));
				double2 dd = make_double2(0.5*(zk.x - zc.x), 0.5*(zk.y - zc.y));
				double2 o = make_double2(dd.y, -dd.x);
				double2 x = cadd(e, cmul(__ldg(&A.d.tw[m]), o));
				if (flip) x.y = -x.y;
				legc[(int64_t)m*A.nring_pad] = cscale(cmulc(x, ring_phase(A, m, bx)), wgt);
			}
		} else
		for (int m = tid; m <= mmax; m += T) {
			int k = m % n; bool fold = k > nf; if (fold) k = n - k;
			int k1 = k == nf ? 0 : k, k2 = k == 0 ? 0 : nf - k;
			double2 zk = s[SI(A.d.rev[k1])], zc = cconj(s[SI(A.d.rev[k2])]);
			double2 e = make_double2(0.5*(zk.x + zc.x), 0.5*(zk.y + zc.y));
			double2 dd = make_double2(0.5*(zk.x - zc.x), 0.5*(zk.y - zc.y));
			double2 o = make_double2(dd.y, -dd.x);
			double2 x = cadd(e, cmul(A.d.tw[k], o));
			if (fold) x.y = -x.y;
			if (A.xdir < 0) x.y = -x.y;
			legc[(int64_t)m*A.nring_pad] = cscale(cmulc(x, ring_phase(A, m, bx)), wgt);
		}
	} else {
		for (int j = tid; j < n; j += T) s[SI(j)] = make_double2(j < A.npix ? (double)row[j] : 0.0, 0.0);
		__syncthreads();
		fft_smem<false>(s, A.d, tid, T, 1, twsm);
		for (int m = tid; m <= mmax; m += T) {
			double2 x = s[SI(A.d.rev[m % n])];
			if (A.xdir < 0) x.y = -x.y;
			legc[(int64_t)m*A.nring_pad] = cscale(cmulc(x, ring_phase(A, m, bx)), wgt);
		}
	}
}

template<typename MapT> __global__ void __launch_bounds__(512) k_leg2map(RingArgs A) { leg2map_body<MapT>(A, blockIdx.x); }
template<typename MapT> __global__ void __launch_bounds__(512) k_map2leg(RingArgs A) { map2leg_body<MapT>(A, blockIdx.x); }

// Packed launches for ring sets with many distinct ring lengths (HEALPix: one group per nphi, two rings each in the caps):
// all groups whose transforms want the same block size share ONE launch; block b looks up its group and its position in
// it, takes the group's FFT description from device memory and runs the same body.  (One launch per group made a HEALPix
// transform at nside 2048 a string of 4096 small launches.)
struct GroupDesc { FftDesc d; int half, nfft, twoff, nring; int64_t nphi, npix; const int *ring_ids; const double *phi0s; };

__device__ __forceinline__ RingArgs pack_args(const RingArgs &base, const GroupDesc *D, const int2 *blk, int &bx)
{
	const int2 b = blk[blockIdx.x];
	const GroupDesc &G = D[b.x];
	RingArgs A = base;
	A.d = G.d; A.half = G.half; A.nfft = G.nfft; A.twoff = G.twoff; A.nring = G.nring; A.nphi = G.nphi; A.npix = G.npix;
	A.ring_ids = G.ring_ids; A.phi0s = G.phi0s;
	bx = b.y;
	return A;
}
template<typename MapT> __global__ void __launch_bounds__(512) k_leg2map_pack(RingArgs base, const GroupDesc *D, const int2 *blk)
{
	int bx; const RingArgs A = pack_args(base, D, blk, bx);
	leg2map_body<MapT>(A, bx);
}
template<typename MapT> __global__ void __launch_bounds__(512) k_map2leg_pack(RingArgs base, const GroupDesc *D, const int2 *blk)
{
	int bx; const RingArgs A = pack_args(base, D, blk, bx);
	map2leg_body<MapT>(A, bx);
}

// ------------------------------------------------------------------------------------ host

static RingArgs ring_args(const RingFft &F, const double2 *leg, int64_t nring_pad, const void *map, int64_t map_cstride, int use_weight)
{
	RingArgs A;
	A.d = F.tab.d; A.twoff = F.twoff; A.half = F.half; A.nfft = F.nfft; A.mmax = F.mmax; A.xdir = F.xdir; A.nring = F.nring;
	A.nphi = F.nphi; A.npix = F.npix; A.nring_pad = nring_pad;
	A.phase = F.phase.p; A.ringstart = F.ringstart.p; A.weight = (use_weight && F.weight.n) ? F.weight.p : nullptr;
	A.ring_ids = F.ring_ids.n ? F.ring_ids.p : nullptr; A.phi0s = F.phi0s.n ? F.phi0s.p : nullptr;
	A.leg = (double2*)leg; A.map = (void*)map; A.map_cstride = map_cstride; A.ring0 = 0;
	return A;
}

// raise the kernel's dynamic shared-memory limit when a launch needs more than any before it (the limit is per
// kernel and device; plans with many ring groups would otherwise pay the driver call on every launch)
template<typename K> static int set_smem(K kern, size_t smem)
{
	static thread_local std::map<std::pair<int, const void*>, size_t> granted;
	if (smem <= 48*1024) return 0;
	int dev; B2_CHECK(cudaGetDevice(&dev));
	size_t &g = granted[std::make_pair(dev, (const void*)kern)];
	if (smem > g) {
		B2_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		g = smem;
	}
	return 0;
}

int ring_leg2map(const RingFft &F, int ncomp, const double2 *leg, int64_t nring_pad,
	void *map, int64_t map_cstride, int dtype, cudaStream_t st, int ring0, int nrings)
{
	RingArgs A = ring_args(F, leg, nring_pad, map, map_cstride, 0);
	dim3 grid(F.nring, ncomp);
	if (nrings > 0) { B2_REQUIRE(!F.ring_ids.n && ring0 >= 0 && ring0 + nrings <= F.nring, "leg2map: bad ring range"); A.ring0 = ring0; grid.x = nrings; }
	if (dtype == 0) { if (set_smem(k_leg2map<double>, F.smem)) return 1; k_leg2map<double><<<grid, F.threads, F.smem, st>>>(A); }
	else            { if (set_smem(k_leg2map<float>,  F.smem)) return 1; k_leg2map<float><<<grid, F.threads, F.smem, st>>>(A); }
	B2_LAUNCH_CHECK();
	return 0;
}

int ring_map2leg(const RingFft &F, int ncomp, double2 *leg, int64_t nring_pad,
	const void *map, int64_t map_cstride, int dtype, int use_weight, cudaStream_t st)
{
	RingArgs A = ring_args(F, leg, nring_pad, map, map_cstride, use_weight);
	dim3 grid(F.nring, ncomp);
	if (dtype == 0) { if (set_smem(k_map2leg<double>, F.smem)) return 1; k_map2leg<double><<<grid, F.threads, F.smem, st>>>(A); }
	else            { if (set_smem(k_map2leg<float>,  F.smem)) return 1; k_map2leg<float><<<grid, F.threads, F.smem, st>>>(A); }
	B2_LAUNCH_CHECK();
	return 0;
}

// ------------------------------------------------------------------------------------ packed groups

int RingPack::build(const std::vector<std::unique_ptr<RingFft>> &groups)
{
	buckets.clear();
	if (groups.empty()) return 0;
	std::vector<GroupDesc> gd(groups.size());
	for (size_t g = 0; g < groups.size(); g++) {
		const RingFft &F = *groups[g];
		GroupDesc &G = gd[g];
		G.d = F.tab.d; G.half = F.half; G.nfft = F.nfft; G.twoff = F.twoff; G.nring = F.nring; G.nphi = F.nphi; G.npix = F.npix;
		G.ring_ids = F.ring_ids.p; G.phi0s = F.phi0s.p;
	}
	if (desc.alloc(sizeof(GroupDesc)*gd.size())) return 1;
	B2_CHECK(cudaMemcpy(desc.p, gd.data(), sizeof(GroupDesc)*gd.size(), cudaMemcpyHostToDevice));
	// one bucket per block size; inside a bucket the long rings go first
	std::map<int, std::vector<int>> by_threads;
	for (size_t g = 0; g < groups.size(); g++) by_threads[groups[g]->threads].push_back((int)g);
	std::vector<int2> all;
	for (auto &bt : by_threads) {
		std::vector<int> &gs = bt.second;
		std::sort(gs.begin(), gs.end(), [&](int a, int b) { return groups[a]->nfft > groups[b]->nfft; });
		Bucket B; B.threads = bt.first; B.first = (int)all.size(); B.smem = 0;
		for (int g : gs) {
			B.smem = std::max(B.smem, groups[g]->smem);
			for (int i = 0; i < groups[g]->nring; i++) all.push_back(make_int2(g, i));
		}
		B.nblocks = (int)all.size() - B.first;
		buckets.push_back(B);
	}
	if (blocks.upload(all)) return 1;
	return 0;
}

int ring_leg2map_pack(const RingPack &P, const RingFft &F0, int ncomp, const double2 *leg, int64_t nring_pad, void *map, int64_t map_cstride, int dtype, cudaStream_t st)
{
	RingArgs A = ring_args(F0, leg, nring_pad, map, map_cstride, 0);
	for (const RingPack::Bucket &B : P.buckets) {
		dim3 grid(B.nblocks, ncomp);
		const int2 *blk = P.blocks.p + B.first;
		if (dtype == 0) { if (set_smem(k_leg2map_pack<double>, B.smem)) return 1; k_leg2map_pack<double><<<grid, B.threads, B.smem, st>>>(A, (const GroupDesc*)P.desc.p, blk); }
		else            { if (set_smem(k_leg2map_pack<float>,  B.smem)) return 1; k_leg2map_pack<float><<<grid, B.threads, B.smem, st>>>(A, (const GroupDesc*)P.desc.p, blk); }
		B2_LAUNCH_CHECK();
	}
	return 0;
}

int ring_map2leg_pack(const RingPack &P, const RingFft &F0, int ncomp, double2 *leg, int64_t nring_pad, const void *map, int64_t map_cstride, int dtype, int use_weight, cudaStream_t st)
{
	RingArgs A = ring_args(F0, leg, nring_pad, map, map_cstride, use_weight);
	for (const RingPack::Bucket &B : P.buckets) {
		dim3 grid(B.nblocks, ncomp);
		const int2 *blk = P.blocks.p + B.first;
		if (dtype == 0) { if (set_smem(k_map2leg_pack<double>, B.smem)) return 1; k_map2leg_pack<double><<<grid, B.threads, B.smem, st>>>(A, (const GroupDesc*)P.desc.p, blk); }
		else            { if (set_smem(k_map2leg_pack<float>,  B.smem)) return 1; k_map2leg_pack<float><<<grid, B.threads, B.smem, st>>>(A, (const GroupDesc*)P.desc.p, blk); }
		B2_LAUNCH_CHECK();
	}
	return 0;
}
