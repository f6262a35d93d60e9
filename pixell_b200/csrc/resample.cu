// resample.cu -- K5, see resample.cuh.  Five batched length-N FFT passes over the (comp, m) columns
// of leg (contiguous in ring), N = size of the grid's circle extension (2 ntheta for F1):
//   S1  ext -> FFT -> half-sample phase shift            A = C_f e^{i pi f/N}
//   S2  A -> IFFT -> x weight at the shifted nodes       B = W(theta_{k+1/2}) F(theta_{k+1/2})
//   S3  W(theta_k) ext -> FFT, keep |f| <= lmax          A = a_f
//   S4  B -> FFT, combine                                A = (a_f + e^{-i pi f/N} b_f)/2, |f| <= lmax
//   S5  A -> IFFT -> B = g on the circle;   S6  leg = 2 mult g(theta_k)
// Two columns share every transform: neighbouring m of one component have opposite parity under
// theta -> 2 pi - theta, sigma = (-1)^(m+s), and every operator above commutes with that mirror map, so the
// pipeline runs on z = ext(col m) + ext(col m+1) and S6 separates the results again by parity,
//   g_m = (z + sigma z o mirror)/2,  g_{m+1} = (z - sigma z o mirror)/2  (half the FFT work per column).
// The two polyphase halves {theta_k}, {theta_{k+1/2}} together are the Clenshaw-Curtis circle grid
// with 2N points on which the weight function W is exact for band limit 2 lmax + 1.
// Transforms longer than the shared-memory capacity are split decimation-in-frequency over P CTAs
// (each CTA forms y_p[j] = w_N^{jp} sum_q x[j + qN/P] w_P^{qp} on load and owns outputs k = p mod P).
#include "resample.cuh"
#include <algorithm>
#include <stdlib.h>

struct ResampArgs {
	FftDesc d;
	int N, P, o2, n, L, nm, npc, spin, twoff, M0;
	int64_t nring_pad;
	const int *src, *pos, *mir; const double *wfine; const double *mult;
	double2 *leg, *A, *B, *C;
	int64_t col0;          // first column pair of this batch (pair index = comp*npc + i, columns m = 2i, 2i+1)
};

// the two leg columns of pair `pidx` (the second is null for the unpaired last m) and the parity of the first
__device__ __forceinline__ void pair_cols(const ResampArgs &R, int64_t pidx, double2 *&ca, double2 *&cb, double &sigma)
{
	int comp = (int)(pidx / R.npc), i = (int)(pidx % R.npc), m = 2*i;
	ca = R.leg + ((int64_t)comp*R.nm + m)*R.nring_pad;
	cb = (m + 1 < R.nm) ? ca + R.nring_pad : nullptr;
	sigma = ((m + R.spin) & 1) ? -1.0 : 1.0;
}

// a ring that is its own mirror image (a pole ring of CC / MW / MWflip)
__device__ __forceinline__ bool self_mirrored(const ResampArgs &R, int r) { return 2*r == R.M0 || 2*r == R.M0 - R.N; }

// sample j of z = ext(col a) + ext(col b) on the circle: slots [0, n) hold the rings themselves, slots [n, N) the
// mirror images, ring M0 - j (M0 = N - 1 for F1 / MW, N for CC / MWflip), with the parity sign.  On a
// self-mirrored ring a column of odd parity must vanish: whatever the data holds there is dropped, which keeps the
// two columns of a pair exactly separable (no leakage between m and m+1 for inconsistent pole-ring input).
__device__ __forceinline__ double2 ext_load(const ResampArgs &R, const double2 *ca, const double2 *cb, int j, double sigma)
{
	const bool mirrored = j >= R.n;
	const int r = mirrored ? R.M0 - j : j;
	double2 a = ca[r], b = cb ? cb[r] : make_double2(0, 0);
	if (mirrored) return make_double2(sigma*(a.x - b.x), sigma*(a.y - b.y));
	if (self_mirrored(R, r)) return sigma > 0 ? a : b;
	return make_double2(a.x + b.x, a.y + b.y);
}

template<int STAGE, int P> __global__ void __launch_bounds__(512) k_resamp(ResampArgs R)
{
	extern __shared__ __align__(16) double2 s[];
	constexpr bool INV = (STAGE == 2 || STAGE == 5);
	const int tid = threadIdx.x, T = blockDim.x, c = blockIdx.y, p = blockIdx.x;      // the P CTAs of a column pair are neighbours in the grid (second reader hits L2)
	const int N = R.N, Nl = N/P;
	double2 *ca, *cb; double sigma;
	pair_cols(R, R.col0 + c, ca, cb, sigma);
	double2 *A = R.A + (int64_t)c*N, *B = R.B + (int64_t)c*N;
	const double2 *twsm = s + R.twoff;
	fft_load_tw(s + R.twoff, R.d, tid, T);
	// w_P^{qp} of the decimation-in-frequency split
	double2 wq[P];
	#pragma unroll
	for (int q = 0; q < P; q++) wq[q] = cj(R.d.tw[(2*N/P)*((q*p) % P)], INV);
	#pragma unroll 4
	for (int j = tid; j < Nl; j += T) {
		double2 v[P];
		#pragma unroll
		for (int q = 0; q < P; q++) {
			const int idx = j + q*Nl;
			if (STAGE == 1) v[q] = ext_load(R, ca, cb, idx, sigma);
			else if (STAGE == 3) { int t = 2*idx + R.o2; v[q] = cscale(ext_load(R, ca, cb, idx, sigma), __ldg(&R.wfine[t >= 2*N ? t - 2*N : t])); }
			else if (STAGE == 4) v[q] = B[idx];
			else v[q] = A[idx];
		}
		double2 acc = v[0];
		#pragma unroll
		for (int q = 1; q < P; q++) acc = cadd(acc, p ? cmul(v[q], wq[q]) : v[q]);
		if (P > 1 && p) acc = cmul(acc, cj(__ldg(&R.d.tw[2*j*p]), INV));
		s[fft_pad(R.d, j)] = acc;
	}
	__syncthreads();
	fft_smem<INV>(s, R.d, tid, T, 1, twsm);
	const double inv = 1.0/N;
	#pragma unroll 8
	for (int kk = tid; kk < Nl; kk += T) {
		const int k = p + P*kk;
		double2 x = s[fft_pad(R.d, fft_rev(R.d, kk))];
		const int f = (2*k <= N) ? k : k - N;
		const bool nyq = (2*k == N);
		if (STAGE == 1) {
			double2 w = __ldg(&R.d.tw[f >= 0 ? f : -f]); if (f >= 0) w.y = -w.y;     // e^{+i pi f/N}
			A[k] = nyq ? make_double2(0, 0) : cscale(cmul(x, w), inv);
		} else if (STAGE == 2) {
			int t = 2*k + R.o2 + 1;
			B[k] = cscale(x, __ldg(&R.wfine[t >= 2*N ? t - 2*N : t]));
		} else if (STAGE == 3) {
			A[k] = (abs(f) <= R.L && !nyq) ? cscale(x, inv) : make_double2(0, 0);
		} else if (STAGE == 4) {
			if (abs(f) <= R.L && !nyq) {
				double2 w = __ldg(&R.d.tw[f >= 0 ? f : -f]); if (f < 0) w.y = -w.y;  // e^{-i pi f/N}
				double2 b = cscale(cmul(x, w), inv), a = A[k];
				A[k] = make_double2(0.5*(a.x + b.x), 0.5*(a.y + b.y));
			}
		} else {
			B[k] = x;      // not A: with P > 1 the other CTAs of this column are still reading A
		}
	}
}

// ---- adjoint of the whole operator (for adjoint_analysis_2d).  With F / Fi the unnormalised forward / inverse
// DFT, D = diag(e^{i pi f/N}), Z the Nyquist projector, LP the |f| <= lmax projector, Wo / Wh the weight function at
// the original / half-shifted nodes, E the parity extension and R the restriction with the factor 2 mult:
//   K  = R Fi (LP/2) [ (1/N) F Wo + D^-1 (1/N) F Wh Fi D Z (1/N) F ] E
//   K^H = E^H [ Wo (1/N) Fi + (1/N) Fi Z D^-1 F Wh (1/N) Fi D ] (LP/2) F R^H
//   T1  R^H leg -> F -> LP/2                      A = c
//   T2  D A -> Fi -> x Wh/N                        B = b1
//   T3  B -> F -> x Z D^-1 / N                     C = b2
//   T4  C -> Fi                                    B = b3
//   T5  A -> Fi -> x Wo/N                          C = a
//   T6  leg = E^H (B + C)
// The column pairing of the forward operator carries over (R^H and E^H act on the pair).
template<int STAGE, int P> __global__ void __launch_bounds__(512) k_resamp_adj(ResampArgs R)
{
	extern __shared__ __align__(16) double2 s[];
	constexpr bool INV = (STAGE == 2 || STAGE == 4 || STAGE == 5);
	const int tid = threadIdx.x, T = blockDim.x, c = blockIdx.y, p = blockIdx.x;      // the P CTAs of a column pair are neighbours in the grid (second reader hits L2)
	const int N = R.N, Nl = N/P;
	double2 *ca, *cb; double sigma;
	pair_cols(R, R.col0 + c, ca, cb, sigma);
	double2 *A = R.A + (int64_t)c*N, *B = R.B + (int64_t)c*N, *C = R.C + (int64_t)c*N;
	const double2 *twsm = s + R.twoff;
	fft_load_tw(s + R.twoff, R.d, tid, T);
	double2 wq[P];
	#pragma unroll
	for (int q = 0; q < P; q++) wq[q] = cj(R.d.tw[(2*N/P)*((q*p) % P)], INV);
	#pragma unroll 4
	for (int j = tid; j < Nl; j += T) {
		double2 v[P];
		#pragma unroll
		for (int q = 0; q < P; q++) {
			const int idx = j + q*Nl;
			if (STAGE == 1) {
				// R^H: ring r feeds its own slot with mult (ya + yb) and its mirror slot with mult sigma (ya - yb)
				const bool mirrored = idx >= R.n;
				const int r = mirrored ? R.M0 - idx : idx;
				double2 a = ca[r], b = cb ? cb[r] : make_double2(0, 0);
				const double mu = R.mult[r];
				double2 sum = make_double2(a.x + b.x, a.y + b.y), dif = make_double2(sigma*(a.x - b.x), sigma*(a.y - b.y));
				if (mirrored) v[q] = cscale(dif, mu);
				else if (self_mirrored(R, r)) v[q] = cadd(sum, dif);      // pole ring: both contributions land here (the odd column cancels)
				else v[q] = cscale(sum, mu);
			} else if (STAGE == 2) {
				const int f = (2*idx <= N) ? idx : idx - N;
				double2 w = __ldg(&R.d.tw[f >= 0 ? f : -f]); if (f >= 0) w.y = -w.y;     // e^{+i pi f/N}
				v[q] = cmul(A[idx], w);
			} else if (STAGE == 3) v[q] = B[idx];
			else if (STAGE == 4) v[q] = C[idx];
			else v[q] = A[idx];
		}
		double2 acc = v[0];
		#pragma unroll
		for (int q = 1; q < P; q++) acc = cadd(acc, p ? cmul(v[q], wq[q]) : v[q]);
		if (P > 1 && p) acc = cmul(acc, cj(__ldg(&R.d.tw[2*j*p]), INV));
		s[fft_pad(R.d, j)] = acc;
	}
	__syncthreads();
	fft_smem<INV>(s, R.d, tid, T, 1, twsm);
	const double inv = 1.0/N;
	#pragma unroll 8
	for (int kk = tid; kk < Nl; kk += T) {
		const int k = p + P*kk;
		double2 x = s[fft_pad(R.d, fft_rev(R.d, kk))];
		const int f = (2*k <= N) ? k : k - N;
		const bool nyq = (2*k == N);
		if (STAGE == 1) {
			A[k] = (abs(f) <= R.L && !nyq) ? cscale(x, 0.5) : make_double2(0, 0);
		} else if (STAGE == 2) {
			int t = 2*k + R.o2 + 1;
			B[k] = cscale(x, inv*__ldg(&R.wfine[t >= 2*N ? t - 2*N : t]));
		} else if (STAGE == 3) {
			double2 w = __ldg(&R.d.tw[f >= 0 ? f : -f]); if (f < 0) w.y = -w.y;      // e^{-i pi f/N}
			C[k] = nyq ? make_double2(0, 0) : cscale(cmul(x, w), inv);
		} else if (STAGE == 4) {
			B[k] = x;
		} else {
			int t = 2*k + R.o2;
			C[k] = cscale(x, inv*__ldg(&R.wfine[t >= 2*N ? t - 2*N : t]));
		}
	}
}

// T6: leg = E^H (B + C) for both columns of each pair
__global__ void k_resamp_adj_gather(ResampArgs R)
{
	double2 *ca, *cb; double sigma;
	pair_cols(R, R.col0 + blockIdx.x, ca, cb, sigma);
	const double2 *B = R.B + (int64_t)blockIdx.x*R.N, *C = R.C + (int64_t)blockIdx.x*R.N;
	for (int r = threadIdx.x; r < R.n; r += blockDim.x) {
		const int ip = R.pos[r], im = R.mir[r];
		double2 up = cadd(B[ip], C[ip]);
		if (im == ip) {      // pole ring: only the column of even parity lives there
			ca[r] = sigma > 0 ? up : make_double2(0, 0);
			if (cb) cb[r] = sigma > 0 ? make_double2(0, 0) : up;
			continue;
		}
		double2 um = cadd(B[im], C[im]);
		ca[r] = make_double2(up.x + sigma*um.x, up.y + sigma*um.y);
		if (cb) cb[r] = make_double2(up.x - sigma*um.x, up.y - sigma*um.y);
	}
}

// S6: separate the two columns of each pair by parity and scale: leg = 2 mult g(theta_k)
__global__ void k_resamp_split(ResampArgs R)
{
	double2 *ca, *cb; double sigma;
	pair_cols(R, R.col0 + blockIdx.x, ca, cb, sigma);
	const double2 *A = R.B + (int64_t)blockIdx.x*R.N;
	for (int r = threadIdx.x; r < R.n; r += blockDim.x) {
		double2 zp = A[R.pos[r]], zm = A[R.mir[r]];
		double mu = R.mult[r];
		ca[r] = make_double2(mu*(zp.x + sigma*zm.x), mu*(zp.y + sigma*zm.y));
		if (cb) cb[r] = make_double2(mu*(zp.x - sigma*zm.x), mu*(zp.y - sigma*zm.y));
	}
}

// weight function of the Clenshaw-Curtis rule with nt = N+1 rings on its 2N-point circle grid:
// W_t = (4 pi / 2N) (1 - sum_j c_j cos(2 j t pi/N)/(4 j^2 - 1)) / nphi, c_j = 2 (1 when 2j == N)
__global__ void k_wfine(double *w, int N, double scale)
{
	int t = blockIdx.x*blockDim.x + threadIdx.x;
	if (t > N) return;
	double sum = 0;
	for (int j = N/2; j >= 1; j--) {
		double c = (2*j == N) ? 1.0 : 2.0;
		long long r = (2LL*j*t) % (2LL*N);
		sum += c*cospi((double)r/(double)N)/(4.0*j*j - 1.0);
	}
	double v = scale*(1.0 - sum);
	w[t] = v;
	if (t > 0 && t < N) w[2*N - t] = v;
}

// shared memory a CTA may use for one transform (measured: one CTA per SM with the transform split over two CTAs is as
// fast as smaller pieces with several CTAs per SM; B2_RESAMP_SMEM_KB overrides, for tuning)
static size_t resamp_smem_limit()
{
	const char *e = getenv("B2_RESAMP_SMEM_KB");
	return (size_t)(e ? atoi(e) : 210)*1024;
}

bool ThetaResampler::needed(const std::string &g, int ntheta, int lmax)
{
	if (g == "DH" || g == "F2") return false;
	return ntheta < 2*lmax + 2;
}

int ThetaResampler::build(const std::string &g, int ntheta, int64_t nphi_, int lmax_, int mmax, int64_t nring_pad_)
{
	n = ntheta; lmax = lmax_; nm = mmax + 1; npc = (nm + 1)/2; nring_pad = nring_pad_; nphi = nphi_;
	std::vector<int> pos(n), mir(n); std::vector<double> mu(n, 2.0);
	M0 = (g == "CC" || g == "MWflip") ? -1 : 0;      // completed below
	if (g == "CC")          { N = 2*(n - 1); o2 = 0; for (int k = 0; k < n; k++) { pos[k] = k; mir[k] = (N - k) % N; } }
	else if (g == "F1")     { N = 2*n;       o2 = 1; for (int k = 0; k < n; k++) { pos[k] = k; mir[k] = N - 1 - k; } }
	else if (g == "MW")     { N = 2*n - 1;   o2 = 1; for (int k = 0; k < n; k++) { pos[k] = k; mir[k] = N - 1 - k; } }
	else if (g == "MWflip") { N = 2*n - 1;   o2 = 0; for (int k = 0; k < n; k++) { pos[k] = k; mir[k] = (N - k) % N; } }
	else { b2_set_error("theta resampling is not defined for geometry %s", g.c_str()); return 1; }
	M0 = (M0 < 0) ? N : N - 1;
	B2_REQUIRE(N > 2*lmax, "grid %s with %d rings cannot carry lmax=%d", g.c_str(), n, lmax);
	std::vector<int> sr(N, -1);
	for (int k = 0; k < n; k++) {
		sr[pos[k]] = k;
		if (mir[k] == pos[k]) mu[k] = 1.0; else sr[mir[k]] = k | 0x40000000;
	}
	P = 1;
	while ((size_t)(FftTables::smem_len(N/P) + 2*N/FFT_TWLO + FFT_TWLO + 1)*sizeof(double2) > resamp_smem_limit()) {
		int np = P*2;
		B2_REQUIRE(np <= 8 && N % np == 0, "theta transform of length %d does not fit in shared memory", N);
		P = np;
	}
	if (tab.build(N/P, 2*N)) return 1;
	twoff = (int)FftTables::smem_len(N/P);
	smem = sizeof(double2)*(size_t)(twoff + tab.twsm_len());
	threads = (int)std::min<int64_t>(512, std::max<int64_t>(64, b2_round_up(N/P/16, 32)));      // one radix-16 butterfly per thread and pass
	if (src.upload(sr) || mult.upload(mu) || wfine.alloc(2*(size_t)N) || dpos.upload(pos) || dmir.upload(mir)) return 1;
	k_wfine<<<(N + 128)/128, 128>>>(wfine.p, N, 4.0*M_PI/(2.0*N)/(double)nphi);
	B2_LAUNCH_CHECK();
	cb = std::max<int64_t>(1, std::min<int64_t>((int64_t)npc*2, (int64_t)(1 << 26)/N));
	if (A.alloc((size_t)cb*N) || B.alloc((size_t)cb*N)) return 1;
	B2_CHECK(cudaDeviceSynchronize());
	return 0;
}

template<int STAGE, int P> static int launch_stage_p(const ResampArgs &R, int ncols, int threads, size_t smem, cudaStream_t st)
{
	if (smem > 48*1024) B2_CHECK(cudaFuncSetAttribute(k_resamp<STAGE, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_resamp<STAGE, P><<<dim3(P, ncols), threads, smem, st>>>(R);
	B2_LAUNCH_CHECK();
	return 0;
}
template<int STAGE> static int launch_stage(const ResampArgs &R, int ncols, int threads, size_t smem, cudaStream_t st)
{
	switch (R.P) {
		case 1: return launch_stage_p<STAGE, 1>(R, ncols, threads, smem, st);
		case 2: return launch_stage_p<STAGE, 2>(R, ncols, threads, smem, st);
		case 4: return launch_stage_p<STAGE, 4>(R, ncols, threads, smem, st);
		default: return launch_stage_p<STAGE, 8>(R, ncols, threads, smem, st);
	}
}

template<int STAGE, int P> static int launch_adj_p(const ResampArgs &R, int ncols, int threads, size_t smem, cudaStream_t st)
{
	if (smem > 48*1024) B2_CHECK(cudaFuncSetAttribute(k_resamp_adj<STAGE, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_resamp_adj<STAGE, P><<<dim3(P, ncols), threads, smem, st>>>(R);
	B2_LAUNCH_CHECK();
	return 0;
}
template<int STAGE> static int launch_adj(const ResampArgs &R, int ncols, int threads, size_t smem, cudaStream_t st)
{
	switch (R.P) {
		case 1: return launch_adj_p<STAGE, 1>(R, ncols, threads, smem, st);
		case 2: return launch_adj_p<STAGE, 2>(R, ncols, threads, smem, st);
		case 4: return launch_adj_p<STAGE, 4>(R, ncols, threads, smem, st);
		default: return launch_adj_p<STAGE, 8>(R, ncols, threads, smem, st);
	}
}

int ThetaResampler::apply(double2 *leg, int ncomp, int spin, cudaStream_t st, bool adjoint)
{
	if (adjoint && C.n == 0 && C.alloc((size_t)cb*N)) return 1;
	ResampArgs R;
	R.d = tab.d; R.twoff = twoff; R.M0 = M0; R.N = N; R.P = P; R.o2 = o2; R.n = n; R.L = lmax; R.nm = nm; R.npc = npc; R.spin = spin;
	R.nring_pad = nring_pad; R.src = src.p; R.pos = dpos.p; R.mir = dmir.p; R.wfine = wfine.p; R.mult = mult.p;
	R.leg = leg; R.A = A.p; R.B = B.p; R.C = C.p;
	int64_t ncol = (int64_t)ncomp*npc;
	for (int64_t c0 = 0; c0 < ncol; c0 += cb) {
		int nc = (int)std::min<int64_t>(cb, ncol - c0);
		R.col0 = c0;
		if (adjoint) {
			if (launch_adj<1>(R, nc, threads, smem, st)) return 1;
			if (launch_adj<2>(R, nc, threads, smem, st)) return 1;
			if (launch_adj<3>(R, nc, threads, smem, st)) return 1;
			if (launch_adj<4>(R, nc, threads, smem, st)) return 1;
			if (launch_adj<5>(R, nc, threads, smem, st)) return 1;
			k_resamp_adj_gather<<<nc, 256, 0, st>>>(R);
			B2_LAUNCH_CHECK();
			continue;
		}
		if (launch_stage<1>(R, nc, threads, smem, st)) return 1;
		if (launch_stage<2>(R, nc, threads, smem, st)) return 1;
		if (launch_stage<3>(R, nc, threads, smem, st)) return 1;
		if (launch_stage<4>(R, nc, threads, smem, st)) return 1;
		if (launch_stage<5>(R, nc, threads, smem, st)) return 1;
		k_resamp_split<<<nc, 256, 0, st>>>(R);
		B2_LAUNCH_CHECK();
	}
	return 0;
}
