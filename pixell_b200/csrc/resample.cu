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

struct ResampArgs {
	FftDesc d;
	int N, P, o2, n, L, nm, npc, spin, twoff;
	int64_t nring_pad;
	const int *src, *pos, *mir; const double *wfine; const double *mult;
	double2 *leg, *A, *B;
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

// sample j of z = ext(col a) + ext(col b) on the circle
__device__ __forceinline__ double2 ext_load(const ResampArgs &R, const double2 *ca, const double2 *cb, int j, double sigma)
{
	int sidx = R.src[j];
	if (sidx < 0) return make_double2(0, 0);
	int r = sidx & 0x3fffffff;
	double2 a = ca[r], b = cb ? cb[r] : make_double2(0, 0);
	if (sidx & 0x40000000) return make_double2(sigma*(a.x - b.x), sigma*(a.y - b.y));
	return make_double2(a.x + b.x, a.y + b.y);
}

template<int STAGE> __global__ void k_resamp(ResampArgs R)
{
	extern __shared__ __align__(16) double2 s[];
	constexpr bool INV = (STAGE == 2 || STAGE == 5);
	const int tid = threadIdx.x, T = blockDim.x, c = blockIdx.x, p = blockIdx.y;
	const int N = R.N, P = R.P, Nl = N/P;
	double2 *ca, *cb; double sigma;
	pair_cols(R, R.col0 + c, ca, cb, sigma);
	double2 *A = R.A + (int64_t)c*N, *B = R.B + (int64_t)c*N;
	const double2 *twsm = s + R.twoff;
	fft_load_tw(s + R.twoff, R.d, tid, T);
	for (int j = tid; j < Nl; j += T) {
		double2 acc = make_double2(0, 0);
		for (int q = 0; q < P; q++) {
			int idx = j + q*Nl;
			double2 v;
			if (STAGE == 1) v = ext_load(R, ca, cb, idx, sigma);
			else if (STAGE == 3) v = cscale(ext_load(R, ca, cb, idx, sigma), R.wfine[(2*idx + R.o2) % (2*N)]);
			else if (STAGE == 4) v = B[idx];
			else v = A[idx];
			int e = (q*p) % P;
			if (e) v = cmul(v, cj(R.d.tw[(2*N/P)*e], INV));
			acc = cadd(acc, v);
		}
		if (p) acc = cmul(acc, cj(R.d.tw[2*j*p], INV));
		s[fft_pad(R.d, j)] = acc;
	}
	__syncthreads();
	fft_smem<INV>(s, R.d, tid, T, 1, twsm);
	const double inv = 1.0/N;
	for (int kk = tid; kk < Nl; kk += T) {
		const int k = p + P*kk;
		double2 x = s[fft_pad(R.d, R.d.rev[kk])];
		const int f = (2*k <= N) ? k : k - N;
		const bool nyq = (2*k == N);
		if (STAGE == 1) {
			double2 w = R.d.tw[f >= 0 ? f : -f]; if (f >= 0) w.y = -w.y;     // e^{+i pi f/N}
			A[k] = nyq ? make_double2(0, 0) : cscale(cmul(x, w), inv);
		} else if (STAGE == 2) {
			B[k] = cscale(x, R.wfine[(2*k + R.o2 + 1) % (2*N)]);
		} else if (STAGE == 3) {
			A[k] = (abs(f) <= R.L && !nyq) ? cscale(x, inv) : make_double2(0, 0);
		} else if (STAGE == 4) {
			if (abs(f) <= R.L && !nyq) {
				double2 w = R.d.tw[f >= 0 ? f : -f]; if (f < 0) w.y = -w.y;  // e^{-i pi f/N}
				double2 b = cscale(cmul(x, w), inv), a = A[k];
				A[k] = make_double2(0.5*(a.x + b.x), 0.5*(a.y + b.y));
			}
		} else {
			B[k] = x;      // not A: with P > 1 the other CTAs of this column are still reading A
		}
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

bool ThetaResampler::needed(const std::string &g, int ntheta, int lmax)
{
	if (g == "DH" || g == "F2") return false;
	return ntheta < 2*lmax + 2;
}

int ThetaResampler::build(const std::string &g, int ntheta, int64_t nphi_, int lmax_, int mmax, int64_t nring_pad_)
{
	n = ntheta; lmax = lmax_; nm = mmax + 1; npc = (nm + 1)/2; nring_pad = nring_pad_; nphi = nphi_;
	std::vector<int> pos(n), mir(n); std::vector<double> mu(n, 2.0);
	if (g == "CC")          { N = 2*(n - 1); o2 = 0; for (int k = 0; k < n; k++) { pos[k] = k; mir[k] = (N - k) % N; } }
	else if (g == "F1")     { N = 2*n;       o2 = 1; for (int k = 0; k < n; k++) { pos[k] = k; mir[k] = N - 1 - k; } }
	else if (g == "MW")     { N = 2*n - 1;   o2 = 1; for (int k = 0; k < n; k++) { pos[k] = k; mir[k] = N - 1 - k; } }
	else if (g == "MWflip") { N = 2*n - 1;   o2 = 0; for (int k = 0; k < n; k++) { pos[k] = k; mir[k] = (N - k) % N; } }
	else { b2_set_error("theta resampling is not defined for geometry %s", g.c_str()); return 1; }
	B2_REQUIRE(N > 2*lmax, "grid %s with %d rings cannot carry lmax=%d", g.c_str(), n, lmax);
	std::vector<int> sr(N, -1);
	for (int k = 0; k < n; k++) {
		sr[pos[k]] = k;
		if (mir[k] == pos[k]) mu[k] = 1.0; else sr[mir[k]] = k | 0x40000000;
	}
	P = 1;
	while ((size_t)(FftTables::smem_len(N/P) + 2*N/FFT_TWLO + FFT_TWLO + 1)*sizeof(double2) > 210*1024) {
		int np = P*2;
		B2_REQUIRE(np <= 8 && N % np == 0, "theta transform of length %d does not fit in shared memory", N);
		P = np;
	}
	if (tab.build(N/P, 2*N)) return 1;
	twoff = (int)FftTables::smem_len(N/P);
	smem = sizeof(double2)*(size_t)(twoff + tab.twsm_len());
	threads = (int)std::min<int64_t>(512, std::max<int64_t>(64, b2_round_up(N/P/4, 32)));
	if (src.upload(sr) || mult.upload(mu) || wfine.alloc(2*(size_t)N) || dpos.upload(pos) || dmir.upload(mir)) return 1;
	k_wfine<<<(N + 128)/128, 128>>>(wfine.p, N, 4.0*M_PI/(2.0*N)/(double)nphi);
	B2_LAUNCH_CHECK();
	cb = std::max<int64_t>(1, std::min<int64_t>((int64_t)npc*2, (int64_t)(1 << 26)/N));
	if (A.alloc((size_t)cb*N) || B.alloc((size_t)cb*N)) return 1;
	B2_CHECK(cudaDeviceSynchronize());
	return 0;
}

template<int STAGE> static int launch_stage(const ResampArgs &R, int ncols, int threads, size_t smem, cudaStream_t st)
{
	if (smem > 48*1024) B2_CHECK(cudaFuncSetAttribute(k_resamp<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_resamp<STAGE><<<dim3(ncols, R.P), threads, smem, st>>>(R);
	B2_LAUNCH_CHECK();
	return 0;
}

int ThetaResampler::apply(double2 *leg, int ncomp, int spin, cudaStream_t st)
{
	ResampArgs R;
	R.d = tab.d; R.twoff = twoff; R.N = N; R.P = P; R.o2 = o2; R.n = n; R.L = lmax; R.nm = nm; R.npc = npc; R.spin = spin;
	R.nring_pad = nring_pad; R.src = src.p; R.pos = dpos.p; R.mir = dmir.p; R.wfine = wfine.p; R.mult = mult.p;
	R.leg = leg; R.A = A.p; R.B = B.p;
	int64_t ncol = (int64_t)ncomp*npc;
	for (int64_t c0 = 0; c0 < ncol; c0 += cb) {
		int nc = (int)std::min<int64_t>(cb, ncol - c0);
		R.col0 = c0;
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
