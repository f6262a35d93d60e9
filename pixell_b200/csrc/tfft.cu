// tfft.cu -- K7t: the TMA-pipelined 2-D FFT (float64) behind the pixell.fft plug-in / enmap.fft
// (reference pixell/fft.py:8-113 engine objects, :133-209 fft/ifft/rfft/irfft; pixell/enmap.py:1307-1337).
//
// Every 1-D transform of length N = N1 N2 along an axis of a [batch][ny][nx] array runs as two "tile-DFT" launches:
//   sub-pass 1   tiles [N1 strided elements][W contiguous columns]: DFT over j1, times w_N^(k1 j2), stored so that
//                sub-pass 2 finds its lines (row axis: transposed inside the tile; column axis: rows permuted)
//   sub-pass 2   tiles [N2][W]: DFT over j2; output k = k1 + N1 k2 lands in natural order
// In both, the transformed dimension is the tile's slow dimension and the lanes of a warp run across the W contiguous
// columns, so every shared-memory access of the butterflies is a contiguous, conflict-free row segment and every
// global access is a TMA box with W*16-byte contiguous rows.  The intermediate array is blocked in slabs small
// enough to stay in the 126 MB L2 between the two launches, so HBM sees one read and one write per axis.
//
// One kernel (k_tfft) serves all sub-passes: persistent CTAs, one producer warp issuing cp.async.bulk.tensor loads
// into a 3-stage mbarrier ring, 8 consumer warps doing register butterflies in shared memory, TMA (tensor or 1-D
// bulk) stores from two alternating output tiles, so the load of tile i+2 and the store of tile i-1 overlap the
// butterflies of tile i.  Real transforms use the packed half-length complex transform along x; the untangling
// X_k = f(Z_k, conj Z_{Nc-k}) (r2c) / its inverse (c2r) is fused into the neighbouring column sub-pass, whose tiles
// then carry the mirrored column blocks [a, a+W) and [Nc-a-W+1, Nc-a] (plus the self-mirrored column Nc/2 once).
#include "../../include/b200sht.h"
#include "fft_smem.cuh"
#include "tma.cuh"
#include "tfft.cuh"
#include <algorithm>
#include <map>
#include <memory>
#include <stdlib.h>

#define TF_MAXFAC 6
#define TF_MAXSTAGES 6
#define TF_NCONS 256
#define TF_THREADS (TF_NCONS + 32)

struct TfMaps { CUtensorMap ld[3], st[3]; };      // regions A, B (mirror), M (self-mirrored column)

struct TfArgs {
	int n, nfac, fac[TF_MAXFAC];
	int W, wshift, nreg, LW;
	int midcol;                  // mirrored tiles: the self-mirrored column (carried by the last column block), else -1
	int ncb;                     // column blocks in the whole array
	int cb0, ncb_l;              // column blocks of this launch
	int G2, g2_0, G3, g3_0;      // tile coordinates 2 and 3: count and offset in this launch
	int mirror;                  // Nc: region B starts at column Nc - a - W + 1
	int tw_mode, inv, op, tstore;
	double scale;
	double2 *tbase; long long t_g2stride; int ncols_valid;      // transposed bulk stores (row axis, sub-pass 1)
	const double2 *twn; const int *natk; const double2 *twN_hi, *twN_lo, *twR_hi, *twR_lo; int nhiN, nhiR;
	int tile_bytes, out_bytes, off_work, off_out, off_tab, nstage;
	long long ntiles;
};

__device__ __forceinline__ double2 tw2(const double2 *hi, const double2 *lo, int e)
{
	const double2 a = hi[e >> 7], b = lo[e & 127];
	return make_double2(fma(a.x, b.x, -a.y*b.y), fma(a.x, b.y, a.y*b.x));
}

// one radix-R decimation-in-frequency pass over the tile (rows = transform index, lanes across columns): reads S, writes D
// (D == S: in place).  LAST: the rows go to their natural positions in D, times the twiddle between the two sub-passes
// and the scale.
template<int R, bool INV, bool LAST> __device__ __forceinline__ void tf_pass(const double2 *__restrict__ S, double2 *__restrict__ D, const TfArgs &A,
	const double2 *__restrict__ twn, const int *__restrict__ natk, const double2 *__restrict__ hiN, const double2 *__restrict__ loN,
	int Ls, int Wcur, int tid, int col0, int g2)
{
	const int n = A.n, m = Ls/R, nb = n/R, W = A.W, LW = A.LW, nrw = A.nreg*W;
	const int tx = tid & (LW - 1), ty = tid/LW, NTY = TF_NCONS/LW;
	const int tws = n/Ls;
	for (int b = ty; b < nb; b += NTY) {
		const int blk = b/m, jj = b - blk*m, base = blk*Ls + jj;
		for (int cc = tx; cc < Wcur; cc += LW) {
			int off, pitch;
			if (cc < nrw) { const int r = cc >> A.wshift; off = r*n*W + (cc & (W - 1)); pitch = W; }
			else { off = nrw*n; pitch = 1; }
			double2 u[R];
			const double2 *sp = S + off + base*pitch;
			const int mp = m*pitch;
			#pragma unroll
			for (int q = 0; q < R; q++) u[q] = sp[q*mp];
			dft_small<R, INV>(u);
			if (!LAST) {
				if (jj) {
					const double2 *tp = twn + tws*jj;
					#pragma unroll
					for (int k = 1; k < R; k++) { double2 w = tp[(k - 1)*tws*jj]; if (INV) w.y = -w.y; u[k] = cmul(u[k], w); }
				}
				double2 *dp = D + off + base*pitch;
				#pragma unroll
				for (int q = 0; q < R; q++) dp[q*mp] = u[q];
			} else {
				const int t = A.tw_mode == 1 ? col0 + cc : g2;
				#pragma unroll
				for (int k = 0; k < R; k++) {
					const int row = natk[base + k];
					double2 v = u[k];
					if (A.tw_mode) {
						double2 w = tw2(hiN, loN, row*t); if (INV) w.y = -w.y;
						v = cmul(v, w);
					}
					v.x *= A.scale; v.y *= A.scale;
					if (A.tstore) D[cc*(n + 1) + row] = v;
					else D[off + row*pitch] = v;
				}
			}
		}
	}
}

template<bool INV, bool LAST> __device__ __forceinline__ void tf_pass_r(int R, const double2 *S, double2 *D, const TfArgs &A,
	const double2 *twn, const int *natk, const double2 *hiN, const double2 *loN, int Ls, int Wcur, int tid, int col0, int g2)
{
	switch (R) {
		case 2:  tf_pass<2, INV, LAST>(S, D, A, twn, natk, hiN, loN, Ls, Wcur, tid, col0, g2); break;
		case 3:  tf_pass<3, INV, LAST>(S, D, A, twn, natk, hiN, loN, Ls, Wcur, tid, col0, g2); break;
		case 4:  tf_pass<4, INV, LAST>(S, D, A, twn, natk, hiN, loN, Ls, Wcur, tid, col0, g2); break;
		case 5:  tf_pass<5, INV, LAST>(S, D, A, twn, natk, hiN, loN, Ls, Wcur, tid, col0, g2); break;
		case 8:  tf_pass<8, INV, LAST>(S, D, A, twn, natk, hiN, loN, Ls, Wcur, tid, col0, g2); break;
		default: tf_pass<16, INV, LAST>(S, D, A, twn, natk, hiN, loN, Ls, Wcur, tid, col0, g2); break;
	}
}

// shared memory: [nstage landing tiles][work tile][output tile][tables]; the landing tile is released to the producer as
// soon as the first pass has moved its contents into the work tile
template<bool INV> __global__ void __launch_bounds__(TF_THREADS, 1) k_tfft(const __grid_constant__ TfMaps M, const TfArgs A)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ __align__(8) uint64_t bars[2*TF_MAXSTAGES];
	__shared__ int sfac[TF_MAXFAC];
	const int NS = A.nstage;
	uint64_t *full = bars, *empty = bars + TF_MAXSTAGES;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int n = A.n, W = A.W;
	double2 *twn = (double2*)(smem + A.off_tab);
	double2 *hiN = twn + n, *loN = hiN + A.nhiN, *hiR = loN + 128, *loR = hiR + A.nhiR;
	int *natk = (int*)(loR + 128);
	if (tid == 0) {
		for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
		mbar_fence_init();
	}
	if (tid < TF_MAXFAC) sfac[tid] = A.fac[tid];
	for (int i = tid; i < n; i += TF_THREADS) { twn[i] = A.twn[i]; natk[i] = A.natk[i]; }
	for (int i = tid; i < 128; i += TF_THREADS) { loN[i] = A.tw_mode ? A.twN_lo[i] : make_double2(1, 0); loR[i] = A.op ? A.twR_lo[i] : make_double2(1, 0); }
	for (int i = tid; i < A.nhiN; i += TF_THREADS) hiN[i] = A.twN_hi[i];
	for (int i = tid; i < A.nhiR; i += TF_THREADS) hiR[i] = A.twR_hi[i];
	__syncthreads();
	const long long tiles_per_g = (long long)A.ncb_l*A.G2;
	const int region_bytes = n*W*16;

	if (warp == TF_NCONS/32) {
		// ---------------- producer: one lane keeps the ring of landing tiles full
		if (lane == 0) {
			tma_prefetch_desc(&M.ld[0]); tma_prefetch_desc(&M.st[0]);
			int s = 0, round = 0;
			for (long long t = blockIdx.x; t < A.ntiles; t += gridDim.x) {
				if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
				const int g3 = A.g3_0 + (int)(t/tiles_per_g); const long long r = t % tiles_per_g;
				const int g2 = A.g2_0 + (int)(r/A.ncb_l), cb = A.cb0 + (int)(r % A.ncb_l);
				const bool mid = A.midcol >= 0 && cb == A.ncb - 1;
				unsigned char *dst = smem + (size_t)s*A.tile_bytes;
				mbar_expect_tx(&full[s], (uint32_t)(A.nreg*region_bytes + (mid ? n*16 : 0)));
				const int a = cb*W;
				tma_load_4d(dst, &M.ld[0], &full[s], 2*a, 0, g2, g3);
				if (A.nreg == 2) tma_load_4d(dst + region_bytes, &M.ld[1], &full[s], 2*(A.mirror - a - W + 1), 0, g2, g3);
				if (mid) tma_load_4d(dst + 2*region_bytes, &M.ld[2], &full[s], 2*A.midcol, 0, g2, g3);
				if (++s == NS) { s = 0; round++; }
			}
		}
		return;
	}

	// ---------------- consumers
	double2 *Wk = (double2*)(smem + A.off_work), *O = (double2*)(smem + A.off_out);
	const bool storer = A.tstore ? warp == 0 : tid == 0;
	const int nfac = A.nfac;
	int s = 0, round = 0;
	for (long long t = blockIdx.x; t < A.ntiles; t += gridDim.x) {
		const int g3 = A.g3_0 + (int)(t/tiles_per_g); const long long r = t % tiles_per_g;
		const int g2 = A.g2_0 + (int)(r/A.ncb_l), cb = A.cb0 + (int)(r % A.ncb_l);
		const bool mid = A.midcol >= 0 && cb == A.ncb - 1;
		const int a = cb*W, Wcur = A.nreg*W + (mid ? 1 : 0);
		double2 *S = (double2*)(smem + (size_t)s*A.tile_bytes);
		mbar_wait(&full[s], round & 1);
		if (A.op == 1) {
			// r2c: packed spectrum Z -> X on the mirrored column blocks (same row), in place in the landing tile
			double2 *SA = S, *SB = S + n*W;
			for (int idx = tid; idx < n*W; idx += TF_NCONS) {
				const int j = idx >> A.wshift, i = idx & (W - 1), k = a + i;
				const double2 zk = SA[j*W + i];
				double2 zp = SB[j*W + (W - 1 - i)];
				if (k == 0) zp = zk;                                     // Z_Nc = Z_0 (the box column beyond the array was zero-filled)
				const double2 sm = make_double2(zk.x + zp.x, zk.y - zp.y), df = make_double2(zk.x - zp.x, zk.y + zp.y);
				const double2 u = cmul(df, tw2(hiR, loR, k));            // w_nx^k (Z_k - conj Z_{Nc-k})
				SA[j*W + i] = make_double2(sm.x + u.y, sm.y - u.x);      // 2 X_k: the factor 1/2 rides on the row pass
				SB[j*W + (W - 1 - i)] = make_double2(sm.x - u.y, -(sm.y + u.x));
			}
			if (mid) for (int j = tid; j < n; j += TF_NCONS) { double2 *p = S + 2*n*W + j; *p = make_double2(2*p->x, -2*p->y); }
			fence_proxy_async();                                         // the landing tile was written by the generic proxy
			bar_sync(1, TF_NCONS);
		}
		int Ls = n;
		if (nfac > 1) {
			tf_pass_r<INV, false>(sfac[0], S, Wk, A, twn, natk, hiN, loN, Ls, Wcur, tid, a, g2);
			Ls /= sfac[0];
			for (int f = 1; f < nfac - 1; f++) {
				bar_sync(1, TF_NCONS);
				if (f == 1 && tid == 0) mbar_arrive(&empty[s]);          // every consumer is past the first pass: landing tile free
				tf_pass_r<INV, false>(sfac[f], Wk, Wk, A, twn, natk, hiN, loN, Ls, Wcur, tid, a, g2);
				Ls /= sfac[f];
			}
			if (storer) bulk_wait_read<0>();                             // the output tile's previous contents have left
			bar_sync(1, TF_NCONS);
			if (nfac == 2 && tid == 0) mbar_arrive(&empty[s]);
			tf_pass_r<INV, true>(sfac[nfac - 1], Wk, O, A, twn, natk, hiN, loN, Ls, Wcur, tid, a, g2);
		} else {
			if (storer) bulk_wait_read<0>();
			bar_sync(1, TF_NCONS);
			tf_pass_r<INV, true>(sfac[0], S, O, A, twn, natk, hiN, loN, Ls, Wcur, tid, a, g2);
		}
		if (A.op == 2) {
			// c2r: X -> packed spectrum Z on the mirrored column blocks of the output tile
			bar_sync(1, TF_NCONS);
			double2 *OA = O, *OB = O + n*W;
			for (int idx = tid; idx < n*W; idx += TF_NCONS) {
				const int j = idx >> A.wshift, i = idx & (W - 1), k = a + i;
				const double2 xa = OA[j*W + i], xb = OB[j*W + (W - 1 - i)];
				const double2 sm = make_double2(xa.x + xb.x, xa.y - xb.y), df = make_double2(xa.x - xb.x, xa.y + xb.y);
				double2 w = tw2(hiR, loR, k); w.y = -w.y;                // conj(w_nx^k)
				const double2 u = cmul(df, w);
				OA[j*W + i] = make_double2(sm.x - u.y, sm.y + u.x);
				OB[j*W + (W - 1 - i)] = make_double2(sm.x + u.y, u.x - sm.y);
			}
			if (mid) for (int j = tid; j < n; j += TF_NCONS) { double2 *p = O + 2*n*W + j; *p = make_double2(2*p->x, -2*p->y); }
		}
		fence_proxy_async();
		bar_sync(1, TF_NCONS);
		if (nfac == 1 && tid == 0) mbar_arrive(&empty[s]);
		if (A.tstore) {
			if (warp == 0) {
				for (int c = lane; c < W; c += 32)
					if (a + c < A.ncols_valid) bulk_store_1d(A.tbase + (long long)g2*A.t_g2stride + (long long)(a + c)*n, O + c*(n + 1), (uint32_t)n*16);
				bulk_commit();
			}
		} else if (tid == 0) {
			tma_store_4d(&M.st[0], O, 2*a, 0, g2, g3);
			if (A.nreg == 2) tma_store_4d(&M.st[1], O + n*W, 2*(A.mirror - a - W + 1), 0, g2, g3);
			if (mid) tma_store_4d(&M.st[2], O + 2*n*W, 2*A.midcol, 0, g2, g3);
			bulk_commit();
		}
		if (++s == NS) { s = 0; round++; }
	}
	if (storer) bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------ specialised kernel
// Power-of-two tile lengths N = R0 R1 with compile-time strides.  Two teams of four warps work on alternate tiles (each
// with its own output tile), so one team's barrier and mbarrier waits are covered by the other team's butterflies.
// Pass 0 reads the landing tile and writes the output tile in transposed digit order (row jj R0 + k0), which frees the
// landing tile at once; pass 1 then runs in place on rows {q R0 + k0} and leaves row k0 + R0 k1 in natural order.

template<int N, int W, int NREG, bool TS> __device__ __forceinline__ int tf2_addr(int c, int row)
{
	if (TS) return c*(N + 1) + row;
	if (NREG == 1) return row*W + c;
	return (c/W)*N*W + row*W + (c % W);
}

// Twiddles come from shared-memory tables, not from products: the FP64 pipe (64 lanes per SM and clock) is the scarce
// unit of these kernels.  twn: w_N^j.  TT (row sub-pass 1, TS): scale w_Ntot^(row (a + c)) for the whole tile, the same
// for every tile of a CTA (all its tiles share the column block).  V (column sub-pass 1): scale w_Ntot^(row g2), one
// vector per tile.
template<bool INV, int N, int R0, int R1, int W, int NREG, bool TS, int NT>
__device__ __forceinline__ void tf2_tile(const double2 *__restrict__ S, double2 *__restrict__ O, const TfArgs &A, const double2 *__restrict__ twn,
	const double2 *__restrict__ TT, const double2 *__restrict__ V, int ttid, int team, bool mid, uint64_t *empty_bar)
{
	constexpr int WT = NREG*W, TF_TEAM = TF_NCONS/NT;
	// ---- pass 0: radix R0 over rows jj + q R1, twiddle w_N^(jj k), to rows jj R0 + k of the output tile
	#pragma unroll 1
	for (int item = ttid; item < R1*WT; item += TF_TEAM) {
		const int c = item % WT, jj = item / WT;
		double2 u[R0];
		const double2 *sp = S + (NREG == 1 ? c : (c/W)*N*W + (c % W)) + jj*W;
		#pragma unroll
		for (int q = 0; q < R0; q++) u[q] = sp[q*R1*W];
		dft_small<R0, INV>(u);
		if (jj) {
			#pragma unroll
			for (int k = 1; k < R0; k++) { double2 w = twn[jj*k]; if (INV) w.y = -w.y; u[k] = cmul(u[k], w); }
		}
		double2 *dp = O + tf2_addr<N, W, NREG, TS>(c, jj*R0);
		#pragma unroll
		for (int k = 0; k < R0; k++) dp[TS ? k : k*W] = u[k];
	}
	if (mid) {
		// the self-mirrored column (pitch 1, after the two regions)
		for (int jj = ttid; jj < R1; jj += TF_TEAM) {
			double2 u[R0];
			const double2 *sp = S + 2*N*W + jj;
			#pragma unroll
			for (int q = 0; q < R0; q++) u[q] = sp[q*R1];
			dft_small<R0, INV>(u);
			if (jj) {
				#pragma unroll
				for (int k = 1; k < R0; k++) { double2 w = twn[jj*k]; if (INV) w.y = -w.y; u[k] = cmul(u[k], w); }
			}
			double2 *dp = O + 2*N*W + jj*R0;
			#pragma unroll
			for (int k = 0; k < R0; k++) dp[k] = u[k];
		}
	}
	bar_sync(1 + team, TF_TEAM);
	if (ttid == 0) mbar_arrive(empty_bar);                       // the landing tile is free
	// ---- pass 1: radix R1 in place over rows q R0 + k0; twiddle between the sub-passes and scale from the tables
	const double sc = A.scale;
	#pragma unroll 1
	for (int item = ttid; item < R0*WT; item += TF_TEAM) {
		const int c = item % WT, k0 = item / WT;
		double2 u[R1];
		const int ad = tf2_addr<N, W, NREG, TS>(c, k0);
		double2 *dp = O + ad;
		#pragma unroll
		for (int q = 0; q < R1; q++) u[q] = dp[(TS ? 1 : W)*q*R0];
		dft_small<R1, INV>(u);
		if (TS) {
			#pragma unroll
			for (int k = 0; k < R1; k++) dp[k*R0] = cmul(u[k], TT[ad + k*R0]);
		} else if (A.tw_mode == 2) {
			#pragma unroll
			for (int k = 0; k < R1; k++) dp[W*k*R0] = cmul(u[k], V[k0 + k*R0]);
		} else if (sc != 1.0) {
			#pragma unroll
			for (int k = 0; k < R1; k++) dp[W*k*R0] = make_double2(u[k].x*sc, u[k].y*sc);
		} else {
			#pragma unroll
			for (int k = 0; k < R1; k++) dp[W*k*R0] = u[k];
		}
	}
	if (mid) {
		for (int k0 = ttid; k0 < R0; k0 += TF_TEAM) {
			double2 u[R1];
			double2 *dp = O + 2*N*W + k0;
			#pragma unroll
			for (int q = 0; q < R1; q++) u[q] = dp[q*R0];
			dft_small<R1, INV>(u);
			#pragma unroll
			for (int k = 0; k < R1; k++) dp[k*R0] = A.tw_mode == 2 ? cmul(u[k], V[k0 + k*R0]) : make_double2(u[k].x*sc, u[k].y*sc);
		}
	}
}

// Tile order.  TS kernels (row sub-pass 1): CTA b owns column block cb0 + b % ncb_l and walks the lines b / ncb_l,
// + gridDim.x / ncb_l, ... (the grid is a multiple of ncb_l), so its twiddle tile never changes.  Others: tile
// t = blockIdx.x + i gridDim.x with the column block fastest.
template<bool INV, int N, int R0, int R1, int W, int NREG, bool TS, int NT>
__global__ void __launch_bounds__(TF_THREADS, 1) k_tfft2(const __grid_constant__ TfMaps M, const TfArgs A)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ __align__(8) uint64_t bars[2*TF_MAXSTAGES];
	const int NS = A.nstage;
	uint64_t *full = bars, *empty = bars + TF_MAXSTAGES;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	double2 *twn = (double2*)(smem + A.off_tab);
	double2 *hiN = twn + N, *loN = hiN + A.nhiN, *hiR = loN + 128, *loR = hiR + A.nhiR;
	double2 *Vbase = loR + 128;                                  // two vectors of N (one per team)
	double2 *TT = (double2*)(smem + A.off_work);                 // TS: the twiddle tile
	if (tid == 0) {
		for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
		mbar_fence_init();
	}
	for (int i = tid; i < N; i += TF_THREADS) twn[i] = A.twn[i];
	for (int i = tid; i < 128; i += TF_THREADS) { loN[i] = A.tw_mode ? A.twN_lo[i] : make_double2(1, 0); loR[i] = A.op ? A.twR_lo[i] : make_double2(1, 0); }
	for (int i = tid; i < A.nhiN; i += TF_THREADS) hiN[i] = A.twN_hi[i];
	for (int i = tid; i < A.nhiR; i += TF_THREADS) hiR[i] = A.twR_hi[i];
	__syncthreads();
	const long long tiles_per_g = (long long)A.ncb_l*A.G2;
	constexpr int region_bytes = N*W*16;
	// tile walk of this CTA
	long long t_first, t_step, t_count;
	int cb_fixed = 0;
	if (TS) {
		const int per = gridDim.x/A.ncb_l;                       // CTAs per column block
		cb_fixed = A.cb0 + (int)(blockIdx.x % A.ncb_l);
		t_first = blockIdx.x/A.ncb_l; t_step = per;
		const long long nl = (long long)A.G2*A.G3;
		t_count = t_first < nl ? (nl - t_first + per - 1)/per : 0;
		// the twiddle tile of this column block: scale w^(row (a + c)), in the layout of the output tile
		const int a = cb_fixed*W;
		for (int idx = tid; idx < N*W; idx += TF_THREADS) {
			const int c = idx / N, row = idx % N;
			double2 w = tw2(hiN, loN, row*(a + c)); if (INV) w.y = -w.y;
			TT[c*(N + 1) + row] = make_double2(w.x*A.scale, w.y*A.scale);
		}
		__syncthreads();
	} else {
		t_first = blockIdx.x; t_step = gridDim.x;
		t_count = t_first < A.ntiles ? (A.ntiles - t_first + t_step - 1)/t_step : 0;
	}

	if (warp == TF_NCONS/32) {
		if (lane == 0) {
			tma_prefetch_desc(&M.ld[0]); tma_prefetch_desc(&M.st[0]);
			int s = 0, round = 0;
			for (long long i = 0; i < t_count; i++) {
				const long long t = t_first + i*t_step;
				if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
				int g2, g3, cb;
				if (TS) { g3 = A.g3_0 + (int)(t/A.G2); g2 = A.g2_0 + (int)(t % A.G2); cb = cb_fixed; }
				else { g3 = A.g3_0 + (int)(t/tiles_per_g); const long long r = t % tiles_per_g; g2 = A.g2_0 + (int)(r/A.ncb_l); cb = A.cb0 + (int)(r % A.ncb_l); }
				const bool mid = NREG == 2 && A.midcol >= 0 && cb == A.ncb - 1;
				unsigned char *dst = smem + (size_t)s*A.tile_bytes;
				mbar_expect_tx(&full[s], (uint32_t)(NREG*region_bytes + (mid ? N*16 : 0)));
				const int a = cb*W;
				tma_load_4d(dst, &M.ld[0], &full[s], 2*a, 0, g2, g3);
				if (NREG == 2) tma_load_4d(dst + region_bytes, &M.ld[1], &full[s], 2*(A.mirror - a - W + 1), 0, g2, g3);
				if (mid) tma_load_4d(dst + 2*region_bytes, &M.ld[2], &full[s], 2*A.midcol, 0, g2, g3);
				if (++s == NS) { s = 0; round++; }
			}
		}
		return;
	}

	constexpr int TF_TEAM = TF_NCONS/NT;
	const int team = NT == 2 ? warp >> 2 : 0, ttid = tid & (TF_TEAM - 1), twarp = NT == 2 ? warp & 3 : warp;
	double2 *O = (double2*)(smem + A.off_out + (size_t)team*A.out_bytes);
	double2 *V = Vbase + team*N;
	const bool storer = TS ? twarp == 0 : ttid == 0;
	for (long long i = team; i < t_count; i += NT) {
		const long long t = t_first + i*t_step;
		const int s = (int)(i % NS), round = (int)(i/NS);
		int g2, g3, cb;
		if (TS) { g3 = A.g3_0 + (int)(t/A.G2); g2 = A.g2_0 + (int)(t % A.G2); cb = cb_fixed; }
		else { g3 = A.g3_0 + (int)(t/tiles_per_g); const long long r = t % tiles_per_g; g2 = A.g2_0 + (int)(r/A.ncb_l); cb = A.cb0 + (int)(r % A.ncb_l); }
		const bool mid = NREG == 2 && A.midcol >= 0 && cb == A.ncb - 1;
		const int a = cb*W;
		double2 *S = (double2*)(smem + (size_t)s*A.tile_bytes);
		if (storer) bulk_wait_read<0>();                         // this team's output tile has left
		if (!TS && A.tw_mode == 2) {
			for (int row = ttid; row < N; row += TF_TEAM) {
				double2 w = tw2(hiN, loN, row*g2); if (INV) w.y = -w.y;
				V[row] = make_double2(w.x*A.scale, w.y*A.scale);
			}
		}
		mbar_wait(&full[s], round & 1);
		if (NREG == 2 && A.op == 1) {
			// r2c: packed spectrum Z -> 2 X on the mirrored column blocks (the factor 1/2 is folded into the scale of the row pass)
			double2 *SA = S, *SB = S + N*W;
			for (int idx = ttid; idx < N*W; idx += TF_TEAM) {
				const int j = idx / W, i2 = idx % W, k = a + i2;
				const double2 zk = SA[j*W + i2];
				double2 zp = SB[j*W + (W - 1 - i2)];
				if (k == 0) zp = zk;
				const double2 sm = make_double2(zk.x + zp.x, zk.y - zp.y), df = make_double2(zk.x - zp.x, zk.y + zp.y);
				const double2 u = cmul(df, tw2(hiR, loR, k));
				SA[j*W + i2] = make_double2(sm.x + u.y, sm.y - u.x);
				SB[j*W + (W - 1 - i2)] = make_double2(sm.x - u.y, -(sm.y + u.x));
			}
			if (mid) for (int j = ttid; j < N; j += TF_TEAM) { double2 *p = S + 2*N*W + j; *p = make_double2(2*p->x, -2*p->y); }
			fence_proxy_async();
		}
		bar_sync(1 + team, TF_TEAM);
		tf2_tile<INV, N, R0, R1, W, NREG, TS, NT>(S, O, A, twn, TT, V, ttid, team, mid, &empty[s]);
		if (NREG == 2 && A.op == 2) {
			bar_sync(1 + team, TF_TEAM);
			double2 *OA = O, *OB = O + N*W;
			for (int idx = ttid; idx < N*W; idx += TF_TEAM) {
				const int j = idx / W, i2 = idx % W, k = a + i2;
				const double2 xa = OA[j*W + i2], xb = OB[j*W + (W - 1 - i2)];
				const double2 sm = make_double2(xa.x + xb.x, xa.y - xb.y), df = make_double2(xa.x - xb.x, xa.y + xb.y);
				double2 w = tw2(hiR, loR, k); w.y = -w.y;
				const double2 u = cmul(df, w);
				OA[j*W + i2] = make_double2(sm.x - u.y, sm.y + u.x);
				OB[j*W + (W - 1 - i2)] = make_double2(sm.x + u.y, u.x - sm.y);
			}
			if (mid) for (int j = ttid; j < N; j += TF_TEAM) { double2 *p = O + 2*N*W + j; *p = make_double2(2*p->x, -2*p->y); }
		}
		fence_proxy_async();
		bar_sync(1 + team, TF_TEAM);
		if (TS) {
			if (twarp == 0) {
				for (int c = lane; c < W; c += 32)
					if (a + c < A.ncols_valid) bulk_store_1d(A.tbase + (long long)g2*A.t_g2stride + (long long)(a + c)*N, O + c*(N + 1), (uint32_t)N*16);
				bulk_commit();
			}
		} else if (ttid == 0) {
			tma_store_4d(&M.st[0], O, 2*a, 0, g2, g3);
			if (NREG == 2) tma_store_4d(&M.st[1], O + N*W, 2*(A.mirror - a - W + 1), 0, g2, g3);
			if (mid) tma_store_4d(&M.st[2], O + 2*N*W, 2*A.midcol, 0, g2, g3);
			bulk_commit();
		}
	}
	if (storer) bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------ host: tables

typedef long double ld_t;
static const ld_t TF_TAU = 6.283185307179586476925286766559005768L;

static bool tf_factor(int n, int *fac, int &nfac)
{
	// 5s and 3s first, then a leftover 2 / 4, then 8s and 16s (the last radix has unit stride)
	nfac = 0;
	int rem = n, a = 0;
	while (rem % 5 == 0) { if (nfac >= TF_MAXFAC) return false; fac[nfac++] = 5; rem /= 5; }
	while (rem % 3 == 0) { if (nfac >= TF_MAXFAC) return false; fac[nfac++] = 3; rem /= 3; }
	while (rem % 2 == 0) { a++; rem /= 2; }
	if (rem != 1) return false;
	int f2[8], n2 = 0;
	if (a == 1) f2[n2++] = 2;
	else if (a == 2) f2[n2++] = 4;
	else if (a == 3) f2[n2++] = 8;
	else if (a == 4) f2[n2++] = 16;
	else if (a == 5) { f2[n2++] = 4; f2[n2++] = 8; }
	else if (a == 6) { f2[n2++] = 8; f2[n2++] = 8; }
	else if (a == 7) { f2[n2++] = 16; f2[n2++] = 8; }
	else if (a == 8) { f2[n2++] = 16; f2[n2++] = 16; }
	else if (a > 8) return false;
	for (int i = 0; i < n2; i++) { if (nfac >= TF_MAXFAC) return false; fac[nfac++] = f2[i]; }
	return nfac >= 1;
}

struct TfLen {      // tables of one in-tile transform length
	int n = 0, nfac = 0, fac[TF_MAXFAC];
	DevBuf<double2> twn; DevBuf<int> natk;
	int build(int n_) {
		n = n_;
		if (!tf_factor(n, fac, nfac)) { b2_set_error("tfft: length %d has no radix plan", n); return 1; }
		std::vector<double2> tw(n); std::vector<int> nk(n);
		for (int k = 0; k < n; k++) { ld_t a = TF_TAU*(ld_t)k/(ld_t)n; tw[k] = make_double2((double)cosl(a), (double)-sinl(a)); }
		for (int k = 0; k < n; k++) {
			int kk = k, pos = 0, len = n;
			for (int f = 0; f < nfac; f++) { int r = fac[f]; len /= r; pos += (kk % r)*len; kk /= r; }
			nk[pos] = k;
		}
		return twn.upload(tw) || natk.upload(nk);
	}
};

struct TfTw2 {      // two-level table of exp(-2 pi i e / mod), e < count
	int nhi = 0; DevBuf<double2> hi, lo;
	int build(long long mod, long long count) {
		nhi = (int)((count + 127)/128) + 1;
		std::vector<double2> h(nhi), l(128);
		for (int i = 0; i < nhi; i++) { ld_t a = TF_TAU*(ld_t)((128LL*i) % mod)/(ld_t)mod; h[i] = make_double2((double)cosl(a), (double)-sinl(a)); }
		for (int i = 0; i < 128; i++) { ld_t a = TF_TAU*(ld_t)(i % mod)/(ld_t)mod; l[i] = make_double2((double)cosl(a), (double)-sinl(a)); }
		return hi.upload(h) || lo.upload(l);
	}
};

struct TfAxis {     // a transform length N = N1 N2 along one axis
	int N = 0, N1 = 0, N2 = 0;
	TfLen L1, L2; TfTw2 tw;
	int build(int N_) {
		N = N_;
		// balanced split into two tile-sized factors with radix plans
		int best = 0;
		for (int a = 1; a <= 256 && a <= N; a++) {
			if (N % a) continue;
			int b = N/a;
			if (a > 256 || b > 256 || a < 4 || b < 4) continue;
			int fa[TF_MAXFAC], fb[TF_MAXFAC], na, nb;
			if (!tf_factor(a, fa, na) || !tf_factor(b, fb, nb)) continue;
			if (!best || std::max(a, b) < std::max(best, N/best) || (std::max(a, b) == std::max(best, N/best) && a > best)) best = a;
		}
		if (!best) { b2_set_error("tfft: no split of %d", N); return 1; }
		N1 = best; N2 = N/best;
		return L1.build(N1) || L2.build(N2) || tw.build(N, (long long)N1*N2);
	}
	static bool ok(int N) {
		for (int a = 4; a <= 256 && a <= N; a++) {
			if (N % a) continue;
			int b = N/a, fa[TF_MAXFAC], fb[TF_MAXFAC], na, nb;
			if (b > 256 || b < 4) continue;
			if (tf_factor(a, fa, na) && tf_factor(b, fb, nb)) return true;
		}
		return false;
	}
};

struct TfPlan {
	int kind = 0;
	int64_t nb = 1, ny = 0, nx = 0;      // nx: real length for r2c / c2r
	int Nc = 0;                           // complex length along x
	int64_t in_pitch = 0, out_pitch = 0, in_bstride = 0, out_bstride = 0;      // elements of the respective type
	TfAxis ax, ay; TfTw2 twR;
	DevBuf<char> work, work2;
	int nsm = 148;
};

static int pow2_floor(int v) { int p = 1; while (2*p <= v) p *= 2; return p; }
// Slabs: run the two sub-passes of an axis on blocks of lines small enough for the intermediate to stay in L2.  Measured
// on B200 (profiles/r2*_tfft_*): the launches this takes cost more than the HBM traffic they save, so the default is one
// slab (every sub-pass streams the whole array); B2_TFFT_SLAB_MB sets a slab size for experiments.
static size_t tf_slab_bytes() { const char *e = getenv("B2_TFFT_SLAB_MB"); return e ? (size_t)atoi(e) << 20 : (size_t)1 << 60; }
static int tf_tile_elems() { const char *e = getenv("B2_TFFT_TILE"); return e ? atoi(e) : 2048; }

// columns per tile for a transform length n with `cols` columns available
static int tf_width(int n, int64_t cols)
{
	int w = pow2_floor(std::max(8, tf_tile_elems()/n));
	w = std::min(w, 128);
	while (w > 8 && w > cols) w /= 2;
	return w;
}

bool tfft_eligible(int kind, int dtype, int ndim, const int64_t *shape, const int64_t *istride, const int64_t *ostride, int naxes, const int *axes)
{
	static const bool off = getenv("B2_FFT_NO_TMA") && atoi(getenv("B2_FFT_NO_TMA"));
	if (off || dtype != B2_F64 || naxes != 2 || ndim < 2 || ndim > 3) return false;
	if (axes[0] != ndim - 2 || axes[1] != ndim - 1) return false;
	if (!b2_get_encode_tiled()) return false;
	const int64_t ny = shape[ndim - 2], nx = shape[ndim - 1];
	if (istride[ndim - 1] != 1 || ostride[ndim - 1] != 1) return false;
	const int64_t Nc = kind == B2_FFT_C2C ? nx : nx/2;
	if (kind != B2_FFT_C2C && (nx % 2 || (Nc/2) % 8 || Nc % 2)) return false;
	if (ny < 64 || Nc < 64 || ny > 65536 || Nc > 65536) return false;
	if (!TfAxis::ok((int)ny) || !TfAxis::ok((int)Nc)) return false;
	// real rows must start on 16-byte boundaries
	if (kind == B2_FFT_R2C && (istride[ndim - 2] % 2 || (ndim == 3 && istride[0] % 2))) return false;
	if (kind == B2_FFT_C2R && (ostride[ndim - 2] % 2 || (ndim == 3 && ostride[0] % 2))) return false;
	const int64_t in_cols = kind == B2_FFT_C2R ? Nc + 1 : nx, out_cols = kind == B2_FFT_R2C ? Nc + 1 : nx;
	if (istride[ndim - 2] < in_cols || ostride[ndim - 2] < out_cols) return false;
	return true;
}

int tfft_plan_create(TfPlan **out, int kind, int ndim, const int64_t *shape, const int64_t *istride, const int64_t *ostride)
{
	std::unique_ptr<TfPlan> p(new TfPlan());
	p->kind = kind; p->ny = shape[ndim - 2]; p->nx = shape[ndim - 1];
	p->nb = ndim == 3 ? shape[0] : 1;
	p->Nc = (int)(kind == B2_FFT_C2C ? p->nx : p->nx/2);
	p->in_pitch = istride[ndim - 2]; p->out_pitch = ostride[ndim - 2];
	p->in_bstride = ndim == 3 ? istride[0] : 0; p->out_bstride = ndim == 3 ? ostride[0] : 0;
	if (p->ax.build(p->Nc) || p->ay.build((int)p->ny)) return 1;
	if (kind != B2_FFT_C2C && p->twR.build(p->nx, p->Nc + 1)) return 1;
	int dev; B2_CHECK(cudaGetDevice(&dev));
	B2_CHECK(cudaDeviceGetAttribute(&p->nsm, cudaDevAttrMultiProcessorCount, dev));
	*out = p.release();
	return 0;
}
void tfft_plan_destroy(TfPlan *p) { delete p; }

// ------------------------------------------------------------------------------------ host: launches

template<bool INV, int N, int R0, int R1, int W, int NREG, bool TS, int NT> static int tf_launch2_k(const TfMaps &M, const TfArgs &A, unsigned grid, size_t smem, cudaStream_t st)
{
	static thread_local std::map<int, size_t> granted;
	int dev; B2_CHECK(cudaGetDevice(&dev));
	size_t &g = granted[dev];
	if (smem > g) { B2_CHECK(cudaFuncSetAttribute(k_tfft2<INV, N, R0, R1, W, NREG, TS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); g = smem; }
	k_tfft2<INV, N, R0, R1, W, NREG, TS, NT><<<grid, TF_THREADS, smem, st>>>(M, A);
	B2_LAUNCH_CHECK();
	static const bool dbg = getenv("B2_TFFT_SYNC") && atoi(getenv("B2_TFFT_SYNC"));
	if (dbg) {
		cudaError_t e = cudaStreamSynchronize(st);
		B2_REQUIRE(e == cudaSuccess, "k_tfft2<%d,%d,%d,%d,%d,%d,%d> grid %u smem %zu ntiles %lld ncb_l %d G2 %d G3 %d nstage %d tw_mode %d op %d: %s",
			(int)INV, N, R0, R1, W, NREG, (int)TS, grid, smem, A.ntiles, A.ncb_l, A.G2, A.G3, A.nstage, A.tw_mode, A.op, cudaGetErrorString(e));
	}
	return 0;
}

// the specialised two-team kernel for power-of-two tiles; returns -1 when there is no instance for this shape
static int tf_launch2(const TfMaps &M, TfArgs &A, int nsm, size_t tabs, cudaStream_t st)
{
	static const bool off = getenv("B2_TFFT_GENERIC") && atoi(getenv("B2_TFFT_GENERIC"));
	if (off || A.nfac != 2) return -1;
	const int Wtot = A.nreg*A.W + (A.midcol >= 0 ? 1 : 0);
	A.tile_bytes = (int)b2_round_up((int64_t)Wtot*A.n*16, 128);
	A.out_bytes = (int)b2_round_up(A.tstore ? (int64_t)A.W*(A.n + 1)*16 : (int64_t)Wtot*A.n*16, 128);
	tabs += 2*(size_t)A.n*16;                                    // the teams' twiddle vectors
	// two teams of four warps on alternate tiles, or (tiles of 64 KB) one team of eight warps with a single output tile
	const int nteam = A.tile_bytes > 40*1024 ? 1 : 2;
	const size_t fixed = (nteam + (A.tstore ? 1 : 0))*(size_t)A.out_bytes + tabs;      // TS: + the twiddle tile
	if (226*1024 < fixed + 2*(size_t)A.tile_bytes) return -1;
	static const int max_stage = getenv("B2_TFFT_STAGES") ? atoi(getenv("B2_TFFT_STAGES")) : TF_MAXSTAGES;
	A.nstage = (int)std::min<size_t>(std::min(max_stage, TF_MAXSTAGES), (226*1024 - fixed)/A.tile_bytes);
	// the two teams take alternate tiles: with an even ring every landing tile belongs to one team, whose waits on its
	// mbarrier are then consecutive phases (an odd ring lets a team test a phase parity that an older fill satisfies)
	if (nteam == 2) A.nstage &= ~1;
	A.off_out = A.nstage*A.tile_bytes; A.off_work = A.off_out + nteam*A.out_bytes; A.off_tab = A.off_work + (A.tstore ? A.out_bytes : 0);
	const size_t smem = (size_t)A.off_tab + tabs;
	unsigned grid = (unsigned)std::min<long long>(A.ntiles, nsm);
	if (A.tstore) {
		// every CTA owns one column block: the grid is a multiple of the number of column blocks
		if (A.ncb_l > nsm) return -1;
		const long long per = std::max<long long>(1, std::min<long long>(nsm/A.ncb_l, (long long)A.G2*A.G3));
		grid = (unsigned)(per*A.ncb_l);
	}
	#define TF2_PLAIN(N_, R0_, R1_, W_) if (A.n == N_ && A.W == W_ && A.nreg == 1 && nteam == 2 && A.fac[0] == R0_ && A.fac[1] == R1_) { \
		if (A.tstore) return A.inv ? tf_launch2_k<true, N_, R0_, R1_, W_, 1, true, 2>(M, A, grid, smem, st) : tf_launch2_k<false, N_, R0_, R1_, W_, 1, true, 2>(M, A, grid, smem, st); \
		return A.inv ? tf_launch2_k<true, N_, R0_, R1_, W_, 1, false, 2>(M, A, grid, smem, st) : tf_launch2_k<false, N_, R0_, R1_, W_, 1, false, 2>(M, A, grid, smem, st); }
	#define TF2_MIRR(N_, R0_, R1_, W_, NT_) if (A.n == N_ && A.W == W_ && A.nreg == 2 && nteam == NT_ && !A.tstore && A.fac[0] == R0_ && A.fac[1] == R1_) \
		return A.inv ? tf_launch2_k<true, N_, R0_, R1_, W_, 2, false, NT_>(M, A, grid, smem, st) : tf_launch2_k<false, N_, R0_, R1_, W_, 2, false, NT_>(M, A, grid, smem, st);
	TF2_PLAIN(32, 4, 8, 64) TF2_PLAIN(32, 4, 8, 32) TF2_PLAIN(64, 8, 8, 32) TF2_PLAIN(64, 8, 8, 16) TF2_PLAIN(128, 16, 8, 16) TF2_PLAIN(256, 16, 16, 8)
	TF2_MIRR(32, 4, 8, 32, 2) TF2_MIRR(32, 4, 8, 16, 2) TF2_MIRR(64, 8, 8, 16, 2) TF2_MIRR(64, 8, 8, 8, 2) TF2_MIRR(128, 16, 8, 8, 2)
	TF2_MIRR(64, 8, 8, 32, 1) TF2_MIRR(128, 16, 8, 16, 1) TF2_MIRR(32, 4, 8, 64, 1)
	#undef TF2_PLAIN
	#undef TF2_MIRR
	return -1;
}

static int tf_launch(TfMaps &M, TfArgs &A, const TfLen &L, const TfTw2 *twN, const TfTw2 *twR, int nsm, cudaStream_t st)
{
	A.n = L.n; A.nfac = L.nfac; for (int i = 0; i < L.nfac; i++) A.fac[i] = L.fac[i];
	A.twn = L.twn.p; A.natk = L.natk.p;
	A.twN_hi = twN ? twN->hi.p : nullptr; A.twN_lo = twN ? twN->lo.p : nullptr; A.nhiN = twN ? twN->nhi : 1;
	A.twR_hi = twR ? twR->hi.p : nullptr; A.twR_lo = twR ? twR->lo.p : nullptr; A.nhiR = twR ? twR->nhi : 1;
	if (!twN) A.nhiN = 0;
	if (!twR) A.nhiR = 0;
	A.wshift = 0; while ((1 << A.wshift) < A.W) A.wshift++;
	const int Wtot = A.nreg*A.W + (A.midcol >= 0 ? 1 : 0);
	A.LW = std::min(32, A.W);
	A.tile_bytes = (int)b2_round_up((int64_t)Wtot*A.n*16, 128);
	A.out_bytes = (int)b2_round_up(A.tstore ? (int64_t)A.W*(A.n + 1)*16 : (int64_t)Wtot*A.n*16, 128);
	const size_t tabs = ((size_t)A.n + A.nhiN + 128 + A.nhiR + 128)*16 + (size_t)A.n*4 + 16;
	A.ntiles = (long long)A.ncb_l*A.G2*A.G3;
	if (A.ntiles <= 0) return 0;
	{ int rc = tf_launch2(M, A, nsm, tabs, st); if (rc >= 0) return rc; }
	const size_t fixed = (A.nfac > 1 ? A.tile_bytes : 0) + A.out_bytes + tabs;
	static const int max_stage = getenv("B2_TFFT_STAGES") ? atoi(getenv("B2_TFFT_STAGES")) : TF_MAXSTAGES;
	A.nstage = (int)std::min<size_t>(std::min(max_stage, TF_MAXSTAGES), (226*1024 - fixed)/A.tile_bytes);
	B2_REQUIRE(A.nstage >= 2, "tfft: tile of %d x %d does not leave room for two landing tiles", A.n, Wtot);
	A.off_work = A.nstage*A.tile_bytes;
	A.off_out = A.off_work + (A.nfac > 1 ? A.tile_bytes : 0);
	A.off_tab = A.off_out + A.out_bytes;
	const size_t smem = (size_t)A.off_tab + tabs;
	A.ntiles = (long long)A.ncb_l*A.G2*A.G3;
	if (A.ntiles <= 0) return 0;
	const unsigned grid = (unsigned)std::min<long long>(A.ntiles, nsm);
	static thread_local std::map<std::pair<int, int>, size_t> granted;
	int dev; B2_CHECK(cudaGetDevice(&dev));
	size_t &g = granted[std::make_pair(dev, A.inv)];
	if (smem > g) {
		if (A.inv) B2_CHECK(cudaFuncSetAttribute(k_tfft<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		else B2_CHECK(cudaFuncSetAttribute(k_tfft<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		g = smem;
	}
	if (A.inv) k_tfft<true><<<grid, TF_THREADS, smem, st>>>(M, A);
	else k_tfft<false><<<grid, TF_THREADS, smem, st>>>(M, A);
	B2_LAUNCH_CHECK();
	return 0;
}

static void tf_args_init(TfArgs &A, int inv)
{
	A = TfArgs();
	A.nreg = 1; A.midcol = -1; A.G2 = A.G3 = 1; A.g2_0 = A.g3_0 = 0; A.cb0 = 0; A.mirror = 0;
	A.tw_mode = 0; A.inv = inv; A.op = 0; A.tstore = 0; A.scale = 1.0; A.tbase = nullptr; A.t_g2stride = 0; A.ncols_valid = 0;
}

#define TF_MAP(m, base, d0, d1, d2, d3, s1, s2, s3, b0, b1) do { \
	uint64_t dims_[4] = {(uint64_t)(d0), (uint64_t)(d1), (uint64_t)(d2), (uint64_t)(d3)}; \
	uint64_t str_[3] = {(uint64_t)(s1), (uint64_t)(s2), (uint64_t)(s3)}; \
	uint32_t box_[4] = {(uint32_t)(b0), (uint32_t)(b1), 1, 1}; \
	int rc_ = b2_make_map_f64(&(m), (base), dims_, str_, box_); \
	B2_REQUIRE(rc_ == 0, "tfft: cuTensorMapEncodeTiled failed (%d) dims %llu %llu %llu %llu strides %llu %llu %llu box %u %u", rc_, \
		(unsigned long long)dims_[0], (unsigned long long)dims_[1], (unsigned long long)dims_[2], (unsigned long long)dims_[3], \
		(unsigned long long)str_[0], (unsigned long long)str_[1], (unsigned long long)str_[2], box_[0], box_[1]); } while (0)

// transform along x (the contiguous axis) of `nlines` lines of Nc complex elements: src line pitch ps / dst pitch pd (bytes)
static int tf_rows(TfPlan *p, const char *src, int64_t ps, char *dst, int64_t pd, int64_t nlines, int inv, double scale, cudaStream_t st)
{
	const TfAxis &X = p->ax;
	const int N1 = X.N1, N2 = X.N2, N = X.N;
	const int w1 = tf_width(N1, N2), w2 = tf_width(N2, N1);
	double2 *work = (double2*)p->work.p;
	const int64_t slab = std::max<int64_t>(1, (int64_t)(tf_slab_bytes()/((size_t)N*16)));
	for (int64_t l0 = 0; l0 < nlines; l0 += slab) {
		const int64_t nl = std::min(slab, nlines - l0);
		TfMaps M; TfArgs A;
		// sub-pass 1: tiles [j1: N1, stride N2][j2: w1] -> work[line][j2][k1]
		tf_args_init(A, inv);
		A.W = w1; A.ncb = (N2 + w1 - 1)/w1; A.ncb_l = A.ncb; A.G2 = (int)nl; A.g2_0 = (int)l0;
		A.tw_mode = 1; A.tstore = 1; A.tbase = work; A.t_g2stride = N; A.ncols_valid = N2;
		TF_MAP(M.ld[0], src, 2*N2, N1, nlines, 1, (int64_t)N2*16, ps, ps*nlines, 2*w1, N1);
		M.st[0] = M.ld[0];
		if (tf_launch(M, A, X.L1, &X.tw, nullptr, p->nsm, st)) return 1;
		// sub-pass 2: tiles [j2: N2, stride N1][k1: w2] of work -> dst[line][k1 + N1 k2]
		tf_args_init(A, inv);
		A.W = w2; A.ncb = (N1 + w2 - 1)/w2; A.ncb_l = A.ncb; A.G2 = (int)nl; A.g2_0 = (int)l0; A.scale = scale;
		TF_MAP(M.ld[0], work, 2*N1, N2, nlines, 1, (int64_t)N1*16, (int64_t)N*16, (int64_t)N*16*nlines, 2*w2, N2);
		TF_MAP(M.st[0], dst, 2*N1, N2, nlines, 1, (int64_t)N1*16, pd, pd*nlines, 2*w2, N2);
		if (tf_launch(M, A, X.L2, nullptr, nullptr, p->nsm, st)) return 1;
	}
	return 0;
}

// columns per region of a mirrored tile (two regions per tile) over a half range of `half` columns
static int tf_width_m(int n, int half)
{
	static const int big = getenv("B2_TFFT_MTILE") ? atoi(getenv("B2_TFFT_MTILE")) : 4096;      // elements of a mirrored tile
	const bool spec = n == 32 || n == 64 || n == 128;            // lengths with a single-team instance of the specialised kernel
	int w = pow2_floor(std::max(8, (spec ? big : tf_tile_elems())/(2*n)));
	w = std::min(w, 64);
	while (w > 8 && half % w) w /= 2;
	return w;
}

// transform along y of [nb][ny][ncols] complex arrays (pitches in bytes).  mode 0: plain; 1: r2c untangle fused into the
// load of sub-pass 1 (src holds Nc columns, work and dst Nc + 1); 2: c2r tangle fused into the store of sub-pass 2 (src and
// work hold Nc + 1 columns, dst Nc).  The fused step needs tiles of mirrored column blocks; with slabs (the intermediate
// array kept in L2 between the two sub-passes) both sub-passes walk the mirrored blocks so that a slab is the same set
// of columns in both, without slabs the other sub-pass uses plain tiles of twice the width.
static int tf_cols(TfPlan *p, const char *src, int64_t ps, int64_t bs_s, char *dst, int64_t pd, int64_t bs_d, char *work, int64_t pw,
	int ncols_src, int ncols_mid, int ncols_dst, int mode, int inv, double scale, cudaStream_t st)
{
	const TfAxis &Y = p->ay;
	const int N1 = Y.N1, N2 = Y.N2, ny = Y.N, Nc = p->Nc;
	const int64_t bs_w = (int64_t)ny*pw;
	const int next = std::max(ncols_src, ncols_dst);
	const bool slabbed = tf_slab_bytes() < (size_t)ny*16*next;
	const bool mir1 = mode == 1 || (slabbed && mode != 0), mir2 = mode == 2 || (slabbed && mode != 0);
	const int ext1 = mir1 ? Nc/2 : next, ext2 = mir2 ? Nc/2 : next;      // the range the column-block index runs over
	const int w1 = mir1 ? tf_width_m(N1, ext1) : tf_width(N1, ext1), w2 = mir2 ? tf_width_m(N2, ext2) : tf_width(N2, ext2);
	const int ncb1 = (ext1 + w1 - 1)/w1, ncb2 = (ext2 + w2 - 1)/w2, wmax = std::max(w1, w2);
	int64_t slab_cols = std::max(ext1, ext2);
	if (slabbed) {
		slab_cols = std::max<int64_t>(wmax, (int64_t)(tf_slab_bytes()/((size_t)ny*16*(mode != 0 ? 2 : 1))));
		slab_cols = slab_cols/wmax*wmax;
	}
	// without slabs one launch per sub-pass covers every batch member
	const int64_t bstep = slabbed ? 1 : p->nb;
	for (int64_t b = 0; b < p->nb; b += bstep) {
		for (int64_t c0 = 0; c0 < std::max(ext1, ext2); c0 += slab_cols) {
			const int64_t c1 = c0 + slab_cols;
			TfMaps M; TfArgs A;
			// ---- sub-pass 1: rows j1 (stride N2 rows) of a column block, fixed j2 -> work rows j2 N1 + k1, times w^(k1 j2)
			tf_args_init(A, inv);
			A.W = w1; A.ncb = ncb1; A.G2 = N2; A.G3 = (int)bstep; A.g3_0 = (int)b; A.tw_mode = 2;
			A.cb0 = (int)(c0/w1); A.ncb_l = (int)((std::min<int64_t>(c1, ext1) + w1 - 1)/w1) - A.cb0;
			if (mir1) { A.nreg = 2; A.mirror = Nc; A.midcol = Nc/2; A.op = mode == 1 ? 1 : 0; }
			TF_MAP(M.ld[0], src, 2*(int64_t)ncols_src, N1, N2, p->nb, (int64_t)N2*ps, ps, bs_s, 2*w1, N1);
			TF_MAP(M.st[0], work, 2*(int64_t)ncols_mid, N1, N2, p->nb, pw, (int64_t)N1*pw, bs_w, 2*w1, N1);
			M.ld[1] = M.ld[0]; M.st[1] = M.st[0]; M.ld[2] = M.ld[0]; M.st[2] = M.st[0];
			if (mir1) {
				TF_MAP(M.ld[2], src, 2*(int64_t)ncols_src, N1, N2, p->nb, (int64_t)N2*ps, ps, bs_s, 2, N1);
				TF_MAP(M.st[2], work, 2*(int64_t)ncols_mid, N1, N2, p->nb, pw, (int64_t)N1*pw, bs_w, 2, N1);
			}
			if (A.ncb_l > 0 && tf_launch(M, A, Y.L1, &Y.tw, mode == 1 ? &p->twR : nullptr, p->nsm, st)) return 1;
			// ---- sub-pass 2: work rows j2 (stride N1 rows), fixed k1 -> dst rows k1 + N1 k2
			tf_args_init(A, inv);
			A.W = w2; A.ncb = ncb2; A.G2 = N1; A.G3 = (int)bstep; A.g3_0 = (int)b; A.scale = scale;
			A.cb0 = (int)(c0/w2); A.ncb_l = (int)((std::min<int64_t>(c1, ext2) + w2 - 1)/w2) - A.cb0;
			if (mir2) { A.nreg = 2; A.mirror = Nc; A.midcol = Nc/2; A.op = mode == 2 ? 2 : 0; }
			TF_MAP(M.ld[0], work, 2*(int64_t)ncols_mid, N2, N1, p->nb, (int64_t)N1*pw, pw, bs_w, 2*w2, N2);
			TF_MAP(M.st[0], dst, 2*(int64_t)ncols_dst, N2, N1, p->nb, (int64_t)N1*pd, pd, bs_d, 2*w2, N2);
			M.ld[1] = M.ld[0]; M.st[1] = M.st[0]; M.ld[2] = M.ld[0]; M.st[2] = M.st[0];
			if (mir2) {
				TF_MAP(M.ld[2], work, 2*(int64_t)ncols_mid, N2, N1, p->nb, (int64_t)N1*pw, pw, bs_w, 2, N2);
				TF_MAP(M.st[2], dst, 2*(int64_t)ncols_dst, N2, N1, p->nb, (int64_t)N1*pd, pd, bs_d, 2, N2);
			}
			if (A.ncb_l > 0 && tf_launch(M, A, Y.L2, nullptr, mode == 2 ? &p->twR : nullptr, p->nsm, st)) return 1;
		}
	}
	return 0;
}

int tfft_execute(TfPlan *p, const void *in, void *out, int forward, double scale, cudaStream_t st)
{
	B2_REQUIRE(((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0), "tfft: arrays must be 16-byte aligned");
	const int inv = forward ? 0 : 1;
	const int64_t ny = p->ny, nb = p->nb; const int Nc = p->Nc;
	const int64_t nlines = nb*ny;
	// intermediate arrays are ours: their rows start on 128-byte boundaries whatever the caller's pitches are
	const int64_t pwe = b2_round_up(Nc + 1, 8), pw = pwe*16;
	const size_t wbytes = (size_t)nlines*pwe*16;
	if (p->work.n < wbytes && p->work.alloc(wbytes)) return 1;
	if (p->kind == B2_FFT_C2C) {
		const int64_t ps = p->in_pitch*16, pd = p->out_pitch*16;
		const bool lines_in = nb == 1 || p->in_bstride == ny*p->in_pitch, lines_out = nb == 1 || p->out_bstride == ny*p->out_pitch;
		if (lines_in && lines_out) { if (tf_rows(p, (const char*)in, ps, (char*)out, pd, nlines, inv, 1.0, st)) return 1; }
		else for (int64_t b = 0; b < nb; b++)
			if (tf_rows(p, (const char*)in + b*p->in_bstride*16, ps, (char*)out + b*p->out_bstride*16, pd, ny, inv, 1.0, st)) return 1;
		return tf_cols(p, (const char*)out, pd, p->out_bstride*16, (char*)out, pd, p->out_bstride*16, p->work.p, pw,
			Nc, Nc, Nc, 0, inv, scale, st);
	}
	if (p->work2.n < wbytes && p->work2.alloc(wbytes)) return 1;
	if (p->kind == B2_FFT_R2C) {
		// rows: in -> (work) -> work2 [line][Nc of pwe]; columns with the untangling: work2 -> work -> out
		const int64_t ps = p->in_pitch*8;
		const bool lines_in = nb == 1 || p->in_bstride == ny*p->in_pitch;
		// the untangling step of the column pass produces 2 X: the factor 1/2 rides on the row pass
		if (lines_in) { if (tf_rows(p, (const char*)in, ps, p->work2.p, pw, nlines, 0, 0.5, st)) return 1; }
		else for (int64_t b = 0; b < nb; b++)
			if (tf_rows(p, (const char*)in + b*p->in_bstride*8, ps, p->work2.p + b*ny*pw, pw, ny, 0, 0.5, st)) return 1;
		return tf_cols(p, p->work2.p, pw, ny*pw, (char*)out, p->out_pitch*16, p->out_bstride*16, p->work.p, pw,
			Nc, Nc + 1, Nc + 1, 1, 0, scale, st);
	}
	// c2r: columns first (on the half spectrum), tangle into the packed spectrum work2 [line][Nc of pwe], then the rows
	if (tf_cols(p, (const char*)in, p->in_pitch*16, p->in_bstride*16, p->work2.p, pw, ny*pw, p->work.p, pw,
		Nc + 1, Nc + 1, Nc, 2, 1, 1.0, st)) return 1;
	const int64_t pd = p->out_pitch*8;
	const bool lines_out = nb == 1 || p->out_bstride == ny*p->out_pitch;
	if (lines_out) return tf_rows(p, p->work2.p, pw, (char*)out, pd, nlines, 1, scale, st);
	for (int64_t b = 0; b < nb; b++)
		if (tf_rows(p, p->work2.p + b*ny*pw, pw, (char*)out + b*p->out_bstride*8, pd, ny, 1, scale, st)) return 1;
	return 0;
}
