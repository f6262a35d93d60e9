// resample.cuh -- K5: exact-quadrature weighting of leg along theta for analysis_2d.
#pragma once
#include "fft_smem.cuh"
#include <string>

// Applies, to every (component, m) column of leg[.][m][ring] on a CC / F1 / MW / MWflip grid with
// fewer than 2 lmax + 2 rings, the operator that ducc0's analysis_2d realises by upsampling to a
// Clenshaw-Curtis grid and weighting there (pixell/curvedsky.py:1033-1046) -- but folded back onto the ORIGINAL rings: the weighted fine
// samples are low-passed to |k| <= lmax and re-evaluated on the coarse grid, which leaves every
// a_lm unchanged (lambda_lm has bandwidth lmax) and lets the Legendre stage run on ntheta rings
// instead of 2 lmax + 2.  See DESIGN.md "theta weighting".
struct ThetaResampler {
	int n = 0, N = 0, o2 = 0, P = 1, lmax = 0, nm = 0, npc = 0;   // npc: column pairs per component
	int64_t nring_pad = 0, nphi = 0;
	FftTables tab;             // length N/P, twiddle table of 2N entries
	DevBuf<int> src;           // [N] circle slot -> ring index, bit 30 = mirrored copy, -1 = empty
	DevBuf<int> dpos, dmir;    // [n] ring -> its circle slot and the slot of its mirror image
	DevBuf<double> wfine;      // [2N] quadrature weight function on the fine circle / nphi
	DevBuf<double> mult;       // [n] 2 (interior ring) or 1 (pole ring)
	DevBuf<double2> A, B;      // scratch [cb][N]
	DevBuf<double2> C;         // third scratch array, allocated by the first adjoint call
	int64_t cb = 0;            // column pairs per batch
	int threads = 256; size_t smem = 0; int twoff = 0;
	int M0 = 0;                // the mirror image of ring r sits in circle slot M0 - r
	static bool needed(const std::string &geom, int ntheta, int lmax);
	int build(const std::string &geom, int ntheta, int64_t nphi, int lmax, int mmax, int64_t nring_pad);
	// in place on leg[ncomp][nm][nring_pad]; adjoint = true applies the conjugate-transposed operator (adjoint_analysis_2d)
	int apply(double2 *leg, int ncomp, int spin, cudaStream_t st, bool adjoint = false);
	size_t bytes() const { return tab.bytes() + src.bytes() + dpos.bytes() + dmir.bytes() + wfine.bytes() + mult.bytes() + A.bytes() + B.bytes() + C.bytes(); }
};
