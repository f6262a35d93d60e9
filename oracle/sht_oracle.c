/* oracle/sht_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU checker, never the product path).
 *
 * CPU restatement of the Legendre stage of the spherical-harmonic transforms that
 * pixell's curvedsky module delegates to ducc0 (third-party, pinned "ducc0>=0.36.0",
 * reference pyproject.toml:27; call sites pixell/curvedsky.py:907-924, 936-960,
 * 1032-1046, 1068-1084).  ducc0 is not vendored in the reference tree, so this file
 * restates the published mathematics (HEALPix / libsharp conventions, SURVEY.md
 * Appendix A) and is pinned against the reference's own golden fixtures
 * (tests/data/MM_unlensed_071123.fits, MM_041121.pkl; see tests/test_oracle_golden.py).
 *
 * Conventions
 *   spin 0 :  f(theta,phi) = sum_{m>=0} w_m Re[ e^{i m phi} sum_l a_lm lambda_lm(theta) ]
 *             lambda_lm = sqrt((2l+1)/4pi) d^l_{m,0}(theta)   (Condon-Shortley phase)
 *   spin s>0: p_l = (-1)^s n_l d^l_{m,-s},  q_l = n_l d^l_{m,+s},  n_l = sqrt((2l+1)/4pi)
 *             W = -(p+q)/2,  X = -(p-q)/2
 *             Q_m = sum_l (E W + i B X),  U_m = sum_l (B W - i E X)
 *   DERIV1  : spin 1 with E_lm = sqrt(l(l+1)) a_lm, B = 0 -> (d/dtheta f, 1/sin(theta) d/dphi f)
 *
 * Only the theta-dependent part lives here: alm <-> leg[comp][ring][m].  The ring FFTs,
 * phi0 phases, theta resampling and quadrature weights are numpy code in sht_oracle.py.
 *
 * The Wigner d-functions are generated with the three-term recurrence in l
 * (Varshalovich 4.8.2) started from the closed form at l0 = max(m,s); the start value is
 * formed in log2-space and carried with an explicit binary exponent so that
 * sin^m(theta/2) underflow (1e-700 and below at lmax=8000) never produces garbage.
 * This is deliberately a different formulation from the CUDA kernels (which use the
 * alpha-normalised two-FMA recurrence and the +/- basis) so that the two check each other.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define VL 8            /* rings processed together (lets gcc vectorise the inner loops) */
/* bench.py's bounded CPU-baseline sample: process only every g_mstride-th m (default: all) */
static int g_mstride = 1;
void orc_set_mstride(int s) { g_mstride = s > 0 ? s : 1; }
#define RS_BITS 256     /* renormalisation step (binary exponent) */

typedef struct { double c0, c1, c2; } rec_t;   /* F_{l+1} = (c0*cos(theta) - c1) F_l - c2 F_{l-1} */

/* recurrence coefficients for F_l = sqrt(2l+1) d^l_{m,n}, l = l0 .. lmax-1 (index l-l0) */
static void make_rec(int lmax, int m, int n, rec_t *rc)
{
	int l0 = abs(m) > abs(n) ? abs(m) : abs(n);
	for (int l = l0; l < lmax; l++) {
		double l1 = l + 1.0;
		double den = sqrt((l1*l1 - (double)m*m) * (l1*l1 - (double)n*n));
		double f1 = sqrt((2*l+3.0)/(2*l+1.0)) * (2*l+1.0) * l1 / den;
		rec_t r;
		r.c0 = f1;
		r.c1 = (l == 0) ? 0.0 : f1 * ((double)m*n) / ((double)l*l1);
		r.c2 = (l == l0 || l == 0) ? 0.0 :
			sqrt((2*l+3.0)/(2*l-1.0)) * l1 * sqrt(((double)l*l - (double)m*m) * ((double)l*l - (double)n*n)) / (l*den);
		rc[l-l0] = r;
	}
}

/* log2|F_{l0}| and its sign for F = sqrt(2l+1)/sqrt(4pi) d^{l0}_{m,n}(theta); m>=0.
 * returns 0 if the value is exactly zero (pole).  sh = sin(theta/2), ch = cos(theta/2). */
static int start_value(int m, int n, double sh, double ch, long double *log2v, int *sign)
{
	int an = abs(n), l0 = m > an ? m : an;
	int pc, ps, sg;   /* powers of cos(theta/2), sin(theta/2) and sign */
	int a;            /* the smaller index, for the binomial prefactor */
	if (m >= an) { pc = m + n; ps = m - n; sg = ((m - n) & 1) ? -1 : 1; a = n; }
	else if (n > 0) { pc = n + m; ps = n - m; sg = 1; a = m; }              /* d^s_{m,s}  */
	else { pc = an - m; ps = an + m; sg = ((an + m) & 1) ? -1 : 1; a = m; }   /* d^s_{m,-s} */
	if ((ps > 0 && sh == 0.0) || (pc > 0 && ch == 0.0)) return 0;
	/* long double keeps the absolute error of the ~1e5-sized logarithm near 1e-14 */
	long double lg = 0.5L*(lgammal(2.0L*l0+1) - lgammal(l0+a+1.0L) - lgammal(l0-a+1.0L))/M_LN2l
	          + 0.5L*log2l((2.0L*l0+1)/(4*M_PIl));
	if (pc > 0) lg += pc*log2l((long double)ch);
	if (ps > 0) lg += ps*log2l((long double)sh);
	*log2v = lg; *sign = sg;
	return 1;
}

static inline void half_angle(double theta, double *sh, double *ch)
{
	/* accurate near both poles */
	if (theta <= M_PI_2) { double t = theta > 0 ? theta : 0; *sh = sin(0.5*t); *ch = cos(0.5*t); }
	else { double t = M_PI - theta; if (t < 0) t = 0; *sh = cos(0.5*t); *ch = sin(0.5*t); }
}

/* Fill out[v][l-l0], l=l0..lmax, v<nv<=VL, with the true values of n_l d^l_{m,n}(theta_v)
 * (times an overall sign `gsign`). */
static void wigner_cols(int lmax, int m, int n, const rec_t *rc, int nv, const double *theta,
                        double gsign, double *out, int ldo)
{
	int an = abs(n), l0 = m > an ? m : an;
	double cur[VL], prev[VL], cth[VL]; int ex[VL];
	int anyscaled = 0;
	for (int v = 0; v < VL; v++) { cur[v] = prev[v] = 0; cth[v] = 0; ex[v] = 0; }
	for (int v = 0; v < nv; v++) {
		double sh, ch; long double lg; int sg;
		half_angle(theta[v], &sh, &ch);
		cth[v] = cos(theta[v]);
		if (!start_value(m, n, sh, ch, &lg, &sg)) { cur[v] = 0; ex[v] = 0; continue; }
		if (lg > -900) { cur[v] = gsign*sg*(double)exp2l(lg); ex[v] = 0; }
		else {
			int k = (int)ceill((-lg - 600)/RS_BITS);   /* lg + k*RS in (-856,-600] */
			ex[v] = -k*RS_BITS;
			cur[v] = gsign*sg*(double)exp2l(lg + (long double)k*RS_BITS);
			anyscaled = 1;
		}
	}
	for (int l = l0; l <= lmax; l++) {
		if (anyscaled) {
			anyscaled = 0;
			for (int v = 0; v < VL; v++) {
				if (ex[v] < 0 && fabs(cur[v]) >= 0x1p+200) {
					cur[v] *= 0x1p-256; prev[v] *= 0x1p-256; ex[v] += RS_BITS;
				}
				out[v*ldo + (l-l0)] = ex[v] == 0 ? cur[v] : ldexp(cur[v], ex[v]);
				anyscaled |= ex[v] < 0;
			}
		} else {
			for (int v = 0; v < VL; v++) out[v*ldo + (l-l0)] = cur[v];
		}
		if (l < lmax) {
			rec_t r = rc[l-l0];
			for (int v = 0; v < VL; v++) {
				double nx = (r.c0*cth[v] - r.c1)*cur[v] - r.c2*prev[v];
				prev[v] = cur[v]; cur[v] = nx;
			}
		}
	}
}

/* Ring bookkeeping: pair rings i<j with theta_j == pi - theta_i so the north/south symmetry
 * lambda(pi-theta) = (-1)^(l+m(+s)) lambda(theta) halves the work.  pair[2k],pair[2k+1] = (north, south or -1). */
static int make_pairs(int nring, const double *theta, int *pair)
{
	char *used = calloc(nring, 1);
	int np = 0;
	for (int i = 0; i < nring; i++) {
		if (used[i]) continue;
		used[i] = 1;
		int mate = -1;
		for (int j = nring-1; j > i; j--) {
			if (used[j]) continue;
			if (fabs(theta[j] - (M_PI - theta[i])) < 1e-14*M_PI) { mate = j; break; }
		}
		if (mate >= 0) used[mate] = 1;
		pair[2*np] = i; pair[2*np+1] = mate; np++;
	}
	free(used);
	return np;
}

/* ---- synthesis: alm -> leg ------------------------------------------------------------
 * alm : ncomp_alm arrays of interleaved complex doubles, component c at alm + 2*c*alm_cstride
 * leg : [ncomp_map][nring][mmax+1] interleaved complex doubles
 * spin 0: 1->1.  spin>0: 2->2.  deriv1: spin must be 1, 1->2.
 */
int orc_alm2leg(int spin, int deriv1, int lmax, int mmax, const int64_t *mstart,
                int nring, const double *theta,
                const double *alm, int64_t alm_cstride, double *leg)
{
	if (spin < 0 || lmax < 0 || mmax > lmax || (deriv1 && spin != 1)) return 1;
	int ncm = spin == 0 ? 1 : 2;
	int nm = mmax + 1;
	int *pair = malloc(sizeof(int)*2*nring);
	int np = make_pairs(nring, theta, pair);
	memset(leg, 0, sizeof(double)*2*(size_t)ncm*nring*nm);
	#pragma omp parallel
	{
		int ldo = lmax + 1;
		rec_t *rcp = malloc(sizeof(rec_t)*(lmax+1)), *rcq = malloc(sizeof(rec_t)*(lmax+1));
		double *P = malloc(sizeof(double)*VL*ldo), *Q = malloc(sizeof(double)*VL*ldo);
		double *a0 = malloc(sizeof(double)*2*(lmax+1)), *a1 = malloc(sizeof(double)*2*(lmax+1));
		#pragma omp for schedule(dynamic,1)
		for (int m = 0; m <= mmax; m += g_mstride) {
			int l0 = m > spin ? m : spin;
			if (l0 > lmax) continue;
			/* gather this m's coefficients */
			for (int l = l0; l <= lmax; l++) {
				int64_t i = mstart[m] + l;
				double f = deriv1 ? sqrt((double)l*(l+1.0)) : 1.0;
				a0[2*(l-l0)] = f*alm[2*i]; a0[2*(l-l0)+1] = f*alm[2*i+1];
				if (spin > 0 && !deriv1) {
					const double *b = alm + 2*alm_cstride;
					a1[2*(l-l0)] = b[2*i]; a1[2*(l-l0)+1] = b[2*i+1];
				} else { a1[2*(l-l0)] = a1[2*(l-l0)+1] = 0; }
			}
			make_rec(lmax, m, -spin, rcp);
			if (spin > 0) make_rec(lmax, m, spin, rcq);
			double sgs = (spin & 1) ? -1.0 : 1.0;
			for (int pb = 0; pb < np; pb += VL) {
				int nv = np - pb < VL ? np - pb : VL;
				double th[VL];
				for (int v = 0; v < nv; v++) th[v] = theta[pair[2*(pb+v)]];
				wigner_cols(lmax, m, -spin, rcp, nv, th, sgs, P, ldo);
				if (spin > 0) wigner_cols(lmax, m, spin, rcq, nv, th, 1.0, Q, ldo);
				for (int v = 0; v < nv; v++) {
					int rn = pair[2*(pb+v)], rs = pair[2*(pb+v)+1];
					const double *p = P + v*ldo, *q = Q + v*ldo;
					if (spin == 0) {
						double er = 0, ei = 0, orr = 0, oi = 0;   /* even / odd (l+m) parts */
						int l = l0;
						for (; l+1 <= lmax; l += 2) {
							er  += a0[2*(l-l0)]*p[l-l0];     ei += a0[2*(l-l0)+1]*p[l-l0];
							orr += a0[2*(l-l0)+2]*p[l-l0+1]; oi += a0[2*(l-l0)+3]*p[l-l0+1];
						}
						if (l <= lmax) { er += a0[2*(l-l0)]*p[l-l0]; ei += a0[2*(l-l0)+1]*p[l-l0]; }
						/* l0 = m so (l-l0) even <=> (l+m) even */
						double *o = leg + 2*((size_t)rn*nm + m);
						o[0] = er + orr; o[1] = ei + oi;
						if (rs >= 0) { o = leg + 2*((size_t)rs*nm + m); o[0] = er - orr; o[1] = ei - oi; }
					} else {
						/* accumulate Q_m, U_m split by the parity of W under theta -> pi-theta */
						double qe[2] = {0,0}, qo[2] = {0,0}, ue[2] = {0,0}, uo[2] = {0,0};
						for (int l = l0; l <= lmax; l++) {
							double W = -0.5*(p[l-l0] + q[l-l0]), X = -0.5*(p[l-l0] - q[l-l0]);
							double Er = a0[2*(l-l0)], Ei = a0[2*(l-l0)+1], Br = a1[2*(l-l0)], Bi = a1[2*(l-l0)+1];
							/* Q += E W + i B X ; U += B W - i E X */
							double qWr = Er*W, qWi = Ei*W, qXr = -Bi*X, qXi = Br*X;
							double uWr = Br*W, uWi = Bi*W, uXr = Ei*X, uXi = -Er*X;
							if (((l + m + spin) & 1) == 0) {   /* W even, X odd */
								qe[0] += qWr; qe[1] += qWi; qo[0] += qXr; qo[1] += qXi;
								ue[0] += uWr; ue[1] += uWi; uo[0] += uXr; uo[1] += uXi;
							} else {
								qo[0] += qWr; qo[1] += qWi; qe[0] += qXr; qe[1] += qXi;
								uo[0] += uWr; uo[1] += uWi; ue[0] += uXr; ue[1] += uXi;
							}
						}
						double *oq = leg + 2*((size_t)rn*nm + m);
						double *ou = leg + 2*(((size_t)nring + rn)*nm + m);
						oq[0] = qe[0]+qo[0]; oq[1] = qe[1]+qo[1]; ou[0] = ue[0]+uo[0]; ou[1] = ue[1]+uo[1];
						if (rs >= 0) {
							oq = leg + 2*((size_t)rs*nm + m); ou = leg + 2*(((size_t)nring + rs)*nm + m);
							oq[0] = qe[0]-qo[0]; oq[1] = qe[1]-qo[1]; ou[0] = ue[0]-uo[0]; ou[1] = ue[1]-uo[1];
						}
					}
				}
			}
		}
		free(rcp); free(rcq); free(P); free(Q); free(a0); free(a1);
	}
	free(pair);
	return 0;
}

/* ---- adjoint synthesis: leg -> alm (plain transpose of orc_alm2leg; no weights) ---------
 * alm_lm = sum_ring conj-free transpose: for spin 0  a_lm = sum_r leg[r][m] lambda_lm(theta_r)
 * Entries of alm outside l0<=l<=lmax, m<=mmax are left untouched; covered entries are overwritten.
 */
int orc_leg2alm(int spin, int deriv1, int lmax, int mmax, const int64_t *mstart,
                int nring, const double *theta,
                const double *leg, double *alm, int64_t alm_cstride)
{
	if (spin < 0 || lmax < 0 || mmax > lmax || (deriv1 && spin != 1)) return 1;
	int nm = mmax + 1;
	int *pair = malloc(sizeof(int)*2*nring);
	int np = make_pairs(nring, theta, pair);
	#pragma omp parallel
	{
		int ldo = lmax + 1;
		rec_t *rcp = malloc(sizeof(rec_t)*(lmax+1)), *rcq = malloc(sizeof(rec_t)*(lmax+1));
		double *P = malloc(sizeof(double)*VL*ldo), *Q = malloc(sizeof(double)*VL*ldo);
		double *a0 = malloc(sizeof(double)*2*(lmax+1)), *a1 = malloc(sizeof(double)*2*(lmax+1));
		#pragma omp for schedule(dynamic,1)
		for (int m = 0; m <= mmax; m += g_mstride) {
			int l0 = m > spin ? m : spin;
			/* spin>0: l < spin entries are zero by definition */
			for (int l = m; l < l0 && l <= lmax; l++) {
				int64_t i = mstart[m] + l;
				alm[2*i] = alm[2*i+1] = 0;
				if (spin > 0 && !deriv1) { double *b = alm + 2*alm_cstride; b[2*i] = b[2*i+1] = 0; }
			}
			if (l0 > lmax) continue;
			memset(a0, 0, sizeof(double)*2*(lmax+1)); memset(a1, 0, sizeof(double)*2*(lmax+1));
			make_rec(lmax, m, -spin, rcp);
			if (spin > 0) make_rec(lmax, m, spin, rcq);
			double sgs = (spin & 1) ? -1.0 : 1.0;
			for (int pb = 0; pb < np; pb += VL) {
				int nv = np - pb < VL ? np - pb : VL;
				double th[VL];
				for (int v = 0; v < nv; v++) th[v] = theta[pair[2*(pb+v)]];
				wigner_cols(lmax, m, -spin, rcp, nv, th, sgs, P, ldo);
				if (spin > 0) wigner_cols(lmax, m, spin, rcq, nv, th, 1.0, Q, ldo);
				for (int v = 0; v < nv; v++) {
					int rn = pair[2*(pb+v)], rs = pair[2*(pb+v)+1];
					const double *p = P + v*ldo, *q = Q + v*ldo;
					if (spin == 0) {
						const double *gn = leg + 2*((size_t)rn*nm + m);
						double sr = 0, si = 0;
						if (rs >= 0) { const double *gs = leg + 2*((size_t)rs*nm + m); sr = gs[0]; si = gs[1]; }
						double er = gn[0]+sr, ei = gn[1]+si, orr = gn[0]-sr, oi = gn[1]-si;
						for (int l = l0; l <= lmax; l++) {
							if (((l-l0)&1) == 0) { a0[2*(l-l0)] += er*p[l-l0];  a0[2*(l-l0)+1] += ei*p[l-l0]; }
							else                 { a0[2*(l-l0)] += orr*p[l-l0]; a0[2*(l-l0)+1] += oi*p[l-l0]; }
						}
					} else {
						const double *qn = leg + 2*((size_t)rn*nm + m);
						const double *un = leg + 2*(((size_t)nring + rn)*nm + m);
						double qs[2] = {0,0}, us[2] = {0,0};
						if (rs >= 0) {
							const double *a = leg + 2*((size_t)rs*nm + m), *b = leg + 2*(((size_t)nring + rs)*nm + m);
							qs[0] = a[0]; qs[1] = a[1]; us[0] = b[0]; us[1] = b[1];
						}
						double qe[2] = {qn[0]+qs[0], qn[1]+qs[1]}, qo[2] = {qn[0]-qs[0], qn[1]-qs[1]};
						double ue[2] = {un[0]+us[0], un[1]+us[1]}, uo[2] = {un[0]-us[0], un[1]-us[1]};
						for (int l = l0; l <= lmax; l++) {
							double W = -0.5*(p[l-l0] + q[l-l0]), X = -0.5*(p[l-l0] - q[l-l0]);
							const double *qw, *qx, *uw, *ux;
							if (((l + m + spin) & 1) == 0) { qw = qe; uw = ue; qx = qo; ux = uo; }
							else                           { qw = qo; uw = uo; qx = qe; ux = ue; }
							/* transpose of: Q += E W + i B X ; U += B W - i E X
							 * E += W Q + (-i X U)^T-> E += W*Q + conj-free: real-linear transpose:
							 * (Er,Ei,Br,Bi) <- [[W,0],[0,W]]Q ... derived component-wise below */
							/* Qr = Er W - Bi X ; Qi = Ei W + Br X ; Ur = Br W + Ei X ; Ui = Bi W - Er X */
							a0[2*(l-l0)]   += W*qw[0] - X*ux[1];
							a0[2*(l-l0)+1] += W*qw[1] + X*ux[0];
							a1[2*(l-l0)]   += X*qx[1] + W*uw[0];
							a1[2*(l-l0)+1] += -X*qx[0] + W*uw[1];
						}
					}
				}
			}
			for (int l = l0; l <= lmax; l++) {
				int64_t i = mstart[m] + l;
				double f = deriv1 ? sqrt((double)l*(l+1.0)) : 1.0;
				alm[2*i] = f*a0[2*(l-l0)]; alm[2*i+1] = f*a0[2*(l-l0)+1];
				if (spin > 0 && !deriv1) { double *b = alm + 2*alm_cstride; b[2*i] = a1[2*(l-l0)]; b[2*i+1] = a1[2*(l-l0)+1]; }
			}
		}
		free(rcp); free(rcq); free(P); free(Q); free(a0); free(a1);
	}
	free(pair);
	return 0;
}

int orc_num_threads(void) { return omp_get_max_threads(); }
