/* oracle/sht_fast.c -- TEST INFRASTRUCTURE ONLY (the CPU arm of bench.py; never the product path).
 *
 * A tuned CPU Legendre stage for the "pixell + ducc0 on the host cores" leg of the measurement
 * (BASELINE.json north_star; SURVEY.md 8d "CPU baseline timed beside it").  ducc0 itself (PyPI
 * "ducc0>=0.36.0", reference pyproject.toml:27; call sites pixell/curvedsky.py:907-924, 936-960,
 * 1032-1046, 1068-1084) cannot be installed in this image, so this file restates the same published
 * algorithm family the way a CPU library runs it (libsharp / ducc structure as described in Reinecke &
 * Seljebotn 2013): OpenMP over m, SIMD over blocks of ring pairs, north/south symmetry, generation and
 * accumulation fused in one pass over l, rings beyond the turning point skipped, extended-exponent
 * scaling only in the prologue of a block.  It is checked against the slow checker (sht_oracle.c) in
 * tests/test_oracle_basic.py; the slow checker, not this file, is what the CUDA kernels are pinned to.
 *
 * leg layout here: [ncomp][nmlist][nring] complex, ring fastest (the m list is explicit so that a
 * bounded sample of the l,m triangle can be timed).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define NV 32                 /* ring pairs per block (4 AVX-512 vectors) */
#define RS 256                /* binary exponent step of the scaled prologue */

typedef struct { double c0, c1, c2; } frec_t;   /* F_{l+1} = (c0 x - c1) F_l - c2 F_{l-1},  F_l = n_l d^l_{m,n} */

static void f_make_rec(int lmax, int m, int n, frec_t *rc)
{
	int l0 = abs(m) > abs(n) ? abs(m) : abs(n);
	for (int l = l0; l < lmax; l++) {
		double l1 = l + 1.0;
		double den = sqrt((l1*l1 - (double)m*m)*(l1*l1 - (double)n*n));
		double f1 = sqrt((2*l + 3.0)/(2*l + 1.0))*(2*l + 1.0)*l1/den;
		rc[l - l0].c0 = f1;
		rc[l - l0].c1 = (l == 0) ? 0.0 : f1*((double)m*n)/((double)l*l1);
		rc[l - l0].c2 = (l == l0 || l == 0) ? 0.0 :
			sqrt((2*l + 3.0)/(2*l - 1.0))*l1*sqrt(((double)l*l - (double)m*m)*((double)l*l - (double)n*n))/(l*den);
	}
}

/* log2 |F_{l0}| and sign of F_{l0} = n_{l0} d^{l0}_{m,n}(theta), m >= 0 (closed form, Varshalovich 4.3.1) */
static int f_start(int m, int n, double sh, double ch, long double *lg, int *sg)
{
	int an = abs(n), l0 = m > an ? m : an, pc, ps, a;
	if (m >= an)    { pc = m + n; ps = m - n; *sg = ((m - n) & 1) ? -1 : 1; a = n; }
	else if (n > 0) { pc = n + m; ps = n - m; *sg = 1; a = m; }
	else            { pc = an - m; ps = an + m; *sg = ((an + m) & 1) ? -1 : 1; a = m; }
	if ((ps > 0 && sh == 0.0) || (pc > 0 && ch == 0.0)) return 0;
	long double v = 0.5L*(lgammal(2.0L*l0 + 1) - lgammal(l0 + a + 1.0L) - lgammal(l0 - a + 1.0L))/M_LN2l
		+ 0.5L*log2l((2.0L*l0 + 1)/(4*M_PIl));
	if (pc > 0) v += pc*log2l((long double)ch);
	if (ps > 0) v += ps*log2l((long double)sh);
	*lg = v;
	return 1;
}

typedef struct { int np; int *rn, *rs; double *x, *sh, *ch, *st; } fgeom_t;

static int cmp_st(const void *a, const void *b) { double d = ((const double*)a)[0] - ((const double*)b)[0]; return d < 0 ? -1 : d > 0; }

/* ring pairs (theta, pi - theta), sorted pole -> equator */
static void f_geom(int nring, const double *theta, fgeom_t *g)
{
	char *used = calloc(nring, 1);
	double *key = malloc(sizeof(double)*3*nring);
	int np = 0;
	/* mate search through a sorted copy: O(n log n) */
	double *srt = malloc(sizeof(double)*2*nring);
	for (int i = 0; i < nring; i++) { srt[2*i] = theta[i]; srt[2*i + 1] = i; }
	qsort(srt, nring, 2*sizeof(double), cmp_st);
	for (int i = 0; i < nring; i++) {
		if (used[i]) continue;
		used[i] = 1;
		double want = M_PI - theta[i];
		int lo = 0, hi = nring - 1, mate = -1;
		while (lo <= hi) { int mid = (lo + hi)/2; if (srt[2*mid] < want - 1e-13*M_PI) lo = mid + 1; else hi = mid - 1; }
		for (int k = lo; k < nring && srt[2*k] <= want + 1e-13*M_PI; k++) {
			int j = (int)srt[2*k + 1];
			if (!used[j]) { mate = j; break; }
		}
		if (mate >= 0) used[mate] = 1;
		int a = i, b = mate;
		if (mate >= 0 && theta[mate] < theta[i]) { a = mate; b = i; }      /* primary = northern ring of the pair */
		key[3*np] = sin(theta[a]); key[3*np + 1] = a; key[3*np + 2] = b; np++;
	}
	qsort(key, np, 3*sizeof(double), cmp_st);
	g->np = np;
	int npad = (np + NV - 1)/NV*NV;
	g->rn = malloc(sizeof(int)*npad); g->rs = malloc(sizeof(int)*npad);
	g->x = malloc(sizeof(double)*npad); g->sh = malloc(sizeof(double)*npad); g->ch = malloc(sizeof(double)*npad); g->st = malloc(sizeof(double)*npad);
	for (int k = 0; k < npad; k++) {
		if (k >= np) { g->rn[k] = g->rs[k] = -1; g->x[k] = 0; g->sh[k] = g->ch[k] = M_SQRT1_2; g->st[k] = 1; continue; }
		int a = (int)key[3*k + 1], b = (int)key[3*k + 2];
		double t = theta[a];
		g->rn[k] = a; g->rs[k] = b; g->x[k] = cos(t); g->st[k] = sin(t);
		if (t <= M_PI_2) { g->sh[k] = sin(0.5*t); g->ch[k] = cos(0.5*t); }
		else { double u = M_PI - t; g->sh[k] = cos(0.5*u); g->ch[k] = sin(0.5*u); }
	}
	free(used); free(key); free(srt);
}
static void f_geom_free(fgeom_t *g) { free(g->rn); free(g->rs); free(g->x); free(g->sh); free(g->ch); free(g->st); }

/* first pair index whose ring can receive anything from order m: beyond the turning point (l + 1/2) sin(theta) < m
 * the functions decay like exp(-0.94 d^1.5 / sqrt(m)), d = m - (lmax + 1/2) sin(theta); a margin of 20 m^(1/3) + 12
 * puts the dropped part below 1e-40 of the kept part */
static int f_first_pair(const fgeom_t *g, int lmax, int m, int spin)
{
	double lim = m - 20.0*cbrt((double)m) - spin - 12.0;
	if (lim <= 0) return 0;
	int lo = 0, hi = g->np;
	while (lo < hi) { int mid = (lo + hi)/2; if ((lmax + 0.5)*g->st[mid] < lim) lo = mid + 1; else hi = mid; }
	return lo/NV*NV;
}

/* state of one block of NV ring pairs for one sequence F_l = n_l d^l_{m,n}: value = cur * 2^ex, ex <= 0 in steps of RS;
 * live = 1 where ex == 0, else 0 (lanes still scaled are below 2^-100 of the unit scale and count as zero) */
typedef struct { double cur[NV], prev[NV], ex[NV], live[NV]; int nlive; } fstate_t;

static void f_init(fstate_t *s, const fgeom_t *g, int pb, int m, int n, double gsign)
{
	s->nlive = 0;
	for (int v = 0; v < NV; v++) {
		long double lg; int sg;
		s->prev[v] = 0; s->ex[v] = 0; s->live[v] = 1;
		if (g->rn[pb + v] < 0 || !f_start(m, n, g->sh[pb + v], g->ch[pb + v], &lg, &sg)) { s->cur[v] = 0; s->nlive++; continue; }
		if (lg > -900) { s->cur[v] = gsign*sg*(double)exp2l(lg); s->nlive++; }
		else {
			int k = (int)ceill((-lg - 600)/RS);
			s->ex[v] = -k*RS; s->live[v] = 0;
			s->cur[v] = gsign*sg*(double)exp2l(lg + (long double)k*RS);
		}
	}
}
/* n recurrence steps from index k (n <= CHK keeps a scaled lane, which is below 2^100 after a rescale and grows by less
 * than 2^7 per step, far from overflow) */
#define CHK 8
static inline void f_steps(fstate_t *restrict s, const double *restrict x, const frec_t *rc, int n)
{
	for (int j = 0; j < n; j++) {
		const frec_t r = rc[j];
		#pragma omp simd
		for (int v = 0; v < NV; v++) {
			double nx = (r.c0*x[v] - r.c1)*s->cur[v] - r.c2*s->prev[v];
			s->prev[v] = s->cur[v]; s->cur[v] = nx;
		}
	}
}
static inline void f_rescale(fstate_t *restrict s)
{
	double nl = 0;
	#pragma omp simd reduction(+:nl)
	for (int v = 0; v < NV; v++) {
		int up = s->ex[v] < 0 && fabs(s->cur[v]) >= 0x1p+100;
		double f = up ? 0x1p-256 : 1.0;
		s->cur[v] *= f; s->prev[v] *= f; s->ex[v] += up ? (double)RS : 0.0;
		s->live[v] = s->ex[v] == 0 ? 1.0 : 0.0;
		nl += s->live[v];
	}
	s->nlive = (int)nl;
}

int fast_num_threads(void) { return omp_get_max_threads(); }

/* spin 0, indices [k0, k1): accumulate a_k F_k by parity of k (k = l - m) and advance; MASKED multiplies by live */
static inline __attribute__((always_inline)) void syn0_run(int k0, int k1, const int masked, const double *restrict ap, const frec_t *rp,
	const double *restrict x, fstate_t *restrict P, double *restrict er, double *restrict ei, double *restrict orr, double *restrict oi)
{
	int k = k0;
	double c0[NV], c1[NV], lv[NV];
	for (int v = 0; v < NV; v++) { c0[v] = P->prev[v]; c1[v] = P->cur[v]; lv[v] = masked ? P->live[v] : 1.0; }
	if (k < k1 && (k & 1)) {
		const double a1r = ap[2*k], a1i = ap[2*k + 1]; const frec_t r0 = rp[k];
		#pragma omp simd
		for (int v = 0; v < NV; v++) {
			double c = masked ? c1[v]*lv[v] : c1[v];
			orr[v] += a1r*c; oi[v] += a1i*c;
			double n1 = (r0.c0*x[v])*c1[v] - r0.c2*c0[v];
			c0[v] = c1[v]; c1[v] = n1;
		}
		k++;
	}
	for (; k + 1 < k1; k += 2) {
		const double a0r = ap[2*k], a0i = ap[2*k + 1], a1r = ap[2*k + 2], a1i = ap[2*k + 3];
		const frec_t r0 = rp[k], r1 = rp[k + 1];
		#pragma omp simd
		for (int v = 0; v < NV; v++) {
			double c = masked ? c1[v]*lv[v] : c1[v];
			er[v] += a0r*c; ei[v] += a0i*c;
			double n1 = (r0.c0*x[v])*c1[v] - r0.c2*c0[v];
			double d = masked ? n1*lv[v] : n1;
			orr[v] += a1r*d; oi[v] += a1i*d;
			double n2 = (r1.c0*x[v])*n1 - r1.c2*c1[v];
			c0[v] = n1; c1[v] = n2;
		}
	}
	if (k < k1) {
		const double a0r = ap[2*k], a0i = ap[2*k + 1]; const frec_t r0 = rp[k];
		#pragma omp simd
		for (int v = 0; v < NV; v++) {
			double c = masked ? c1[v]*lv[v] : c1[v];
			er[v] += a0r*c; ei[v] += a0i*c;
			double n1 = (r0.c0*x[v])*c1[v] - r0.c2*c0[v];
			c0[v] = c1[v]; c1[v] = n1;
		}
	}
	for (int v = 0; v < NV; v++) { P->prev[v] = c0[v]; P->cur[v] = c1[v]; }
}

/* spin s: S+ = sum A+ p, S- = sum A- q (north); T+[par] = sum A+ q, T-[par] = sum A- p by parity of l + m + s (south) */
typedef struct { double spr[NV], spi[NV], smr[NV], smi[NV], tpr[2][NV], tpi[2][NV], tmr[2][NV], tmi[2][NV]; } facc2_t;
static inline __attribute__((always_inline)) void syn2_run(int k0, int k1, int par0, const int masked, const double *restrict ap, const double *restrict am,
	const frec_t *rp, const frec_t *rq, const double *restrict x, fstate_t *restrict P, fstate_t *restrict Q, facc2_t *restrict A)
{
	for (int k = k0; k < k1; k++) {
		const int par = (par0 + k) & 1;
		const double apr = ap[2*k], api = ap[2*k + 1], amr = am[2*k], ami = am[2*k + 1];
		const frec_t r0 = rp[k], r1 = rq[k];
		double *restrict t0 = A->tpr[par], *restrict t1 = A->tpi[par], *restrict t2 = A->tmr[par], *restrict t3 = A->tmi[par];
		#pragma omp simd
		for (int v = 0; v < NV; v++) {
			double p0 = P->cur[v], q0 = Q->cur[v];
			double p = masked ? p0*P->live[v] : p0, q = masked ? q0*Q->live[v] : q0;
			A->spr[v] += apr*p; A->spi[v] += api*p; A->smr[v] += amr*q; A->smi[v] += ami*q;
			t0[v] += apr*q; t1[v] += api*q; t2[v] += amr*p; t3[v] += ami*p;
			double np_ = (r0.c0*x[v] - r0.c1)*p0 - r0.c2*P->prev[v];
			double nq_ = (r1.c0*x[v] - r1.c1)*q0 - r1.c2*Q->prev[v];
			P->prev[v] = p0; P->cur[v] = np_; Q->prev[v] = q0; Q->cur[v] = nq_;
		}
	}
}

/* ---------------------------------------------------------------- synthesis: alm -> leg */
int fast_alm2leg(int spin, int lmax, int mmax, const int64_t *mstart, int nml, const int *mlist,
	int nring, const double *theta, const double *alm, int64_t alm_cstride, double *leg)
{
	if (spin < 0 || lmax < 0 || mmax > lmax) return 1;
	const int ncm = spin == 0 ? 1 : 2;
	fgeom_t g; f_geom(nring, theta, &g);
	memset(leg, 0, sizeof(double)*2*(size_t)ncm*nml*nring);
	#pragma omp parallel
	{
		frec_t *rp = calloc(lmax + 2, sizeof(frec_t)), *rq = calloc(lmax + 2, sizeof(frec_t));
		double *ap = malloc(sizeof(double)*2*(lmax + 2)), *am = malloc(sizeof(double)*2*(lmax + 2));
		#pragma omp for schedule(dynamic,1)
		for (int im = 0; im < nml; im++) {
			const int m = mlist[im], l0 = m > spin ? m : spin;
			if (m > mmax || l0 > lmax) continue;
			const int nl = lmax - l0 + 1;
			f_make_rec(lmax, m, -spin, rp); memset(rp + nl - 1, 0, sizeof(frec_t));      /* stepping past lmax is harmless */
			if (spin) { f_make_rec(lmax, m, spin, rq); memset(rq + nl - 1, 0, sizeof(frec_t)); }
			/* spin 0: ap = a_lm.  spin s: ap = A+ = -(E + iB)/2, am = A- = -(E - iB)/2 */
			for (int k = 0; k < nl; k++) {
				int64_t i = mstart[m] + l0 + k;
				if (!spin) { ap[2*k] = alm[2*i]; ap[2*k + 1] = alm[2*i + 1]; }
				else {
					const double *B = alm + 2*alm_cstride;
					double er = alm[2*i], ei = alm[2*i + 1], br = B[2*i], bi = B[2*i + 1];
					ap[2*k] = -0.5*(er - bi); ap[2*k + 1] = -0.5*(ei + br);
					am[2*k] = -0.5*(er + bi); am[2*k + 1] = -0.5*(ei - br);
				}
			}
			const double sgs = (spin & 1) ? -1.0 : 1.0;
			for (int pb = f_first_pair(&g, lmax, m, spin); pb < g.np; pb += NV) {
				const double *x = g.x + pb;
				fstate_t P, Q;
				f_init(&P, &g, pb, m, -spin, sgs);
				if (spin) f_init(&Q, &g, pb, m, spin, 1.0); else Q.nlive = NV;
				int k = 0;
				/* prologue 1: nothing live yet -- recurrence only, range check every CHK steps */
				while (k < nl && P.nlive == 0 && Q.nlive == 0) {
					int n = nl - k < CHK ? nl - k : CHK;
					f_steps(&P, x, rp + k, n); f_rescale(&P);
					if (spin) { f_steps(&Q, x, rq + k, n); f_rescale(&Q); }
					k += n;
				}
				if (k >= nl) continue;
				if (!spin) {
					double er[NV] = {0}, ei[NV] = {0}, orr[NV] = {0}, oi[NV] = {0};      /* by parity of l - m: north = e + o, south = e - o */
					/* prologue 2: some lanes still scaled */
					while (k < nl && P.nlive < NV) {
						int n = nl - k < CHK ? nl - k : CHK;
						syn0_run(k, k + n, 1, ap, rp, x, &P, er, ei, orr, oi); f_rescale(&P);
						k += n;
					}
					if (k < nl) syn0_run(k, nl, 0, ap, rp, x, &P, er, ei, orr, oi);
					double *o = leg + 2*(size_t)im*nring;
					for (int v = 0; v < NV; v++) {
						int rn = g.rn[pb + v], rs = g.rs[pb + v];
						if (rn >= 0) { o[2*rn] = er[v] + orr[v]; o[2*rn + 1] = ei[v] + oi[v]; }
						if (rs >= 0) { o[2*rs] = er[v] - orr[v]; o[2*rs + 1] = ei[v] - oi[v]; }
					}
				} else {
					facc2_t A; memset(&A, 0, sizeof(A));
					const int par0 = (l0 + m + spin) & 1;
					while (k < nl && (P.nlive < NV || Q.nlive < NV)) {
						int n = nl - k < CHK ? nl - k : CHK;
						syn2_run(k, k + n, par0, 1, ap, am, rp, rq, x, &P, &Q, &A); f_rescale(&P); f_rescale(&Q);
						k += n;
					}
					if (k < nl) syn2_run(k, nl, par0, 0, ap, am, rp, rq, x, &P, &Q, &A);
					/* Q = S+ + S-, U = i (S- - S+) */
					double *oq = leg + 2*(size_t)im*nring, *ou = leg + 2*((size_t)nml + im)*nring;
					for (int v = 0; v < NV; v++) {
						int rn = g.rn[pb + v], rs = g.rs[pb + v];
						if (rn >= 0) {
							oq[2*rn] = A.spr[v] + A.smr[v]; oq[2*rn + 1] = A.spi[v] + A.smi[v];
							ou[2*rn] = -(A.smi[v] - A.spi[v]); ou[2*rn + 1] = A.smr[v] - A.spr[v];
						}
						if (rs >= 0) {
							double ar = A.tpr[0][v] - A.tpr[1][v], ai = A.tpi[0][v] - A.tpi[1][v];      /* S+ south */
							double br = A.tmr[0][v] - A.tmr[1][v], bi = A.tmi[0][v] - A.tmi[1][v];      /* S- south */
							oq[2*rs] = ar + br; oq[2*rs + 1] = ai + bi;
							ou[2*rs] = -(bi - ai); ou[2*rs + 1] = br - ar;
						}
					}
				}
			}
		}
		free(rp); free(rq); free(ap); free(am);
	}
	f_geom_free(&g);
	return 0;
}

/* spin 0 transpose, indices [k0, k1): a_k += sum_v z[par(k)]_v F_k,v */
static inline __attribute__((always_inline)) void adj0_run(int k0, int k1, const int masked, double *restrict ap, const frec_t *rp, const double *restrict x,
	fstate_t *restrict P, const double *restrict e_r, const double *restrict e_i, const double *restrict o_r, const double *restrict o_i)
{
	for (int k = k0; k < k1; k++) {
		const double *restrict zr = (k & 1) ? o_r : e_r, *restrict zi = (k & 1) ? o_i : e_i;
		const frec_t r0 = rp[k];
		double s0 = 0, s1 = 0;
		#pragma omp simd reduction(+:s0,s1)
		for (int v = 0; v < NV; v++) {
			double c0 = P->cur[v], c = masked ? c0*P->live[v] : c0;
			s0 += zr[v]*c; s1 += zi[v]*c;
			double n1 = (r0.c0*x[v])*c0 - r0.c2*P->prev[v];
			P->prev[v] = c0; P->cur[v] = n1;
		}
		ap[2*k] += s0; ap[2*k + 1] += s1;
	}
}
typedef struct { double znr[NV], zni[NV], zsr[NV], zsi[NV], ynr[NV], yni[NV], ysr[NV], ysi[NV]; } fin2_t;
static inline __attribute__((always_inline)) void adj2_run(int k0, int k1, int par0, const int masked, double *restrict ap, double *restrict am,
	const frec_t *rp, const frec_t *rq, const double *restrict x, fstate_t *restrict P, fstate_t *restrict Q, const fin2_t *restrict Z)
{
	for (int k = k0; k < k1; k++) {
		const double sg = ((par0 + k) & 1) ? -1.0 : 1.0;
		const frec_t r0 = rp[k], r1 = rq[k];
		double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
		#pragma omp simd reduction(+:a0,a1,b0,b1)
		for (int v = 0; v < NV; v++) {
			double p0 = P->cur[v], q0 = Q->cur[v];
			double p = masked ? p0*P->live[v] : p0, q = masked ? q0*Q->live[v] : q0, sq = sg*q, sp = sg*p;
			a0 += p*Z->znr[v] + sq*Z->zsr[v]; a1 += p*Z->zni[v] + sq*Z->zsi[v];
			b0 += q*Z->ynr[v] + sp*Z->ysr[v]; b1 += q*Z->yni[v] + sp*Z->ysi[v];
			double np_ = (r0.c0*x[v] - r0.c1)*p0 - r0.c2*P->prev[v];
			double nq_ = (r1.c0*x[v] - r1.c1)*q0 - r1.c2*Q->prev[v];
			P->prev[v] = p0; P->cur[v] = np_; Q->prev[v] = q0; Q->cur[v] = nq_;
		}
		ap[2*k] += a0; ap[2*k + 1] += a1; am[2*k] += b0; am[2*k + 1] += b1;
	}
}

/* ---------------------------------------------------------------- adjoint synthesis: leg -> alm (plain transpose) */
int fast_leg2alm(int spin, int lmax, int mmax, const int64_t *mstart, int nml, const int *mlist,
	int nring, const double *theta, const double *leg, double *alm, int64_t alm_cstride)
{
	if (spin < 0 || lmax < 0 || mmax > lmax) return 1;
	fgeom_t g; f_geom(nring, theta, &g);
	#pragma omp parallel
	{
		frec_t *rp = calloc(lmax + 2, sizeof(frec_t)), *rq = calloc(lmax + 2, sizeof(frec_t));
		double *ap = malloc(sizeof(double)*2*(lmax + 2)), *am = malloc(sizeof(double)*2*(lmax + 2));
		#pragma omp for schedule(dynamic,1)
		for (int im = 0; im < nml; im++) {
			const int m = mlist[im], l0 = m > spin ? m : spin;
			if (m > mmax) continue;
			for (int l = m; l < l0 && l <= lmax; l++) {
				int64_t i = mstart[m] + l;
				alm[2*i] = alm[2*i + 1] = 0;
				if (spin) { double *B = alm + 2*alm_cstride; B[2*i] = B[2*i + 1] = 0; }
			}
			if (l0 > lmax) continue;
			const int nl = lmax - l0 + 1;
			memset(ap, 0, sizeof(double)*2*(lmax + 2)); memset(am, 0, sizeof(double)*2*(lmax + 2));
			f_make_rec(lmax, m, -spin, rp); memset(rp + nl - 1, 0, sizeof(frec_t));
			if (spin) { f_make_rec(lmax, m, spin, rq); memset(rq + nl - 1, 0, sizeof(frec_t)); }
			const double sgs = (spin & 1) ? -1.0 : 1.0;
			for (int pb = f_first_pair(&g, lmax, m, spin); pb < g.np; pb += NV) {
				const double *x = g.x + pb;
				fstate_t P, Q;
				f_init(&P, &g, pb, m, -spin, sgs);
				if (spin) f_init(&Q, &g, pb, m, spin, 1.0); else Q.nlive = NV;
				int k = 0;
				while (k < nl && P.nlive == 0 && Q.nlive == 0) {
					int n = nl - k < CHK ? nl - k : CHK;
					f_steps(&P, x, rp + k, n); f_rescale(&P);
					if (spin) { f_steps(&Q, x, rq + k, n); f_rescale(&Q); }
					k += n;
				}
				if (k >= nl) continue;
				if (!spin) {
					const double *gi = leg + 2*(size_t)im*nring;
					double e_r[NV], e_i[NV], o_r[NV], o_i[NV];
					for (int v = 0; v < NV; v++) {
						int rn = g.rn[pb + v], rs = g.rs[pb + v];
						double nr = rn >= 0 ? gi[2*rn] : 0, ni = rn >= 0 ? gi[2*rn + 1] : 0, sr = rs >= 0 ? gi[2*rs] : 0, si = rs >= 0 ? gi[2*rs + 1] : 0;
						e_r[v] = nr + sr; e_i[v] = ni + si; o_r[v] = nr - sr; o_i[v] = ni - si;
					}
					while (k < nl && P.nlive < NV) {
						int n = nl - k < CHK ? nl - k : CHK;
						adj0_run(k, k + n, 1, ap, rp, x, &P, e_r, e_i, o_r, o_i); f_rescale(&P);
						k += n;
					}
					if (k < nl) adj0_run(k, nl, 0, ap, rp, x, &P, e_r, e_i, o_r, o_i);
				} else {
					/* the real-linear transpose conjugates the coefficients:
					 * alpha_l = sum p Z+_n + sigma q Z+_s ; beta_l = sum q Z-_n + sigma p Z-_s ; Z+- = Q +- iU
					 * E = -(alpha + beta)/2, B = i (alpha - beta)/2 */
					const double *gq = leg + 2*(size_t)im*nring, *gu = leg + 2*((size_t)nml + im)*nring;
					fin2_t Z;
					for (int v = 0; v < NV; v++) {
						int rn = g.rn[pb + v], rs = g.rs[pb + v];
						double qr = rn >= 0 ? gq[2*rn] : 0, qi = rn >= 0 ? gq[2*rn + 1] : 0, ur = rn >= 0 ? gu[2*rn] : 0, ui = rn >= 0 ? gu[2*rn + 1] : 0;
						Z.znr[v] = qr - ui; Z.zni[v] = qi + ur; Z.ynr[v] = qr + ui; Z.yni[v] = qi - ur;      /* z = Z+ = Q + iU, y = Z- = Q - iU */
						qr = rs >= 0 ? gq[2*rs] : 0; qi = rs >= 0 ? gq[2*rs + 1] : 0; ur = rs >= 0 ? gu[2*rs] : 0; ui = rs >= 0 ? gu[2*rs + 1] : 0;
						Z.zsr[v] = qr - ui; Z.zsi[v] = qi + ur; Z.ysr[v] = qr + ui; Z.ysi[v] = qi - ur;
					}
					const int par0 = (l0 + m + spin) & 1;
					while (k < nl && (P.nlive < NV || Q.nlive < NV)) {
						int n = nl - k < CHK ? nl - k : CHK;
						adj2_run(k, k + n, par0, 1, ap, am, rp, rq, x, &P, &Q, &Z); f_rescale(&P); f_rescale(&Q);
						k += n;
					}
					if (k < nl) adj2_run(k, nl, par0, 0, ap, am, rp, rq, x, &P, &Q, &Z);
				}
			}
			for (int k = 0; k < nl; k++) {
				int64_t i = mstart[m] + l0 + k;
				if (!spin) { alm[2*i] = ap[2*k]; alm[2*i + 1] = ap[2*k + 1]; }
				else {
					double *B = alm + 2*alm_cstride;
					double ar = ap[2*k], ai = ap[2*k + 1], br = am[2*k], bi = am[2*k + 1];
					alm[2*i] = -0.5*(ar + br); alm[2*i + 1] = -0.5*(ai + bi);
					B[2*i] = -0.5*(ai - bi); B[2*i + 1] = 0.5*(ar - br);      /* i (alpha - beta)/2 */
				}
			}
		}
		free(rp); free(rq); free(ap); free(am);
	}
	f_geom_free(&g);
	return 0;
}
