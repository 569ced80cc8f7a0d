/* TEST INFRASTRUCTURE ONLY -- see oxdna_oracle.h.
 *
 * CPU restatement of the oxDNA2 force field, Verlet list and velocity-Verlet step, double precision.
 * It follows the *algorithm* of the reference's CPU classes (cited per function) but is organised
 * differently: every angular/radial factor is differentiated through one generic chain-rule helper
 * instead of the reference's hand-expanded expressions, which makes it an independent check.
 */
#include "oxdna_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* the reference defines PI as a *float* literal (src/defs.h:14); every t0 derived from it inherits that rounding */
#define OXO_PI 3.141592653589793238462643f
#define SQ(x) ((x) * (x))

/* ------------------------------------------------------------------ small vector helpers */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double *a, const double *b, double *o) {
	o[0] = a[1] * b[2] - a[2] * b[1];
	o[1] = a[2] * b[0] - a[0] * b[2];
	o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void axpy3(double s, const double *x, double *y) { y[0] += s * x[0]; y[1] += s * x[1]; y[2] += s * x[2]; }

/* ------------------------------------------------------------------ parameters */
static void set_f4(oxo_f4 *f, double a, double b, double t0, double ts, double tc) {
	f->a = a; f->b = b; f->t0 = t0; f->ts = ts; f->tc = tc;
}

static void fill_f1(oxo_f1 *f, double eps) {
	/* shift = eps * (1 - exp(-(rc - r0) a))^2, src/Interactions/DNA2Interaction.cpp:114-121.  (rc - r0)*a is float
	 * arithmetic in the reference (float macros); exp() there resolves to the double overload. */
	for(int i = 0; i < 5; i++) for(int j = 0; j < 5; j++) {
		f->eps[i][j] = eps;
		f->shift[i][j] = eps * SQ(1 - exp(-(double) ((float) (f->rc - f->r0) * (float) f->a)));
	}
}

void oxo_dna2_params_init(oxo_dna2_params *P, double T, double salt, int dh_half, int use_mbf, double mbf_fmax, double mbf_finf) {
	memset(P, 0, sizeof(*P));
	P->T = T;
	/* sites: src/model.h:14-18 (oxDNA2 "major-minor grooving" backbone), DNANucleotide.cpp:76-80 */
	P->back_a1 = -0.3400f; P->back_a2 = 0.3408f; P->stack_a1 = 0.34f;
	P->base_a1 = P->stack_a1 * ((double) 0.4f / (double) 0.34f);
	P->backref_a1 = -0.4f;
	/* FENE: model.h:46-50; max_backbone_force: DNAInteraction.cpp:258-276 */
	P->fene_eps = 2.0f; P->fene_r0 = 0.7564f; P->fene_delta = 0.25f; P->fene_delta2 = 0.0625f;
	P->use_mbf = use_mbf;
	if(use_mbf) {
		P->mbf_fmax = mbf_fmax;
		P->mbf_finf = mbf_finf;
		P->mbf_xmax = (-P->fene_eps + sqrt(P->fene_eps * P->fene_eps + 4.f * mbf_fmax * mbf_fmax * P->fene_delta2)) / (2.f * mbf_fmax);
	}
	/* excluded volume: model.h:55-88 */
	P->excl_eps = 2.0f;
	P->excl[0] = (oxo_excl){ 0.70f, 0.675f, 892.016223343f, 0.711879214356f };
	P->excl[1] = (oxo_excl){ 0.33f, 0.32f, 4119.70450017f, 0.335388426126f };
	P->excl[2] = (oxo_excl){ 0.515f, 0.50f, 1707.30627298f, 0.52329943261f };
	P->excl[3] = (oxo_excl){ 0.515f, 0.50f, 1707.30627298f, 0.52329943261f };
	/* hydrogen bonding radial part: model.h:93-106; oxDNA2 eps: DNA2Interaction.cpp:119 */
	P->hb.a = 8.f; P->hb.rc = 0.75f; P->hb.r0 = 0.4f; P->hb.blow = -126.243f; P->hb.bhigh = -7.87708f;
	P->hb.rlow = 0.34f; P->hb.rhigh = 0.7f; P->hb.rclow = 0.276908f; P->hb.rchigh = 0.783775f;
	fill_f1(&P->hb, 1.0678f);
	/* stacking radial part: model.h:153-166; eps(T): DNA2Interaction.cpp:115 */
	P->stck.a = 6.f; P->stck.rc = 0.9f; P->stck.r0 = 0.4f; P->stck.blow = -68.1857f; P->stck.bhigh = -3.12992f;
	P->stck.rlow = 0.32f; P->stck.rhigh = 0.75f; P->stck.rclow = 0.23239f; P->stck.rchigh = 0.956f;
	fill_f1(&P->stck, 1.3523f + 2.6717f * T);
	/* cross stacking / coaxial stacking radial parts: model.h:202-211, 369-381; oxDNA2 K: DNA2Interaction.cpp:10 */
	P->crst = (oxo_f2){ 47.5f, 0.675f, 0.575f, -0.888889f, 0.495f, 0.45f, -0.888889f, 0.655f, 0.7f };
	P->cxst = (oxo_f2){ 58.5f, 0.6f, 0.400f, -2.13158f, 0.22f, 0.177778f, -2.13158f, 0.58f, 0.6222222f };
	/* angular parts */
	set_f4(&P->stck_t4, 1.3f, 6.4381f, 0.f, 0.8f, 0.961538f);
	set_f4(&P->stck_t5, 0.9f, 3.89361f, 0.f, 0.95f, 1.16959f);
	set_f4(&P->hb_t1, 1.5f, 4.16038f, 0.f, 0.7f, 0.952381f);
	set_f4(&P->hb_t2, 1.5f, 4.16038f, 0.f, 0.7f, 0.952381f);
	set_f4(&P->hb_t4, 0.46f, 0.133855f, OXO_PI, 0.7f, 3.10559f);
	set_f4(&P->hb_t7, 4.f, 17.0526f, (OXO_PI * 0.5f), 0.45f, 0.555556f);
	set_f4(&P->crst_t1, 2.25f, 7.00545f, (OXO_PI - 2.35f), 0.58f, 0.766284f);
	set_f4(&P->crst_t2, 1.70f, 6.2469f, 1.f, 0.68f, 0.865052f);
	set_f4(&P->crst_t4, 1.50f, 2.59556f, 0.f, 0.65f, 1.02564f);
	set_f4(&P->crst_t7, 1.70f, 6.2469f, 0.875f, 0.68f, 0.865052f);
	set_f4(&P->cxst_t1, 2.f, 10.9032f, (OXO_PI - 0.25f), 0.65f, 0.769231f);
	set_f4(&P->cxst_t4, 1.3f, 6.4381f, 0.f, 0.8f, 0.961538f);
	set_f4(&P->cxst_t5, 0.9f, 3.89361f, 0.f, 0.95f, 1.16959f);
	P->cxst_t1_sa = 20.f;
	P->cxst_t1_sb = (OXO_PI - 0.1f * (OXO_PI - (OXO_PI - 0.25f)));
	P->stck_phi1 = (oxo_f5){ 2.0f, 10.9032f, -0.769231f, -0.65f };
	P->stck_phi2 = P->stck_phi1;
	/* Debye-Hueckel: DNA2Interaction.cpp:66-84 (get_settings) and :124-149 (init).  The smoothing onset RHIGH is
	 * computed in get_settings with a *float* 0.1f, lambda in init with a double 0.1 -- reproduced as is. */
	const double lfac = 0.3616455, q = 0.0543;
	salt = (double) (float) salt; /* DNA2Interaction.h:31: the salt concentration is stored in a float */
	double lambda_gs = lfac * sqrt(T / 0.1f) / sqrt(salt);
	double lambda = lfac * sqrt(T / 0.1) / sqrt(salt);
	P->dh_rhigh = 3.0 * lambda_gs;
	P->dh_minus_kappa = -1.0 / lambda;
	P->dh_prefactor = q;
	double x = P->dh_rhigh, l = lambda;
	P->dh_b = -(exp(-x / l) * q * q * (x + l) * (x + l)) / (-4. * x * x * x * l * l * q);
	P->dh_rc = x * (q * x + 3. * q * l) / (q * (x + l));
	P->dh_half_charged_ends = dh_half;
	P->hb_multiplier = 1.0;
	/* cutoff: DNAInteraction.cpp:296-311, DNA2Interaction.cpp:137-149 (float macros => float products) */
	double rcutback = 2 * sqrt((double) ((-0.3400f) * (-0.3400f) + (0.3408f) * (0.3408f))) + (double) 0.711879214356f;
	double rcutbase = 2 * fabs((double) 0.4f) + (double) 0.783775f;
	P->rcut = fmax(rcutback, rcutbase);
	double debyecut = 2.0 * sqrt((double) (SQ(-0.3400f) + SQ(0.3408f))) + P->dh_rc;
	if(debyecut > P->rcut) P->rcut = debyecut;
}

void oxo_dna1_params_init(oxo_dna2_params *P, double T, int grooving, int use_mbf, double mbf_fmax, double mbf_finf) {
	/* first-generation oxDNA (interaction_type = DNA, class DNAInteraction): src/Interactions/DNAInteraction.cpp:12-228,295-328 */
	oxo_dna2_params_init(P, T, 1.0, 0, use_mbf, mbf_fmax, mbf_finf);
	P->v1 = 1;
	if(!grooving) {
		/* DNANucleotide.cpp:83-87: STACK = BACK * (POS_STACK / POS_BACK), BASE = STACK * (POS_BASE / POS_STACK), float ratios */
		P->back_a1 = -0.4f; P->back_a2 = 0.;
		P->stack_a1 = (double) -0.4f * (double) (0.34f / -0.4f);
		P->base_a1 = P->stack_a1 * (double) (0.4f / 0.34f);
	}
	P->fene_r0 = 0.7525f;
	fill_f1(&P->hb, 1.077f);
	fill_f1(&P->stck, 1.3448f + 2.6568f * T);
	P->cxst.k = 46.0f;
	set_f4(&P->cxst_t1, 2.f, 10.9032f, (OXO_PI - 0.60f), 0.65f, 0.769231f);
	P->cxst_phi3 = (oxo_f5){ 2.0f, 10.9032f, -0.769231f, -0.65f };
	P->dh_prefactor = 0; P->dh_b = 0; P->dh_rc = 0; P->dh_rhigh = 0; P->dh_minus_kappa = 0;
	double rcutback = grooving ? 2 * sqrt((double) ((-0.3400f) * (-0.3400f) + (0.3408f) * (0.3408f))) + (double) 0.711879214356f
			: 2 * fabs((double) -0.4f) + (double) 0.711879214356f;
	double rcutbase = 2 * fabs((double) 0.4f) + (double) 0.783775f;
	P->rcut = fmax(rcutback, rcutbase);
}

void oxo_dna2_params_seqdep(oxo_dna2_params *P, const double *stck_raw16, double stck_fact_eps, double hb_AT, double hb_GC) {
	/* DNAInteraction.cpp:329-375; base order A=0 G=1 C=2 T=3 (src/defs.h) */
	double sh_st = SQ(1 - exp(-(double) ((float) (P->stck.rc - P->stck.r0) * (float) P->stck.a)));
	double sh_hb = SQ(1 - exp(-(double) ((float) (P->hb.rc - P->hb.r0) * (float) P->hb.a)));
	for(int i = 0; i < 4; i++) for(int j = 0; j < 4; j++) {
		P->stck.eps[i][j] = stck_raw16[4 * i + j] * (1.0 - stck_fact_eps + (P->T * 9.0 * stck_fact_eps));
		P->stck.shift[i][j] = P->stck.eps[i][j] * sh_st;
	}
	/* dummy base (type 4): the optional keys STCK_D_X / STCK_X_D are read into a variable that still holds the last mandatory key, STCK_T_T
	 * (DNAInteraction.cpp:349-359): with the stock parameter file every dummy entry gets the T-T strength */
	for(int i = 0; i < 5; i++) for(int j = 0; j < 5; j++) if(i == 4 || j == 4) { P->stck.eps[i][j] = P->stck.eps[3][3]; P->stck.shift[i][j] = P->stck.shift[3][3]; }
	P->hb.eps[0][3] = P->hb.eps[3][0] = hb_AT;
	P->hb.eps[1][2] = P->hb.eps[2][1] = hb_GC;
	P->hb.shift[0][3] = P->hb.shift[3][0] = hb_AT * sh_hb;
	P->hb.shift[1][2] = P->hb.shift[2][1] = hb_GC * sh_hb;
}

/* ------------------------------------------------------------------ modulation functions */
/* value and d/dr of f1 (DNAInteraction.cpp:1249-1283) */
static void f1_eval(const oxo_f1 *f, double r, int n3t, int n5t, double *v, double *d) {
	double eps = f->eps[n3t][n5t];
	*v = 0; *d = 0;
	if(r < f->rchigh) {
		if(r > f->rhigh) { *v = eps * f->bhigh * SQ(r - f->rchigh); *d = eps * 2 * f->bhigh * (r - f->rchigh); }
		else if(r > f->rlow) {
			double e = exp(-(r - f->r0) * f->a);
			*v = eps * SQ(1 - e) - f->shift[n3t][n5t];
			*d = eps * 2 * (1 - e) * e * f->a;
		}
		else if(r > f->rclow) { *v = eps * f->blow * SQ(r - f->rclow); *d = eps * 2 * f->blow * (r - f->rclow); }
	}
}

/* f2 (DNAInteraction.cpp:1285-1315) */
static void f2_eval(const oxo_f2 *f, double r, double *v, double *d) {
	*v = 0; *d = 0;
	if(r < f->rchigh) {
		if(r > f->rhigh) { *v = f->k * f->bhigh * SQ(r - f->rchigh); *d = 2. * f->k * f->bhigh * (r - f->rchigh); }
		else if(r > f->rlow) { *v = (f->k / 2.) * (SQ(r - f->r0) - SQ(f->rc - f->r0)); *d = f->k * (r - f->r0); }
		else if(r > f->rclow) { *v = f->k * f->blow * SQ(r - f->rclow); *d = 2. * f->k * f->blow * (r - f->rclow); }
	}
}

static double clamp_acos(double c) { return acos(fmax(-1.0, fmin(1.0, c))); }

/* f4 as a function of cos(theta): value and derivative with respect to the cosine.
 * d f4(acos c)/dc = -f4'(theta)/sin(theta), with the reference's small-angle guard (DNAInteraction.cpp:1351-1420). */
static void f4_eval(const oxo_f4 *f, double c, double *v, double *dc) {
	double t = clamp_acos(c);
	double x = t - f->t0, m = 1;
	if(x < 0) { x = -x; m = -1; }
	*v = 0; *dc = 0;
	if(x < f->tc) {
		double s = sin(t), dsin;
		if(x > f->ts) { *v = f->b * SQ(f->tc - x); dsin = m * 2 * f->b * (x - f->tc) / s; }
		else {
			*v = 1. - f->a * SQ(x);
			dsin = (SQ(s) > 1e-8) ? -m * 2 * f->a * x / s : -m * 2 * f->a;
		}
		*dc = -dsin;
	}
}

/* symmetrised version F(c) = f4(c) + f4(-c), used by cross and coaxial stacking */
static void f4_sym(const oxo_f4 *f, double c, double *v, double *dc) {
	double v1, d1, v2, d2;
	f4_eval(f, c, &v1, &d1);
	f4_eval(f, -c, &v2, &d2);
	*v = v1 + v2;
	*dc = d1 - d2;
}

/* oxDNA2 coaxial theta1: f4 plus a pure harmonic beyond t = sb (DNA2Interaction.cpp:308-363) */
static void f4_cxst_t1(const oxo_dna2_params *P, double c, double *v, double *dc) {
	if(c * c > 1) c = copysign(1, c);
	f4_eval(&P->cxst_t1, c, v, dc);
	double t = acos(c), x = t - P->cxst_t1_sb;
	if(x >= 0) {
		double s = sin(t);
		*v += P->cxst_t1_sa * SQ(x);
		double dsin = (SQ(s) > 1e-8) ? 2 * P->cxst_t1_sa * x / s : 2 * P->cxst_t1_sa;
		*dc += -dsin;
	}
}

/* f5 (DNAInteraction.cpp:1422-1454) */
static void f5_eval(const oxo_f5 *f, double c, double *v, double *d) {
	*v = 0; *d = 0;
	if(c > f->xc) {
		if(c < f->xs) { *v = f->b * SQ(f->xc - c); *d = 2 * f->b * (c - f->xc); }
		else if(c < 0) { *v = 1. - f->a * SQ(c); *d = -2. * f->a * c; }
		else { *v = 1; }
	}
}

/* ------------------------------------------------------------------ pair accumulator + generic chain rule */
typedef struct { double Fq[3], Tp[3], Tq[3]; } pacc; /* force on q (p gets -Fq), lab-frame torques */

/* a force F acting on q at lever sq and -F acting on p at lever sp */
static void site_force(pacc *A, const double *sp, const double *sq, const double *F) {
	double x[3];
	axpy3(1, F, A->Fq);
	cross3(sp, F, x); axpy3(-1, x, A->Tp);
	cross3(sq, F, x); axpy3(1, x, A->Tq);
}

/* c = u.v, u rigidly attached to p and v to q, g = dE/dc */
static void chain_body_body(pacc *A, double g, const double *u, const double *v) {
	double x[3];
	cross3(u, v, x);
	axpy3(-g, x, A->Tp);
	axpy3(g, x, A->Tq);
}

/* c = u.rhat with rhat the unit vector from site sp (on p) to site sq (on q), |r| = rmod; u attached to p (onq=0) or q (onq=1) */
static void chain_body_dir(pacc *A, double g, const double *u, const double *rhat, double rmod, double c, int onq,
		const double *sp, const double *sq) {
	double F[3], x[3];
	for(int k = 0; k < 3; k++) F[k] = -g * (u[k] - c * rhat[k]) / rmod;
	site_force(A, sp, sq, F);
	cross3(u, rhat, x);
	axpy3(-g, x, onq ? A->Tq : A->Tp);
}

/* ------------------------------------------------------------------ angular factors, angle based (RNAInteraction.cpp:1304-1371) */
static double rna_acos(double c) { return (c > 1) ? 0. : ((c < -1) ? (double) OXO_PI : acos(c)); } /* LRACOS, src/defs.h:17 */

static double f4_val(const oxo_f4 *f, double t) {
	t -= f->t0;
	if(t < 0) t = -t;
	if(t < f->tc) return (t > f->ts) ? f->b * SQ(f->tc - t) : 1. - f->a * SQ(t);
	return 0;
}

/* f4'(t) / sin(t) with the reference's small-angle guard */
static double f4_dsin(const oxo_f4 *f, double t) {
	double x = t - f->t0, m = 1;
	if(x < 0) { x = -x; m = -1; }
	if(x < f->tc) {
		double s = sin(t);
		if(x > f->ts) return m * 2 * f->b * (x - f->tc) / s;
		return (SQ(s) > 1e-8) ? -m * 2 * f->a * x / s : -m * 2 * f->a;
	}
	return 0;
}


/* ------------------------------------------------------------------ the cubic mesh of the reference (src/Interactions/Mesh.{h,cpp})
 * The CPU classes interpolate some f4 factors in cos(theta) on meshes built from the analytic function and its derivative
 * (DNAInteraction.cpp:199-213, RNAInteraction.cpp:311-325); the CUDA kernels -- the reference's and ours -- evaluate the analytic form.
 * Restated so that the oracle can reproduce the CPU class to rounding where a test wants that (oxDNA3: cxst_mesh, oxRNA: cpu_quirks bit 2). */
typedef struct { int n; double xlow, xupp, delta, inv_sqr_delta, A[256], B[256], C[256], D[256]; } oxo_mesh;

/* mirror2pi: the first-generation coaxial theta1, f4(t) + f4(2 PI - t) (DNAInteraction.cpp:1335-1351) */
static void mesh_build_mode(oxo_mesh *m, const oxo_f4 *f, int npoints, int mirror2pi) {
	const double xupp = cos(fmax(0., f->t0 - f->tc)), xlow = cos(fmin((double) OXO_PI, f->t0 + f->tc));
	const double dx = (xupp - xlow) / (double) npoints;
	m->n = npoints; m->xlow = xlow; m->xupp = xupp; m->delta = dx; m->inv_sqr_delta = 1 / SQ(dx);
	for(int i = 0; i < npoints + 1; i++) {
		const double x = xlow + i * dx;
		/* _fakef4 / _fakef4D: f4(acos x) and -f4Dsin(acos x); beyond x = 1 (the unused far side of the last entry) the argument is clamped */
		const double t0 = clamp_acos(x), t1 = clamp_acos(x + dx);
		double fx0 = f4_val(f, t0), fx1 = f4_val(f, t1), d0 = -f4_dsin(f, t0), d1 = -f4_dsin(f, t1);
		if(mirror2pi) {
			fx0 += f4_val(f, 2 * OXO_PI - t0); fx1 += f4_val(f, 2 * OXO_PI - t1);
			d0 -= f4_dsin(f, 2 * OXO_PI - t0); d1 -= f4_dsin(f, 2 * OXO_PI - t1);
		}
		m->A[i] = fx0; m->B[i] = d0;
		m->D[i] = (2 * (fx0 - fx1) + (d0 + d1) * dx) / dx;
		m->C[i] = (fx1 - fx0 + (-d0 - m->D[i]) * dx);
	}
}
static void mesh_build(oxo_mesh *m, const oxo_f4 *f, int npoints) { mesh_build_mode(m, f, npoints, 0); }
static double mesh_query(const oxo_mesh *m, double x) {
	if(x <= m->xlow) return m->A[0];
	if(x >= m->xupp) x = m->xupp - FLT_EPSILON;
	const int i = (int) ((x - m->xlow) / m->delta);
	const double dx = x - m->xlow - m->delta * i;
	return m->A[i] + dx * (m->B[i] + dx * (m->C[i] + dx * m->D[i]) * m->inv_sqr_delta);
}
static double mesh_query_derivative(const oxo_mesh *m, double x) {
	if(x < m->xlow) return m->B[0];
	if(x >= m->xupp) x = m->xupp - FLT_EPSILON;
	const int i = (int) ((x - m->xlow) / m->delta);
	const double dx = x - m->xlow - m->delta * i;
	return m->B[i] + (2 * dx * m->C[i] + 3 * dx * dx * m->D[i]) * m->inv_sqr_delta;
}

/* f4 through the CPU class's mesh when P->mesh is set (DNAInteraction::_custom_f4, DNAInteraction.cpp:1173-1180; mesh sizes: model.h:410-433),
 * analytic otherwise.  Meshes are cached by their parameters. */
static const oxo_mesh *mesh_for(const oxo_f4 *f, int npoints) {
	static struct { oxo_f4 key; int n; oxo_mesh m; } cache[24];
	static int used = 0;
	for(int i = 0; i < used; i++) if(cache[i].n == npoints && memcmp(&cache[i].key, f, sizeof(oxo_f4)) == 0) return &cache[i].m;
	const int slot = used < 24 ? used++ : 0;
	cache[slot].key = *f; cache[slot].n = npoints;
	mesh_build(&cache[slot].m, f, npoints);
	return &cache[slot].m;
}
static void f4m_eval(int mesh, const oxo_f4 *f, int npoints, double c, double *v, double *dc) {
	if(!mesh) { f4_eval(f, c, v, dc); return; }
	const oxo_mesh *m = mesh_for(f, npoints);
	*v = mesh_query(m, c); *dc = mesh_query_derivative(m, c);
}
static void f4m_sym(int mesh, const oxo_f4 *f, int npoints, double c, double *v, double *dc) {
	double v1, d1, v2, d2;
	f4m_eval(mesh, f, npoints, c, &v1, &d1);
	f4m_eval(mesh, f, npoints, -c, &v2, &d2);
	*v = v1 + v2; *dc = d1 - d2;
}

/* c = shat . (bhat x u), u a body vector of p (onq = 0) or q (onq = 1); shat between the stacking sites, bhat between the
 * backbone sites; g = dE/dc.  RNAInteraction.cpp:1093-1142 */
static void chain_triple(pacc *A, double g, const double *u, int onq, const double *sh, double sm, const double *ssp, const double *ssq,
		const double *bh, double bm, const double *bsp, const double *bsq) {
	double bu[3], us[3], sb[3], F[3], x[3];
	cross3(bh, u, bu);  /* c = shat . bu */
	cross3(u, sh, us);  /* c = bhat . us */
	cross3(sh, bh, sb); /* c = u . sb    */
	double c = dot3(sh, bu);
	for(int k = 0; k < 3; k++) F[k] = -g * (bu[k] - c * sh[k]) / sm;
	site_force(A, ssp, ssq, F);
	for(int k = 0; k < 3; k++) F[k] = -g * (us[k] - c * bh[k]) / bm;
	site_force(A, bsp, bsq, F);
	cross3(u, sb, x);
	axpy3(-g, x, onq ? A->Tq : A->Tp);
}


static double excl_eval(const oxo_dna2_params *P, const oxo_excl *e, const double *r, double *F) {
	/* DNAInteraction.cpp:1182-1205; F = force on q */
	double r2 = dot3(r, r), en = 0;
	F[0] = F[1] = F[2] = 0;
	if(r2 < SQ(e->rc)) {
		if(r2 > SQ(e->rstar)) {
			double rm = sqrt(r2), rrc = rm - e->rc;
			en = P->excl_eps * e->b * SQ(rrc);
			double s = -(2 * P->excl_eps * e->b * rrc / rm);
			for(int k = 0; k < 3; k++) F[k] = s * r[k];
		}
		else {
			double t = SQ(e->sigma) / r2, lj = t * t * t;
			en = 4 * P->excl_eps * (SQ(lj) - lj);
			double s = -(24 * P->excl_eps * (lj - 2 * SQ(lj)) / r2);
			for(int k = 0; k < 3; k++) F[k] = s * r[k];
		}
	}
	if(en == 0) F[0] = F[1] = F[2] = 0;
	return en;
}

typedef struct { double back[3], stack[3], base[3], backref[3]; const double *a1, *a2, *a3; } sites_t;

static void make_sites(const oxo_dna2_params *P, const double *ax, sites_t *S) {
	S->a1 = ax; S->a2 = ax + 3; S->a3 = ax + 6;
	for(int k = 0; k < 3; k++) {
		S->back[k] = S->a1[k] * P->back_a1 + S->a2[k] * P->back_a2;
		S->stack[k] = S->a1[k] * P->stack_a1;
		S->base[k] = S->stack[k] * (P->base_a1 / P->stack_a1);
		S->backref[k] = S->a1[k] * P->backref_a1;
	}
}

static void site_sep(const double *r, const double *sp, const double *sq, double *out, double *mod, double *hat) {
	for(int k = 0; k < 3; k++) out[k] = r[k] + sq[k] - sp[k];
	*mod = sqrt(dot3(out, out));
	if(hat) for(int k = 0; k < 3; k++) hat[k] = out[k] / *mod;
}

static void neg3(const double *a, double *o) { o[0] = -a[0]; o[1] = -a[1]; o[2] = -a[2]; }

/* ------------------------------------------------------------------ bonded pair p -> q = n3(p) */
static void bonded_pair(const oxo_dna2_params *P, const double *r, const sites_t *sp, const sites_t *sq, int tp, int tq,
		pacc *A, double *e) {
	double v[3], m, F[3];
	/* FENE on the backbone sites: DNAInteraction.cpp:415-466 */
	site_sep(r, sp->back, sq->back, v, &m, NULL);
	double x = m - P->fene_r0, en, s;
	if(P->use_mbf && fabs(x) > P->mbf_xmax) {
		double fene_xmax = -(P->fene_eps / 2.f) * log(1.f - SQ(P->mbf_xmax) / P->fene_delta2);
		double long_xmax = (P->mbf_fmax - P->mbf_finf) * P->mbf_xmax * log(P->mbf_xmax) + P->mbf_finf * P->mbf_xmax;
		en = (P->mbf_fmax - P->mbf_finf) * P->mbf_xmax * log(fabs(x)) + P->mbf_finf * fabs(x) - long_xmax + fene_xmax;
		s = -copysign(1.f, x) * ((P->mbf_fmax - P->mbf_finf) * P->mbf_xmax / fabs(x) + P->mbf_finf) / m;
	}
	else if(fabs(x) > P->fene_delta - DBL_EPSILON) {
		en = 1.e12; s = 0;
	}
	else {
		en = -(P->fene_eps / 2.f) * log(1.f - SQ(x) / P->fene_delta2);
		s = -(P->fene_eps * x / (P->fene_delta2 - SQ(x))) / m;
	}
	for(int k = 0; k < 3; k++) F[k] = s * v[k];
	site_force(A, sp->back, sq->back, F);
	e[OXO_FENE] += en;

	/* bonded excluded volume: DNAInteraction.cpp:468-528 */
	site_sep(r, sp->base, sq->base, v, &m, NULL);
	e[OXO_BEXC] += excl_eval(P, &P->excl[1], v, F); site_force(A, sp->base, sq->base, F);
	site_sep(r, sp->base, sq->back, v, &m, NULL);
	e[OXO_BEXC] += excl_eval(P, &P->excl[2], v, F); site_force(A, sp->base, sq->back, F);
	site_sep(r, sp->back, sq->base, v, &m, NULL);
	e[OXO_BEXC] += excl_eval(P, &P->excl[3], v, F); site_force(A, sp->back, sq->base, F);

	/* stacking: DNAInteraction.cpp:530-705.  Angles: t4 = (a3,b3), t5 = (a3,-rhat), t6 = (b3,-rhat); phi1, phi2 are
	 * cosines of a2, b2 against the unit vector between the *ungrooved* backbone sites (-0.4 a1). */
	double rst[3], rstm, rsth[3], w[3], wm, wh[3];
	site_sep(r, sp->stack, sq->stack, rst, &rstm, rsth);
	site_sep(r, sp->backref, sq->backref, w, &wm, wh);
	double ma3[3], mb3[3];
	neg3(sp->a3, ma3); neg3(sq->a3, mb3);
	double c4 = dot3(sp->a3, sq->a3), c5 = dot3(ma3, rsth), c6 = dot3(mb3, rsth);
	double cp1 = dot3(sp->a2, wh), cp2 = dot3(sq->a2, wh);
	double f1, f1d, g4, g4d, g5, g5d, g6, g6d, h1, h1d, h2, h2d;
	f1_eval(&P->stck, rstm, tq, tp, &f1, &f1d);
	f4m_eval(P->mesh, &P->stck_t4, 250, c4, &g4, &g4d);
	f4m_eval(P->mesh, &P->stck_t5, 250, c5, &g5, &g5d);
	f4m_eval(P->mesh, &P->stck_t5, 250, c6, &g6, &g6d);
	f5_eval(&P->stck_phi1, cp1, &h1, &h1d);
	f5_eval(&P->stck_phi2, cp2, &h2, &h2d);
	double E = f1 * g4 * g5 * g6 * h1 * h2;
	e[OXO_STCK] += E;
	if(E != 0.) {
		for(int k = 0; k < 3; k++) F[k] = -rsth[k] * (f1d * g4 * g5 * g6 * h1 * h2);
		site_force(A, sp->stack, sq->stack, F);
		chain_body_body(A, f1 * g4d * g5 * g6 * h1 * h2, sp->a3, sq->a3);
		chain_body_dir(A, f1 * g4 * g5d * g6 * h1 * h2, ma3, rsth, rstm, c5, 0, sp->stack, sq->stack);
		chain_body_dir(A, f1 * g4 * g5 * g6d * h1 * h2, mb3, rsth, rstm, c6, 1, sp->stack, sq->stack);
		chain_body_dir(A, f1 * g4 * g5 * g6 * h1d * h2, sp->a2, wh, wm, cp1, 0, sp->backref, sq->backref);
		chain_body_dir(A, f1 * g4 * g5 * g6 * h1 * h2d, sq->a2, wh, wm, cp2, 1, sp->backref, sq->backref);
	}
}

/* ------------------------------------------------------------------ non-bonded pair */
static void nonbonded_pair(const oxo_dna2_params *P, const double *r, const sites_t *sp, const sites_t *sq, int btp, int btq,
		int tp, int tq, int p_end, int q_end, pacc *A, double *e) {
	double v[3], m, h[3], F[3];
	/* excluded volume, four site pairs: DNAInteraction.cpp:707-777 */
	site_sep(r, sp->base, sq->base, v, &m, NULL);
	e[OXO_NEXC] += excl_eval(P, &P->excl[1], v, F); site_force(A, sp->base, sq->base, F);
	site_sep(r, sp->back, sq->base, v, &m, NULL);
	e[OXO_NEXC] += excl_eval(P, &P->excl[3], v, F); site_force(A, sp->back, sq->base, F);
	site_sep(r, sp->base, sq->back, v, &m, NULL);
	e[OXO_NEXC] += excl_eval(P, &P->excl[2], v, F); site_force(A, sp->base, sq->back, F);
	site_sep(r, sp->back, sq->back, v, &m, NULL);
	e[OXO_NEXC] += excl_eval(P, &P->excl[0], v, F); site_force(A, sp->back, sq->back, F);

	/* base-base vector shared by HB and cross stacking */
	site_sep(r, sp->base, sq->base, v, &m, h);
	double ma1[3], mb1[3], mb3[3];
	neg3(sp->a1, ma1); neg3(sq->a1, mb1); neg3(sq->a3, mb3);
	double c1 = dot3(ma1, sq->a1), c2 = dot3(mb1, h), c3 = dot3(sp->a1, h);
	double c4 = dot3(sp->a3, sq->a3), c7 = dot3(mb3, h), c8 = dot3(sp->a3, h);

	/* hydrogen bonding: DNAInteraction.cpp:779-902 */
	if(btp + btq == 3 && P->hb.rclow < m && m < P->hb.rchigh) {
		double mult = (abs(btq) >= 300 && abs(btp) >= 300) ? P->hb_multiplier : 1.;
		double f1, f1d, g[6], d[6];
		f1_eval(&P->hb, m, tq, tp, &f1, &f1d);
		f1 *= mult; f1d *= mult;
		f4m_eval(P->mesh, &P->hb_t1, 6, c1, &g[0], &d[0]);
		f4m_eval(P->mesh, &P->hb_t2, 6, c2, &g[1], &d[1]);
		f4m_eval(P->mesh, &P->hb_t2, 6, c3, &g[2], &d[2]);
		f4m_eval(P->mesh, &P->hb_t4, 250, c4, &g[3], &d[3]);
		f4m_eval(P->mesh, &P->hb_t7, 12, c7, &g[4], &d[4]);
		f4m_eval(P->mesh, &P->hb_t7, 12, c8, &g[5], &d[5]);
		double E = f1 * g[0] * g[1] * g[2] * g[3] * g[4] * g[5];
		e[OXO_HB] += E;
		if(E != 0.) {
			double rest[6];
			for(int i = 0; i < 6; i++) { rest[i] = f1 * d[i]; for(int j = 0; j < 6; j++) if(j != i) rest[i] *= g[j]; }
			for(int k = 0; k < 3; k++) F[k] = -h[k] * (f1d * g[0] * g[1] * g[2] * g[3] * g[4] * g[5]);
			site_force(A, sp->base, sq->base, F);
			chain_body_body(A, rest[0], ma1, sq->a1);
			chain_body_dir(A, rest[1], mb1, h, m, c2, 1, sp->base, sq->base);
			chain_body_dir(A, rest[2], sp->a1, h, m, c3, 0, sp->base, sq->base);
			chain_body_body(A, rest[3], sp->a3, sq->a3);
			chain_body_dir(A, rest[4], mb3, h, m, c7, 1, sp->base, sq->base);
			chain_body_dir(A, rest[5], sp->a3, h, m, c8, 0, sp->base, sq->base);
		}
	}

	/* cross stacking: DNAInteraction.cpp:904-1021 */
	if(P->crst.rclow < m && m < P->crst.rchigh) {
		double f2, f2d, g[6], d[6];
		f2_eval(&P->crst, m, &f2, &f2d);
		f4m_eval(P->mesh, &P->crst_t1, 250, c1, &g[0], &d[0]);
		f4m_eval(P->mesh, &P->crst_t2, 250, c2, &g[1], &d[1]);
		f4m_eval(P->mesh, &P->crst_t2, 250, c3, &g[2], &d[2]);
		f4m_sym(P->mesh, &P->crst_t4, 6, c4, &g[3], &d[3]);
		f4m_sym(P->mesh, &P->crst_t7, 250, c7, &g[4], &d[4]);
		f4m_sym(P->mesh, &P->crst_t7, 250, c8, &g[5], &d[5]);
		double E = f2 * g[0] * g[1] * g[2] * g[3] * g[4] * g[5];
		e[OXO_CRST] += E;
		if(E != 0.) {
			double rest[6];
			for(int i = 0; i < 6; i++) { rest[i] = f2 * d[i]; for(int j = 0; j < 6; j++) if(j != i) rest[i] *= g[j]; }
			for(int k = 0; k < 3; k++) F[k] = -h[k] * (f2d * g[0] * g[1] * g[2] * g[3] * g[4] * g[5]);
			site_force(A, sp->base, sq->base, F);
			chain_body_body(A, rest[0], ma1, sq->a1);
			chain_body_dir(A, rest[1], mb1, h, m, c2, 1, sp->base, sq->base);
			chain_body_dir(A, rest[2], sp->a1, h, m, c3, 0, sp->base, sq->base);
			chain_body_body(A, rest[3], sp->a3, sq->a3);
			chain_body_dir(A, rest[4], mb3, h, m, c7, 1, sp->base, sq->base);
			chain_body_dir(A, rest[5], sp->a3, h, m, c8, 0, sp->base, sq->base);
		}
	}

	/* coaxial stacking, oxDNA2 form: DNA2Interaction.cpp:212-306 */
	site_sep(r, sp->stack, sq->stack, v, &m, h);
	if(P->cxst.rclow < m && m < P->cxst.rchigh) {
		double c5 = dot3(sp->a3, h), c6 = dot3(mb3, h);
		double f2, f2d, g[4], d[4];
		f2_eval(&P->cxst, m, &f2, &f2d);
		if(P->v1 && P->mesh) {
			/* the meshed class interpolates this combination too (250 intervals, DNAInteraction.cpp:210-212) */
			static oxo_mesh M; static oxo_f4 built; static int have = 0;
			if(!have || memcmp(&built, &P->cxst_t1, sizeof(oxo_f4))) { mesh_build_mode(&M, &P->cxst_t1, 250, 1); built = P->cxst_t1; have = 1; }
			g[0] = mesh_query(&M, c1); d[0] = mesh_query_derivative(&M, c1);
		}
		else if(P->v1) {
			/* oxDNA1: f4(t1) + f4(2 PI - t1) (DNAInteraction.cpp:1331-1352) */
			double t = rna_acos(c1);
			g[0] = f4_val(&P->cxst_t1, t) + f4_val(&P->cxst_t1, 2 * OXO_PI - t);
			d[0] = -f4_dsin(&P->cxst_t1, t) - f4_dsin(&P->cxst_t1, 2 * OXO_PI - t);
		}
		else f4_cxst_t1(P, c1, &g[0], &d[0]);
		f4m_eval(P->mesh, &P->cxst_t4, 6, c4, &g[1], &d[1]);
		f4m_sym(P->mesh, &P->cxst_t5, 6, c5, &g[2], &d[2]);
		f4m_sym(P->mesh, &P->cxst_t5, 6, c6, &g[3], &d[3]);
		/* oxDNA1: times f5(cos phi3)^2, cos phi3 = shat . (bhat_ref x a1) with bhat_ref between the UNGROOVED backbone
		 * reference sites (DNAInteraction.cpp:1049-1062) */
		double wv[3], wm = 1, wh[3] = { 0, 0, 0 }, f5v = 1, f5d = 0;
		if(P->v1) {
			double x3[3];
			site_sep(r, sp->backref, sq->backref, wv, &wm, wh);
			cross3(wh, sp->a1, x3);
			f5_eval(&P->cxst_phi3, dot3(h, x3), &f5v, &f5d);
			f2 *= f5v * f5v; f2d *= f5v * f5v;
		}
		double E = f2 * g[0] * g[1] * g[2] * g[3];
		e[OXO_CXST] += E;
		if(E != 0.) {
			double rest[4];
			for(int i = 0; i < 4; i++) { rest[i] = f2 * d[i]; for(int j = 0; j < 4; j++) if(j != i) rest[i] *= g[j]; }
			for(int k = 0; k < 3; k++) F[k] = -h[k] * (f2d * g[0] * g[1] * g[2] * g[3]);
			site_force(A, sp->stack, sq->stack, F);
			chain_body_body(A, rest[0], ma1, sq->a1);
			chain_body_body(A, rest[1], sp->a3, sq->a3);
			chain_body_dir(A, rest[2], sp->a3, h, m, c5, 0, sp->stack, sq->stack);
			chain_body_dir(A, rest[3], mb3, h, m, c6, 1, sp->stack, sq->stack);
			if(P->v1 && f5d != 0.) chain_triple(A, (E / f5v) * 2 * f5d, sp->a1, 0, h, m, sp->stack, sq->stack, wh, wm, sp->backref, sq->backref);
		}
	}

	/* Debye-Hueckel on the backbone sites: DNA2Interaction.cpp:157-210 */
	site_sep(r, sp->back, sq->back, v, &m, h);
	if(m < P->dh_rc) {
		double cut = 1.0f;
		if(P->dh_half_charged_ends && p_end) cut *= 0.5f;
		if(P->dh_half_charged_ends && q_end) cut *= 0.5f;
		double en, fs;
		if(m < P->dh_rhigh) {
			en = exp(m * P->dh_minus_kappa) * (P->dh_prefactor / m);
			fs = -1.0f * (P->dh_prefactor * exp(P->dh_minus_kappa * m)) * (P->dh_minus_kappa / m - 1.0f / SQ(m));
		}
		else {
			en = P->dh_b * SQ(m - P->dh_rc);
			fs = -(2.0f * P->dh_b * (m - P->dh_rc));
		}
		e[OXO_DH] += en * cut;
		for(int k = 0; k < 3; k++) F[k] = h[k] * fs * cut;
		site_force(A, sp->back, sq->back, F);
	}
}

/* base 'D' of the topology: btype = type = N_DUMMY = 4 (src/Utilities/TopologyParser.cpp:96-99, Utils.cpp:33-34); numeric bases: TopologyParser.cpp:101-109 */
static inline int type_of(int btype) { return (btype == 4) ? 4 : ((btype < 0) ? 3 - ((3 - btype) % 4) : btype % 4); }

static void min_image(const double *box, const double *p, const double *q, double *r) {
	/* src/Boxes/CubicBox.cpp:51-57 (and OrthogonalBox) */
	for(int k = 0; k < 3; k++) r[k] = q[k] - p[k] - rint((q[k] - p[k]) / box[k]) * box[k];
}

static void scatter(const pacc *A, int p, int q, double half_e, double *force, double *tl, double *epart) {
	for(int k = 0; k < 3; k++) {
		if(force) { force[3 * p + k] -= A->Fq[k]; force[3 * q + k] += A->Fq[k]; }
		if(tl) { tl[3 * p + k] += A->Tp[k]; tl[3 * q + k] += A->Tq[k]; }
	}
	if(epart) { epart[p] += half_e; epart[q] += half_e; }
}

void oxo_dna2_forces(const oxo_dna2_params *P, int N, const double *pos, const double *axes, const int *btype,
		const int *n3, const int *n5, const double *box, const int *pairs, long long npairs,
		double *force, double *torque_lab, double *torque_body, double *eterms, double *epart) {
	double *tl = torque_lab ? torque_lab : (torque_body ? (double *) calloc(3 * (size_t) N, sizeof(double)) : NULL);
	double et[OXO_NTERMS] = { 0 };
	if(force) memset(force, 0, 3 * (size_t) N * sizeof(double));
	if(torque_lab) memset(torque_lab, 0, 3 * (size_t) N * sizeof(double));
	if(epart) memset(epart, 0, (size_t) N * sizeof(double));
	sites_t *S = (sites_t *) malloc((size_t) N * sizeof(sites_t));
	for(int i = 0; i < N; i++) make_sites(P, axes + 9 * (size_t) i, &S[i]);

	for(int p = 0; p < N; p++) {
		int q = n3[p];
		if(q < 0) continue;
		double r[3] = { pos[3 * q] - pos[3 * p], pos[3 * q + 1] - pos[3 * p + 1], pos[3 * q + 2] - pos[3 * p + 2] };
		pacc A; memset(&A, 0, sizeof(A));
		double e[OXO_NTERMS] = { 0 };
		bonded_pair(P, r, &S[p], &S[q], type_of(btype[p]), type_of(btype[q]), &A, e);
		double tot = 0;
		for(int t = 0; t < OXO_NTERMS; t++) { et[t] += e[t]; tot += e[t]; }
		scatter(&A, p, q, 0.5 * tot, force, tl, epart);
	}
	double rc2 = SQ(P->rcut);
	for(long long i = 0; i < npairs; i++) {
		/* the reference evaluates pair (p, q) with p the higher index (Cells.cpp:163) */
		int a = pairs[2 * i], b = pairs[2 * i + 1];
		int p = a > b ? a : b, q = a > b ? b : a;
		if(n3[p] == q || n5[p] == q) continue;
		double r[3];
		min_image(box, pos + 3 * (size_t) p, pos + 3 * (size_t) q, r);
		if(dot3(r, r) >= rc2) continue;
		pacc A; memset(&A, 0, sizeof(A));
		double e[OXO_NTERMS] = { 0 };
		nonbonded_pair(P, r, &S[p], &S[q], btype[p], btype[q], type_of(btype[p]), type_of(btype[q]),
				n3[p] < 0 || n5[p] < 0, n3[q] < 0 || n5[q] < 0, &A, e);
		double tot = 0;
		for(int t = 0; t < OXO_NTERMS; t++) { et[t] += e[t]; tot += e[t]; }
		scatter(&A, p, q, 0.5 * tot, force, tl, epart);
	}
	if(torque_body) {
		for(int i = 0; i < N; i++) {
			const double *ax = axes + 9 * (size_t) i;
			for(int k = 0; k < 3; k++) torque_body[3 * i + k] = dot3(ax + 3 * k, tl + 3 * (size_t) i);
		}
	}
	if(eterms) memcpy(eterms, et, sizeof(et));
	if(tl && tl != torque_lab) free(tl);
	free(S);
}

/* ------------------------------------------------------------------ external forces */
static void ext_one(const oxo_ext_force *f, const double *pos, int p, const double *box, long long step, double *force) {
	const double *pp = pos + 3 * (size_t) p;
	double *F = force + 3 * (size_t) p;
	if(f->type == OXO_EXT_STRING) {
		/* src/Forces/ConstantRateForce.cpp:52-61 */
		double s = f->F0 + f->rate * step;
		if(f->pbc) {
			/* dir_as_centre: the force points from the particle to the point pos0 (ConstantRateForce.cpp:54-61) */
			double d[3] = { f->pos0[0] - pp[0], f->pos0[1] - pp[1], f->pos0[2] - pp[2] };
			double m = sqrt(dot3(d, d));
			for(int k = 0; k < 3; k++) F[k] += s * d[k] / m;
		}
		else axpy3(s, f->dir, F);
	}
	else if(f->type == OXO_EXT_TRAP) {
		/* src/Forces/MovingTrap.cpp:50-64 */
		for(int k = 0; k < 3; k++) F[k] += -f->stiff * (pp[k] - (f->pos0[k] + (f->rate * step) * f->dir[k]));
	}
	else if(f->type == OXO_EXT_LOWDIM) {
		/* src/Forces/LowdimMovingTrap.cpp:68-82 */
		for(int k = 0; k < 3; k++) {
			double trap = ((f->iaux >> k) & 1) ? f->pos0[k] + (f->rate * step) * f->dir[k] : pp[k];
			F[k] += -f->stiff * (pp[k] - trap);
		}
	}
	else if(f->type == OXO_EXT_MUTUAL) {
		/* src/Forces/MutualTrap.cpp:54-66 */
		const double *qq = pos + 3 * (size_t) f->ref;
		double dr[3];
		if(f->pbc) min_image(box, pp, qq, dr);
		else for(int k = 0; k < 3; k++) dr[k] = qq[k] - pp[k];
		double m = sqrt(dot3(dr, dr));
		double s = (m - (f->r0 + (f->rate * step))) * (f->stiff + (f->stiff_rate * step));
		for(int k = 0; k < 3; k++) F[k] += (dr[k] / m) * s;
	}
	else if(f->type == OXO_EXT_REPULSION_PLANE) {
		/* src/Forces/RepulsionPlane.cpp:44-58 */
		double position = f->aux[0] + f->aux[1] * step, end = f->aux[2], start = f->aux[0];
		if(end > start && position > end) position = end;
		if(end < start && position < end) position = end;
		double d = dot3(f->dir, pp) + position;
		if(d < 0.) axpy3(-(d * f->stiff), f->dir, F);
	}
	else if(f->type == OXO_EXT_ATTRACTION_PLANE) {
		/* src/Forces/AttractionPlane.cpp:45-55 */
		double d = dot3(f->dir, pp) + f->aux[0];
		if(d >= 0.) axpy3(-f->stiff * 1.0, f->dir, F);
		else axpy3(-(d * f->stiff), f->dir, F);
	}
	else if(f->type == OXO_EXT_SPHERE) {
		/* src/Forces/RepulsiveSphere.cpp:46-53 */
		double d[3];
		min_image(box, f->pos0, pp, d);
		double m = sqrt(dot3(d, d)), radius = f->r0 + f->rate * (double) step;
		if(!(m <= radius || m >= f->aux[0])) axpy3(-f->stiff * (1. - radius / m), d, F);
	}
	else if(f->type == OXO_EXT_LJ_WALL) {
		/* src/Forces/LJWall.cpp:62-68 */
		double d = dot3(f->dir, pp) + f->aux[0], rel = d / f->aux[1];
		if(!(rel > f->aux[2])) {
			double lj = pow(rel, -f->iaux);
			axpy3(4 * f->iaux * f->stiff * (2 * SQ(lj) - lj) / d, f->dir, F);
		}
	}
	else if(f->type == OXO_EXT_TWIST) {
		/* src/Forces/ConstantRateTorque.cpp:76-103 */
		double t = f->F0 + f->rate * (double) step, sn = sin(t), cs = cos(t), oc = 1. - cs;
		const double *a = f->dir, *c = f->aux, *mask = f->aux + 3;
		double v[3] = { f->pos0[0] - c[0], f->pos0[1] - c[1], f->pos0[2] - c[2] };
		double R[3][3] = { { a[0] * a[0] * oc + cs, a[0] * a[1] * oc - a[2] * sn, a[0] * a[2] * oc + a[1] * sn },
				{ a[0] * a[1] * oc + a[2] * sn, a[1] * a[1] * oc + cs, a[1] * a[2] * oc - a[0] * sn },
				{ a[0] * a[2] * oc - a[1] * sn, a[1] * a[2] * oc + a[0] * sn, a[2] * a[2] * oc + cs } };
		for(int k = 0; k < 3; k++) {
			double trap = R[k][0] * v[0] + R[k][1] * v[1] + R[k][2] * v[2] + c[k];
			F[k] += -f->stiff * (pp[k] - trap) * mask[k];
		}
	}
	else if(f->type == OXO_EXT_SPHERE_SMOOTH) {
		/* src/Forces/RepulsiveSphereSmooth.cpp:48-63 */
		double d[3];
		min_image(box, f->pos0, pp, d);
		double m = sqrt(dot3(d, d)), r_ext = f->aux[0], smooth = f->aux[1], alpha = f->aux[2];
		if(!(m < f->r0 || m > r_ext)) {
			if(m >= alpha && m <= r_ext) axpy3(-(f->stiff * 0.5 * exp((m - alpha) / smooth)) / m, d, F);
			else axpy3(-(f->stiff * m - f->stiff * 0.5 * exp(-(m - alpha) / smooth)) / m, d, F);
		}
	}
	else if(f->type == OXO_EXT_ELLIPSOID) {
		/* src/Forces/RepulsiveEllipsoid.cpp:55-67 */
		double d[3];
		min_image(box, f->pos0, pp, d);
		const double *r2 = f->aux, *r1 = f->aux + 3;
		double in = SQ(d[0]) / SQ(r2[0]) + SQ(d[1]) / SQ(r2[1]) + SQ(d[2]) / SQ(r2[2]);
		double out = SQ(d[0]) / SQ(r1[0]) + SQ(d[1]) / SQ(r1[1]) + SQ(d[2]) / SQ(r1[2]);
		if(!(in < 1. && out > 1.)) axpy3(-f->stiff / sqrt(dot3(d, d)), d, F);
	}
}

/* further single-particle types */
static void ext_more(const oxo_ext_force *f, const double *pos, int p, const double *box, long long step, double *force) {
	const double *pp = pos + 3 * (size_t) p;
	double *F = force + 3 * (size_t) p;
	if(f->type == OXO_EXT_REPULSION_PLANE_MOVING) {
		/* src/Forces/RepulsionPlaneMoving.cpp:58-66 */
		for(int idx = f->ref; idx <= f->iaux; idx++) {
			const double *qq = pos + 3 * (size_t) idx;
			double d = (pp[0] - qq[0]) * f->dir[0] + (pp[1] - qq[1]) * f->dir[1] + (pp[2] - qq[2]) * f->dir[2];
			axpy3(-f->stiff * (d < 0. ? d : 0.), f->dir, F);
		}
	}
	else if(f->type == OXO_EXT_GENERIC_CENTRAL) {
		/* src/Forces/GenericCentralForce.cpp:174-193, gravity */
		double d[3] = { f->pos0[0] - pp[0], f->pos0[1] - pp[1], f->pos0[2] - pp[2] };
		double d2 = dot3(d, d);
		if(d2 < f->aux[0]) return;
		if(f->aux[1] > 0. && d2 > f->aux[1]) return;
		axpy3(f->F0 / sqrt(d2), d, F);
	}
	else if(f->type == OXO_EXT_LJ_CONE) {
		/* src/Forces/LJCone.cpp:69-92 */
		double sigma = f->aux[0], cutoff = f->aux[1], alpha = f->aux[2];
		double va[3] = { pp[0] - f->pos0[0], pp[1] - f->pos0[1], pp[2] - f->pos0[2] };
		double d_along = dot3(va, f->dir);
		double vfa[3] = { f->dir[0] * d_along - va[0], f->dir[1] * d_along - va[1], f->dir[2] * d_along - va[2] };
		double d_from_axis = sqrt(dot3(vfa, vfa));
		double d_from_cone = d_along * sin(alpha) - d_from_axis * cos(alpha);
		double rel = d_from_cone / sigma;
		if(rel > cutoff) return;
		double C = d_from_axis * tan(alpha);
		double nrm[3] = { (d_along + C) * f->dir[0] - va[0], (d_along + C) * f->dir[1] - va[1], (d_along + C) * f->dir[2] - va[2] };
		double nn = sqrt(dot3(nrm, nrm));
		double lj = pow(rel, -f->iaux);
		axpy3(4 * f->iaux * f->stiff * (2 * SQ(lj) - lj) / d_from_cone / nn, nrm, F);
	}
	else if(f->type == OXO_EXT_YUKAWA_SPHERE) {
		/* src/Forces/YukawaSphere.cpp:53-74 */
		double d[3];
		min_image(box, f->pos0, pp, d);
		double m = sqrt(dot3(d, d)), ds = f->r0 - m;
		if(ds < f->aux[4]) {
			double s = (f->aux[3] * exp(-ds / f->aux[2])) * (1.0 / (ds * f->aux[2]) + 1.0 / SQ(ds));
			if(ds < f->aux[1]) {
				double w = pow(f->aux[0] / ds, 6);
				s += 4 * f->stiff * f->iaux * (2 * SQ(w) - w) / ds;
			}
			axpy3(-s / m, d, F);
		}
	}
	else if(f->type == OXO_EXT_SPHERE_MOVING) {
		/* src/Forces/RepulsiveSphereMoving.cpp:85-131 */
		double c[3] = { f->pos0[0], f->pos0[1], f->pos0[2] };
		if(f->aux[4] > 0.) {
			double t = (double) step / (double) (long long) f->aux[4];
			t = t < 0. ? 0. : (t > 1. ? 1. : t);
			for(int k = 0; k < 3; k++) c[k] = f->pos0[k] + (f->aux[1 + k] - f->pos0[k]) * t;
		}
		double d[3];
		min_image(box, c, pp, d);
		double m = sqrt(dot3(d, d)), r = m - (f->r0 + f->rate * (double) step);
		if(r >= f->aux[0] || m <= 0. || r >= pow(2.0, 0.5)) return;
		double rs = r > 1e-9 ? r : 1e-9;
		double A = pow(1. / rs, 2);
		double dUdr = 4.0 * f->stiff * (2.0 * A - 1.0) * (-(2. / rs) * A);
		axpy3(-dUdr / m, d, F);
	}
}

static const int *ext_pool = NULL;
/* COMForce::_compute_coms caches both centres of mass by step index (COMForce.cpp:48-62, _last_step = -1 initially), so the force
 * evaluation of the first MD step -- same step index as the one done at initialisation -- sees the centres of mass of the initial
 * configuration.  Reproduced here (cache keyed by table position, cleared by oxo_set_ext_pool); the reference's CUDA kernel and ours
 * recompute the sums at every evaluation (CUDA_MD.cuh:441-468). */
#define OXO_MAX_COM 64
static struct { long long last_step; double com[3], ref[3]; } com_cache[OXO_MAX_COM];
/* index lists of the COM forces (COMForce::_com_list, _ref_list) */
void oxo_set_ext_pool(const int *pool) {
	ext_pool = pool;
	for(int k = 0; k < OXO_MAX_COM; k++) com_cache[k].last_step = -1;
}

static void ext_com(const oxo_ext_force *f, int slot, const double *pos, long long step, double *force) {
	/* src/Forces/COMForce.cpp:46-71: every particle of com_list feels the spring between the two centres of mass / n_com */
	const int *cl = ext_pool + f->ref, *rl = cl + f->iaux;
	double *com = com_cache[slot % OXO_MAX_COM].com, *ref = com_cache[slot % OXO_MAX_COM].ref, d[3];
	if(step != com_cache[slot % OXO_MAX_COM].last_step) {
		for(int k = 0; k < 3; k++) com[k] = ref[k] = 0.;
		for(int k = 0; k < f->iaux; k++) axpy3(1., pos + 3 * (size_t) cl[k], com);
		for(int k = 0; k < f->pbc; k++) axpy3(1., pos + 3 * (size_t) rl[k], ref);
		for(int k = 0; k < 3; k++) { com[k] /= f->iaux; ref[k] /= f->pbc; }
		com_cache[slot % OXO_MAX_COM].last_step = step;
	}
	for(int k = 0; k < 3; k++) d[k] = ref[k] - com[k];
	double m = sqrt(dot3(d, d));
	double s = (m - (f->r0 + f->rate * step)) * f->stiff / f->iaux;
	for(int k = 0; k < f->iaux; k++) axpy3(s / m, d, force + 3 * (size_t) cl[k]);
}

static const double *ext_grid = NULL;
/* tabulated bias potentials of the metadynamics COM traps (LTCOMTrap::potential_grid) */
void oxo_set_ext_grid(const double *grid) { ext_grid = grid; }

static void ext_meta_com(const oxo_ext_force *f, const double *pos, const double *box, double *force) {
	/* src/Forces/Metadynamics/LTCOMTrap.cpp:52-78, meta_utils.h:30-34, meta_utils.cpp:18-21 */
	const int *l1 = ext_pool + f->ref, *l2 = l1 + f->iaux;
	const int n1 = f->iaux, n2 = f->pbc, mode = (int) f->aux[3], n_grid = (int) f->aux[2];
	const double xmin = f->aux[0], dX = f->aux[1], *grid = ext_grid + (int) f->aux[4];
	double c1[3] = { 0, 0, 0 }, c2[3] = { 0, 0, 0 }, dra[3];
	for(int k = 0; k < n1; k++) axpy3(1., pos + 3 * (size_t) l1[k], c1);
	for(int k = 0; k < n2; k++) axpy3(1., pos + 3 * (size_t) l2[k], c2);
	for(int k = 0; k < 3; k++) { c1[k] /= n1; c2[k] /= n2; }
	if(f->aux[5] != 0.) min_image(box, c2, c1, dra);
	else for(int k = 0; k < 3; k++) dra[k] = c1[k] - c2[k];
	double x = sqrt(dot3(dra, dra));
	int il = (int) floor((x - xmin) / dX), ir = il + 1;
	double fx = 0.;
	if(!(il < 0 || ir > n_grid - 1)) fx = -(grid[ir] - grid[il]) / dX;
	if(mode == 1) for(int k = 0; k < n1; k++) axpy3(fx / x / n1, dra, force + 3 * (size_t) l1[k]);
	else for(int k = 0; k < n2; k++) axpy3(fx / x / (-1. * n2), dra, force + 3 * (size_t) l2[k]);
}

void oxo_ext_forces(int nf, const oxo_ext_force *ef, int N, const double *pos, const double *box, long long step, double *force) {
	for(int i = 0; i < nf; i++) {
		if(ef[i].type == OXO_EXT_META_COM_TRAP) ext_meta_com(&ef[i], pos, box, force);
		else if(ef[i].type == OXO_EXT_COM) ext_com(&ef[i], i, pos, step, force);
		else if(ef[i].type > OXO_EXT_ELLIPSOID) {
			if(ef[i].particle >= 0) ext_more(&ef[i], pos, ef[i].particle, box, step, force);
			else for(int p = 0; p < N; p++) ext_more(&ef[i], pos, p, box, step, force);
		}
		else if(ef[i].particle >= 0) ext_one(&ef[i], pos, ef[i].particle, box, step, force);
		else for(int p = 0; p < N; p++) ext_one(&ef[i], pos, p, box, step, force);
	}
}

/* ------------------------------------------------------------------ metadynamics coordination bias
 * LTCoordination (src/Forces/Metadynamics/LTCoordination.cpp:94-221) over the helpers of src/Forces/Metadynamics/meta_utils.{h,cpp}:
 * a smooth count of formed base pairs among a list of candidate pairs, biased by a tabulated potential.  The hydrogen-bond energy is
 * the unsmoothed oxDNA2 expression of meta_utils.cpp:252-366 (its constants are the float literals of src/model.h). */
static double mc_f1(double r) {
	/* meta_utils.cpp:215-232; `shift` is evaluated with a float argument to exp there */
	const double shift = 1.0678f * SQ(1.0 - (double) expf(-(0.75f - 0.4f) * 8.f));
	if(!(r < 0.783775f)) return 0.;
	double tmp = 1.0 - exp(-(r - 0.4f) * 8.f);
	return 1.0678f * SQ(tmp) - shift;
}
static double mc_f1D(double r) {
	if(!(r < 0.783775f)) return 0.;
	double tmp = exp(-(r - 0.4f) * 8.f);
	return 2.0 * 1.0678f * (1 - tmp) * tmp * 8.f;
}
static double mc_f4(double t, double t0, double a) {
	t -= t0;
	if(t < 0.0) t *= -1.0;
	return (t < 1.0 / sqrt(a)) ? 1.0 - a * SQ(t) : 0.;
}
static double mc_f4Dsin(double t, double t0, double a) {
	/* meta_utils.cpp:246-267 */
	double m = 1.0, tt0 = t - t0;
	if(tt0 < 0.0) { tt0 *= -1.0; m = -1.0; }
	if(!(tt0 < 1.0 / sqrt(a))) return 0.;
	double sint = sin(t);
	return (sint > 1e-10) ? m * 2.0 * a * tt0 / sint : m * 2.0 * a;
}
static double safe_acos_(double x) { return acos(fmax(-1.0, fmin(1.0, x))); }

/* energy of the pair (p, q); force on p and lab-frame torque on p if wanted (meta_utils.cpp:269-366) */
static double mc_hb(const double *pos, const double *axes, const int *btype, const double *box, int p, int q, double *force, double *torque, int want) {
	const float PIf = 3.141592653589793238462643f;
	const double T0[6] = { 0.f, 0.f, 0.f, PIf, PIf * 0.5f, PIf * 0.5f }, A[6] = { 1.5f, 1.5f, 1.5f, 0.46f, 4.f, 4.f };
	for(int k = 0; k < 3; k++) force[k] = torque[k] = 0.;
	if(btype[p] + btype[q] != 3) return 0.;
	const double *a1 = axes + 9 * (size_t) p, *a3 = a1 + 6, *b1 = axes + 9 * (size_t) q, *b3 = b1 + 6;
	double r[3], rh[3];
	min_image(box, pos + 3 * (size_t) p, pos + 3 * (size_t) q, r);
	for(int k = 0; k < 3; k++) rh[k] = r[k] + 0.4f * b1[k] - 0.4f * a1[k];
	double m = sqrt(dot3(rh, rh));
	if(!(0.276908f < m && m < 0.783775f)) return 0.;
	double h[3] = { rh[0] / m, rh[1] / m, rh[2] / m };
	double c[6] = { -dot3(a1, b1), -dot3(b1, h), dot3(a1, h), dot3(a3, b3), -dot3(b3, h), dot3(a3, h) };
	double t[6], f4[6], f1 = mc_f1(m), e = f1;
	for(int k = 0; k < 6; k++) { t[k] = safe_acos_(c[k]); f4[k] = mc_f4(t[k], T0[k], A[k]); e *= f4[k]; }
	if(!want || e == 0.) return e;
	const double sgn[6] = { 1., 1., -1., -1., 1., -1. }; /* f4t1Dsin, f4t2Dsin, -f4t3Dsin, -f4t4Dsin, f4t7Dsin, -f4t8Dsin */
	double D[6], prod_wo[6];
	for(int k = 0; k < 6; k++) {
		D[k] = sgn[k] * mc_f4Dsin(t[k], T0[k], A[k]);
		prod_wo[k] = f1;
		for(int j = 0; j < 6; j++) prod_wo[k] *= (j == k) ? D[k] : f4[j];
	}
	double all4 = f4[0] * f4[1] * f4[2] * f4[3] * f4[4] * f4[5], dir[3];
	axpy3(-(mc_f1D(m) * all4), h, force);                          /* radial */
	cross3(a3, b3, dir); axpy3(-prod_wo[3], dir, torque);          /* theta4 */
	cross3(a1, b1, dir); axpy3(-prod_wo[0], dir, torque);          /* theta1 */
	for(int k = 0; k < 3; k++) force[k] += (b1[k] + h[k] * c[1]) * (prod_wo[1] / m);      /* theta2 (torque on q only) */
	for(int k = 0; k < 3; k++) force[k] += (a1[k] - h[k] * c[2]) * (prod_wo[2] / m);      /* theta3 */
	cross3(h, a1, dir); axpy3(prod_wo[2], dir, torque);
	for(int k = 0; k < 3; k++) force[k] += (b3[k] + h[k] * c[4]) * (prod_wo[4] / m);      /* theta7 (torque on q only) */
	for(int k = 0; k < 3; k++) force[k] += (a3[k] - h[k] * c[5]) * (prod_wo[5] / m);      /* theta8 */
	cross3(h, a3, dir); axpy3(prod_wo[5], dir, torque);
	double base[3] = { 0.4f * a1[0], 0.4f * a1[1], 0.4f * a1[2] };
	cross3(base, force, dir); axpy3(1., dir, torque);
	return e;
}
static double mc_smooth(double cut, double width, double e) {
	double x = (cut - e) / width;
	if(x > 10.0) return 1.0;
	if(x < -10.0) return 0.0;
	return 0.5 * (1.0 + tanh(x));
}
static double mc_dsmooth(double cut, double width, double e) {
	double x = (cut - e) / width;
	if(x > 10.0 || x < -10.0) return 0.0;
	double th = tanh(x);
	return -0.5 * (1.0 - SQ(th)) / width;
}
static void mc_base_vec(const double *pos, const double *axes, const double *box, int p, int q, double *r) {
	double bp[3], bq[3];
	for(int k = 0; k < 3; k++) { bp[k] = pos[3 * (size_t) p + k] + 0.4f * axes[9 * (size_t) p + k]; bq[k] = pos[3 * (size_t) q + k] + 0.4f * axes[9 * (size_t) q + k]; }
	min_image(box, bp, bq, r);
}
static double mc_pair_contribution(const oxo_coord *C, const double *pos, const double *axes, const int *btype, const double *box, int p, int q) {
	/* meta_utils.cpp:127-157 */
	double f[3], t[3], hbc = 0., sw = 0.;
	if(C->mode != 1) hbc = mc_smooth(C->hb_energy_cutoff, C->hb_transition_width, mc_hb(pos, axes, btype, box, p, q, f, t, 0));
	if(C->mode != 0) {
		double r[3];
		mc_base_vec(pos, axes, box, p, q, r);
		sw = 1.0 / (1.0 + pow((sqrt(dot3(r, r)) - C->d0) / C->r0, C->n));
	}
	return C->mode == 0 ? hbc : (C->mode == 1 ? sw : C->mixed_weight * hbc + (1.0 - C->mixed_weight) * sw);
}

double oxo_meta_coordination(const oxo_coord *C, int N, const double *pos, const double *axes, const int *btype, const double *box,
		double *force, double *torque_lab) {
	(void) N;
	double coord = 0.;
	for(int k = 0; k < C->n_pairs; k++) coord += mc_pair_contribution(C, pos, axes, btype, box, C->pairs[2 * k], C->pairs[2 * k + 1]);
	/* LTCoordination::_coordination clamps; ::force reads -dV/dcoord off the grid by finite difference (meta_utils.h:30-34) */
	double cl = coord < C->coord_min ? C->coord_min : (coord > C->coord_max ? C->coord_max : coord);
	double dc = (C->coord_max - C->coord_min) / (C->N_grid - 1.0);
	int il = (int) floor((cl - C->coord_min) / dc), ir = il + 1;
	double df = 0.;
	if(!(il < 0 || ir > C->N_grid - 1)) df = -(C->grid[ir] - C->grid[il]) / dc;
	for(int k = 0; k < C->n_pairs; k++) for(int side = 0; side < 2; side++) {
		/* meta_utils.cpp:159-205 with current = p, other = q */
		int p = C->pairs[2 * k + side], q = C->pairs[2 * k + 1 - side];
		double f[3] = { 0, 0, 0 }, t[3] = { 0, 0, 0 }, fs[3] = { 0, 0, 0 }, ts[3] = { 0, 0, 0 };
		if(C->mode != 1) {
			double e = mc_hb(pos, axes, btype, box, p, q, f, t, 1), d = mc_dsmooth(C->hb_energy_cutoff, C->hb_transition_width, e);
			for(int x = 0; x < 3; x++) { f[x] *= d; t[x] *= d; }
		}
		if(C->mode != 0) {
			double r[3], base[3];
			mc_base_vec(pos, axes, box, p, q, r);
			double rm = sqrt(dot3(r, r)), xx = (rm - C->d0) / C->r0, xn = pow(xx, C->n);
			double dcdr = (C->n / C->r0) * pow(xx, C->n - 1) / SQ(1.0 + xn);
			for(int x = 0; x < 3; x++) { fs[x] = r[x] / rm * dcdr; base[x] = 0.4f * axes[9 * (size_t) p + x]; }
			cross3(base, fs, ts);
		}
		double w = C->mode == 2 ? C->mixed_weight : (C->mode == 0 ? 1. : 0.);
		for(int x = 0; x < 3; x++) {
			force[3 * (size_t) p + x] += df * (w * f[x] + (1. - w) * fs[x]);
			torque_lab[3 * (size_t) p + x] += df * (w * t[x] + (1. - w) * ts[x]);
		}
	}
	return coord;
}

/* ------------------------------------------------------------------ Verlet list */
static int cell_of(const double *box, const int *nc, const double *p) {
	/* src/Lists/Cells.h:60-65 */
	int c[3];
	for(int k = 0; k < 3; k++) c[k] = (int) ((p[k] / box[k] - floor(p[k] / box[k])) * (1. - DBL_EPSILON) * nc[k]);
	return c[0] + nc[0] * (c[1] + nc[1] * c[2]);
}

long long oxo_verlet_pairs(int N, const double *pos, const int *n3, const int *n5, const double *box, double rv,
		int *pairs, long long max_pairs) {
	int nc[3];
	for(int k = 0; k < 3; k++) {
		/* Cells.cpp:47-58 without cells_auto_optimisation: does not change the pair set */
		nc[k] = (int) (floor(box[k] / rv) + 0.1);
		if(nc[k] < 3) nc[k] = 3;
		if(nc[k] > 256) nc[k] = 256; /* memory guard only */
	}
	int ncells = nc[0] * nc[1] * nc[2];
	int *head = (int *) malloc((size_t) ncells * sizeof(int));
	int *next = (int *) malloc((size_t) N * sizeof(int));
	int *cell = (int *) malloc((size_t) N * sizeof(int));
	for(int c = 0; c < ncells; c++) head[c] = -1;
	for(int i = 0; i < N; i++) {
		int c = cell_of(box, nc, pos + 3 * (size_t) i);
		cell[i] = c; next[i] = head[c]; head[c] = i;
	}
	double rv2 = rv * rv;
	long long n = 0;
	for(int p = 0; p < N; p++) {
		int c = cell[p];
		int ind[3] = { c % nc[0], (c / nc[0]) % nc[1], c / (nc[0] * nc[1]) };
		int seen[27], nseen = 0;
		for(int dz = -1; dz <= 1; dz++) for(int dy = -1; dy <= 1; dy++) for(int dx = -1; dx <= 1; dx++) {
			int cc = ((ind[0] + dx + nc[0]) % nc[0]) + nc[0] * (((ind[1] + dy + nc[1]) % nc[1]) + nc[1] * ((ind[2] + dz + nc[2]) % nc[2]));
			int dup = 0;
			for(int s = 0; s < nseen; s++) if(seen[s] == cc) dup = 1;
			if(dup) continue;
			seen[nseen++] = cc;
			for(int q = head[cc]; q != -1; q = next[q]) {
				if(q >= p) continue;
				if(n3[p] == q || n5[p] == q) continue;
				double r[3];
				min_image(box, pos + 3 * (size_t) p, pos + 3 * (size_t) q, r);
				if(dot3(r, r) < rv2) {
					if(n < max_pairs) { pairs[2 * n] = q; pairs[2 * n + 1] = p; }
					n++;
				}
			}
		}
	}
	free(head); free(next); free(cell);
	return n;
}

void oxo_axes_from_a1a3(int N, const double *a1, const double *a3, double *axes) {
	/* orthonormalisation of the configuration reader, src/Backends/SimBackend.cpp:623-629 */
	for(int i = 0; i < N; i++) {
		double v1[3], v3[3], v2[3];
		double n1 = sqrt(dot3(a1 + 3 * i, a1 + 3 * i)), n3 = sqrt(dot3(a3 + 3 * i, a3 + 3 * i));
		for(int k = 0; k < 3; k++) { v1[k] = a1[3 * i + k] / n1; v3[k] = a3[3 * i + k] / n3; }
		double d = dot3(v1, v3);
		for(int k = 0; k < 3; k++) v1[k] -= v3[k] * d;
		n1 = sqrt(dot3(v1, v1));
		for(int k = 0; k < 3; k++) v1[k] /= n1;
		cross3(v3, v1, v2);
		double n2 = sqrt(dot3(v2, v2));
		for(int k = 0; k < 3; k++) v2[k] /= n2;
		memcpy(axes + 9 * (size_t) i, v1, sizeof(v1));
		memcpy(axes + 9 * (size_t) i + 3, v2, sizeof(v2));
		memcpy(axes + 9 * (size_t) i + 6, v3, sizeof(v3));
	}
}

/* ------------------------------------------------------------------ MD step */
typedef void (*forces_cb)(const void *P, int N, const double *pos, const double *axes, const int *btype, const int *n3, const int *n5,
		const double *box, const int *pairs, long long npairs, double *force, double *tl, double *et);

static void dna_forces_cb(const void *P, int N, const double *pos, const double *axes, const int *btype, const int *n3, const int *n5,
		const double *box, const int *pairs, long long npairs, double *force, double *tl, double *et) {
	oxo_dna2_forces((const oxo_dna2_params *) P, N, pos, axes, btype, n3, n5, box, pairs, npairs, force, tl, NULL, et, NULL);
}

static void md_compute_forces_generic(const void *P, forces_cb cb, oxo_md *S) {
	double et[OXO_NTERMS];
	double *tl = (double *) calloc(3 * (size_t) S->N, sizeof(double));
	cb(P, S->N, S->pos, S->axes, S->btype, S->n3, S->n5, S->box, S->pairs, S->npairs, S->force, tl, et);
	/* external forces enter before the interactions in the reference (MD_CPUBackend.cpp:149-153); addition commutes */
	if(S->nf > 0) oxo_ext_forces(S->nf, S->ef, S->N, S->pos, S->box, S->step, S->force);
	for(int i = 0; i < S->N; i++) {
		const double *ax = S->axes + 9 * (size_t) i;
		for(int k = 0; k < 3; k++) S->torque_body[3 * i + k] = dot3(ax + 3 * k, tl + 3 * (size_t) i);
	}
	S->U = 0;
	for(int t = 0; t < OXO_NTERMS; t++) S->U += et[t];
	free(tl);
}

void oxo_md_compute_forces(const oxo_dna2_params *P, oxo_md *S) { md_compute_forces_generic(P, dna_forces_cb, S); }

static void rebuild(double rcut, oxo_md *S) {
	S->npairs = oxo_verlet_pairs(S->N, S->pos, S->n3, S->n5, S->box, rcut + 2 * S->skin, S->pairs, S->max_pairs);
	memcpy(S->list_pos, S->pos, 3 * (size_t) S->N * sizeof(double));
}

static int md_steps_generic(const void *P, double rcut, forces_cb cb, oxo_md *S, int nsteps) {
	int rebuilds = 0;
	const double dt = S->dt;
	for(int it = 0; it < nsteps; it++) {
		int stale = 0;
		/* first half: MD_CPUBackend.cpp:67-144 */
		for(int i = 0; i < S->N; i++) {
			double *v = S->vel + 3 * i, *x = S->pos + 3 * i, *L = S->L + 3 * i, *ax = S->axes + 9 * (size_t) i;
			for(int k = 0; k < 3; k++) { v[k] += S->force[3 * i + k] * (dt * 0.5); x[k] += v[k] * dt; }
			for(int k = 0; k < 3; k++) L[k] += S->torque_body[3 * i + k] * (dt * 0.5);
			double norm = sqrt(dot3(L, L));
			double u[3] = { L[0] / norm, L[1] / norm, L[2] / norm };
			double s = sin(dt * norm), c = cos(dt * norm), o = 1. - c;
			double R[3][3] = {
				{ u[0] * u[0] * o + c, u[0] * u[1] * o - u[2] * s, u[0] * u[2] * o + u[1] * s },
				{ u[0] * u[1] * o + u[2] * s, u[1] * u[1] * o + c, u[1] * u[2] * o - u[0] * s },
				{ u[0] * u[2] * o - u[1] * s, u[1] * u[2] * o + u[0] * s, u[2] * u[2] * o + c } };
			/* orientation (columns = body axes) <- orientation * R, i.e. new axis j = sum_k old axis k * R[k][j] */
			double na[9];
			for(int j = 0; j < 3; j++) for(int k = 0; k < 3; k++)
				na[3 * j + k] = ax[k] * R[0][j] + ax[3 + k] * R[1][j] + ax[6 + k] * R[2][j];
			memcpy(ax, na, sizeof(na));
			double d[3] = { x[0] - S->list_pos[3 * i], x[1] - S->list_pos[3 * i + 1], x[2] - S->list_pos[3 * i + 2] };
			if(dot3(d, d) > SQ(S->skin)) stale = 1;
		}
		if(stale) { rebuild(rcut, S); rebuilds++; }
		md_compute_forces_generic(P, cb, S);
		for(int i = 0; i < S->N; i++) for(int k = 0; k < 3; k++) {
			S->vel[3 * i + k] += S->force[3 * i + k] * dt * 0.5;
			S->L[3 * i + k] += S->torque_body[3 * i + k] * dt * 0.5;
		}
		S->step++;
	}
	return rebuilds;
}

int oxo_md_steps(const oxo_dna2_params *P, oxo_md *S, int nsteps) { return md_steps_generic(P, P->rcut, dna_forces_cb, S, nsteps); }

/* ------------------------------------------------------------------ thermostat parameters */
void oxo_brownian_params(double T, double dt_in, int ns, double pt_in, double diff_coeff, double *pt, double *pr, double *rescale) {
	/* BrownianThermostat.cpp:27-54: pt, diff_coeff and dt are read as *floats* */
	double dt = (float) dt_in, D = (float) diff_coeff, p = (float) pt_in;
	if(p == 0.) p = (2 * T * ns * dt) / (T * ns * dt + 2 * D);
	D = T * ns * dt * (1. / p - 1. / 2.);
	*pt = p;
	*pr = (2 * T * ns * dt) / (T * ns * dt + 2 * 3 * D);
	*rescale = sqrt(T);
}

void oxo_langevin_params(double T, double dt, double gamma_in, double diff_in, double *gt, double *gr, double *rt, double *rr) {
	/* LangevinThermostat.cpp:27-76: gamma_trans / diff_coeff are read as floats, dt as a number */
	double g = (float) gamma_in, D = (float) diff_in;
	if(D == 0.) D = T / g; else g = T / D;
	double Dr = 3. * D;
	*gt = g; *gr = T / Dr;
	*rt = sqrt(2. * g * T / dt);
	*rr = sqrt(2. * (*gr) * T / dt);
}

#include "oxrna_oracle.inc"
#include "oxdna3_oracle.inc"
