// oxDNA2 pair potential, FP32 device functions (hand-written for sm_100a; no tensor cores: the work is
// transcendental/SFU + FP32-pipe bound).
//
// What it evaluates is the reference's model (src/CUDA/Interactions/CUDA_DNA.cuh:43-725, CPU mirror
// src/Interactions/DNAInteraction.cpp:415-1172 + DNA2Interaction.cpp:157-306); how it evaluates it is different:
//  * every factor f(cos theta) is differentiated with respect to the COSINE and the chain rule is applied through
//    two generic moves (axis.axis and axis.direction), so hydrogen bonding and cross stacking -- which share their
//    six angles -- accumulate their derivatives first and touch the vectors once;
//  * sin(theta) is taken from the cross product that the torque needs anyway, and theta = atan2(sin, cos): unlike
//    acosf(cos) this keeps FP32 relative accuracy when theta approaches 0 or pi (coaxial stacking lives near pi);
//  * site forces are accumulated per lever arm (a1-collinear sites share one accumulator), two cross products per
//    particle per pair instead of one per site interaction.
#pragma once

#include "common.cuh"

// ---- fast elementary functions.  On the device: SFU reciprocal / rsqrt / ex2 (1-2 ulp); on the host (unit test build):
// the plain libm equivalents.  The arctangent is our own (both sides): Cephes-style range reduction + degree-9 odd
// polynomial, |error| < 1.5e-7 rad on [0, pi].
#ifdef __CUDA_ARCH__
#define OXB_DIV(a, b) __fdividef((a), (b))
#define OXB_RSQRT(x) rsqrtf(x)
#define OXB_EXP(x) __expf(x)
#else
#define OXB_DIV(a, b) ((a) / (b))
#define OXB_RSQRT(x) (1.0f / sqrtf(x))
#define OXB_EXP(x) expf(x)
#endif

// theta = atan2(s, c) for s >= 0
OXB_HD float atan2_pos(float s, float c) {
	float ax = fabsf(c);
	float mx = fmaxf(ax, s), mn = fminf(ax, s);
	if(mx == 0.f) return 0.f;
	float a = OXB_DIV(mn, mx);
	float base = 0.f;
	if(a > 0.41421356f) {
		a = OXB_DIV(a - 1.f, a + 1.f);
		base = 0.78539816339744831f;
	}
	float z = a * a;
	float p = ((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f;
	float r = base + (a + a * z * p);
	if(s > ax) r = 1.57079632679489662f - r;
	if(c < 0.f) r = 3.14159265358979324f - r;
	return r;
}

struct AngVal {
	float v;  // f(theta)
	float dc; // d f / d cos(theta)
};

// f4 as a function of theta with the derivative taken with respect to cos(theta); s = sin(theta) >= 0 comes from the
// cross product the torque needs anyway.
OXB_HD AngVal f4_ts(const oxb_f4 &f, float t, float s) {
	AngVal r;
	r.v = 0.f;
	r.dc = 0.f;
	float x = t - f.t0;
	float m = 1.f;
	if(x < 0.f) {
		x = -x;
		m = -1.f;
	}
	if(x < f.tc) {
		if(x > f.ts) {
			float d = f.tc - x;
			r.v = f.b * d * d;
			// d/dtheta = m 2 b (x - tc);  d/dcos = -(d/dtheta) / sin
			r.dc = OXB_DIV(m * 2.f * f.b * d, fmaxf(s, 1e-12f));
		}
		else {
			r.v = 1.f - f.a * x * x;
			// same small-angle limit as the CPU code (DNAInteraction.cpp:1411-1414)
			r.dc = (s * s > 1e-8f) ? OXB_DIV(m * 2.f * f.a * x, s) : m * 2.f * f.a;
		}
	}
	return r;
}

#define OXB_PI_F 3.14159265358979f
// F(theta) = f4(theta) + f4(pi - theta)
OXB_HD AngVal f4_ts_sym(const oxb_f4 &f, float t, float s) {
	AngVal p = f4_ts(f, t, s), n = f4_ts(f, OXB_PI_F - t, s);
	AngVal r;
	r.v = p.v + n.v;
	r.dc = p.dc - n.dc;
	return r;
}

// oxDNA2 coaxial-stacking theta1: f4 plus a pure harmonic beyond theta = sb (DNA2Interaction.cpp:308-363)
OXB_HD AngVal f4_ts_cxst_t1(const oxb_dna2_params &M, float t, float s) {
	AngVal r = f4_ts(M.f4[OXB_F4_CXST_T1], t, s);
	float x = t - M.cxst_t1_sb;
	if(x >= 0.f) {
		r.v += M.cxst_t1_sa * x * x;
		r.dc -= (s * s > 1e-8f) ? OXB_DIV(2.f * M.cxst_t1_sa * x, s) : 2.f * M.cxst_t1_sa;
	}
	return r;
}

// coaxial theta1 of oxDNA (first generation) and oxRNA: f4(theta) + f4(2 pi - theta)
// (DNAInteraction.cpp:1331-1352, RNAInteraction.cpp:1026,1298-1303); sin(2 pi - theta) = -sin(theta)
OXB_HD AngVal f4_ts_rna_cxst_t1(const oxb_f4 &f, float t, float s) {
	AngVal p = f4_ts(f, t, s), n = f4_ts(f, 2.f * OXB_PI_F - t, s);
	AngVal r;
	r.v = p.v + n.v;
	r.dc = p.dc - n.dc;
	return r;
}

OXB_HD AngVal f5_c(const oxb_f5 &f, float c) {
	AngVal r;
	r.v = 0.f;
	r.dc = 0.f;
	if(c > f.xc) {
		if(c < f.xs) {
			float d = f.xc - c;
			r.v = f.b * d * d;
			r.dc = -2.f * f.b * d;
		}
		else if(c < 0.f) {
			r.v = 1.f - f.a * c * c;
			r.dc = -2.f * f.a * c;
		}
		else {
			r.v = 1.f;
		}
	}
	return r;
}

struct RadVal {
	float v, d;
};

OXB_HD RadVal f1_r(const oxb_f1 &f, float eps, float shift, float r) {
	RadVal o;
	o.v = 0.f;
	o.d = 0.f;
	if(r < f.rchigh) {
		if(r > f.rhigh) {
			float x = r - f.rchigh;
			o.v = eps * f.bhigh * x * x;
			o.d = 2.f * eps * f.bhigh * x;
		}
		else if(r > f.rlow) {
			float e = OXB_EXP(-(r - f.r0) * f.a);
			float t = 1.f - e;
			o.v = eps * t * t - shift;
			o.d = 2.f * eps * t * e * f.a;
		}
		else if(r > f.rclow) {
			float x = r - f.rclow;
			o.v = eps * f.blow * x * x;
			o.d = 2.f * eps * f.blow * x;
		}
	}
	return o;
}

OXB_HD RadVal f2_r(const oxb_f2 &f, float r) {
	RadVal o;
	o.v = 0.f;
	o.d = 0.f;
	if(r < f.rchigh) {
		if(r > f.rhigh) {
			float x = r - f.rchigh;
			o.v = f.k * f.bhigh * x * x;
			o.d = 2.f * f.k * f.bhigh * x;
		}
		else if(r > f.rlow) {
			float x = r - f.r0, y = f.rc - f.r0;
			o.v = 0.5f * f.k * (x * x - y * y);
			o.d = f.k * x;
		}
		else if(r > f.rclow) {
			float x = r - f.rclow;
			o.v = f.k * f.blow * x * x;
			o.d = 2.f * f.k * f.blow * x;
		}
	}
	return o;
}

// repulsive LJ + quadratic smoothing: returns energy, writes the scalar s with force-on-q = s * r
OXB_HD float excl_s(const oxb_excl &e, float eps, v3 r, float &s) {
	float r2 = dot(r, r);
	float en = 0.f;
	s = 0.f;
	if(r2 < e.rc2) {
		if(r2 > e.rstar2) {
			float inv = OXB_RSQRT(r2);
			float rrc = r2 * inv - e.rc;
			en = eps * e.b * rrc * rrc;
			s = -2.f * eps * e.b * rrc * inv;
		}
		else {
			float ir2 = OXB_DIV(1.f, r2);
			float t = e.sigma2 * ir2;
			float lj = t * t * t;
			en = 4.f * eps * (lj * lj - lj);
			s = -24.f * eps * (lj - 2.f * lj * lj) * ir2;
		}
	}
	return en;
}

// Accumulator for one pair (p, q).  F is the force on q (p receives -F).  Forces acting at a1-collinear sites
// (base, stack, ungrooved backbone reference) are summed pre-multiplied by their offset along a1; forces at the
// (grooved) backbone site are summed separately.  Tp/Tq collect the pure (non lever-arm) torques, lab frame.
// Excluded volume is the other stiff piece of the model: d2V/dr2 = 2 eps b ~ 3.6e3 (backbone-backbone) to 1.8e4 (base-base) in the
// quadratic smoothing shell and about as much on the Lennard-Jones side of r*.  An FP32 site-site distance (ulp 1.2e-7 at 1 sigma,
// orientation in FP32) then carries ~1e-3 of absolute force error, i.e. more than 1e-5 of max|F| in every large system.  Kernels that
// can reach the FP64 state hand it in through PairAcc::refine; an ACTIVE excluded-volume term (rare between non-bonded and between
// bonded nucleotides alike) is then re-evaluated in double from positions and quaternions, and the difference is added.
struct ExclRefine {
	const double4 *posd, *quatd; // FP64 state, slot-indexed
	int sp, sq;                  // slots of p and q
	double L[3];                 // box sides
	double b1, b2, b3;           // backbone-site coefficients on a1, a2, a3
};
enum { OXB_SITE_KK = 0, OXB_SITE_AA = 1, OXB_SITE_AK = 2, OXB_SITE_KA = 3 }; // (site of p)(site of q): k = backbone, a = on the a1 axis

struct PairAcc {
	v3 F, Pa, Pk, Qa, Qk, Tp, Tq;
	const ExclRefine *refine;
	const oxb_replica_consts *rep; // replica batching: the temperature-dependent constants of this pair's replica (null = those of the model block)
	OXB_HD void clear() {
		refine = nullptr;
		rep = nullptr;
		F = Pa = Pk = Qa = Qk = Tp = Tq = mk3(0.f, 0.f, 0.f);
	}
	// site codes: coefficient along a1, or "backbone" handled by the *_k variants
	OXB_HD void site_aa(v3 f, float cp, float cq) { F += f; axpy(Pa, cp, f); axpy(Qa, cq, f); }
	OXB_HD void site_ak(v3 f, float cp) { F += f; axpy(Pa, cp, f); Qk += f; }
	OXB_HD void site_ka(v3 f, float cq) { F += f; Pk += f; axpy(Qa, cq, f); }
	OXB_HD void site_kk(v3 f) { F += f; Pk += f; Qk += f; }
	// arbitrary lever arms (oxRNA's 3'/5' stacking sites)
	OXB_HD void site_gg(v3 f, v3 sp, v3 sq) { F += f; Tp -= cross(sp, f); Tq += cross(sq, f); }
	// lab-frame torques including lever arms
	OXB_HD v3 torque_p(const Axes &A, v3 pback) const { return Tp - cross(A.a1, Pa) - cross(pback, Pk); }
	OXB_HD v3 torque_q(const Axes &B, v3 qback) const { return Tq + cross(B.a1, Qa) + cross(qback, Qk); }
};

#ifdef __CUDACC__
// the force on q of one excluded-volume site pair evaluated in double from the FP64 state.  Deliberately NOT inlined: it runs for the
// rare active terms only and must not add to the register footprint of the kernels' common path.
static __device__ __noinline__ float3 excl_force_double(const ExclRefine *R, const oxb_excl *e, float eps, int kind, float cbp, float cbq) {
	const double4 pp = R->posd[R->sp], pq = R->posd[R->sq], qp = R->quatd[R->sp], qq = R->quatd[R->sq];
	double rd[3] = { pq.x - pp.x, pq.y - pp.y, pq.z - pp.z };
	for(int k = 0; k < 3; k++) rd[k] -= R->L[k] * rint(rd[k] / R->L[k]);
	double a1p[3], a2p[3], a3p[3], a1q[3], a2q[3], a3q[3];
	quatd Qp = { qp.x, qp.y, qp.z, qp.w }, Qq = { qq.x, qq.y, qq.z, qq.w };
	axes_from_quatd(Qp, a1p, a2p, a3p);
	axes_from_quatd(Qq, a1q, a2q, a3q);
	const bool p_back = (kind == OXB_SITE_KK || kind == OXB_SITE_KA), q_back = (kind == OXB_SITE_KK || kind == OXB_SITE_AK);
	double dd[3];
	for(int k = 0; k < 3; k++) {
		const double sp = p_back ? R->b1 * a1p[k] + R->b2 * a2p[k] + R->b3 * a3p[k] : (double) cbp * a1p[k];
		const double sq = q_back ? R->b1 * a1q[k] + R->b2 * a2q[k] + R->b3 * a3q[k] : (double) cbq * a1q[k];
		dd[k] = rd[k] + sq - sp;
	}
	const double r2 = dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2];
	double sd = 0.;
	if(r2 < (double) e->rc2) {
		if(r2 > (double) e->rstar2) {
			const double m = sqrt(r2);
			sd = -2. * (double) eps * (double) e->b * (m - (double) e->rc) / m;
		}
		else {
			const double t = (double) e->sigma2 / r2, lj = t * t * t;
			sd = -24. * (double) eps * (lj - 2. * lj * lj) / r2;
		}
	}
	return make_float3((float) (dd[0] * sd), (float) (dd[1] * sd), (float) (dd[2] * sd));
}
#endif

// adds (double-precision force) - (the FP32 force d * s already accumulated) of one active excluded-volume site pair; cbp / cbq: offsets of
// the base sites of p and q along a1 (oxDNA3: per type)
OXB_HD void excl_fix2(PairAcc &acc, const oxb_excl &e, float eps, v3 d, float s, int kind, float cbp, float cbq) {
#ifdef __CUDA_ARCH__
	if(acc.refine == nullptr) return;
	const float3 fd = excl_force_double(acc.refine, &e, eps, kind, cbp, cbq);
	const v3 delta = mk3(fd.x - d.x * s, fd.y - d.y * s, fd.z - d.z * s);
	if(kind == OXB_SITE_KK) acc.site_kk(delta);
	else if(kind == OXB_SITE_AA) acc.site_aa(delta, cbp, cbq);
	else if(kind == OXB_SITE_AK) acc.site_ak(delta, cbp);
	else acc.site_ka(delta, cbq);
#endif
}
OXB_HD void excl_fix(PairAcc &acc, const oxb_excl &e, float eps, v3 d, float s, int kind, float cb) { excl_fix2(acc, e, eps, d, s, kind, cb, cb); }


struct Angle {
	float c, s, t; // cos, sin (>= 0), theta = atan2(sin, cos)
	v3 x;          // u cross v
};

OXB_HD Angle make_angle(v3 u, v3 v) {
	Angle a;
	a.c = dot(u, v);
	a.x = cross(u, v);
	float s2 = dot(a.x, a.x);
	a.s = (s2 > 0.f) ? s2 * OXB_RSQRT(s2) : 0.f;
	a.t = atan2_pos(a.s, a.c);
	return a;
}

// cosine + cross product only (for the f5 factors, which are functions of the cosine itself)
OXB_HD Angle make_angle_cos(v3 u, v3 v) {
	Angle a;
	a.c = dot(u, v);
	a.x = cross(u, v);
	a.s = 0.f;
	a.t = 0.f;
	return a;
}

template<class PB> OXB_HD bool in_window(const PB &M, int k, float c) { return c > M.f4_cmin[k] && c < M.f4_cmax[k]; }
template<class PB> OXB_HD bool in_window_sym(const PB &M, int k, float c) { return in_window(M, k, c) || in_window(M, k, -c); }

// chain rule, c = u.v with u on p and v on q, g = dE/dc
OXB_HD void chain_bb(PairAcc &A, float g, const Angle &a) {
	axpy(A.Tp, -g, a.x);
	axpy(A.Tq, g, a.x);
}
// chain rule, c = u.rhat; returns the site force (on q) -g (u - c rhat) / r; pure torque goes to the owner of u
template<bool ON_Q>
OXB_HD v3 chain_bd(PairAcc &A, float g, v3 u, v3 rhat, float inv_r, const Angle &a) {
	if(ON_Q) axpy(A.Tq, -g, a.x);
	else axpy(A.Tp, -g, a.x);
	return (u - rhat * a.c) * (-g * inv_r);
}

// chain rule for c = shat . (bhat x u): u a body axis of p (ON_Q = false) or q; shat between the stacking sites (a1-collinear,
// coefficient cs), bhat between the backbone sites -- the true (grooved / a3-displaced) ones, or with BACK_A1 the a1-collinear
// reference sites of coefficient cr (oxDNA).  g = dE/dc.  RNAInteraction.cpp:1093-1142, DNAInteraction.cpp:1105-1160
template<bool ON_Q, bool BACK_A1>
OXB_HD void chain_triple(PairAcc &acc, float g, v3 u, v3 sh, float sinv, float cs, v3 bh, float binv, float cr) {
	v3 bu = cross(bh, u);
	float c = dot(sh, bu);
	acc.site_aa((bu - sh * c) * (-g * sinv), cs, cs);
	v3 us = cross(u, sh);
	v3 fb = (us - bh * c) * (-g * binv);
	if(BACK_A1) acc.site_aa(fb, cr, cr);
	else acc.site_kk(fb);
	v3 t = cross(u, cross(sh, bh));
	if(ON_Q) axpy(acc.Tq, -g, t);
	else axpy(acc.Tp, -g, t);
}

struct PairEnergy {
	float total, hb;
};

// ---------------------------------------------------------------------------------------------------------------
// Non-bonded pair, split into the pieces the staged edge pipeline launches separately (forces.cu):
//   dna2_dh        Debye-Hueckel on the backbone-backbone vector (the only term beyond rcut_near)
//   dna2_excl      the four excluded-volume site pairs
//   dna2_hbcr      hydrogen bonding + cross stacking (same six angles on the base-base vector)
//   dna2_cxst      coaxial stacking (oxDNA2 form: no phi3 term, harmonic add-on to theta1)
// r = min-image(q - p) between centres of mass; forces are "on q".
// ---------------------------------------------------------------------------------------------------------------

// the Debye-Hueckel constants as a value: either the model block's or a replica's row (same field names: dna2_dh / dna2_dh_fast take both)
struct DhView {
	float dh_minus_kappa, dh_prefactor, dh_rhigh, dh_rc, dh_b;
	int dh_half_charged_ends;
};
template<class PB> OXB_HD DhView dh_view(const PB &M, const oxb_replica_consts *rep) {
	DhView D;
	D.dh_half_charged_ends = M.dh_half_charged_ends;
	if(rep != nullptr) { D.dh_minus_kappa = rep->dh_minus_kappa; D.dh_prefactor = rep->dh_prefactor; D.dh_rhigh = rep->dh_rhigh; D.dh_rc = rep->dh_rc; D.dh_b = rep->dh_b; }
	else { D.dh_minus_kappa = M.dh_minus_kappa; D.dh_prefactor = M.dh_prefactor; D.dh_rhigh = M.dh_rhigh; D.dh_rc = M.dh_rc; D.dh_b = M.dh_b; }
	return D;
}

// returns the DH energy; fs such that force-on-q = fs * rbb (zero outside the range)
template<class PB> OXB_HD float dna2_dh(const PB &M, float rbb2, bool p_end, bool q_end, float &fs) {
	fs = 0.f;
	if(rbb2 >= M.dh_rc * M.dh_rc) return 0.f;
	float m = sqrtf(rbb2);
	float cut = 1.f;
	if(M.dh_half_charged_ends) {
		if(p_end) cut *= 0.5f;
		if(q_end) cut *= 0.5f;
	}
	float en, f; // force on q = f * rhat
	if(m < M.dh_rhigh) {
		float ex = expf(m * M.dh_minus_kappa) * M.dh_prefactor / m;
		en = ex;
		f = -ex * (M.dh_minus_kappa - 1.f / m);
	}
	else {
		float x = m - M.dh_rc;
		en = M.dh_b * x * x;
		f = -2.f * M.dh_b * x;
	}
	fs = f * cut / m;
	return en * cut;
}

#ifdef __CUDACC__
// device-only variant with SFU intrinsics (rsqrt, ex2): no IEEE division / square root on the most frequent path.
// Relative error ~2e-7, far inside the 1e-5 force tolerance.
template<class PB> __device__ __forceinline__ float dna2_dh_fast(const PB &M, float rbb2, bool p_end, bool q_end, float &fs) {
	// both radial regimes are evaluated and selected (ex2 and rsqrt are single SFU instructions): no divergent branch in
	// the most frequently executed loop of the step
	float inv = rsqrtf(rbb2);
	float m = rbb2 * inv;
	float cut = (rbb2 < M.dh_rc * M.dh_rc) ? 1.f : 0.f;
	if(M.dh_half_charged_ends) {
		if(p_end) cut *= 0.5f;
		if(q_end) cut *= 0.5f;
	}
	float ex = __expf(m * M.dh_minus_kappa) * M.dh_prefactor * inv;
	float x = m - M.dh_rc;
	bool inner = m < M.dh_rhigh;
	float en = inner ? ex : M.dh_b * x * x;
	float f = inner ? -ex * (M.dh_minus_kappa - inv) : -2.f * M.dh_b * x;
	fs = f * cut * inv;
	return en * cut;
}
#endif

template<class PB> OXB_HD float dna2_excl(const PB &M, v3 r, v3 rbb, v3 rb, const Axes &A, const Axes &B, v3 pback, v3 qback, PairAcc &acc) {
	const float cb = M.base_a1;
	float s, E = 0.f;
	float en = excl_s(M.excl[0], M.excl_eps, rbb, s);
	if(en != 0.f) { E += en; acc.site_kk(rbb * s); excl_fix(acc, M.excl[0], M.excl_eps, rbb, s, OXB_SITE_KK, cb); }
	en = excl_s(M.excl[1], M.excl_eps, rb, s);
	if(en != 0.f) { E += en; acc.site_aa(rb * s, cb, cb); excl_fix(acc, M.excl[1], M.excl_eps, rb, s, OXB_SITE_AA, cb); }
	v3 d = r + B.a1 * cb - pback; // back(p) - base(q)
	en = excl_s(M.excl[3], M.excl_eps, d, s);
	if(en != 0.f) { E += en; acc.site_ka(d * s, cb); excl_fix(acc, M.excl[3], M.excl_eps, d, s, OXB_SITE_KA, cb); }
	d = r + qback - A.a1 * cb; // base(p) - back(q)
	en = excl_s(M.excl[2], M.excl_eps, d, s);
	if(en != 0.f) { E += en; acc.site_ak(d * s, cb); excl_fix(acc, M.excl[2], M.excl_eps, d, s, OXB_SITE_AK, cb); }
	return E;
}

// the same with the families the list builder has ruled out until the next rebuild skipped (cls: OXB_CLS_* bits of the near edge)
template<class PB> OXB_HD float dna2_excl_cls(const PB &M, int cls, v3 r, v3 rbb, v3 rb, const Axes &A, const Axes &B, v3 pback, v3 qback, PairAcc &acc) {
	const float cb = M.base_a1;
	float s, E = 0.f, en;
	if(cls & OXB_CLS_BB) {
		en = excl_s(M.excl[0], M.excl_eps, rbb, s);
		if(en != 0.f) { E += en; acc.site_kk(rbb * s); excl_fix(acc, M.excl[0], M.excl_eps, rbb, s, OXB_SITE_KK, cb); }
	}
	if(cls & OXB_CLS_EB) {
		en = excl_s(M.excl[1], M.excl_eps, rb, s);
		if(en != 0.f) { E += en; acc.site_aa(rb * s, cb, cb); excl_fix(acc, M.excl[1], M.excl_eps, rb, s, OXB_SITE_AA, cb); }
	}
	if(cls & OXB_CLS_BK) {
		v3 d = r + B.a1 * cb - pback; // back(p) - base(q)
		en = excl_s(M.excl[3], M.excl_eps, d, s);
		if(en != 0.f) { E += en; acc.site_ka(d * s, cb); excl_fix(acc, M.excl[3], M.excl_eps, d, s, OXB_SITE_KA, cb); }
		d = r + qback - A.a1 * cb; // base(p) - back(q)
		en = excl_s(M.excl[2], M.excl_eps, d, s);
		if(en != 0.f) { E += en; acc.site_ak(d * s, cb); excl_fix(acc, M.excl[2], M.excl_eps, d, s, OXB_SITE_AK, cb); }
	}
	return E;
}

// which of the four site pairs are inside their excluded-volume range (bit = OXB_SITE_*)
template<class PB> OXB_HD int dna2_excl_mask(const PB &M, v3 r, v3 rbb, v3 rb, const Axes &A, const Axes &B, v3 pback, v3 qback) {
	const float cb = M.base_a1;
	int m = 0;
	if(dot(rbb, rbb) < M.excl[0].rc2) m |= 1 << OXB_SITE_KK;
	if(dot(rb, rb) < M.excl[1].rc2) m |= 1 << OXB_SITE_AA;
	v3 d = r + B.a1 * cb - pback;
	if(dot(d, d) < M.excl[3].rc2) m |= 1 << OXB_SITE_KA;
	d = r + qback - A.a1 * cb;
	if(dot(d, d) < M.excl[2].rc2) m |= 1 << OXB_SITE_AK;
	return m;
}

OXB_HD bool dna2_hb_in_range(const oxb_dna2_params &M, float rbm2, int btp, int btq) {
	return (btp + btq == 3) && rbm2 > M.hb.rclow * M.hb.rclow && rbm2 < M.hb.rchigh * M.hb.rchigh;
}
OXB_HD bool dna2_crst_in_range(const oxb_dna2_params &M, float rbm2) {
	return rbm2 > M.crst.rclow * M.crst.rclow && rbm2 < M.crst.rchigh * M.crst.rchigh;
}
OXB_HD bool dna2_cxst_in_range(const oxb_dna2_params &M, float rs2) {
	return rs2 > M.cxst.rclow * M.cxst.rclow && rs2 < M.cxst.rchigh * M.cxst.rchigh;
}

// Can the hydrogen-bonding / cross-stacking product be non-zero?  Only dot products: every factor f4 vanishes outside a
// cosine window.  h = rb / |rb|.
OXB_HD bool dna2_hbcr_may_act(const oxb_dna2_params &M, v3 h, const Axes &A, const Axes &B, bool hb_on, bool cr_on) {
	float c1 = -dot(A.a1, B.a1), c2 = -dot(B.a1, h), c3 = dot(A.a1, h), c4 = dot(A.a3, B.a3), c7 = -dot(B.a3, h), c8 = dot(A.a3, h);
	bool hb = hb_on && in_window(M, OXB_F4_HB_T1, c1) && in_window(M, OXB_F4_HB_T2, c2) && in_window(M, OXB_F4_HB_T2, c3) &&
			in_window(M, OXB_F4_HB_T4, c4) && in_window(M, OXB_F4_HB_T7, c7) && in_window(M, OXB_F4_HB_T7, c8);
	bool cr = cr_on && in_window(M, OXB_F4_CRST_T1, c1) && in_window(M, OXB_F4_CRST_T2, c2) && in_window(M, OXB_F4_CRST_T2, c3) &&
			in_window_sym(M, OXB_F4_CRST_T4, c4) && in_window_sym(M, OXB_F4_CRST_T7, c7) && in_window_sym(M, OXB_F4_CRST_T7, c8);
	return hb || cr;
}

OXB_HD bool dna2_cxst_may_act(const oxb_dna2_params &M, v3 h, const Axes &A, const Axes &B) {
	float c1 = -dot(A.a1, B.a1), c4 = dot(A.a3, B.a3), c5 = dot(A.a3, h), c6 = -dot(B.a3, h);
	return in_window(M, OXB_F4_CXST_T1, c1) && in_window(M, OXB_F4_CXST_T4, c4) && in_window_sym(M, OXB_F4_CXST_T5, c5) &&
			in_window_sym(M, OXB_F4_CXST_T5, c6);
}

// rb = base-base vector.  Returns total energy (HB + cross stacking), ehb = the HB part.
template<bool WITH_HB = true>
OXB_HD float dna2_hbcr(const oxb_dna2_params &M, v3 rb, float rbm2, const Axes &A, const Axes &B, int btp, int btq, bool hb_on, bool cr_on,
		PairAcc &acc, float &ehb) {
	const float cb = M.base_a1;
	float E = 0.f;
	ehb = 0.f;
	float inv = OXB_RSQRT(rbm2);
	float m = rbm2 * inv;
	v3 h = rb * inv;
	Angle t1 = make_angle(-A.a1, B.a1);
	Angle t2 = make_angle(-B.a1, h);
	Angle t3 = make_angle(A.a1, h);
	Angle t4 = make_angle(A.a3, B.a3);
	Angle t7 = make_angle(-B.a3, h);
	Angle t8 = make_angle(A.a3, h);
	float g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f, g7 = 0.f, g8 = 0.f, grad = 0.f;
	if(WITH_HB && hb_on) {
		int ti = btype_to_type(btq) * 5 + btype_to_type(btp);
		float mult = (abs(btq) >= 300 && abs(btp) >= 300) ? M.hb_multiplier : 1.f;
		RadVal f1 = f1_r(M.hb, M.hb_eps[ti], M.hb_shift[ti], m);
		f1.v *= mult;
		f1.d *= mult;
		AngVal a1 = f4_ts(M.f4[OXB_F4_HB_T1], t1.t, t1.s);
		AngVal a2 = f4_ts(M.f4[OXB_F4_HB_T2], t2.t, t2.s);
		AngVal a3 = f4_ts(M.f4[OXB_F4_HB_T2], t3.t, t3.s);
		AngVal a4 = f4_ts(M.f4[OXB_F4_HB_T4], t4.t, t4.s);
		AngVal a7 = f4_ts(M.f4[OXB_F4_HB_T7], t7.t, t7.s);
		AngVal a8 = f4_ts(M.f4[OXB_F4_HB_T7], t8.t, t8.s);
		float p12 = a1.v * a2.v, p34 = a3.v * a4.v, p78 = a7.v * a8.v;
		float ang = p12 * p34 * p78;
		float e = f1.v * ang;
		if(e != 0.f) {
			E += e;
			ehb += e;
			grad += f1.d * ang;
			float f34_78 = f1.v * p34 * p78, f12_78 = f1.v * p12 * p78, f12_34 = f1.v * p12 * p34;
			g1 += f34_78 * a1.dc * a2.v;
			g2 += f34_78 * a1.v * a2.dc;
			g3 += f12_78 * a3.dc * a4.v;
			g4 += f12_78 * a3.v * a4.dc;
			g7 += f12_34 * a7.dc * a8.v;
			g8 += f12_34 * a7.v * a8.dc;
		}
	}
	if(cr_on) {
		RadVal f2 = f2_r(M.crst, m);
		AngVal a1 = f4_ts(M.f4[OXB_F4_CRST_T1], t1.t, t1.s);
		AngVal a2 = f4_ts(M.f4[OXB_F4_CRST_T2], t2.t, t2.s);
		AngVal a3 = f4_ts(M.f4[OXB_F4_CRST_T2], t3.t, t3.s);
		AngVal a4 = f4_ts_sym(M.f4[OXB_F4_CRST_T4], t4.t, t4.s);
		AngVal a7 = f4_ts_sym(M.f4[OXB_F4_CRST_T7], t7.t, t7.s);
		AngVal a8 = f4_ts_sym(M.f4[OXB_F4_CRST_T7], t8.t, t8.s);
		float p12 = a1.v * a2.v, p34 = a3.v * a4.v, p78 = a7.v * a8.v;
		float ang = p12 * p34 * p78;
		float e = f2.v * ang;
		if(e != 0.f) {
			E += e;
			grad += f2.d * ang;
			float f34_78 = f2.v * p34 * p78, f12_78 = f2.v * p12 * p78, f12_34 = f2.v * p12 * p34;
			g1 += f34_78 * a1.dc * a2.v;
			g2 += f34_78 * a1.v * a2.dc;
			g3 += f12_78 * a3.dc * a4.v;
			g4 += f12_78 * a3.v * a4.dc;
			g7 += f12_34 * a7.dc * a8.v;
			g8 += f12_34 * a7.v * a8.dc;
		}
	}
	if(E != 0.f) {
		v3 f = h * (-grad);
		chain_bb(acc, g1, t1);
		f += chain_bd<true>(acc, g2, -B.a1, h, inv, t2);
		f += chain_bd<false>(acc, g3, A.a1, h, inv, t3);
		chain_bb(acc, g4, t4);
		f += chain_bd<true>(acc, g7, -B.a3, h, inv, t7);
		f += chain_bd<false>(acc, g8, A.a3, h, inv, t8);
		acc.site_aa(f, cb, cb);
	}
	return E;
}

// rs = stack-stack vector
OXB_HD float dna2_cxst(const oxb_dna2_params &M, v3 rs, float rs2, const Axes &A, const Axes &B, PairAcc &acc) {
	const float cs = M.stack_a1;
	float inv = OXB_RSQRT(rs2);
	float m = rs2 * inv;
	v3 h = rs * inv;
	Angle t1 = make_angle(-A.a1, B.a1);
	Angle t4 = make_angle(A.a3, B.a3);
	Angle t5 = make_angle(A.a3, h);
	Angle t6 = make_angle(-B.a3, h);
	RadVal f2 = f2_r(M.cxst, m);
	AngVal a1 = M.v1 ? f4_ts_rna_cxst_t1(M.f4[OXB_F4_CXST_T1], t1.t, t1.s) : f4_ts_cxst_t1(M, t1.t, t1.s);
	AngVal a4 = f4_ts(M.f4[OXB_F4_CXST_T4], t4.t, t4.s);
	AngVal a5 = f4_ts_sym(M.f4[OXB_F4_CXST_T5], t5.t, t5.s);
	AngVal a6 = f4_ts_sym(M.f4[OXB_F4_CXST_T5], t6.t, t6.s);
	float p14 = a1.v * a4.v, p56 = a5.v * a6.v;
	float e = f2.v * p14 * p56;
	if(M.v1 && e != 0.f) {
		// oxDNA: times f5(cos phi3)^2, cos phi3 = shat . (bhat_ref x a1), bhat_ref between the ungrooved backbone reference sites
		v3 w = rs + (B.a1 - A.a1) * (M.backref_a1 - cs);
		float winv = OXB_RSQRT(dot(w, w));
		v3 wh = w * winv;
		AngVal b3 = f5_c(M.phi3, dot(h, cross(wh, A.a1)));
		float e0 = e;
		e *= b3.v * b3.v;
		f2.v *= b3.v * b3.v;
		f2.d *= b3.v * b3.v;
		if(e != 0.f && b3.dc != 0.f) chain_triple<false, true>(acc, e0 * 2.f * b3.v * b3.dc, A.a1, h, inv, cs, wh, winv, M.backref_a1);
	}
	if(e != 0.f) {
		v3 f = h * (-(f2.d * p14 * p56));
		chain_bb(acc, f2.v * p56 * a1.dc * a4.v, t1);
		chain_bb(acc, f2.v * p56 * a1.v * a4.dc, t4);
		f += chain_bd<false>(acc, f2.v * p14 * a5.dc * a6.v, A.a3, h, inv, t5);
		f += chain_bd<true>(acc, f2.v * p14 * a5.v * a6.dc, -B.a3, h, inv, t6);
		acc.site_aa(f, cs, cs);
	}
	return e;
}

// the whole non-bonded interaction of one pair (particle-centric kernel, host-side unit test)
OXB_HD PairEnergy dna2_nonbonded(const oxb_dna2_params &M, v3 r, const Axes &A, const Axes &B, int btp, int btq,
		bool p_end, bool q_end, v3 pback, v3 qback, PairAcc &acc) {
	PairEnergy E;
	E.total = 0.f;
	E.hb = 0.f;
	float r2 = dot(r, r);
	if(r2 >= (acc.rep ? acc.rep->rcut2 : M.rcut * M.rcut)) return E; // DNA2Interaction.cpp:46-48
	v3 rbb = r + qback - pback;
	float fs;
	float en = acc.rep ? dna2_dh(dh_view(M, acc.rep), dot(rbb, rbb), p_end, q_end, fs) : dna2_dh(M, dot(rbb, rbb), p_end, q_end, fs);
	if(en != 0.f) { E.total += en; acc.site_kk(rbb * fs); }
	if(r2 >= M.rcut_near * M.rcut_near) return E;
	v3 rb = r + (B.a1 - A.a1) * M.base_a1;
	E.total += dna2_excl(M, r, rbb, rb, A, B, pback, qback, acc);
	float rbm2 = dot(rb, rb);
	bool hb_on = dna2_hb_in_range(M, rbm2, btp, btq), cr_on = dna2_crst_in_range(M, rbm2);
	if(hb_on || cr_on) {
		float ehb;
		E.total += dna2_hbcr(M, rb, rbm2, A, B, btp, btq, hb_on, cr_on, acc, ehb);
		E.hb += ehb;
	}
	v3 rs = r + (B.a1 - A.a1) * M.stack_a1;
	float rs2 = dot(rs, rs);
	if(dna2_cxst_in_range(M, rs2)) E.total += dna2_cxst(M, rs, rs2, A, B, acc);
	return E;
}

// FENE backbone + bonded excluded volume (3 site pairs): the same functional form in oxDNA2 and oxRNA2
// (DNAInteraction.cpp:415-528, RNAInteraction.cpp:431-531)
// The FENE spring is the stiffest term of the model (dF/dr of several thousand for a strained bond): an FP32 backbone-backbone
// distance (ulp 6e-8 at 0.75) limits its force to ~1e-5 of max|F| in boxes of a few hundred sigma.  The kernels therefore take that one
// distance in double from the fixed-point backbone sites (exact integer difference x L / 2^32) and hand the result in here.
struct FeneSite {
	v3 d;        // backbone(q) - backbone(p)
	float s, en; // force on q = d * s; energy
	int has_fene; // 0: d, s, en are not set, the FP32 FENE expression is used (bonds far from the stiff end of the FENE range)
	int excl_deferred; // bits OXB_SITE_*: bonded excluded-volume site pairs the caller evaluates itself (in double)
};

#ifdef __CUDACC__
template<class PB> __device__ __forceinline__ FeneSite fene_from_sites(const PB &M, const BoxF &box, int4 ibp, int4 ibq, bool &broken) {
	const double dx = (double) (int) ((unsigned) ibq.x - (unsigned) ibp.x) * box.dsx;
	const double dy = (double) (int) ((unsigned) ibq.y - (unsigned) ibp.y) * box.dsy;
	const double dz = (double) (int) ((unsigned) ibq.z - (unsigned) ibp.z) * box.dsz;
	const double m = sqrt(dx * dx + dy * dy + dz * dz);
	const double x = m - (double) M.fene_r0;
	double en, s;
	if(M.use_mbf && fabs(x) > (double) M.mbf_xmax) {
		const double ax = fabs(x);
		const double k = ((double) M.mbf_fmax - (double) M.mbf_finf) * (double) M.mbf_xmax;
		en = k * log(ax) + (double) M.mbf_finf * ax + (double) M.mbf_e0;
		s = -copysign(1., x) * (k / ax + (double) M.mbf_finf) / m;
	}
	else {
		double den = (double) M.fene_delta2 - x * x;
		if(den <= 0.) {
			broken = true;
			den = 1e-6;
		}
		en = -0.5 * (double) M.fene_eps * log(den / (double) M.fene_delta2);
		s = -(double) M.fene_eps * x / den / m;
	}
	FeneSite f;
	f.d = mk3((float) dx, (float) dy, (float) dz);
	f.s = (float) s;
	f.en = (float) en;
	f.has_fene = 1;
	f.excl_deferred = 0;
	return f;
}
#endif

// bonded pair: which of the three excluded-volume site pairs (base-base, base-back, back-base) are in range
template<class PB> OXB_HD int bonded_excl_mask(const PB &M, v3 r, const Axes &A, const Axes &B, v3 pback, v3 qback) {
	const float cb = M.base_a1;
	v3 pbase = A.a1 * cb, qbase = B.a1 * cb;
	int m = 0;
	v3 d = r + qbase - pbase;
	if(dot(d, d) < M.excl[1].rc2) m |= 1 << OXB_SITE_AA;
	d = r + qback - pbase;
	if(dot(d, d) < M.excl[2].rc2) m |= 1 << OXB_SITE_AK;
	d = r + qbase - pback;
	if(dot(d, d) < M.excl[3].rc2) m |= 1 << OXB_SITE_KA;
	return m;
}

template<class PB>
OXB_HD float bonded_fene_excl(const PB &M, v3 r, const Axes &A, const Axes &B, v3 pback, v3 qback, PairAcc &acc, bool &broken, float *esplit = nullptr,
		const FeneSite *fene = nullptr) {
	float E = 0.f;
	const float cb = M.base_a1;
	// FENE
	if(fene != nullptr && fene->has_fene) {
		E += fene->en;
		if(esplit) esplit[0] += fene->en;
		acc.site_kk(fene->d * fene->s);
	}
	else {
		v3 d = r + qback - pback;
		float d2 = dot(d, d);
		float invm = OXB_RSQRT(d2);
		float m = d2 * invm;
		float x = m - M.fene_r0;
		float en, s;
		if(M.use_mbf && fabsf(x) > M.mbf_xmax) {
			float ax = fabsf(x);
			float k = (M.mbf_fmax - M.mbf_finf) * M.mbf_xmax;
			en = k * logf(ax) + M.mbf_finf * ax + M.mbf_e0;
			s = -copysignf(1.f, x) * (OXB_DIV(k, ax) + M.mbf_finf) * invm;
		}
		else {
			float den = M.fene_delta2 - x * x;
			if(den <= 0.f) {
				broken = true;
				den = 1e-6f;
			}
			en = -0.5f * M.fene_eps * logf(den / M.fene_delta2);
			s = -OXB_DIV(M.fene_eps * x, den) * invm;
		}
		E += en;
		if(esplit) esplit[0] += en;
		acc.site_kk(d * s);
	}
	// bonded excluded volume
	const int deferred = fene != nullptr ? fene->excl_deferred : 0;
	if(deferred == 0) {
		float s;
		v3 pbase = A.a1 * cb, qbase = B.a1 * cb;
		v3 d = r + qbase - pbase;
		float en = excl_s(M.excl[1], M.excl_eps, d, s);
		if(en != 0.f) { E += en; if(esplit) esplit[1] += en; acc.site_aa(d * s, cb, cb); excl_fix(acc, M.excl[1], M.excl_eps, d, s, OXB_SITE_AA, cb); }
		d = r + qback - pbase;
		en = excl_s(M.excl[2], M.excl_eps, d, s);
		if(en != 0.f) { E += en; if(esplit) esplit[1] += en; acc.site_ak(d * s, cb); excl_fix(acc, M.excl[2], M.excl_eps, d, s, OXB_SITE_AK, cb); }
		d = r + qbase - pback;
		en = excl_s(M.excl[3], M.excl_eps, d, s);
		if(en != 0.f) { E += en; if(esplit) esplit[1] += en; acc.site_ka(d * s, cb); excl_fix(acc, M.excl[3], M.excl_eps, d, s, OXB_SITE_KA, cb); }
	}
	return E;
}

// ---------------------------------------------------------------------------------------------------------------
// Bonded pair p -> q = n3(p).  r = q - p.  Terms: FENE backbone, bonded excluded volume (3 site pairs), stacking.
// Returns energy; sets *broken when the bond is outside the FENE range (reference throws; we flag).
// ---------------------------------------------------------------------------------------------------------------
OXB_HD float dna2_bonded(const oxb_dna2_params &M, v3 r, const Axes &A, const Axes &B, int btp, int btq, v3 pback,
		v3 qback, PairAcc &acc, bool &broken, float *esplit = nullptr, const FeneSite *fene = nullptr) {
	float E = 0.f;
	const float cb = M.base_a1, cs = M.stack_a1, cr = M.backref_a1;

	E += bonded_fene_excl(M, r, A, B, pback, qback, acc, broken, esplit, fene);
	// stacking
	{
		v3 rs = r + B.a1 * cs - A.a1 * cs;
		float rs2 = dot(rs, rs);
		float inv = OXB_RSQRT(rs2);
		float m = rs2 * inv;
		int ti = btype_to_type(btq) * 5 + btype_to_type(btp);
		RadVal f1 = acc.rep ? f1_r(M.stck, acc.rep->stck_eps[ti], acc.rep->stck_shift[ti], m) : f1_r(M.stck, M.stck_eps[ti], M.stck_shift[ti], m);
		if(f1.v != 0.f || f1.d != 0.f) {
			v3 h = rs * inv;
			v3 w = r + B.a1 * cr - A.a1 * cr;
			float winv = OXB_RSQRT(dot(w, w));
			v3 wh = w * winv;
			Angle t4 = make_angle(A.a3, B.a3);
			Angle t5 = make_angle(-A.a3, h);
			Angle t6 = make_angle(-B.a3, h);
			Angle p1 = make_angle_cos(A.a2, wh);
			Angle p2 = make_angle_cos(B.a2, wh);
			AngVal a4 = f4_ts(M.f4[OXB_F4_STCK_T4], t4.t, t4.s);
			AngVal a5 = f4_ts(M.f4[OXB_F4_STCK_T5], t5.t, t5.s);
			AngVal a6 = f4_ts(M.f4[OXB_F4_STCK_T5], t6.t, t6.s);
			AngVal b1 = f5_c(M.phi1, p1.c);
			AngVal b2 = f5_c(M.phi2, p2.c);
			float p456 = a4.v * a5.v * a6.v, pb = b1.v * b2.v;
			float e = f1.v * p456 * pb;
			if(e != 0.f) {
				E += e;
				if(esplit) esplit[2] += e;
				v3 f = h * (-(f1.d * p456 * pb));
				float fb = f1.v * pb;
				chain_bb(acc, fb * a4.dc * a5.v * a6.v, t4);
				f += chain_bd<false>(acc, fb * a4.v * a5.dc * a6.v, -A.a3, h, inv, t5);
				f += chain_bd<true>(acc, fb * a4.v * a5.v * a6.dc, -B.a3, h, inv, t6);
				acc.site_aa(f, cs, cs);
				float fa = f1.v * p456;
				v3 fw = chain_bd<false>(acc, fa * b1.dc * b2.v, A.a2, wh, winv, p1);
				fw += chain_bd<true>(acc, fa * b1.v * b2.dc, B.a2, wh, winv, p2);
				acc.site_aa(fw, cr, cr);
			}
		}
	}
	return E;
}
