// oxDNA3 pair potential, FP32 device functions (sm_100a).  What it evaluates is the reference's model
// (src/CUDA/Interactions/CUDA_DNA3.cuh:1-1085, CPU mirror src/Interactions/DNA3Interaction.cpp:1057-2059): the oxDNA2 functional forms with
// every radial / angular / excluded-volume / FENE parameter looked up per "tetramer" (n3_2, n3_1, n5_1, n5_2) -- the types of the two
// interacting nucleotides and of one flanking neighbour each -- and with per-type stacking / base site offsets.
//
// How it is laid out is different.  The reference keeps 214 separate 900-entry tables in global memory and reads them entry by entry
// (one function evaluates f1 with eleven scattered loads, CUDA_DNA3.cuh:119-160).  Here the host packs them once into RECORDS: everything a
// term needs for one tetramer is contiguous and 16-byte aligned (bonded: 48 floats, cross stacking: 32 per diagonal, coaxial: 12; hydrogen
// bonding and the non-bonded excluded volume depend on the two types only: 25 records), fetched with 128-bit read-only loads.  The whole
// set is 0.45 MB and lives in L2.  The neighbour types of a particle come from one packed word per particle (type | n3 type << 3 |
// n5 type << 6 | dummy << 9), indexed by original id, so a pair costs one extra 4-byte gather.
//
// The chain rule is the cosine-space machinery of dna_model.cuh -- except for the phi1 / phi2 factors of the stacking term, where the
// reference's force is NOT the gradient of its energy (it differentiates through the stacking-site vector with the oxDNA2 lever
// gamma = POS_STACK - POS_BACK = 0.74 while the oxDNA3 site sits at 0.37: DNA3Interaction.cpp:1363, CUDA_DNA3.cuh:510).  Parity is with
// the reference, so that expression is evaluated as the reference writes it.
#pragma once

#include "dna_model.cuh"
#include "kernels.h"

struct Nuc3 {
	int type, n3t, n5t, si; // si: index into pos_stack / pos_base (0 = dummy base)
	bool has_n3, has_n5;
};
OXB_HD Nuc3 nuc3_from_code(int code) {
	Nuc3 n;
	n.type = code & 7; n.n3t = (code >> 3) & 7; n.n5t = (code >> 6) & 7;
	n.si = ((code >> 9) & 1) ? 0 : 1 + n.type;
	n.has_n3 = n.n3t != 5; n.has_n5 = n.n5t != 5;
	return n;
}
OXB_HD int ix4(int a, int b, int c, int d) { return ((a * 5 + b) * 5 + c) * 6 + d; }

// NV float4 of a record starting at p
template<int NV>
OXB_HD void load_rec(const float4 *p, float *out) {
#pragma unroll
	for(int k = 0; k < NV; k++) {
#ifdef __CUDA_ARCH__
		const float4 v = __ldg(p + k);
#else
		const float4 v = p[k];
#endif
		out[4 * k] = v.x; out[4 * k + 1] = v.y; out[4 * k + 2] = v.z; out[4 * k + 3] = v.w;
	}
}

// record pieces.  f1 (12 floats): a, rc, r0, blow | bhigh, rlow, rhigh, rclow | rchigh, eps, shift, -
OXB_HD RadVal f1_rec(const float *r, float x) {
	oxb_f1 f;
	f.a = r[0]; f.rc = r[1]; f.r0 = r[2]; f.blow = r[3]; f.bhigh = r[4]; f.rlow = r[5]; f.rhigh = r[6]; f.rclow = r[7]; f.rchigh = r[8];
	return f1_r(f, r[9], r[10], x);
}
// f2 (12 floats): k, rc, r0, blow | rlow, rclow, bhigh, rhigh | rchigh, k_symm, -, -
OXB_HD oxb_f2 f2_rec(const float *r) {
	oxb_f2 f;
	f.k = r[0]; f.rc = r[1]; f.r0 = r[2]; f.blow = r[3]; f.rlow = r[4]; f.rclow = r[5]; f.bhigh = r[6]; f.rhigh = r[7]; f.rchigh = r[8];
	return f;
}
OXB_HD oxb_f4 f4_rec(const float *r) {
	oxb_f4 f;
	f.a = r[0]; f.b = r[1]; f.t0 = r[2]; f.ts = r[3]; f.tc = r[4];
	return f;
}
OXB_HD oxb_f5 f5_rec(const float *r) {
	oxb_f5 f;
	f.a = r[0]; f.b = r[1]; f.xc = r[2]; f.xs = r[3];
	return f;
}
OXB_HD oxb_excl excl_rec(const float *r) {
	oxb_excl e;
	e.sigma2 = r[0]; e.rstar2 = r[1]; e.b = r[2]; e.rc = r[3]; e.rc2 = r[3] * r[3];
	return e;
}

// one excluded-volume site pair with its own lever coefficients (cp, cq: offsets along a1 when the site is a base, unused for a backbone site)
OXB_HD float excl3(const oxb_dna3_dev &M, const oxb_excl &e, v3 d, int kind, float cp, float cq, PairAcc &acc) {
	float s;
	const float en = excl_s(e, M.excl_eps, d, s);
	if(en != 0.f) {
		if(kind == OXB_SITE_KK) acc.site_kk(d * s);
		else if(kind == OXB_SITE_AA) acc.site_aa(d * s, cp, cq);
		else if(kind == OXB_SITE_AK) acc.site_ak(d * s, cp);
		else acc.site_ka(d * s, cq);
		excl_fix2(acc, e, M.excl_eps, d, s, kind, cp, cq);
	}
	return en;
}

// the six angles of the base-base vector and the accumulation of one product f(r) f4(t1) f4(t2) f4(t3) f4(t4) f4(t7) f4(t8)
struct Six {
	Angle t1, t2, t3, t4, t7, t8;
	float g1, g2, g3, g4, g7, g8, grad, E;
};
// f4 block of a record: theta1, theta2 (= theta3), theta4, theta7 (= theta8), 5 floats each, fetched where it is used (the records are not
// held in registers across the evaluation: the kernel is register-bound)
OXB_HD void six_add(Six &S, RadVal f, const float4 *f4blk, float &e_out) {
	float q[20];
	load_rec<5>(f4blk, q);
	const oxb_f4 p1 = f4_rec(q), p2 = f4_rec(q + 5), p4 = f4_rec(q + 10), p7 = f4_rec(q + 15);
	const AngVal a1 = f4_ts(p1, S.t1.t, S.t1.s), a2 = f4_ts(p2, S.t2.t, S.t2.s), a3 = f4_ts(p2, S.t3.t, S.t3.s);
	const AngVal a4 = f4_ts(p4, S.t4.t, S.t4.s), a7 = f4_ts(p7, S.t7.t, S.t7.s), a8 = f4_ts(p7, S.t8.t, S.t8.s);
	const float p12 = a1.v * a2.v, p34 = a3.v * a4.v, p78 = a7.v * a8.v;
	const float ang = p12 * p34 * p78;
	const float e = f.v * ang;
	e_out = e;
	if(e != 0.f) {
		S.E += e;
		S.grad += f.d * ang;
		const float f34_78 = f.v * p34 * p78, f12_78 = f.v * p12 * p78, f12_34 = f.v * p12 * p34;
		S.g1 += f34_78 * a1.dc * a2.v;
		S.g2 += f34_78 * a1.v * a2.dc;
		S.g3 += f12_78 * a3.dc * a4.v;
		S.g4 += f12_78 * a3.v * a4.dc;
		S.g7 += f12_34 * a7.dc * a8.v;
		S.g8 += f12_34 * a7.v * a8.dc;
	}
}

// ---- non-bonded pair (p, q), in the three pieces the force kernel runs as separate uniform loops (forces_dna3.cu); r = min-image(q - p)
// between the centres, forces are "on q".  esplit (optional): per-term energies in the order of OXB_TERM_*.

// the four excluded-volume site pairs, parameters of the tetramer (-, q, p, -): DNA3Interaction.cpp:1137-1228
OXB_HD float dna3_excl4(const oxb_dna3_dev &M, v3 r, v3 rbb, v3 rb, const Axes &A, const Axes &B, const Nuc3 &np, const Nuc3 &nq, v3 pback, v3 qback,
		PairAcc &acc) {
	const float cbp = M.pos_base[np.si], cbq = M.pos_base[nq.si];
	float rec[OXB3_REC_NEXCL];
	load_rec<OXB3_REC_NEXCL / 4>(M.nexcl + (nq.type * 5 + np.type) * (OXB3_REC_NEXCL / 4), rec);
	float en = excl3(M, excl_rec(rec + 4), rb, OXB_SITE_AA, cbp, cbq, acc);
	en += excl3(M, excl_rec(rec + 12), r + B.a1 * cbq - pback, OXB_SITE_KA, cbp, cbq, acc);
	en += excl3(M, excl_rec(rec + 8), r + qback - A.a1 * cbp, OXB_SITE_AK, cbp, cbq, acc);
	en += excl3(M, excl_rec(rec), rbb, OXB_SITE_KK, cbp, cbq, acc);
	return en;
}

// hydrogen bonding (parameters of (0, q, p, 0), DNA3Interaction.cpp:1477-1598) + cross stacking (3'3' diagonal: tetramer (n3(q), q, p, n3(p)),
// 5'5' diagonal: (n5(q), q, p, n5(p)); DNA3Interaction.cpp:1600-1755) on the base-base vector rb; returns the total, ehb = the HB part
OXB_HD float dna3_hbcr(const oxb_dna3_dev &M, v3 rb, float rbm2, const Axes &A, const Axes &B, int btp, int btq, const Nuc3 &np, const Nuc3 &nq,
		PairAcc &acc, float &ehb, float *esplit = nullptr) {
	ehb = 0.f;
	const float cbp = M.pos_base[np.si], cbq = M.pos_base[nq.si];
	const float inv = OXB_RSQRT(rbm2);
	const float m = rbm2 * inv;
	const v3 h = rb * inv;
	const float c7 = -dot(B.a3, h), c8 = dot(A.a3, h);
	// gates first, from the two float4 of each record that hold rclow / rchigh (floats 7, 8 of an f1 record, 5, 8 of an f2 record)
	const float4 *hbr = M.hb + (nq.type * 5 + np.type) * (OXB3_REC_HB / 4);
	bool hb_on = (btp + btq == 3);
	if(hb_on) {
		float g[8];
		load_rec<2>(hbr + 1, g);
		hb_on = g[3] < m && m < g[4];
	}
	const float4 *c33 = M.crst + ix4(nq.n3t, nq.type, np.type, np.n3t) * (OXB3_REC_CRST / 4);
	const float4 *c55 = M.crst + (900 + ix4(nq.n5t, nq.type, np.type, np.n5t)) * (OXB3_REC_CRST / 4);
	bool in33 = c7 > 0.f && c8 > 0.f, in55 = c7 < 0.f && c8 < 0.f;
	if(in33) { float g[8]; load_rec<2>(c33 + 1, g); in33 = g[1] < m && m < g[4]; }
	if(in55) { float g[8]; load_rec<2>(c55 + 1, g); in55 = g[1] < m && m < g[4]; }
	if(!(hb_on || in33 || in55)) return 0.f;
	Six S;
	S.t1 = make_angle(-A.a1, B.a1); S.t2 = make_angle(-B.a1, h); S.t3 = make_angle(A.a1, h);
	S.t4 = make_angle(A.a3, B.a3); S.t7 = make_angle(-B.a3, h); S.t8 = make_angle(A.a3, h);
	S.g1 = S.g2 = S.g3 = S.g4 = S.g7 = S.g8 = S.grad = S.E = 0.f;
	if(hb_on) {
		const float mult = (abs(btq) >= 300 && abs(btp) >= 300) ? M.hb_multiplier : 1.f;
		float fr[12];
		load_rec<3>(hbr, fr);
		RadVal f1 = f1_rec(fr, m);
		f1.v *= mult; f1.d *= mult;
		float e;
		six_add(S, f1, hbr + 3, e);
		ehb += e;
		if(esplit) esplit[4] += e;
	}
	if(in33 || in55) {
		// both diagonals are evaluated once either gate is open, as the reference does (DNA3Interaction.cpp:1640-1665): each f2 has its own range
		float fr[12], e;
		load_rec<3>(c33, fr);
		six_add(S, f2_r(f2_rec(fr), m), c33 + 3, e);
		if(esplit) esplit[5] += e;
		load_rec<3>(c55, fr);
		six_add(S, f2_r(f2_rec(fr), m), c55 + 3, e);
		if(esplit) esplit[5] += e;
	}
	if(S.E != 0.f) {
		v3 f = h * (-S.grad);
		chain_bb(acc, S.g1, S.t1);
		f += chain_bd<true>(acc, S.g2, -B.a1, h, inv, S.t2);
		f += chain_bd<false>(acc, S.g3, A.a1, h, inv, S.t3);
		chain_bb(acc, S.g4, S.t4);
		f += chain_bd<true>(acc, S.g7, -B.a3, h, inv, S.t7);
		f += chain_bd<false>(acc, S.g8, A.a3, h, inv, S.t8);
		acc.site_aa(f, cbp, cbq);
	}
	return S.E;
}

// coaxial stacking on the stack-stack vector rs: radial part per tetramer (three K branches, DNA3Interaction.cpp:1814-1825), angular part the
// scalar oxDNA2 set (DNA3Interaction.cpp:1758-1883)
OXB_HD float dna3_cxst(const oxb_dna3_dev &M, v3 rs, float rs2, const Axes &A, const Axes &B, const Nuc3 &np, const Nuc3 &nq, PairAcc &acc) {
	const float csp = M.pos_stack[np.si], csq = M.pos_stack[nq.si];
	float r0[OXB3_REC_CXST];
	load_rec<OXB3_REC_CXST / 4>(M.cxst + ix4(0, nq.type, np.type, 0) * (OXB3_REC_CXST / 4), r0);
	const float inv = OXB_RSQRT(rs2);
	const float m = rs2 * inv;
	if(!(r0[5] < m && m < r0[8])) return 0.f;
	int ix;
	bool symm = false;
	if(!np.has_n3 && !nq.has_n5) ix = ix4(nq.n3t, nq.type, np.type, np.n5t);
	else if(!np.has_n5 && !nq.has_n3) ix = ix4(np.n5t, np.type, nq.type, nq.n3t);
	else { ix = ix4(nq.n3t, nq.type, np.type, np.n5t); symm = true; }
	float rc[OXB3_REC_CXST];
	load_rec<OXB3_REC_CXST / 4>(M.cxst + ix * (OXB3_REC_CXST / 4), rc);
	oxb_f2 fp = f2_rec(rc);
	if(symm) fp.k = rc[9];
	const RadVal f2 = f2_r(fp, m);
	const v3 h = rs * inv;
	const Angle t1 = make_angle(-A.a1, B.a1), t4 = make_angle(A.a3, B.a3), t5 = make_angle(A.a3, h), t6 = make_angle(-B.a3, h);
	AngVal a1 = f4_ts(M.cxst_t1, t1.t, t1.s);
	{
		const float x = t1.t - M.cxst_t1_sb;
		if(x >= 0.f) {
			a1.v += M.cxst_t1_sa * x * x;
			a1.dc -= (t1.s * t1.s > 1e-8f) ? OXB_DIV(2.f * M.cxst_t1_sa * x, t1.s) : 2.f * M.cxst_t1_sa;
		}
	}
	const AngVal a4 = f4_ts(M.cxst_t4, t4.t, t4.s), a5 = f4_ts_sym(M.cxst_t5, t5.t, t5.s), a6 = f4_ts_sym(M.cxst_t5, t6.t, t6.s);
	const float p14 = a1.v * a4.v, p56 = a5.v * a6.v;
	const float e = f2.v * p14 * p56;
	if(e != 0.f) {
		v3 f = h * (-(f2.d * p14 * p56));
		chain_bb(acc, f2.v * p56 * a1.dc * a4.v, t1);
		chain_bb(acc, f2.v * p56 * a1.v * a4.dc, t4);
		f += chain_bd<false>(acc, f2.v * p14 * a5.dc * a6.v, A.a3, h, inv, t5);
		f += chain_bd<true>(acc, f2.v * p14 * a5.v * a6.dc, -B.a3, h, inv, t6);
		acc.site_aa(f, csp, csq);
	}
	return e;
}

// the whole non-bonded interaction of one pair (split-energy kernel, host-side unit test); Debye-Hueckel inherited from DNA2Interaction.cpp:157-210
OXB_HD PairEnergy dna3_nonbonded(const oxb_dna3_dev &M, v3 r, const Axes &A, const Axes &B, int btp, int btq, const Nuc3 &np, const Nuc3 &nq,
		v3 pback, v3 qback, PairAcc &acc, float *esplit = nullptr) {
	PairEnergy E;
	E.total = 0.f;
	E.hb = 0.f;
	const float r2 = dot(r, r);
	if(r2 >= M.rcut2) return E;
	const v3 rbb = r + qback - pback;
	{
		float fs;
		const float en = dna2_dh(M, dot(rbb, rbb), !(np.has_n3 && np.has_n5), !(nq.has_n3 && nq.has_n5), fs);
		if(en != 0.f) { E.total += en; acc.site_kk(rbb * fs); if(esplit) esplit[7] += en; }
	}
	if(r2 >= M.r2_near_max) return E; // no site pair of any other term can be in range
	const v3 rb = r + B.a1 * M.pos_base[nq.si] - A.a1 * M.pos_base[np.si];
	const float rbm2 = dot(rb, rb);
	if(r2 < M.r2_excl_max) {
		const float en = dna3_excl4(M, r, rbb, rb, A, B, np, nq, pback, qback, acc);
		E.total += en;
		if(esplit) esplit[3] += en;
	}
	if(rbm2 < M.r2_base_max) E.total += dna3_hbcr(M, rb, rbm2, A, B, btp, btq, np, nq, acc, E.hb, esplit);
	const v3 rs = r + B.a1 * M.pos_stack[nq.si] - A.a1 * M.pos_stack[np.si];
	const float rs2 = dot(rs, rs);
	if(rs2 < M.r2_stack_max) {
		const float en = dna3_cxst(M, rs, rs2, A, B, np, nq, acc);
		E.total += en;
		if(esplit) esplit[6] += en;
	}
	return E;
}

// FENE constants of one bond as the generic helpers of dna_model.cuh expect them (fene_from_sites, FP32 fallback)
struct Fene3 {
	float fene_r0, fene_delta2, fene_eps, mbf_xmax, mbf_fmax, mbf_finf, mbf_e0;
	int use_mbf;
};
OXB_HD Fene3 fene3_of(const oxb_dna3_dev &M, const float4 *rec4) {
	float rec[4];
	load_rec<1>(rec4, rec);
	Fene3 f;
	f.fene_r0 = rec[0]; f.fene_delta2 = rec[1]; f.mbf_xmax = rec[2]; f.mbf_e0 = rec[3];
	f.fene_eps = M.fene_eps; f.mbf_fmax = M.mbf_fmax; f.mbf_finf = M.mbf_finf; f.use_mbf = M.use_mbf;
	return f;
}

// Bonded pair p -> q = n3(p): DNA3Interaction.cpp:1230-1287 (FENE), 1057-1135 (excluded volume), 1289-1475 (stacking).
// rec: the bonded record of the tetramer (n3(q), q, p, n5(p)).
OXB_HD float dna3_bonded(const oxb_dna3_dev &M, const float4 *rec4, v3 r, const Axes &A, const Axes &B, const Nuc3 &np, const Nuc3 &nq, v3 pback, v3 qback,
		PairAcc &acc, bool &broken, float *esplit = nullptr, const FeneSite *fene = nullptr) {
	float E = 0.f;
	const float cbp = M.pos_base[np.si], cbq = M.pos_base[nq.si], csp = M.pos_stack[np.si], csq = M.pos_stack[nq.si], cr = M.backref_a1;
	if(fene != nullptr && fene->has_fene) {
		E += fene->en;
		if(esplit) esplit[0] += fene->en;
		acc.site_kk(fene->d * fene->s);
	}
	else {
		const Fene3 F = fene3_of(M, rec4);
		const v3 d = r + qback - pback;
		const float d2 = dot(d, d);
		const float invm = OXB_RSQRT(d2);
		const float x = d2 * invm - F.fene_r0;
		float en, s;
		if(F.use_mbf && fabsf(x) > F.mbf_xmax) {
			const float ax = fabsf(x), k = (F.mbf_fmax - F.mbf_finf) * F.mbf_xmax;
			en = k * logf(ax) + F.mbf_finf * ax + F.mbf_e0;
			s = -copysignf(1.f, x) * (OXB_DIV(k, ax) + F.mbf_finf) * invm;
		}
		else {
			float den = F.fene_delta2 - x * x;
			if(den <= 0.f) { broken = true; den = 1e-6f; }
			en = -0.5f * F.fene_eps * logf(den / F.fene_delta2);
			s = -OXB_DIV(F.fene_eps * x, den) * invm;
		}
		E += en;
		if(esplit) esplit[0] += en;
		acc.site_kk(d * s);
	}
	{
		float ex[12];
		load_rec<3>(rec4 + 1, ex);
		float en = excl3(M, excl_rec(ex), r + B.a1 * cbq - A.a1 * cbp, OXB_SITE_AA, cbp, cbq, acc);
		en += excl3(M, excl_rec(ex + 4), r + qback - A.a1 * cbp, OXB_SITE_AK, cbp, cbq, acc);
		en += excl3(M, excl_rec(ex + 8), r + B.a1 * cbq - pback, OXB_SITE_KA, cbp, cbq, acc);
		E += en;
		if(esplit) esplit[1] += en;
	}
	const v3 rs = r + B.a1 * csq - A.a1 * csp;
	const float rs2 = dot(rs, rs);
	const float inv = OXB_RSQRT(rs2);
	const float m = rs2 * inv;
	RadVal f1;
	{
		float fr[12];
		load_rec<3>(rec4 + 4, fr);
		f1 = f1_rec(fr, m);
	}
	if(f1.v != 0.f || f1.d != 0.f) {
		float rec[20]; // f4 theta4 (0..4), theta5 (5..9), - , f5 phi1 (12..15), phi2 (16..19)
		load_rec<5>(rec4 + 7, rec);
		const v3 h = rs * inv;
		const v3 w = r + (B.a1 - A.a1) * cr;
		const float w2 = dot(w, w);
		const float winv = OXB_RSQRT(w2);
		const v3 wh = w * winv;
		const Angle t4 = make_angle(A.a3, B.a3), t5 = make_angle(-A.a3, h), t6 = make_angle(-B.a3, h);
		const float cp1 = dot(A.a2, wh), cp2 = dot(B.a2, wh);
		const oxb_f4 p5 = f4_rec(rec + 5);
		const AngVal a4 = f4_ts(f4_rec(rec), t4.t, t4.s), a5 = f4_ts(p5, t5.t, t5.s), a6 = f4_ts(p5, t6.t, t6.s);
		const AngVal b1 = f5_c(f5_rec(rec + 12), cp1), b2 = f5_c(f5_rec(rec + 16), cp2);
		const float p456 = a4.v * a5.v * a6.v, pb = b1.v * b2.v;
		const float e = f1.v * p456 * pb;
		if(e != 0.f) {
			E += e;
			if(esplit) esplit[2] += e;
			v3 f = h * (-(f1.d * p456 * pb));
			const float fb = f1.v * pb;
			chain_bb(acc, fb * a4.dc * a5.v * a6.v, t4);
			f += chain_bd<false>(acc, fb * a4.v * a5.dc * a6.v, -A.a3, h, inv, t5);
			f += chain_bd<true>(acc, fb * a4.v * a5.v * a6.dc, -B.a3, h, inv, t6);
			// phi1, phi2 as the reference writes them (see the head of this file): functions of rstack with the lever gamma
			const float g = M.gamma, wm = w2 * winv, icub = winv * winv * winv;
			v3 tp = mk3(0.f, 0.f, 0.f), tq = mk3(0.f, 0.f, 0.f);
			{
				const float ra2 = dot(h, A.a2), ra1 = dot(h, A.a1), rb1 = dot(h, B.a1), a2b1 = dot(A.a2, B.a1);
				const float par = m * ra2 - a2b1 * g;
				const float ddr = (m * m * ra2 - ra2 * wm * wm - m * (a2b1 + ra2 * (rb1 - ra1)) * g + a2b1 * (rb1 - ra1) * g * g) * icub;
				const float dra1 = m * g * par * icub, dra2 = -m * winv, drb1 = -dra1;
				const float da1b1 = -g * g * par * icub, da2b1 = g * winv;
				const float fp = -f1.v * p456 * b1.dc * b2.v;
				f -= (h * ddr + ((A.a2 - h * ra2) * dra2 + (A.a1 - h * ra1) * dra1 + (B.a1 - h * rb1) * drb1) * inv) * fp;
				tp += cross(h, A.a2) * (fp * dra2) + cross(h, A.a1) * (fp * dra1);
				tq += cross(h, B.a1) * (fp * drb1);
				const v3 pure = cross(A.a2, B.a1) * (fp * da2b1) + cross(A.a1, B.a1) * (fp * da1b1);
				tp -= pure; tq += pure;
			}
			{
				const float ra2 = dot(h, B.a2), ra1 = dot(h, B.a1), rb1 = dot(h, A.a1), a2b1 = dot(B.a2, A.a1);
				const float par = m * ra2 + a2b1 * g;
				const float ddr = (par * (m + (rb1 - ra1) * g) - ra2 * wm * wm) * icub;
				const float dra1 = -m * g * par * icub, dra2 = -m * winv, drb1 = m * g * par * icub;
				const float da1b1 = -g * g * par * icub, da2b1 = -g * winv;
				const float fp = -f1.v * p456 * b1.v * b2.dc;
				f -= (h * ddr + ((B.a2 - h * ra2) * dra2 + (B.a1 - h * ra1) * dra1 + (A.a1 - h * rb1) * drb1) * inv) * fp;
				tp += cross(h, A.a1) * (fp * drb1);
				tq += cross(h, B.a2) * (fp * dra2) + cross(h, B.a1) * (fp * dra1);
				const v3 pure = cross(A.a1, B.a2) * (fp * da2b1) + cross(A.a1, B.a1) * (fp * da1b1);
				tp -= pure; tq += pure;
			}
			acc.Tp += tp; acc.Tq += tq;
			acc.site_aa(f, csp, csq);
		}
	}
	return E;
}
