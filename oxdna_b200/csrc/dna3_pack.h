// Host side of oxDNA3: repacks the 215 tetramer-indexed tables of the reference (include/oxdna_b200.h: oxb_set_model_dna3) into the per-term
// records of dna3_model.cuh and fills the scalar part of the kernel argument.  Used by the context (context.cu) and by the host-compiled
// unit test of the FP32 formulation (tests/support/host_model.cu).
#pragma once

#include "kernels.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

// h: the packed records (floats; 16-byte units); off[k]: offset in float4 of bonded, crst, cxst, hb, nexcl inside h.  The pointers of D are
// left null: the caller points them into its copy of h.
inline void dna3_pack(const double *tab, const oxb_dna3_scalars *S, std::vector<float> &h, oxb_dna3_dev &D, size_t off[5]) {
	auto T = [&](int id, int ix) { return tab[(size_t) id * OXB_DNA3_TSIZE + ix]; };
	const int NB = 900 * (OXB3_REC_BONDED / 4), NC = 2 * 900 * (OXB3_REC_CRST / 4), NX = 900 * (OXB3_REC_CXST / 4), NH = 25 * (OXB3_REC_HB / 4), NE = 25 * (OXB3_REC_NEXCL / 4);
	h.assign((size_t) 4 * (NB + NC + NX + NH + NE), 0.f);
	float *hb_ = h.data(), *hc = hb_ + 4 * NB, *hx = hc + 4 * NC, *hh = hx + 4 * NX, *he = hh + 4 * NH;
	auto put_excl = [&](float *o, int which, int ix) {
		const double s = T(OXB_DNA3_EXCL_S + which, ix), r = T(OXB_DNA3_EXCL_R + which, ix);
		o[0] = (float) (s * s); o[1] = (float) (r * r); o[2] = (float) T(OXB_DNA3_EXCL_B + which, ix); o[3] = (float) T(OXB_DNA3_EXCL_RC + which, ix);
	};
	// f1 record: a, rc, r0, blow | bhigh, rlow, rhigh, rclow | rchigh, eps, shift, -   (table order EPS, A, RC, R0, BLOW, BHIGH, RLOW, RHIGH, RCLOW, RCHIGH, SHIFT)
	auto put_f1 = [&](float *o, int ty, int ix) {
		auto F = [&](int par) { return (float) T(OXB_DNA3_F1 + par * 2 + ty, ix); };
		o[0] = F(1); o[1] = F(2); o[2] = F(3); o[3] = F(4); o[4] = F(5); o[5] = F(6); o[6] = F(7); o[7] = F(8); o[8] = F(9); o[9] = F(0); o[10] = F(10); o[11] = 0.f;
	};
	// f2 record: k, rc, r0, blow | rlow, rclow, bhigh, rhigh | rchigh, k_symm, -, -   (table order K, K_SYMM, RC, R0, BLOW, RLOW, RCLOW, BHIGH, RCHIGH, RHIGH)
	auto put_f2 = [&](float *o, int ty, int ix) {
		auto F = [&](int par) { return (float) T(OXB_DNA3_F2 + par * 4 + ty, ix); };
		o[0] = F(0); o[1] = F(2); o[2] = F(3); o[3] = F(4); o[4] = F(5); o[5] = F(6); o[6] = F(7); o[7] = F(9); o[8] = F(8); o[9] = F(1); o[10] = o[11] = 0.f;
	};
	auto put_f4 = [&](float *o, int ty, int ix) { for(int par = 0; par < 5; par++) o[par] = (float) T(OXB_DNA3_F4 + par * 21 + ty, ix); };
	auto put_f5 = [&](float *o, int ty, int ix) { for(int par = 0; par < 4; par++) o[par] = (float) T(OXB_DNA3_F5 + par * 4 + ty, ix); };
	const bool use_mbf = S->use_mbf != 0.;
	double max_excl_rc = 0., max_base = 0., max_stack = 0., max_bb = 0., max_eb = 0., max_bk = 0.;
	for(int ix = 0; ix < 900; ix++) {
		float *o = hb_ + (size_t) ix * OXB3_REC_BONDED;
		const double xmax = T(OXB_DNA3_MBF_XMAX, ix), d2 = T(OXB_DNA3_FENE_DELTA2, ix);
		o[0] = (float) T(OXB_DNA3_FENE_R0, ix); o[1] = (float) d2; o[2] = (float) xmax;
		// constant of the far branch of max_backbone_force (DNA3Interaction.cpp:1252-1256): fene(xmax) - long(xmax)
		o[3] = (use_mbf && xmax > 0.) ? (float) (-(S->fene_eps / 2.) * std::log(1. - xmax * xmax / d2) - ((S->mbf_fmax - S->mbf_finf) * xmax * std::log(xmax) + S->mbf_finf * xmax)) : 0.f;
		put_excl(o + 4, 4, ix); put_excl(o + 8, 5, ix); put_excl(o + 12, 6, ix);
		put_f1(o + 16, 1, ix); // STCK_F1
		put_f4(o + 28, 0, ix); put_f4(o + 33, 1, ix); // STCK_F4_THETA4, THETA5 = THETA6
		put_f5(o + 40, 0, ix); put_f5(o + 44, 1, ix); // STCK_F5_PHI1, PHI2
		for(int br = 0; br < 2; br++) {
			float *q = hc + ((size_t) br * 900 + ix) * OXB3_REC_CRST;
			put_f2(q, 2 + br, ix); // CRST_F2_33, CRST_F2_55
			const int t0 = 13 + 4 * br; // CRST_F4_THETA1_33 = 13, _55 = 17: theta1, theta2 = theta3, theta4, theta7 = theta8
			put_f4(q + 12, t0, ix); put_f4(q + 17, t0 + 1, ix); put_f4(q + 22, t0 + 2, ix); put_f4(q + 27, t0 + 3, ix);
			max_base = std::max(max_base, T(OXB_DNA3_F2 + 8 * 4 + 2 + br, ix));
		}
		put_f2(hx + (size_t) ix * OXB3_REC_CXST, 1, ix); // CXST_F2
		max_stack = std::max(max_stack, T(OXB_DNA3_F2 + 8 * 4 + 1, ix));
		for(int w = 0; w < 7; w++) max_excl_rc = std::max(max_excl_rc, T(OXB_DNA3_EXCL_RC + w, ix));
	}
	for(int tq = 0; tq < 5; tq++) for(int tp = 0; tp < 5; tp++) {
		const int ix0 = ((0 * 5 + tq) * 5 + tp) * 6 + 0, ix5 = ((5 * 5 + tq) * 5 + tp) * 6 + 5;
		float *o = hh + (size_t) (tq * 5 + tp) * OXB3_REC_HB;
		put_f1(o, 0, ix0); // HYDR_F1; HYDR_F4_THETA1 = 2, THETA2 = THETA3 = 3, THETA4 = 4, THETA7 = THETA8 = 5
		put_f4(o + 12, 2, ix0); put_f4(o + 17, 3, ix0); put_f4(o + 22, 4, ix0); put_f4(o + 27, 5, ix0);
		max_base = std::max(max_base, T(OXB_DNA3_F1 + 9 * 2 + 0, ix0));
		float *e = he + (size_t) (tq * 5 + tp) * OXB3_REC_NEXCL;
		for(int w = 0; w < 4; w++) put_excl(e + 4 * w, w, ix5);
		max_bb = std::max(max_bb, T(OXB_DNA3_EXCL_RC + 0, ix5)); max_eb = std::max(max_eb, T(OXB_DNA3_EXCL_RC + 1, ix5));
		max_bk = std::max(max_bk, std::max(T(OXB_DNA3_EXCL_RC + 2, ix5), T(OXB_DNA3_EXCL_RC + 3, ix5)));
	}
	std::memset(&D, 0, sizeof(D));
	off[0] = 0; off[1] = NB; off[2] = (size_t) NB + NC; off[3] = (size_t) NB + NC + NX; off[4] = (size_t) NB + NC + NX + NH;
	D.fene_eps = (float) S->fene_eps; D.use_mbf = use_mbf ? 1 : 0; D.mbf_fmax = (float) S->mbf_fmax; D.mbf_finf = (float) S->mbf_finf;
	D.hb_multiplier = (float) S->hb_multiplier; D.excl_eps = 2.0f; // EXCL_EPS, src/model.h
	D.dh_minus_kappa = (float) S->dh_minus_kappa; D.dh_prefactor = (float) S->dh_prefactor; D.dh_rhigh = (float) S->dh_rhigh; D.dh_rc = (float) S->dh_rc;
	D.dh_b = (float) S->dh_b; D.dh_half_charged_ends = S->dh_half_charged_ends != 0. ? 1 : 0;
	D.rcut2 = (float) (S->rcut * S->rcut);
	auto f4 = [](const double *v) { oxb_f4 f; f.a = (float) v[0]; f.b = (float) v[1]; f.t0 = (float) v[2]; f.ts = (float) v[3]; f.tc = (float) v[4]; return f; };
	D.cxst_t1 = f4(S->cxst_t1); D.cxst_t4 = f4(S->cxst_t4); D.cxst_t5 = f4(S->cxst_t5);
	D.cxst_t1_sa = (float) S->cxst_t1_sa; D.cxst_t1_sb = (float) S->cxst_t1_sb;
	// site offsets, src/model.h:15-39, DNANucleotide.cpp:14-41 (index 0: dummy base, 1 + type otherwise; the backbone site is the same for all)
	D.back_a1 = -0.3400f; D.back_a2 = 0.3408f; D.backref_a1 = -0.4f; D.gamma = 0.34f + 0.4f; // POS_MM_BACK1 / 2, POS_BACK, POS_STACK - POS_BACK
	const float st[5] = { 0.34f, 0.37f, 0.37f, 0.37f, 0.37f }, ba[5] = { 0.4f, 0.43f, 0.43f, 0.37f, 0.37f };
	for(int i = 0; i < 5; i++) { D.pos_stack[i] = st[i]; D.pos_base[i] = ba[i]; }
	const double lever = std::max(std::sqrt(0.34 * 0.34 + 0.3408 * 0.3408), 0.43);
	D.r2_excl_max = (float) std::pow(max_excl_rc + 2. * lever + 0.01, 2);
	D.r2_base_max = (float) std::pow(max_base + 1e-3, 2);
	D.r2_stack_max = (float) std::pow(max_stack + 1e-3, 2);
	D.range_bb = (float) max_bb; D.range_eb = (float) max_eb; D.range_bk = (float) max_bk;
	D.r2_near_max = (float) std::pow(std::max(max_excl_rc + 2. * lever, std::max(max_base + 2. * 0.43, max_stack + 2. * 0.37)) + 0.01, 2);
}
