// oxRNA2 pair potential, FP32 device functions (sm_100a).  Same evaluation scheme as dna_model.cuh (derivatives with
// respect to the cosine, generic chain-rule moves, theta = atan2(|u x v|, u.v), lever arms folded once per pair).
//
// What it evaluates: src/CUDA/Interactions/CUDA_RNA.cuh:358-967 (CPU mirror src/Interactions/RNAInteraction.cpp:431-1160 and
// RNAInteraction2.cpp:121-322).  Where the CPU class and the CUDA kernels of the reference disagree, this follows the CUDA
// kernels, whose force is the gradient of the energy: full product in the phi2 stacking term (CUDA_RNA.cuh:626 against
// RNAInteraction.cpp:620) and the sign of the mirrored coaxial theta1 term (CUDA_RNA.cuh:896 against RNAInteraction.cpp:1046).
// Superset: the sequence-dependent cross-stacking multiplier of the CPU class (RNAInteraction.cpp:901-905), which the
// reference's CUDA kernels ignore, is applied (it is 1 with the shipped parameter file).
#pragma once

#include "dna_model.cuh"

OXB_HD v3 rna2_back(const oxb_rna2_params &M, const Axes &A) { return A.a1 * M.back_a1 + A.a2 * M.back_a2 + A.a3 * M.back_a3; }

// f4(pi - theta) as a function of cos(theta)
OXB_HD AngVal f4_ts_mirror(const oxb_f4 &f, float t, float s) {
	AngVal n = f4_ts(f, OXB_PI_F - t, s);
	n.dc = -n.dc;
	return n;
}

// 0: no hydrogen-bonding-type term, 1: Watson-Crick or (sequence-dependent model) G-U wobble pair, 2: mismatch repulsion
OXB_HD int rna2_hb_kind(const oxb_rna2_params &M, int btp, int btq) {
	bool pair = (btp + btq == 3);
	if(!M.average && btp + btq == 4) {
		int tp = btype_to_type(btp), tq = btype_to_type(btq);
		if((tq == 3 && tp == 1) || (tq == 1 && tp == 3)) pair = true;
	}
	return pair ? 1 : (M.mismatch_repulsion ? 2 : 0);
}

OXB_HD bool rna2_hb_in_range(const oxb_rna2_params &M, float rbm2, int btp, int btq) {
	int k = rna2_hb_kind(M, btp, btq);
	if(k == 0 || rbm2 >= M.hb.rchigh * M.hb.rchigh) return false;
	return k == 2 || rbm2 > M.hb.rclow * M.hb.rclow;
}
OXB_HD bool rna2_crst_in_range(const oxb_rna2_params &M, float rbm2) {
	return rbm2 > M.crst.rclow * M.crst.rclow && rbm2 < M.crst.rchigh * M.crst.rchigh;
}
OXB_HD bool rna2_cxst_in_range(const oxb_rna2_params &M, float rs2) {
	return rs2 > M.cxst.rclow * M.cxst.rclow && rs2 < M.cxst.rchigh * M.cxst.rchigh;
}

OXB_HD bool rna2_hbcr_may_act(const oxb_rna2_params &M, v3 h, const Axes &A, const Axes &B, bool hb_on, bool cr_on) {
	float c1 = -dot(A.a1, B.a1), c2 = -dot(B.a1, h), c3 = dot(A.a1, h), c4 = dot(A.a3, B.a3), c7 = -dot(B.a3, h), c8 = dot(A.a3, h);
	bool hb = hb_on && in_window(M, OXB_RF4_HB_T1, c1) && in_window(M, OXB_RF4_HB_T2, c2) && in_window(M, OXB_RF4_HB_T3, c3) &&
			in_window(M, OXB_RF4_HB_T4, c4) && in_window(M, OXB_RF4_HB_T7, c7) && in_window(M, OXB_RF4_HB_T8, c8);
	bool cr = cr_on && in_window(M, OXB_RF4_CRST_T1, c1) && in_window(M, OXB_RF4_CRST_T2, c2) && in_window(M, OXB_RF4_CRST_T3, c3) &&
			in_window_sym(M, OXB_RF4_CRST_T7, c7) && in_window_sym(M, OXB_RF4_CRST_T8, c8);
	return hb || cr;
}

OXB_HD bool rna2_cxst_may_act(const oxb_rna2_params &M, v3 h, const Axes &A, const Axes &B) {
	// the window of f4(theta1) contains the support of its mirror f4(2 pi - theta1) whenever t0 + tc >= pi
	float c1 = -dot(A.a1, B.a1), c4 = dot(A.a3, B.a3), c5 = dot(A.a3, h), c6 = -dot(B.a3, h);
	return in_window(M, OXB_RF4_CXST_T1, c1) && in_window(M, OXB_RF4_CXST_T4, c4) && in_window_sym(M, OXB_RF4_CXST_T5, c5) &&
			in_window_sym(M, OXB_RF4_CXST_T6, c6);
}

// repulsive radial part of the mismatch potential (RNAInteraction2.cpp:167-201)
OXB_HD RadVal rna2_fX(const oxb_rna2_params &M, float r) {
	const oxb_f1 &f = M.hb;
	RadVal o;
	o.v = 0.f;
	o.d = 0.f;
	if(r < f.rchigh) {
		if(r > f.rhigh) {
			float x = r - f.rchigh;
			o.v = M.mis_eps * f.bhigh * x * x;
			o.d = 2.f * M.mis_eps * f.bhigh * x;
		}
		else if(r > f.r0) {
			float e = OXB_EXP(-(r - f.r0) * f.a);
			float t = 1.f - e;
			o.v = M.mis_eps * t * t - M.mis_shift;
			o.d = 2.f * M.mis_eps * t * e * f.a;
		}
		else o.v = -M.mis_shift;
	}
	return o;
}

// hydrogen bonding (or mismatch repulsion) + cross stacking on the base-base vector rb
template<bool WITH_HB = true>
OXB_HD float rna2_hbcr(const oxb_rna2_params &M, v3 rb, float rbm2, const Axes &A, const Axes &B, int btp, int btq, bool hb_on, bool cr_on,
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
	Angle t7 = make_angle(-B.a3, h);
	Angle t8 = make_angle(A.a3, h);
	float g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f, g7 = 0.f, g8 = 0.f, grad = 0.f;
	Angle t4;
	t4.c = t4.s = t4.t = 0.f;
	t4.x = mk3(0.f, 0.f, 0.f);
	if(WITH_HB && hb_on) {
		t4 = make_angle(A.a3, B.a3);
		RadVal f1;
		if(rna2_hb_kind(M, btp, btq) == 2) f1 = rna2_fX(M, m);
		else {
			int ti = btype_to_type(btq) * 5 + btype_to_type(btp);
			float mult = (abs(btq) >= 300 && abs(btp) >= 300) ? M.hb_multiplier : 1.f;
			f1 = f1_r(M.hb, M.hb_eps[ti], M.hb_shift[ti], m);
			f1.v *= mult;
			f1.d *= mult;
		}
		AngVal a1 = f4_ts(M.f4[OXB_RF4_HB_T1], t1.t, t1.s);
		AngVal a2 = f4_ts(M.f4[OXB_RF4_HB_T2], t2.t, t2.s);
		AngVal a3 = f4_ts(M.f4[OXB_RF4_HB_T3], t3.t, t3.s);
		AngVal a4 = f4_ts(M.f4[OXB_RF4_HB_T4], t4.t, t4.s);
		AngVal a7 = f4_ts(M.f4[OXB_RF4_HB_T7], t7.t, t7.s);
		AngVal a8 = f4_ts(M.f4[OXB_RF4_HB_T8], t8.t, t8.s);
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
		float kf = M.average ? 1.f : M.crst_kfac[btype_to_type(btp) * 5 + btype_to_type(btq)];
		f2.v *= kf;
		f2.d *= kf;
		AngVal a1 = f4_ts(M.f4[OXB_RF4_CRST_T1], t1.t, t1.s);
		AngVal a2 = f4_ts(M.f4[OXB_RF4_CRST_T2], t2.t, t2.s);
		AngVal a3 = f4_ts(M.f4[OXB_RF4_CRST_T3], t3.t, t3.s);
		AngVal a7 = f4_ts_sym(M.f4[OXB_RF4_CRST_T7], t7.t, t7.s);
		AngVal a8 = f4_ts_sym(M.f4[OXB_RF4_CRST_T8], t8.t, t8.s);
		float p12 = a1.v * a2.v, p78 = a7.v * a8.v;
		float e = f2.v * p12 * a3.v * p78;
		if(e != 0.f) {
			E += e;
			grad += f2.d * p12 * a3.v * p78;
			float f3_78 = f2.v * a3.v * p78;
			g1 += f3_78 * a1.dc * a2.v;
			g2 += f3_78 * a1.v * a2.dc;
			g3 += f2.v * p12 * p78 * a3.dc;
			g7 += f2.v * p12 * a3.v * a7.dc * a8.v;
			g8 += f2.v * p12 * a3.v * a7.v * a8.dc;
		}
	}
	if(E != 0.f) {
		v3 f = h * (-grad);
		chain_bb(acc, g1, t1);
		f += chain_bd<true>(acc, g2, -B.a1, h, inv, t2);
		f += chain_bd<false>(acc, g3, A.a1, h, inv, t3);
		if(WITH_HB) chain_bb(acc, g4, t4);
		f += chain_bd<true>(acc, g7, -B.a3, h, inv, t7);
		f += chain_bd<false>(acc, g8, A.a3, h, inv, t8);
		acc.site_aa(f, cb, cb);
	}
	return E;
}

// coaxial stacking on the stack-stack vector rs; rbk = backbone-backbone vector (for phi3 / phi4)
OXB_HD float rna2_cxst(const oxb_rna2_params &M, v3 rs, float rs2, v3 rbk, const Axes &A, const Axes &B, PairAcc &acc) {
	const float cs = M.stack_a1;
	float inv = OXB_RSQRT(rs2);
	float m = rs2 * inv;
	v3 h = rs * inv;
	RadVal f2 = f2_r(M.cxst, m);
	Angle t1 = make_angle(-A.a1, B.a1);
	Angle t4 = make_angle(A.a3, B.a3);
	Angle t5 = make_angle(A.a3, h);
	Angle t6 = make_angle(-B.a3, h);
	AngVal a1 = f4_ts_rna_cxst_t1(M.f4[OXB_RF4_CXST_T1], t1.t, t1.s);
	AngVal a4 = f4_ts(M.f4[OXB_RF4_CXST_T4], t4.t, t4.s);
	AngVal a5 = f4_ts_sym(M.f4[OXB_RF4_CXST_T5], t5.t, t5.s);
	AngVal a6 = f4_ts_sym(M.f4[OXB_RF4_CXST_T6], t6.t, t6.s);
	float p14 = a1.v * a4.v, p56 = a5.v * a6.v;
	float e0 = f2.v * p14 * p56;
	if(e0 == 0.f) return 0.f;
	float binv = OXB_RSQRT(dot(rbk, rbk));
	v3 bh = rbk * binv;
	AngVal b3 = f5_c(M.phi3, dot(h, cross(bh, A.a1)));
	AngVal b4 = f5_c(M.phi4, dot(h, cross(bh, B.a1)));
	float pb = b3.v * b4.v;
	float e = e0 * pb;
	if(e != 0.f) {
		v3 f = h * (-(f2.d * p14 * p56 * pb));
		float fb = f2.v * pb;
		chain_bb(acc, fb * p56 * a1.dc * a4.v, t1);
		chain_bb(acc, fb * p56 * a1.v * a4.dc, t4);
		f += chain_bd<false>(acc, fb * p14 * a5.dc * a6.v, A.a3, h, inv, t5);
		f += chain_bd<true>(acc, fb * p14 * a5.v * a6.dc, -B.a3, h, inv, t6);
		acc.site_aa(f, cs, cs);
		if(b3.dc != 0.f) chain_triple<false, false>(acc, e0 * b3.dc * b4.v, A.a1, h, inv, cs, bh, binv, 0.f);
		if(b4.dc != 0.f) chain_triple<true, false>(acc, e0 * b3.v * b4.dc, B.a1, h, inv, cs, bh, binv, 0.f);
	}
	return e;
}

// the whole non-bonded interaction of one pair (particle-centric kernel, host-side unit test)
OXB_HD PairEnergy rna2_nonbonded(const oxb_rna2_params &M, v3 r, const Axes &A, const Axes &B, int btp, int btq, bool p_end, bool q_end,
		v3 pback, v3 qback, PairAcc &acc) {
	PairEnergy E;
	E.total = 0.f;
	E.hb = 0.f;
	float r2 = dot(r, r);
	if(r2 >= (acc.rep ? acc.rep->rcut2 : M.rcut * M.rcut)) return E; // RNAInteraction2.cpp:20-22
	v3 rbb = r + qback - pback;
	float fs;
	float en = acc.rep ? dna2_dh(dh_view(M, acc.rep), dot(rbb, rbb), p_end, q_end, fs) : dna2_dh(M, dot(rbb, rbb), p_end, q_end, fs);
	if(en != 0.f) { E.total += en; acc.site_kk(rbb * fs); }
	if(r2 >= M.rcut_near * M.rcut_near) return E;
	v3 rb = r + (B.a1 - A.a1) * M.base_a1;
	E.total += dna2_excl(M, r, rbb, rb, A, B, pback, qback, acc);
	float rbm2 = dot(rb, rb);
	bool hb_on = rna2_hb_in_range(M, rbm2, btp, btq), cr_on = rna2_crst_in_range(M, rbm2);
	if(hb_on || cr_on) {
		float ehb;
		E.total += rna2_hbcr(M, rb, rbm2, A, B, btp, btq, hb_on, cr_on, acc, ehb);
		E.hb += ehb;
	}
	v3 rs = r + (B.a1 - A.a1) * M.stack_a1;
	float rs2 = dot(rs, rs);
	if(rna2_cxst_in_range(M, rs2)) E.total += rna2_cxst(M, rs, rs2, rbb, A, B, acc);
	return E;
}

// bonded pair p -> q = n3(p): FENE, bonded excluded volume, stacking between STACK_3(p) and STACK_5(q)
// (RNAInteraction.cpp:431-667)
OXB_HD float rna2_bonded(const oxb_rna2_params &M, v3 r, const Axes &A, const Axes &B, int btp, int btq, v3 pback, v3 qback, PairAcc &acc,
		bool &broken, float *esplit = nullptr, const FeneSite *fene = nullptr) {
	float E = bonded_fene_excl(M, r, A, B, pback, qback, acc, broken, esplit, fene);
	v3 sp = A.a1 * M.stack3_a1 + A.a2 * M.stack3_a2, sq = B.a1 * M.stack5_a1 + B.a2 * M.stack5_a2;
	v3 rs = r + sq - sp;
	float rs2 = dot(rs, rs);
	float inv = OXB_RSQRT(rs2);
	float m = rs2 * inv;
	int ti = btype_to_type(btq) * 5 + btype_to_type(btp);
	RadVal f1 = acc.rep ? f1_r(M.stck, acc.rep->stck_eps[ti], acc.rep->stck_shift[ti], m) : f1_r(M.stck, M.stck_eps[ti], M.stck_shift[ti], m);
	if(f1.v != 0.f || f1.d != 0.f) {
		v3 h = rs * inv;
		v3 rbk = r + qback - pback;
		float binv = OXB_RSQRT(dot(rbk, rbk));
		v3 bh = rbk * binv;
		v3 bb3 = A.a1 * M.p3[0] + A.a2 * M.p3[1] + A.a3 * M.p3[2];
		v3 bb5 = B.a1 * M.p5[0] + B.a2 * M.p5[1] + B.a3 * M.p5[2];
		Angle t5 = make_angle(A.a3, h);
		Angle t6 = make_angle(-B.a3, h);
		Angle tb1 = make_angle(-bb3, bh);
		Angle tb2 = make_angle(-bb5, bh);
		Angle p1 = make_angle_cos(A.a2, bh);
		Angle p2 = make_angle_cos(B.a2, bh);
		AngVal a5 = f4_ts_mirror(M.f4[OXB_RF4_STCK_T5], t5.t, t5.s);
		AngVal a6 = f4_ts(M.f4[OXB_RF4_STCK_T6], t6.t, t6.s);
		AngVal ab1 = f4_ts(M.f4[OXB_RF4_STCK_TB1], tb1.t, tb1.s);
		AngVal ab2 = f4_ts(M.f4[OXB_RF4_STCK_TB2], tb2.t, tb2.s);
		AngVal b1 = f5_c(M.phi1, p1.c);
		AngVal b2 = f5_c(M.phi2, p2.c);
		float p56 = a5.v * a6.v, pbb = ab1.v * ab2.v, pph = b1.v * b2.v;
		float e = f1.v * p56 * pbb * pph;
		if(e != 0.f) {
			E += e;
			if(esplit) esplit[2] += e;
			v3 f = h * (-(f1.d * p56 * pbb * pph));
			float fa = f1.v * pbb * pph;
			f += chain_bd<false>(acc, fa * a5.dc * a6.v, A.a3, h, inv, t5);
			f += chain_bd<true>(acc, fa * a5.v * a6.dc, -B.a3, h, inv, t6);
			acc.site_gg(f, sp, sq);
			float fb = f1.v * p56 * pph;
			v3 fk = chain_bd<false>(acc, fb * ab1.dc * ab2.v, -bb3, bh, binv, tb1);
			fk += chain_bd<true>(acc, fb * ab1.v * ab2.dc, -bb5, bh, binv, tb2);
			float fc = f1.v * p56 * pbb;
			fk += chain_bd<false>(acc, fc * b1.dc * b2.v, A.a2, bh, binv, p1);
			fk += chain_bd<true>(acc, fc * b1.v * b2.dc, B.a2, bh, binv, p2);
			acc.site_kk(fk);
		}
	}
	return E;
}
