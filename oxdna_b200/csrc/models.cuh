// Compile-time model traits: the force kernels (forces.cu) are written once and instantiated for oxDNA2 and oxRNA2.
#pragma once

#include "rna_model.cuh"

struct DnaModel {
	typedef oxb_dna2_params Params;
	static OXB_HD v3 back(const Params &M, const Axes &A) { return A.a1 * M.back_a1 + A.a2 * M.back_a2; }
	static OXB_HD float back_a3(const Params &M) { return 0.f; }
	static OXB_HD bool hb_in_range(const Params &M, float rbm2, int btp, int btq) { return dna2_hb_in_range(M, rbm2, btp, btq); }
	static OXB_HD bool crst_in_range(const Params &M, float rbm2) { return dna2_crst_in_range(M, rbm2); }
	static OXB_HD bool cxst_in_range(const Params &M, float rs2) { return dna2_cxst_in_range(M, rs2); }
	static OXB_HD bool hbcr_may_act(const Params &M, v3 h, const Axes &A, const Axes &B, bool hb_on, bool cr_on) { return dna2_hbcr_may_act(M, h, A, B, hb_on, cr_on); }
	static OXB_HD bool cxst_may_act(const Params &M, v3 h, const Axes &A, const Axes &B) { return dna2_cxst_may_act(M, h, A, B); }
	template<bool WITH_HB>
	static OXB_HD float hbcr(const Params &M, v3 rb, float rbm2, const Axes &A, const Axes &B, int btp, int btq, bool hb_on, bool cr_on, PairAcc &acc, float &ehb) {
		return dna2_hbcr<WITH_HB>(M, rb, rbm2, A, B, btp, btq, hb_on, cr_on, acc, ehb);
	}
	static OXB_HD float cxst(const Params &M, v3 rs, float rs2, v3, const Axes &A, const Axes &B, PairAcc &acc) { return dna2_cxst(M, rs, rs2, A, B, acc); }
	static OXB_HD float bonded(const Params &M, v3 r, const Axes &A, const Axes &B, int btp, int btq, v3 pback, v3 qback, PairAcc &acc, bool &broken, float *esplit = nullptr, const FeneSite *fene = nullptr) {
		return dna2_bonded(M, r, A, B, btp, btq, pback, qback, acc, broken, esplit, fene);
	}
	static OXB_HD PairEnergy nonbonded(const Params &M, v3 r, const Axes &A, const Axes &B, int btp, int btq, bool p_end, bool q_end, v3 pback, v3 qback, PairAcc &acc) {
		return dna2_nonbonded(M, r, A, B, btp, btq, p_end, q_end, pback, qback, acc);
	}
};

struct RnaModel {
	typedef oxb_rna2_params Params;
	static OXB_HD v3 back(const Params &M, const Axes &A) { return rna2_back(M, A); }
	static OXB_HD float back_a3(const Params &M) { return M.back_a3; }
	static OXB_HD bool hb_in_range(const Params &M, float rbm2, int btp, int btq) { return rna2_hb_in_range(M, rbm2, btp, btq); }
	static OXB_HD bool crst_in_range(const Params &M, float rbm2) { return rna2_crst_in_range(M, rbm2); }
	static OXB_HD bool cxst_in_range(const Params &M, float rs2) { return rna2_cxst_in_range(M, rs2); }
	static OXB_HD bool hbcr_may_act(const Params &M, v3 h, const Axes &A, const Axes &B, bool hb_on, bool cr_on) { return rna2_hbcr_may_act(M, h, A, B, hb_on, cr_on); }
	static OXB_HD bool cxst_may_act(const Params &M, v3 h, const Axes &A, const Axes &B) { return rna2_cxst_may_act(M, h, A, B); }
	template<bool WITH_HB>
	static OXB_HD float hbcr(const Params &M, v3 rb, float rbm2, const Axes &A, const Axes &B, int btp, int btq, bool hb_on, bool cr_on, PairAcc &acc, float &ehb) {
		return rna2_hbcr<WITH_HB>(M, rb, rbm2, A, B, btp, btq, hb_on, cr_on, acc, ehb);
	}
	static OXB_HD float cxst(const Params &M, v3 rs, float rs2, v3 rbk, const Axes &A, const Axes &B, PairAcc &acc) { return rna2_cxst(M, rs, rs2, rbk, A, B, acc); }
	static OXB_HD float bonded(const Params &M, v3 r, const Axes &A, const Axes &B, int btp, int btq, v3 pback, v3 qback, PairAcc &acc, bool &broken, float *esplit = nullptr, const FeneSite *fene = nullptr) {
		return rna2_bonded(M, r, A, B, btp, btq, pback, qback, acc, broken, esplit, fene);
	}
	static OXB_HD PairEnergy nonbonded(const Params &M, v3 r, const Axes &A, const Axes &B, int btp, int btq, bool p_end, bool q_end, v3 pback, v3 qback, PairAcc &acc) {
		return rna2_nonbonded(M, r, A, B, btp, btq, p_end, q_end, pback, qback, acc);
	}
};
