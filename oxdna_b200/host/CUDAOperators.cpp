// Implementation of the host-side operator classes (see CUDAOperators.h).  Nothing here computes physics on the CPU:
// settings are parsed with the reference's own input helpers, constants come from the reference's CPU classes, and all
// work is forwarded to liboxdna_b200.so.
#include "CUDAOperators.h"

#include "PluginManagement/PluginManager.h"

#include "Utilities/ConfigInfo.h"
#include "Utilities/Logger.h"
#include "Utilities/Utils.h"
#include "model.h"

#include <cmath>
#include <cstring>

void oxb_check(oxb_ctx *ctx, int rc, const char *what) {
	if(rc != 0) throw oxDNAException("%s: %s (oxdna_b200 status %d)", what, ctx ? oxb_last_error(ctx) : "no context", rc);
}

// ---------------------------------------------------------------------------------------------------- interactions
void CUDABaseInteraction::get_cuda_settings(input_file &inp) {
	// src/CUDA/Interactions/CUDABaseInteraction.cu:79-98
	int tmpi;
	if(getInputBoolAsInt(&inp, "use_edge", &tmpi, 0) == KEY_FOUND) {
		if(tmpi > 0) {
			_use_edge = true;
			getInputInt(&inp, "edge_n_forces", &_n_forces, 0);
			if(_n_forces < 1) throw oxDNAException("edge_n_forces must be > 0");
		}
	}
}

void CUDABaseInteraction::cuda_init(oxb_ctx *ctx, int N) {
	_ctx = ctx;
	_N = N;
}

void CUDABaseInteraction::compute_forces(oxb_ctx *ctx) {
	oxb_check(ctx, oxb_compute_forces(ctx), "compute_forces");
}

void CUDABaseInteraction::compute_forces_views(const oxb_force_views &) {
	throw oxDNAException("this CUDA interaction does not implement compute_forces_views()");
}

namespace {
// C trampoline of the plugin seam: exceptions must not unwind through the C frames of liboxdna_b200.so
int plugin_force_pass(void *user, const oxb_force_views *views) {
	try {
		static_cast<CUDABaseInteraction *>(user)->compute_forces_views(*views);
	}
	catch(oxDNAException &e) {
		OX_LOG(Logger::LOG_ERROR, "plugged-in CUDA interaction: %s", e.what());
		return 1;
	}
	return 0;
}
} // namespace

void CUDABaseInteraction::attach_as_plugin(oxb_ctx *ctx) {
	if(_use_edge) throw oxDNAException("use_edge is not available with a plugged-in CUDA interaction");
	oxb_check(ctx, oxb_set_force_callback(ctx, plugin_force_pass, this, (double) get_cuda_rcut()), "set_force_callback");
}

CUDADNAInteraction::CUDADNAInteraction() {}

CUDADNAInteraction::~CUDADNAInteraction() {}

void CUDADNAInteraction::get_settings(input_file &inp) {
	std::string inter_type("DNA");
	getInputString(&inp, "interaction_type", inter_type, 0);
	if(inter_type != "DNA2") {
		throw oxDNAException("interaction_type = %s is not available in the oxdna_b200 CUDA backend (DNA2 only in this build)", inter_type.c_str());
	}
	// every key of the CPU class: salt_concentration, dh_lambda, dh_strength, dh_half_charged_ends, use_average_seq,
	// seq_dep_file, hb_multiplier, max_backbone_force[_far], major-minor grooving
	DNA2Interaction::get_settings(inp);
	if(!this->_grooving) throw oxDNAException("major_minor_grooving = false is not available in the oxdna_b200 CUDA backend");
}

void CUDADNAInteraction::cuda_init(oxb_ctx *ctx, int N) {
	CUDABaseInteraction::cuda_init(ctx, N);
	Logger::instance()->disable_log("CUDADNAInteraction");
	DNA2Interaction::init();
	Logger::instance()->enable_log("CUDADNAInteraction");
	_upload();
}

void CUDADNAInteraction::_upload() {
	if(_ctx == nullptr) return;
	oxb_dna2_params P;
	double rcut = 0.;
	int rc = oxb_dna2_params_init(&P, (double) this->_T, (double) _salt_concentration, _debye_huckel_half_charged_ends ? 1 : 0, _use_mbf ? 1 : 0,
			(double) _mbf_fmax, (double) _mbf_finf, &rcut);
	if(rc != 0) throw oxDNAException("oxb_dna2_params_init failed (T = %lf, salt = %f)", (double) this->_T, _salt_concentration);
	// what depends on the input file comes from the CPU class that parsed it (CUDADNAInteraction.cu:67-153 copies the same members)
	for(int i = 0; i < 5; i++) {
		for(int j = 0; j < 5; j++) {
			P.hb_eps[5 * i + j] = (float) F1_EPS[HYDR_F1][i][j];
			P.hb_shift[5 * i + j] = (float) F1_SHIFT[HYDR_F1][i][j];
			P.stck_eps[5 * i + j] = (float) F1_EPS[STCK_F1][i][j];
			P.stck_shift[5 * i + j] = (float) F1_SHIFT[STCK_F1][i][j];
		}
	}
	P.hb_multiplier = (float) _hb_multiplier;
	P.dh_minus_kappa = (float) _minus_kappa;
	P.dh_prefactor = (float) _debye_huckel_prefactor;
	P.dh_rhigh = (float) _debye_huckel_RHIGH;
	P.dh_rc = (float) _debye_huckel_RC;
	P.dh_b = (float) _debye_huckel_B;
	if(_use_mbf) {
		P.mbf_xmax = (float) _mbf_xmax;
	}
	// the Verlet radius must be the CPU class's own double-precision cutoff (bit-exact pair sets)
	rcut = (double) this->_rcut;
	P.rcut = (float) rcut;
	oxb_check(_ctx, oxb_set_model_dna2(_ctx, &P, rcut), "set_model_dna2");
}

void CUDADNAInteraction::_on_T_update() {
	// the reference's GPU class re-runs cuda_init -> DNAInteraction::init() (CUDADNAInteraction.cu:156-158)
	this->_T = CONFIG_INFO->temperature();
	if(_ctx != nullptr) {
		Logger::instance()->disable_log("CUDADNAInteraction");
		DNA2Interaction::init();
		Logger::instance()->enable_log("CUDADNAInteraction");
		_upload();
	}
}

void CUDADNA1Interaction::cuda_init(oxb_ctx *ctx, int N) {
	CUDABaseInteraction::cuda_init(ctx, N);
	Logger::instance()->disable_log("CUDADNAInteraction");
	DNAInteraction::init();
	Logger::instance()->enable_log("CUDADNAInteraction");
	_upload();
}

void CUDADNA1Interaction::_upload() {
	if(_ctx == nullptr) return;
	oxb_dna2_params P;
	double rcut = 0.;
	int rc = oxb_dna1_params_init(&P, (double) this->_T, _grooving ? 1 : 0, _use_mbf ? 1 : 0, (double) _mbf_fmax, (double) _mbf_finf, &rcut);
	if(rc != 0) throw oxDNAException("oxb_dna1_params_init failed (T = %lf)", (double) this->_T);
	for(int i = 0; i < 5; i++) {
		for(int j = 0; j < 5; j++) {
			P.hb_eps[5 * i + j] = (float) F1_EPS[HYDR_F1][i][j];
			P.hb_shift[5 * i + j] = (float) F1_SHIFT[HYDR_F1][i][j];
			P.stck_eps[5 * i + j] = (float) F1_EPS[STCK_F1][i][j];
			P.stck_shift[5 * i + j] = (float) F1_SHIFT[STCK_F1][i][j];
		}
	}
	P.hb_multiplier = (float) _hb_multiplier;
	if(_use_mbf) P.mbf_xmax = (float) _mbf_xmax;
	rcut = (double) this->_rcut;
	P.rcut = (float) rcut;
	P.rcut_near = (float) rcut;
	oxb_check(_ctx, oxb_set_model_dna2(_ctx, &P, rcut), "set_model_dna2");
}

void CUDADNA1Interaction::_on_T_update() {
	this->_T = CONFIG_INFO->temperature();
	if(_ctx != nullptr) {
		Logger::instance()->disable_log("CUDADNAInteraction");
		DNAInteraction::init();
		Logger::instance()->enable_log("CUDADNAInteraction");
		_upload();
	}
}

// ---- oxDNA3
void CUDADNA3Interaction::get_settings(input_file &inp) {
	Logger::instance()->disable_log("CUDADNA3Interaction");
	DNA3Interaction::get_settings(inp);
	Logger::instance()->enable_log("CUDADNA3Interaction");
}

void CUDADNA3Interaction::cuda_init(oxb_ctx *ctx, int N) {
	CUDABaseInteraction::cuda_init(ctx, N);
	Logger::instance()->disable_log("CUDADNA3Interaction");
	DNA3Interaction::init();
	Logger::instance()->enable_log("CUDADNA3Interaction");
	_upload();
}

void CUDADNA3Interaction::_upload() {
	if(_ctx == nullptr) return;
	typedef MultiDimArray<TETRAMER_DIM_A, TETRAMER_DIM_B, TETRAMER_DIM_B, TETRAMER_DIM_A> Tab;
	static_assert(Tab::total_size == OXB_DNA3_TSIZE, "tetramer table size");
	std::vector<double> tables((size_t) OXB_DNA3_NTAB * OXB_DNA3_TSIZE);
	double *o = tables.data();
	auto put = [&o](const Tab *t, int n) {
		for(int i = 0; i < n; i++) {
			for(size_t k = 0; k < Tab::total_size; k++) o[k] = (double) t[i].data[k];
			o += Tab::total_size;
		}
	};
	// the order of include/oxdna_b200.h (OXB_DNA3_*)
	put(&_fene_r0_SD, 1); put(&_fene_delta_SD, 1); put(&_fene_delta2_SD, 1); put(&_mbf_xmax_SD, 1);
	put(_excl_s, 7); put(_excl_r, 7); put(_excl_b, 7); put(_excl_rc, 7);
	put(F1_SD_EPS, 2); put(F1_SD_A, 2); put(F1_SD_RC, 2); put(F1_SD_R0, 2); put(F1_SD_BLOW, 2); put(F1_SD_BHIGH, 2); put(F1_SD_RLOW, 2);
	put(F1_SD_RHIGH, 2); put(F1_SD_RCLOW, 2); put(F1_SD_RCHIGH, 2); put(F1_SD_SHIFT, 2);
	put(F2_SD_K, 4); put(F2_SD_K_SYMM, 4); put(F2_SD_RC, 4); put(F2_SD_R0, 4); put(F2_SD_BLOW, 4); put(F2_SD_RLOW, 4); put(F2_SD_RCLOW, 4);
	put(F2_SD_BHIGH, 4); put(F2_SD_RCHIGH, 4); put(F2_SD_RHIGH, 4);
	put(F4_SD_THETA_A, 21); put(F4_SD_THETA_B, 21); put(F4_SD_THETA_T0, 21); put(F4_SD_THETA_TS, 21); put(F4_SD_THETA_TC, 21);
	put(F5_SD_PHI_A, 4); put(F5_SD_PHI_B, 4); put(F5_SD_PHI_XC, 4); put(F5_SD_PHI_XS, 4);
	oxb_dna3_scalars S;
	S.fene_eps = (double) _fene_eps;
	S.use_mbf = _use_mbf ? 1. : 0.; S.mbf_fmax = (double) _mbf_fmax; S.mbf_finf = (double) _mbf_finf;
	S.hb_multiplier = (double) _hb_multiplier;
	S.dh_rc = (double) _debye_huckel_RC; S.dh_rhigh = (double) _debye_huckel_RHIGH; S.dh_prefactor = (double) _debye_huckel_prefactor;
	S.dh_b = (double) _debye_huckel_B; S.dh_minus_kappa = (double) _minus_kappa; S.dh_half_charged_ends = _debye_huckel_half_charged_ends ? 1. : 0.;
	S.rcut = (double) this->_rcut;
	const int ids[3] = { CXST_F4_THETA1, CXST_F4_THETA4, CXST_F4_THETA5 };
	double *dst[3] = { S.cxst_t1, S.cxst_t4, S.cxst_t5 };
	for(int i = 0; i < 3; i++) {
		dst[i][0] = (double) F4_THETA_A[ids[i]]; dst[i][1] = (double) F4_THETA_B[ids[i]]; dst[i][2] = (double) F4_THETA_T0[ids[i]];
		dst[i][3] = (double) F4_THETA_TS[ids[i]]; dst[i][4] = (double) F4_THETA_TC[ids[i]];
	}
	S.cxst_t1_sa = (double) F4_THETA_SA[CXST_F4_THETA1]; S.cxst_t1_sb = (double) F4_THETA_SB[CXST_F4_THETA1];
	oxb_check(_ctx, oxb_set_model_dna3(_ctx, tables.data(), &S), "set_model_dna3");
}

void CUDADNA3Interaction::_on_T_update() {
	// CUDADNA3Interaction.cu:152-154 re-runs cuda_init -> DNA3Interaction::init()
	this->_T = CONFIG_INFO->temperature();
	if(_ctx != nullptr) {
		Logger::instance()->disable_log("CUDADNA3Interaction");
		DNA3Interaction::init();
		Logger::instance()->enable_log("CUDADNA3Interaction");
		_upload();
	}
}

void CUDARNAInteraction::get_settings(input_file &inp) {
	if(_v1) RNAInteraction::get_settings(inp);
	else RNA2Interaction::get_settings(inp);
}

void CUDARNAInteraction::cuda_init(oxb_ctx *ctx, int N) {
	CUDABaseInteraction::cuda_init(ctx, N);
	Logger::instance()->disable_log("CUDARNAInteraction");
	if(_v1) RNAInteraction::init();
	else RNA2Interaction::init();
	Logger::instance()->enable_log("CUDARNAInteraction");
	_upload();
}

static void put_f4(oxb_rna2_params &P, int k, float a, float b, float t0, float ts, float tc) {
	P.f4[k] = oxb_f4 { a, b, t0, ts, tc };
	double lo = std::fmax(0., (double) t0 - tc), hi = std::fmin(3.14159265358979323846, (double) t0 + tc);
	P.f4_cmin[k] = (float) (std::cos(hi) - 1e-4);
	P.f4_cmax[k] = (float) (std::cos(lo) + 1e-4);
}

static void put_excl(oxb_excl &e, float sigma, float rstar, float b, float rc) {
	e.sigma2 = (float) ((double) sigma * sigma);
	e.rstar2 = (float) ((double) rstar * rstar);
	e.b = b;
	e.rc = rc;
	e.rc2 = (float) ((double) rc * rc);
}

void CUDARNAInteraction::_upload() {
	if(_ctx == nullptr) return;
	oxb_rna2_params P;
	double rcut = 0.;
	const bool mismatch = !_v1 && _mismatch_repulsion;
	int rc = oxb_rna2_params_init(&P, (double) this->_T, _v1 ? 0. : (double) _salt_concentration, (!_v1 && _debye_huckel_half_charged_ends) ? 1 : 0, _use_mbf ? 1 : 0,
			(double) _mbf_fmax, (double) _mbf_finf, mismatch ? 1 : 0, mismatch ? (double) _RNA_HYDR_MIS : 0., &rcut);
	if(rc != 0) throw oxDNAException("oxb_rna2_params_init failed (T = %lf, salt = %f)", (double) this->_T, _salt_concentration);
	// everything the input file (or an `external_model` file) can change comes from the CPU class that parsed it
	const Model &m = *model;
	P.back_a1 = m.RNA_POS_BACK_a1; P.back_a2 = m.RNA_POS_BACK_a2; P.back_a3 = m.RNA_POS_BACK_a3;
	P.stack_a1 = m.RNA_POS_STACK; P.base_a1 = m.RNA_POS_BASE;
	P.stack3_a1 = m.RNA_POS_STACK_3_a1; P.stack3_a2 = m.RNA_POS_STACK_3_a2;
	P.stack5_a1 = m.RNA_POS_STACK_5_a1; P.stack5_a2 = m.RNA_POS_STACK_5_a2;
	P.p3[0] = m.p3_x; P.p3[1] = m.p3_y; P.p3[2] = m.p3_z;
	P.p5[0] = m.p5_x; P.p5[1] = m.p5_y; P.p5[2] = m.p5_z;
	P.fene_eps = m.RNA_FENE_EPS; P.fene_r0 = m.RNA_FENE_R0; P.fene_delta = m.RNA_FENE_DELTA; P.fene_delta2 = m.RNA_FENE_DELTA2;
	if(_use_mbf) {
		double xmax = (double) _mbf_xmax, fmax = (double) _mbf_fmax, finf = (double) _mbf_finf;
		P.mbf_xmax = (float) xmax;
		P.mbf_e0 = (float) (-(m.RNA_FENE_EPS / 2.) * std::log(1. - xmax * xmax / m.RNA_FENE_DELTA2) - ((fmax - finf) * xmax * std::log(xmax) + finf * xmax));
	}
	P.excl_eps = m.RNA_EXCL_EPS;
	put_excl(P.excl[0], m.RNA_EXCL_S1, m.RNA_EXCL_R1, m.RNA_EXCL_B1, m.RNA_EXCL_RC1);
	put_excl(P.excl[1], m.RNA_EXCL_S2, m.RNA_EXCL_R2, m.RNA_EXCL_B2, m.RNA_EXCL_RC2);
	put_excl(P.excl[2], m.RNA_EXCL_S3, m.RNA_EXCL_R3, m.RNA_EXCL_B3, m.RNA_EXCL_RC3);
	put_excl(P.excl[3], m.RNA_EXCL_S4, m.RNA_EXCL_R4, m.RNA_EXCL_B4, m.RNA_EXCL_RC4);
	auto put_f1 = [&](oxb_f1 &d, int t) {
		d.a = (float) F1_A[t]; d.rc = (float) F1_RC[t]; d.r0 = (float) F1_R0[t]; d.blow = (float) F1_BLOW[t]; d.bhigh = (float) F1_BHIGH[t];
		d.rlow = (float) F1_RLOW[t]; d.rhigh = (float) F1_RHIGH[t]; d.rclow = (float) F1_RCLOW[t]; d.rchigh = (float) F1_RCHIGH[t];
	};
	put_f1(P.hb, RNA_HYDR_F1);
	put_f1(P.stck, RNA_STCK_F1);
	auto put_f2 = [&](oxb_f2 &d, int t) {
		d.k = (float) F2_K[t]; d.rc = (float) F2_RC[t]; d.r0 = (float) F2_R0[t]; d.blow = (float) F2_BLOW[t]; d.rlow = (float) F2_RLOW[t];
		d.rclow = (float) F2_RCLOW[t]; d.bhigh = (float) F2_BHIGH[t]; d.rhigh = (float) F2_RHIGH[t]; d.rchigh = (float) F2_RCHIGH[t];
	};
	put_f2(P.crst, RNA_CRST_F2);
	put_f2(P.cxst, RNA_CXST_F2);
	// mismatch repulsion rescaled entry [0][0] of the HB table in init() (RNAInteraction2.cpp:96-101): keep it apart
	const float hb_eps_default = m.RNA_HYDR_EPS;
	const double hb_shift_unit = F1_SHIFT[RNA_HYDR_F1][1][1] / F1_EPS[RNA_HYDR_F1][1][1];
	for(int i = 0; i < 5; i++) {
		for(int j = 0; j < 5; j++) {
			P.hb_eps[5 * i + j] = (float) F1_EPS[RNA_HYDR_F1][i][j];
			P.hb_shift[5 * i + j] = (float) F1_SHIFT[RNA_HYDR_F1][i][j];
			P.stck_eps[5 * i + j] = (float) F1_EPS[RNA_STCK_F1][i][j];
			P.stck_shift[5 * i + j] = (float) F1_SHIFT[RNA_STCK_F1][i][j];
			P.crst_kfac[5 * i + j] = (float) _cross_seq_dep_K[i][j];
		}
	}
	P.mismatch_repulsion = mismatch ? 1 : 0;
	if(mismatch) {
		P.mis_eps = (float) F1_EPS[RNA_HYDR_F1][0][0];
		P.mis_shift = (float) F1_SHIFT[RNA_HYDR_F1][0][0];
		P.hb_eps[0] = hb_eps_default; // A-A never hydrogen-bonds; restore the unscaled default
		P.hb_shift[0] = (float) (hb_eps_default * hb_shift_unit);
	}
	P.average = _average ? 1 : 0;
	put_f4(P, OXB_RF4_STCK_T5, m.RNA_STCK_THETA5_A, m.RNA_STCK_THETA5_B, m.RNA_STCK_THETA5_T0, m.RNA_STCK_THETA5_TS, m.RNA_STCK_THETA5_TC);
	put_f4(P, OXB_RF4_STCK_T6, m.RNA_STCK_THETA6_A, m.RNA_STCK_THETA6_B, m.RNA_STCK_THETA6_T0, m.RNA_STCK_THETA6_TS, m.RNA_STCK_THETA6_TC);
	put_f4(P, OXB_RF4_STCK_TB1, m.STCK_THETAB1_A, m.STCK_THETAB1_B, m.STCK_THETAB1_T0, m.STCK_THETAB1_TS, m.STCK_THETAB1_TC);
	put_f4(P, OXB_RF4_STCK_TB2, m.STCK_THETAB2_A, m.STCK_THETAB2_B, m.STCK_THETAB2_T0, m.STCK_THETAB2_TS, m.STCK_THETAB2_TC);
	put_f4(P, OXB_RF4_HB_T1, m.RNA_HYDR_THETA1_A, m.RNA_HYDR_THETA1_B, m.RNA_HYDR_THETA1_T0, m.RNA_HYDR_THETA1_TS, m.RNA_HYDR_THETA1_TC);
	put_f4(P, OXB_RF4_HB_T2, m.RNA_HYDR_THETA2_A, m.RNA_HYDR_THETA2_B, m.RNA_HYDR_THETA2_T0, m.RNA_HYDR_THETA2_TS, m.RNA_HYDR_THETA2_TC);
	put_f4(P, OXB_RF4_HB_T3, m.RNA_HYDR_THETA3_A, m.RNA_HYDR_THETA3_B, m.RNA_HYDR_THETA3_T0, m.RNA_HYDR_THETA3_TS, m.RNA_HYDR_THETA3_TC);
	put_f4(P, OXB_RF4_HB_T4, m.RNA_HYDR_THETA4_A, m.RNA_HYDR_THETA4_B, m.RNA_HYDR_THETA4_T0, m.RNA_HYDR_THETA4_TS, m.RNA_HYDR_THETA4_TC);
	put_f4(P, OXB_RF4_HB_T7, m.RNA_HYDR_THETA7_A, m.RNA_HYDR_THETA7_B, m.RNA_HYDR_THETA7_T0, m.RNA_HYDR_THETA7_TS, m.RNA_HYDR_THETA7_TC);
	put_f4(P, OXB_RF4_HB_T8, m.RNA_HYDR_THETA8_A, m.RNA_HYDR_THETA8_B, m.RNA_HYDR_THETA8_T0, m.RNA_HYDR_THETA8_TS, m.RNA_HYDR_THETA8_TC);
	put_f4(P, OXB_RF4_CRST_T1, m.RNA_CRST_THETA1_A, m.RNA_CRST_THETA1_B, m.RNA_CRST_THETA1_T0, m.RNA_CRST_THETA1_TS, m.RNA_CRST_THETA1_TC);
	put_f4(P, OXB_RF4_CRST_T2, m.RNA_CRST_THETA2_A, m.RNA_CRST_THETA2_B, m.RNA_CRST_THETA2_T0, m.RNA_CRST_THETA2_TS, m.RNA_CRST_THETA2_TC);
	put_f4(P, OXB_RF4_CRST_T3, m.RNA_CRST_THETA3_A, m.RNA_CRST_THETA3_B, m.RNA_CRST_THETA3_T0, m.RNA_CRST_THETA3_TS, m.RNA_CRST_THETA3_TC);
	put_f4(P, OXB_RF4_CRST_T7, m.RNA_CRST_THETA7_A, m.RNA_CRST_THETA7_B, m.RNA_CRST_THETA7_T0, m.RNA_CRST_THETA7_TS, m.RNA_CRST_THETA7_TC);
	put_f4(P, OXB_RF4_CRST_T8, m.RNA_CRST_THETA8_A, m.RNA_CRST_THETA8_B, m.RNA_CRST_THETA8_T0, m.RNA_CRST_THETA8_TS, m.RNA_CRST_THETA8_TC);
	put_f4(P, OXB_RF4_CXST_T1, m.RNA_CXST_THETA1_A, m.RNA_CXST_THETA1_B, m.RNA_CXST_THETA1_T0, m.RNA_CXST_THETA1_TS, m.RNA_CXST_THETA1_TC);
	put_f4(P, OXB_RF4_CXST_T4, m.RNA_CXST_THETA4_A, m.RNA_CXST_THETA4_B, m.RNA_CXST_THETA4_T0, m.RNA_CXST_THETA4_TS, m.RNA_CXST_THETA4_TC);
	put_f4(P, OXB_RF4_CXST_T5, m.RNA_CXST_THETA5_A, m.RNA_CXST_THETA5_B, m.RNA_CXST_THETA5_T0, m.RNA_CXST_THETA5_TS, m.RNA_CXST_THETA5_TC);
	put_f4(P, OXB_RF4_CXST_T6, m.RNA_CXST_THETA6_A, m.RNA_CXST_THETA6_B, m.RNA_CXST_THETA6_T0, m.RNA_CXST_THETA6_TS, m.RNA_CXST_THETA6_TC);
	P.phi1 = oxb_f5 { m.RNA_STCK_PHI1_A, m.RNA_STCK_PHI1_B, m.RNA_STCK_PHI1_XC, m.RNA_STCK_PHI1_XS };
	P.phi2 = oxb_f5 { m.RNA_STCK_PHI2_A, m.RNA_STCK_PHI2_B, m.RNA_STCK_PHI2_XC, m.RNA_STCK_PHI2_XS };
	P.phi3 = oxb_f5 { m.RNA_CXST_PHI3_A, m.RNA_CXST_PHI3_B, m.RNA_CXST_PHI3_XC, m.RNA_CXST_PHI3_XS };
	P.phi4 = oxb_f5 { m.RNA_CXST_PHI4_A, m.RNA_CXST_PHI4_B, m.RNA_CXST_PHI4_XC, m.RNA_CXST_PHI4_XS };
	if(!_v1) {
		P.dh_minus_kappa = (float) _minus_kappa;
		P.dh_prefactor = (float) _debye_huckel_prefactor;
		P.dh_rhigh = (float) _debye_huckel_RHIGH;
		P.dh_rc = (float) _debye_huckel_RC;
		P.dh_b = (float) _debye_huckel_B;
	}
	// cutoffs: the Verlet radius must be the CPU class's own double-precision cutoff (bit-exact pair sets)
	const double back_len = std::sqrt((double) (m.RNA_POS_BACK_a1 * m.RNA_POS_BACK_a1 + m.RNA_POS_BACK_a2 * m.RNA_POS_BACK_a2 + m.RNA_POS_BACK_a3 * m.RNA_POS_BACK_a3));
	const double near_back = 2. * back_len + m.RNA_EXCL_RC1;
	const double near_base = 2. * std::fabs((double) m.RNA_POS_BASE) + std::fmax(std::fmax((double) m.RNA_HYDR_RCHIGH, (double) m.RNA_CRST_RCHIGH), (double) m.RNA_EXCL_RC2);
	const double near_mixed = back_len + std::fabs((double) m.RNA_POS_BASE) + std::fmax((double) m.RNA_EXCL_RC3, (double) m.RNA_EXCL_RC4);
	const double near_stack = 2. * std::fabs((double) m.RNA_POS_STACK) + m.RNA_CXST_RCHIGH;
	P.rcut_near = (float) (std::fmax(std::fmax(near_back, near_base), std::fmax(near_mixed, near_stack)) * (1. + 1e-6));
	rcut = (double) this->_rcut;
	P.rcut = (float) rcut;
	if(P.rcut_near > P.rcut) P.rcut_near = P.rcut;
	oxb_check(_ctx, oxb_set_model_rna2(_ctx, &P, rcut), "set_model_rna2");
}

void CUDARNAInteraction::_on_T_update() {
	// CUDARNAInteraction.cu:421-423 re-runs cuda_init -> RNA2Interaction::init()
	this->_T = CONFIG_INFO->temperature();
	if(_ctx != nullptr) {
		Logger::instance()->disable_log("CUDARNAInteraction");
		if(_v1) RNAInteraction::init();
		else RNA2Interaction::init();
		Logger::instance()->enable_log("CUDARNAInteraction");
		_upload();
	}
}

std::shared_ptr<CUDABaseInteraction> CUDAInteractionFactory::make_interaction(input_file &inp) {
	std::string inter_type("DNA");
	getInputString(&inp, "interaction_type", inter_type, 0);
	if(inter_type == "DNA2") return std::make_shared<CUDADNAInteraction>();
	if(inter_type == "RNA2") return std::make_shared<CUDARNAInteraction>();
	if(inter_type == "RNA") return std::make_shared<CUDARNAInteraction>(true);
	if(inter_type == "DNA" || inter_type == "DNA_nomesh") return std::make_shared<CUDADNA1Interaction>();
	if(inter_type == "DNA3") return std::make_shared<CUDADNA3Interaction>(); // CUDAInteractionFactory.cu:41
	// anything else: CUDA<type>.so on the plugin search path, entry point make_CUDA<type> (or make / make_interaction), exactly as the
	// reference looks it up (CUDAInteractionFactory.cu:44-51, PluginManager.cpp:89-180); the object must be one of OUR CUDABaseInteraction
	std::string cuda_name = "CUDA" + inter_type;
	std::shared_ptr<CUDABaseInteraction> res = std::dynamic_pointer_cast<CUDABaseInteraction>(PluginManager::instance()->get_interaction(cuda_name));
	if(res == nullptr) throw oxDNAException("CUDA interaction '%s' not found. Aborting", cuda_name.c_str());
	return res;
}

// ---------------------------------------------------------------------------------------------------------- lists
void CUDABaseList::init(oxb_ctx *ctx, int N, number rcut, int sort_every) {
	_ctx = ctx;
	_N = N;
}

void CUDASimpleVerletList::get_settings(input_file &inp) {
	// src/CUDA/Lists/CUDASimpleVerletList.cu:47-56
	getInputBool(&inp, "cells_auto_optimisation", &_auto_optimisation, 0);
	getInputBool(&inp, "print_problematic_ids", &_print_problematic_ids, 0);
	getInputNumber(&inp, "verlet_skin", &_verlet_skin, 1);
	getInputNumber(&inp, "max_density_multiplier", &_max_density_multiplier, 0);
	getInputBool(&inp, "use_edge", &_use_edge, 0);
	if(_use_edge) {
		OX_LOG(Logger::LOG_INFO, "Using edge-based approach");
	}
}

void CUDASimpleVerletList::init(oxb_ctx *ctx, int N, number rcut, int sort_every) {
	CUDABaseList::init(ctx, N, rcut, sort_every);
	oxb_check(ctx, oxb_set_lists(ctx, (double) _verlet_skin, _use_edge ? 1 : 0, sort_every, (double) _max_density_multiplier), "set_lists");
}

void CUDASimpleVerletList::update() {
	int rc = oxb_update_lists(_ctx);
	if(rc != 0) throw oxDNAException("A cell or neighbour row contains too many particles (%s). Please increase the value of max_density_multiplier (which defaults to 3) in the input file", oxb_last_error(_ctx));
}

std::shared_ptr<CUDABaseList> CUDAListFactory::make_list(input_file &inp) {
	// src/CUDA/Lists/CUDAListFactory.cu
	std::string list_type("verlet");
	getInputString(&inp, "CUDA_list", list_type, 0);
	if(list_type == "verlet") return std::make_shared<CUDASimpleVerletList>();
	if(list_type == "no") {
		bool use_edge = false;
		getInputBool(&inp, "use_edge", &use_edge, 0);
		if(use_edge) throw oxDNAException("'CUDA_list = no' and 'use_edge = true' are incompatible");
		// CUDANoList evaluates every pair of particles against the interaction cutoff (src/CUDA/Lists/CUDANoList.cu,
		// dna_forces with NO_LIST, CUDA_DNA.cuh:832-904): forces and energies depend on the cutoff, not on how the pairs were found,
		// so the cell-binned list (bit-exact pair set inside rcut + 2 skin) serves the key with identical results in O(N) work
		OX_LOG(Logger::LOG_INFO, "CUDA_list = no: served by the cell-binned Verlet list (same forces, O(N) pair search)");
		return std::make_shared<CUDASimpleVerletList>();
	}
	throw oxDNAException("CUDA_list '%s' is not supported", list_type.c_str());
}

// ----------------------------------------------------------------------------------------------------- thermostats
void CUDABaseThermostat::apply_cuda(llint curr_step) {
	oxb_check(_ctx, oxb_set_step(_ctx, curr_step), "set_step");
	oxb_check(_ctx, oxb_thermostat(_ctx), "thermostat");
}

void CUDANoThermostat::_upload() {
	if(_ctx) oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_NONE, 1, 0., 0., 0., 0., (unsigned long long) _seed), "set_thermostat");
}

void CUDABrownianThermostat::get_settings(input_file &inp) {
	BrownianThermostat::get_settings(inp);
}

void CUDABrownianThermostat::init() {
	BrownianThermostat::init();
	_upload();
}

void CUDABrownianThermostat::_upload() {
	if(_ctx) oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_BROWNIAN, _newtonian_steps, (double) _pt, (double) _pr, (double) _rescale_factor, 0., (unsigned long long) _seed), "set_thermostat");
}

bool CUDABrownianThermostat::would_activate(llint curr_step) {
	return (curr_step % _newtonian_steps == 0);
}

void CUDALangevinThermostat::get_settings(input_file &inp) {
	LangevinThermostat::get_settings(inp);
}

void CUDALangevinThermostat::init() {
	LangevinThermostat::init();
	_upload();
}

void CUDALangevinThermostat::_upload() {
	if(_ctx) oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_LANGEVIN, 1, (double) _gamma_trans, (double) _gamma_rot, (double) _rescale_factor_trans, (double) _rescale_factor_rot, (unsigned long long) _seed), "set_thermostat");
}

void CUDABussiThermostat::get_settings(input_file &inp) {
	BussiThermostat::get_settings(inp);
}

void CUDABussiThermostat::init() {
	BussiThermostat::init();
	_upload();
}

void CUDABussiThermostat::_upload() {
	if(_ctx) oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_BUSSI, _newtonian_steps, (double) this->_T, (double) _exp_dt_tau, 0., 0., (unsigned long long) _seed), "set_thermostat");
}

bool CUDABussiThermostat::would_activate(llint curr_step) {
	return (curr_step % _newtonian_steps == 0);
}

std::shared_ptr<CUDABaseThermostat> CUDAThermostatFactory::make_thermostat(input_file &inp, BaseBox *box) {
	// src/CUDA/Thermostats/CUDAThermostatFactory.cu:18-45
	char thermostat_type[512] = "no";
	getInputString(&inp, "thermostat", thermostat_type, 0);
	if(!strncmp(thermostat_type, "john", 512) || !strncmp(thermostat_type, "brownian", 512)) return std::make_shared<CUDABrownianThermostat>();
	else if(!strncmp(thermostat_type, "bussi", 512) || !strncmp(thermostat_type, "Bussi", 512)) return std::make_shared<CUDABussiThermostat>();
	else if(!strncmp(thermostat_type, "langevin", 512)) return std::make_shared<CUDALangevinThermostat>();
	else if(!strncmp(thermostat_type, "no", 512)) return std::make_shared<CUDANoThermostat>();
	else throw oxDNAException("Invalid CUDA thermostat '%s'", thermostat_type);
}
