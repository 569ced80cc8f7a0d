// Host-side derivation of the oxDNA2 parameter block (product code; does not touch oracle/).
//
// Mirrors what the reference does in two hops -- DNA2Interaction::get_settings/init on the CPU
// (src/Interactions/DNA2Interaction.cpp:61-149, src/Interactions/DNAInteraction.cpp:12-185,290-375) and the
// float down-conversion of CUDADNAInteraction::cuda_init (src/CUDA/Interactions/CUDADNAInteraction.cu:61-154).
// Numbers are the published oxDNA2 model constants (src/model.h).  The cutoff is derived in double precision
// with the reference's exact operation order because the Verlet radius rcut + 2*skin enters a bit-exact
// neighbour predicate.
#include "../../include/oxdna_b200.h"

#include <cmath>
#include <cstring>

namespace {

// the reference's PI is a float literal (src/defs.h:14)
constexpr float kPi = 3.141592653589793238462643f;

struct WellF1 { float a, rc, r0, blow, bhigh, rlow, rhigh, rclow, rchigh; };
struct AngF4 { float a, b, t0, ts, tc; };

constexpr WellF1 kHB = { 8.f, 0.75f, 0.4f, -126.243f, -7.87708f, 0.34f, 0.7f, 0.276908f, 0.783775f };
constexpr WellF1 kSTCK = { 6.f, 0.9f, 0.4f, -68.1857f, -3.12992f, 0.32f, 0.75f, 0.23239f, 0.956f };

double morse_shift(const WellF1 &w) {
	// (1 - exp(-(rc - r0) a))^2 with the float product the reference's macros produce
	double e = std::exp(-(double) ((w.rc - w.r0) * w.a));
	return (1 - e) * (1 - e);
}

void put_f1(oxb_f1 &dst, const WellF1 &w) {
	dst.a = w.a; dst.rc = w.rc; dst.r0 = w.r0; dst.blow = w.blow; dst.bhigh = w.bhigh;
	dst.rlow = w.rlow; dst.rhigh = w.rhigh; dst.rclow = w.rclow; dst.rchigh = w.rchigh;
}

void put_excl(oxb_excl &e, float sigma, float rstar, float b, float rc) {
	e.sigma2 = (float) ((double) sigma * sigma);
	e.rstar2 = (float) ((double) rstar * rstar);
	e.b = b;
	e.rc = rc;
	e.rc2 = (float) ((double) rc * rc);
}

} // namespace

extern "C" int oxb_dna2_params_init(oxb_dna2_params *P, double T, double salt, int dh_half_charged_ends, int use_mbf,
		double mbf_fmax, double mbf_finf, double *rcut_out) {
	if(P == nullptr || !(T > 0) || !(salt > 0)) return 1;
	std::memset(P, 0, sizeof(*P));

	// interaction sites (oxDNA2 grooved backbone), src/model.h:14-18, src/Particles/DNANucleotide.cpp:76-80
	const float back1 = -0.3400f, back2 = 0.3408f, stack = 0.34f, base = 0.4f;
	P->back_a1 = back1; P->back_a2 = back2; P->stack_a1 = stack;
	P->base_a1 = (float) ((double) stack * ((double) base / (double) stack));
	P->backref_a1 = -0.4f;

	P->fene_eps = 2.0f; P->fene_r0 = 0.7564f; P->fene_delta = 0.25f; P->fene_delta2 = 0.0625f;
	P->use_mbf = use_mbf ? 1 : 0;
	if(use_mbf) {
		double eps = 2.0, d2 = 0.0625;
		double xmax = (-eps + std::sqrt(eps * eps + 4. * mbf_fmax * mbf_fmax * d2)) / (2. * mbf_fmax);
		double fene_xmax = -(eps / 2.) * std::log(1. - xmax * xmax / d2);
		double long_xmax = (mbf_fmax - mbf_finf) * xmax * std::log(xmax) + mbf_finf * xmax;
		P->mbf_xmax = (float) xmax; P->mbf_fmax = (float) mbf_fmax; P->mbf_finf = (float) mbf_finf;
		P->mbf_e0 = (float) (fene_xmax - long_xmax);
	}

	P->excl_eps = 2.0f;
	put_excl(P->excl[0], 0.70f, 0.675f, 892.016223343f, 0.711879214356f);
	put_excl(P->excl[1], 0.33f, 0.32f, 4119.70450017f, 0.335388426126f);
	put_excl(P->excl[2], 0.515f, 0.50f, 1707.30627298f, 0.52329943261f);
	put_excl(P->excl[3], 0.515f, 0.50f, 1707.30627298f, 0.52329943261f);

	put_f1(P->hb, kHB);
	put_f1(P->stck, kSTCK);
	const double eps_hb = 1.0678f;
	const double eps_st = 1.3523f + 2.6717f * T; // DNA2Interaction.cpp:115
	for(int i = 0; i < 25; i++) {
		P->hb_eps[i] = (float) eps_hb;
		P->hb_shift[i] = (float) (eps_hb * morse_shift(kHB));
		P->stck_eps[i] = (float) eps_st;
		P->stck_shift[i] = (float) (eps_st * morse_shift(kSTCK));
	}

	P->crst = oxb_f2{ 47.5f, 0.675f, 0.575f, -0.888889f, 0.495f, 0.45f, -0.888889f, 0.655f, 0.7f };
	P->cxst = oxb_f2{ 58.5f, 0.6f, 0.400f, -2.13158f, 0.22f, 0.177778f, -2.13158f, 0.58f, 0.6222222f };

	const AngF4 f4tab[OXB_NF4] = {
		{ 1.3f, 6.4381f, 0.f, 0.8f, 0.961538f },              // STCK theta4
		{ 0.9f, 3.89361f, 0.f, 0.95f, 1.16959f },             // STCK theta5 / theta6
		{ 1.5f, 4.16038f, 0.f, 0.7f, 0.952381f },             // HB theta1
		{ 1.5f, 4.16038f, 0.f, 0.7f, 0.952381f },             // HB theta2 / theta3
		{ 0.46f, 0.133855f, kPi, 0.7f, 3.10559f },            // HB theta4
		{ 4.f, 17.0526f, kPi * 0.5f, 0.45f, 0.555556f },      // HB theta7 / theta8
		{ 2.25f, 7.00545f, kPi - 2.35f, 0.58f, 0.766284f },   // CRST theta1
		{ 1.70f, 6.2469f, 1.f, 0.68f, 0.865052f },            // CRST theta2 / theta3
		{ 1.50f, 2.59556f, 0.f, 0.65f, 1.02564f },            // CRST theta4
		{ 1.70f, 6.2469f, 0.875f, 0.68f, 0.865052f },         // CRST theta7 / theta8
		{ 2.f, 10.9032f, kPi - 0.25f, 0.65f, 0.769231f },     // CXST theta1 (oxDNA2 t0)
		{ 1.3f, 6.4381f, 0.f, 0.8f, 0.961538f },              // CXST theta4
		{ 0.9f, 3.89361f, 0.f, 0.95f, 1.16959f },             // CXST theta5 / theta6
	};
	for(int i = 0; i < OXB_NF4; i++) {
		P->f4[i] = oxb_f4{ f4tab[i].a, f4tab[i].b, f4tab[i].t0, f4tab[i].ts, f4tab[i].tc };
		// support of f4 in cosine space, slightly widened: theta in (t0 - tc, t0 + tc) intersected with [0, pi]
		double lo = std::fmax(0., (double) f4tab[i].t0 - f4tab[i].tc), hi = std::fmin(3.14159265358979323846, (double) f4tab[i].t0 + f4tab[i].tc);
		P->f4_cmin[i] = (float) (std::cos(hi) - 1e-4);
		P->f4_cmax[i] = (float) (std::cos(lo) + 1e-4);
	}
	P->cxst_t1_sa = 20.f;
	P->cxst_t1_sb = kPi - 0.1f * (kPi - (kPi - 0.25f));
	P->phi1 = oxb_f5{ 2.0f, 10.9032f, -0.769231f, -0.65f };
	P->phi2 = P->phi1;

	// Debye-Hueckel.  RHIGH is fixed in get_settings with a float 0.1f, lambda in init with a double 0.1.
	const double lfac = 0.3616455, q = 0.0543;
	salt = (double) (float) salt; // the reference keeps the salt concentration in a float (DNA2Interaction.h:31)
	const double lambda_gs = lfac * std::sqrt(T / 0.1f) / std::sqrt(salt);
	const double lambda = lfac * std::sqrt(T / 0.1) / std::sqrt(salt);
	const double x = 3.0 * lambda_gs, l = lambda;
	const double B = -(std::exp(-x / l) * q * q * (x + l) * (x + l)) / (-4. * x * x * x * l * l * q);
	const double RC = x * (q * x + 3. * q * l) / (q * (x + l));
	P->dh_minus_kappa = (float) (-1.0 / lambda);
	P->dh_prefactor = (float) q;
	P->dh_rhigh = (float) x;
	P->dh_rc = (float) RC;
	P->dh_b = (float) B;
	P->dh_half_charged_ends = dh_half_charged_ends ? 1 : 0;
	P->hb_multiplier = 1.f;

	// cutoffs
	const double back_len = std::sqrt((double) (back1 * back1 + back2 * back2));
	const double rcutback = 2 * back_len + (double) 0.711879214356f;
	const double rcutbase = 2 * std::fabs((double) base) + (double) kHB.rchigh;
	double rcut_near = std::fmax(rcutback, rcutbase);
	double rcut = rcut_near;
	const double debyecut = 2.0 * back_len + RC;
	if(debyecut > rcut) rcut = debyecut;
	P->rcut = (float) rcut;
	P->rcut_near = (float) rcut_near;
	if(rcut_out != nullptr) *rcut_out = rcut;
	return 0;
}

extern "C" int oxb_dna1_params_init(oxb_dna2_params *P, double T, int grooving, int use_mbf, double mbf_fmax, double mbf_finf, double *rcut_out) {
	// everything oxDNA and oxDNA2 share comes from the oxDNA2 block; then the first-generation differences (src/model.h:48,91,152-154,362,377)
	int rc = oxb_dna2_params_init(P, T, 1.0, 0, use_mbf, mbf_fmax, mbf_finf, nullptr);
	if(rc != 0) return rc;
	P->v1 = 1;
	if(!grooving) {
		// DNANucleotide.cpp:83-87: STACK = BACK * (POS_STACK / POS_BACK), BASE = STACK * (POS_BASE / POS_STACK)
		P->back_a1 = -0.4f; P->back_a2 = 0.f;
		const double stack = (double) -0.4f * (double) (0.34f / -0.4f);
		P->stack_a1 = (float) stack;
		P->base_a1 = (float) (stack * (double) (0.4f / 0.34f));
	}
	P->fene_r0 = 0.7525f;
	if(use_mbf) {
		// the energy offset of the log tail does not depend on r0
	}
	const double eps_hb = 1.077f;
	const double eps_st = 1.3448f + 2.6568f * T; // DNAInteraction.cpp:317
	for(int i = 0; i < 25; i++) {
		P->hb_eps[i] = (float) eps_hb;
		P->hb_shift[i] = (float) (eps_hb * morse_shift(kHB));
		P->stck_eps[i] = (float) eps_st;
		P->stck_shift[i] = (float) (eps_st * morse_shift(kSTCK));
	}
	P->cxst.k = 46.0f;
	{
		const float t0 = kPi - 0.60f, tc = 0.769231f;
		P->f4[OXB_F4_CXST_T1] = oxb_f4{ 2.f, 10.9032f, t0, 0.65f, tc };
		// f4(theta) + f4(2 pi - theta): support theta > t0 - tc (the mirror's support lies inside)
		P->f4_cmin[OXB_F4_CXST_T1] = -1.f - 1e-4f;
		P->f4_cmax[OXB_F4_CXST_T1] = (float) (std::cos((double) t0 - tc) + 1e-4);
	}
	P->phi3 = oxb_f5{ 2.0f, 10.9032f, -0.769231f, -0.65f };
	P->dh_minus_kappa = 0.f; P->dh_prefactor = 0.f; P->dh_rhigh = 0.f; P->dh_rc = 0.f; P->dh_b = 0.f;
	P->dh_half_charged_ends = 0;
	const double rcutback = grooving ? 2 * std::sqrt((double) ((-0.3400f) * (-0.3400f) + (0.3408f) * (0.3408f))) + (double) 0.711879214356f
			: 2 * std::fabs((double) -0.4f) + (double) 0.711879214356f;
	const double rcutbase = 2 * std::fabs((double) 0.4f) + (double) kHB.rchigh;
	const double rcut = std::fmax(rcutback, rcutbase);
	P->rcut = (float) rcut;
	P->rcut_near = (float) rcut;
	if(rcut_out != nullptr) *rcut_out = rcut;
	return 0;
}

extern "C" int oxb_dna2_params_seqdep(oxb_dna2_params *P, double T, const double *stck_raw16, double stck_fact_eps, double hb_AT,
		double hb_GC) {
	if(P == nullptr || stck_raw16 == nullptr) return 1;
	// base order A=0, G=1, C=2, T=3; table index = type_n3 * 5 + type_n5
	for(int i = 0; i < 4; i++) {
		for(int j = 0; j < 4; j++) {
			double eps = stck_raw16[4 * i + j] * (1.0 - stck_fact_eps + (T * 9.0 * stck_fact_eps));
			P->stck_eps[5 * i + j] = (float) eps;
			P->stck_shift[5 * i + j] = (float) (eps * morse_shift(kSTCK));
		}
	}
	// stacking with the dummy base 'D' (type 4): the reference reads the OPTIONAL keys STCK_D_X / STCK_X_D with a value variable that still
	// holds the last mandatory key, STCK_T_T (DNAInteraction.cpp:349-359: getInputFloat(..., 0) leaves it untouched when the key is absent),
	// so with the stock parameter file every dummy entry gets the T-T strength.  Restated as is.
	for(int i = 0; i < 5; i++) {
		for(int j = 0; j < 5; j++) {
			if(i == 4 || j == 4) { P->stck_eps[5 * i + j] = P->stck_eps[5 * 3 + 3]; P->stck_shift[5 * i + j] = P->stck_shift[5 * 3 + 3]; }
		}
	}
	const int A = 0, G = 1, C = 2, Tt = 3;
	P->hb_eps[5 * A + Tt] = P->hb_eps[5 * Tt + A] = (float) hb_AT;
	P->hb_eps[5 * G + C] = P->hb_eps[5 * C + G] = (float) hb_GC;
	P->hb_shift[5 * A + Tt] = P->hb_shift[5 * Tt + A] = (float) (hb_AT * morse_shift(kHB));
	P->hb_shift[5 * G + C] = P->hb_shift[5 * C + G] = (float) (hb_GC * morse_shift(kHB));
	return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// oxRNA2.  Constants: src/Interactions/rna_model.h (struct Model, all members float); derivations:
// RNAInteraction::init (src/Interactions/RNAInteraction.cpp:194-397), RNA2Interaction::get_settings/init
// (src/Interactions/RNAInteraction2.cpp:31-102).
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr WellF1 kRnaHB = { 8.f, 0.75f, 0.4f, -126.243f, -7.87708f, 0.34f, 0.7f, 0.276908f, 0.783775f };
constexpr WellF1 kRnaSTCK = { 6.f, 0.93f, 0.43f, -68.1857f, -3.12992f, 0.35f, 0.78f, 0.26239f, 0.986f };
constexpr float kRnaHbEps = 0.870439f;

void put_f4(oxb_rna2_params *P, int k, float a, float b, float t0, float ts, float tc) {
	P->f4[k] = oxb_f4{ a, b, t0, ts, tc };
	double lo = std::fmax(0., (double) t0 - tc), hi = std::fmin(3.14159265358979323846, (double) t0 + tc);
	P->f4_cmin[k] = (float) (std::cos(hi) - 1e-4);
	P->f4_cmax[k] = (float) (std::cos(lo) + 1e-4);
}

} // namespace

extern "C" int oxb_rna2_params_init(oxb_rna2_params *P, double T, double salt, int dh_half_charged_ends, int use_mbf, double mbf_fmax,
		double mbf_finf, int mismatch_repulsion, double mismatch_strength, double *rcut_out) {
	if(P == nullptr || !(T > 0)) return 1;
	const bool with_dh = salt > 0; // salt <= 0: first-generation oxRNA (interaction_type = RNA), no Debye-Hueckel term
	std::memset(P, 0, sizeof(*P));
	P->average = 1;

	// interaction sites, rna_model.h:313-350
	P->back_a1 = -0.4f; P->back_a2 = 0.0f; P->back_a3 = 0.2f;
	P->stack_a1 = 0.34f; P->base_a1 = 0.4f;
	P->stack3_a1 = 0.4f; P->stack3_a2 = 0.1f;
	P->stack5_a1 = 0.124906078525f; P->stack5_a2 = -0.00866274917473f;
	P->p5[0] = -0.104402f; P->p5[1] = -0.841783f; P->p5[2] = 0.529624f;
	P->p3[0] = -0.462510f; P->p3[1] = -0.528218f; P->p3[2] = 0.712089f;

	P->fene_eps = 2.0f; P->fene_r0 = 0.761070781051f; P->fene_delta = 0.25f; P->fene_delta2 = 0.0625f;
	P->use_mbf = use_mbf ? 1 : 0;
	if(use_mbf) {
		double eps = 2.0, d2 = 0.0625;
		double xmax = (-eps + std::sqrt(eps * eps + 4. * mbf_fmax * mbf_fmax * d2)) / (2. * mbf_fmax);
		double fene_xmax = -(eps / 2.) * std::log(1. - xmax * xmax / d2);
		double long_xmax = (mbf_fmax - mbf_finf) * xmax * std::log(xmax) + mbf_finf * xmax;
		P->mbf_xmax = (float) xmax; P->mbf_fmax = (float) mbf_fmax; P->mbf_finf = (float) mbf_finf;
		P->mbf_e0 = (float) (fene_xmax - long_xmax);
	}

	P->excl_eps = 2.0f;
	put_excl(P->excl[0], 0.70f, 0.675f, 892.016223343f, 0.711879214356f);
	put_excl(P->excl[1], 0.33f, 0.32f, 4119.70450017f, 0.335388426126f);
	put_excl(P->excl[2], 0.515f, 0.50f, 1707.30627298f, 0.52329943261f);
	put_excl(P->excl[3], 0.515f, 0.50f, 1707.30627298f, 0.52329943261f);

	put_f1(P->hb, kRnaHB);
	put_f1(P->stck, kRnaSTCK);
	const double eps_hb = kRnaHbEps;
	const double eps_st = 1.40206f + 2.77f * T; // RNAInteraction.cpp:331
	for(int i = 0; i < 25; i++) {
		P->hb_eps[i] = (float) eps_hb;
		P->hb_shift[i] = (float) (eps_hb * morse_shift(kRnaHB));
		P->stck_eps[i] = (float) eps_st;
		P->stck_shift[i] = (float) (eps_st * morse_shift(kRnaSTCK));
		P->crst_kfac[i] = 1.f;
	}
	P->crst = oxb_f2{ 59.9626f, 0.6f, 0.5f, -0.888889f, 0.42f, 0.375f, -0.888889f, 0.58f, 0.625f };
	P->cxst = oxb_f2{ 80.f, 0.6f, 0.5f, -0.888889f, 0.42f, 0.375f, -0.888889f, 0.58f, 0.625f };

	put_f4(P, OXB_RF4_STCK_T5, 0.9f, 3.89361f, 0.f, 0.95f, 1.16959f);
	put_f4(P, OXB_RF4_STCK_T6, 0.9f, 3.89361f, 0.f, 0.95f, 1.16959f);
	put_f4(P, OXB_RF4_STCK_TB1, 1.3f, 6.4381f, 0.f, 0.8f, 0.961538f);
	put_f4(P, OXB_RF4_STCK_TB2, 1.3f, 6.4381f, 0.f, 0.8f, 0.961538f);
	put_f4(P, OXB_RF4_HB_T1, 1.5f, 4.16038f, 0.f, 0.7f, 0.952381f);
	put_f4(P, OXB_RF4_HB_T2, 1.5f, 4.16038f, 0.f, 0.7f, 0.952381f);
	put_f4(P, OXB_RF4_HB_T3, 1.5f, 4.16038f, 0.f, 0.7f, 0.952381f);
	put_f4(P, OXB_RF4_HB_T4, 0.46f, 0.133855f, kPi, 0.7f, 3.10559f);
	put_f4(P, OXB_RF4_HB_T7, 4.f, 17.0526f, kPi * 0.5f, 0.45f, 0.555556f);
	put_f4(P, OXB_RF4_HB_T8, 4.f, 17.0526f, kPi * 0.5f, 0.45f, 0.555556f);
	put_f4(P, OXB_RF4_CRST_T1, 2.25f, 7.00545f, 0.505f, 0.58f, 0.766284f);
	put_f4(P, OXB_RF4_CRST_T2, 1.70f, 6.2469f, 1.266f, 0.68f, 0.865052f);
	put_f4(P, OXB_RF4_CRST_T3, 1.70f, 6.2469f, 1.266f, 0.68f, 0.865052f);
	put_f4(P, OXB_RF4_CRST_T7, 1.70f, 6.2469f, 0.309f, 0.68f, 0.865052f);
	put_f4(P, OXB_RF4_CRST_T8, 1.70f, 6.2469f, 0.309f, 0.68f, 0.865052f);
	put_f4(P, OXB_RF4_CXST_T1, 2.f, 10.9032f, 2.592f, 0.65f, 0.769231f);
	put_f4(P, OXB_RF4_CXST_T4, 1.3f, 6.4381f, 0.151f, 0.8f, 0.961538f);
	put_f4(P, OXB_RF4_CXST_T5, 0.9f, 3.89361f, 0.685f, 0.95f, 1.16959f);
	put_f4(P, OXB_RF4_CXST_T6, 0.9f, 3.89361f, 0.685f, 0.95f, 1.16959f);
	P->phi1 = oxb_f5{ 2.0f, 10.9032f, -0.769231f, -0.65f };
	P->phi2 = P->phi1; P->phi3 = P->phi1; P->phi4 = P->phi1;

	// Debye-Hueckel: both get_settings and init use the float 0.1f here (unlike DNA2)
	const double lfac = 0.3667258, q = 0.0858;
	double RC = 0.;
	if(with_dh) {
		salt = (double) (float) salt;
		const double lambda = lfac * std::sqrt(T / 0.1f) / std::sqrt(salt);
		const double x = 3.0 * lambda, l = lambda;
		const double B = -(std::exp(-x / l) * q * q * (x + l) * (x + l)) / (4. * x * x * x * l * l * (-q));
		RC = x * (q * x + 3. * q * l) / (q * (x + l));
		P->dh_minus_kappa = (float) (-1.0 / lambda);
		P->dh_prefactor = (float) q;
		P->dh_rhigh = (float) x;
		P->dh_rc = (float) RC;
		P->dh_b = (float) B;
	}
	P->dh_half_charged_ends = dh_half_charged_ends ? 1 : 0;
	P->hb_multiplier = 1.f;

	P->mismatch_repulsion = mismatch_repulsion ? 1 : 0;
	if(mismatch_repulsion) {
		const float temp = -1.0f * (float) mismatch_strength / kRnaHbEps; // RNAInteraction2.cpp:97
		P->mis_eps = (float) (eps_hb * temp);
		P->mis_shift = (float) (eps_hb * morse_shift(kRnaHB) * temp);
	}

	const float b2 = P->back_a1 * P->back_a1 + P->back_a2 * P->back_a2 + P->back_a3 * P->back_a3;
	const double back_len = std::sqrt((double) b2);
	const double rcutback = 2 * back_len + (double) 0.711879214356f;
	const double rcutbase = 2 * std::fabs((double) P->base_a1) + std::fmax((double) kRnaHB.rchigh, (double) P->crst.rchigh);
	double rcut_near = std::fmax(rcutback, rcutbase);
	double rcut = rcut_near;
	const double debyecut = 2. * back_len + RC;
	if(with_dh && debyecut > rcut) rcut = debyecut;
	P->rcut = (float) rcut;
	P->rcut_near = (float) rcut_near;
	if(rcut_out != nullptr) *rcut_out = rcut;
	return 0;
}

extern "C" int oxb_rna2_params_seqdep(oxb_rna2_params *P, double T, const double *stck_raw16, double st_t_dep, const double *cross_raw16,
		double hb_AT, double hb_GC, double hb_GT) {
	if(P == nullptr || stck_raw16 == nullptr || cross_raw16 == nullptr) return 1;
	P->average = 0;
	for(int i = 0; i < 4; i++) {
		for(int j = 0; j < 4; j++) {
			double eps = (double) (float) stck_raw16[4 * i + j] * (1.0 + T * (double) (float) st_t_dep);
			P->stck_eps[5 * i + j] = (float) eps;
			P->stck_shift[5 * i + j] = (float) (eps * morse_shift(kRnaSTCK));
			P->crst_kfac[5 * i + j] = (float) cross_raw16[4 * i + j] / P->crst.k;
		}
	}
	const int A = 0, G = 1, C = 2, U = 3;
	const int ij[3][2] = { { A, U }, { G, C }, { G, U } };
	const double v[3] = { (double) (float) hb_AT, (double) (float) hb_GC, (double) (float) hb_GT };
	for(int k = 0; k < 3; k++) {
		int i = ij[k][0], j = ij[k][1];
		P->hb_eps[5 * i + j] = P->hb_eps[5 * j + i] = (float) v[k];
		P->hb_shift[5 * i + j] = P->hb_shift[5 * j + i] = (float) (v[k] * morse_shift(kRnaHB));
	}
	return 0;
}

extern "C" int oxb_sizeof(int which) {
	switch(which) {
	case 0: return (int) sizeof(oxb_dna2_params);
	case 1: return (int) sizeof(oxb_rna2_params);
	case 2: return (int) sizeof(oxb_ext_force);
	case 3: return (int) sizeof(oxb_replica_consts);
	default: return -1;
	}
}

// the temperature-dependent subset of a parameter block (see oxb_replica_consts): stacking strength and Debye-Hueckel
template<class PB>
static void replica_consts_from(const PB *P, double a, double b, double c, double d, oxb_replica_consts *out) {
	for(int i = 0; i < 25; i++) { out->stck_eps[i] = P->stck_eps[i]; out->stck_shift[i] = P->stck_shift[i]; }
	out->dh_minus_kappa = P->dh_minus_kappa; out->dh_prefactor = P->dh_prefactor; out->dh_rhigh = P->dh_rhigh; out->dh_rc = P->dh_rc; out->dh_b = P->dh_b;
	out->rcut2 = P->rcut * P->rcut;
	out->th_a = (float) a; out->th_b = (float) b; out->th_c = (float) c; out->th_d = (float) d;
}
extern "C" void oxb_replica_consts_dna2(const oxb_dna2_params *P, double a, double b, double c, double d, oxb_replica_consts *out) { replica_consts_from(P, a, b, c, d, out); }
extern "C" void oxb_replica_consts_rna2(const oxb_rna2_params *P, double a, double b, double c, double d, oxb_replica_consts *out) { replica_consts_from(P, a, b, c, d, out); }
