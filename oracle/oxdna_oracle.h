/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, double precision) of the reference's algorithm for
 * the oxDNA2 / oxRNA2 MD step.  Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker; the product (oxdna_b200/) never includes, links or calls anything in this directory.
 *
 * Pinning: validated against (i) the reference's golden vector test/DNA/FORCE_FIELD/AVG_SEQ/reference.dat
 * and (ii) the unmodified reference compiled into oracle/_ref (forces, torques, per-term energies, Verlet
 * pair sets, NVE trajectories) -- see tests/test_oracle.py.
 */
#ifndef OXDNA_ORACLE_H
#define OXDNA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { OXO_FENE = 0, OXO_BEXC, OXO_STCK, OXO_NEXC, OXO_HB, OXO_CRST, OXO_CXST, OXO_DH, OXO_NTERMS };

/* radial well ("f1", Morse-like) -- src/Interactions/DNAInteraction.cpp:1249-1283 */
typedef struct { double a, rc, r0, blow, bhigh, rlow, rhigh, rclow, rchigh; double eps[5][5], shift[5][5]; } oxo_f1;
/* radial well ("f2", harmonic) -- DNAInteraction.cpp:1285-1315 */
typedef struct { double k, rc, r0, blow, rlow, rclow, bhigh, rhigh, rchigh; } oxo_f2;
/* angular modulation ("f4") -- DNAInteraction.cpp:1351-1420 */
typedef struct { double a, b, t0, ts, tc; } oxo_f4;
/* angular modulation in cos space ("f5") -- DNAInteraction.cpp:1422-1454 */
typedef struct { double a, b, xc, xs; } oxo_f5;
/* repulsive LJ with quadratic smoothing -- DNAInteraction.cpp:1182-1205 */
typedef struct { double sigma, rstar, b, rc; } oxo_excl;

typedef struct {
	/* site geometry, src/Particles/DNANucleotide.cpp:66-88, src/model.h:14-18 */
	double back_a1, back_a2, stack_a1, base_a1, backref_a1;
	double T;
	/* bonded */
	double fene_eps, fene_r0, fene_delta, fene_delta2;
	int use_mbf;
	double mbf_xmax, mbf_fmax, mbf_finf;
	oxo_excl excl[4]; /* 0: back-back, 1: base-base, 2: base(p)-back(q), 3: back(p)-base(q) */
	double excl_eps;
	oxo_f1 hb, stck;
	oxo_f2 crst, cxst;
	oxo_f4 stck_t4, stck_t5, hb_t1, hb_t2, hb_t4, hb_t7, crst_t1, crst_t2, crst_t4, crst_t7, cxst_t1, cxst_t4, cxst_t5;
	double cxst_t1_sa, cxst_t1_sb; /* pure-harmonic branch of the oxDNA2 coaxial theta1, DNA2Interaction.cpp:324-363 */
	oxo_f5 stck_phi1, stck_phi2;
	/* Debye-Hueckel, DNA2Interaction.cpp:104-149 */
	double dh_minus_kappa, dh_prefactor, dh_rhigh, dh_rc, dh_b;
	int dh_half_charged_ends;
	double hb_multiplier;
	double rcut;
	/* first-generation oxDNA (interaction_type = DNA): mirrored coaxial theta1 and the phi3 factor, no Debye-Hueckel */
	int v1;
	oxo_f5 cxst_phi3;
	/* 1: the f4 factors through the cubic meshes of the CPU classes DNAInteraction / DNA2Interaction (interaction_type = DNA / DNA2; model.h:410-433;
	 * the oxDNA2 coaxial theta1 stays analytic, DNA2Interaction.h:61-64); 0 (default): analytic, as DNA2_nomesh and the CUDA kernels */
	int mesh;
} oxo_dna2_params;

/* average-sequence oxDNA2 parameters at temperature T (simulation units) and molar salt.
 * seq_dep: 0 = average sequence; 1 = stck_eps[4][4] (already multiplied by the T factor is NOT assumed: raw file
 * values STCK_X_Y) + stck_fact_eps + hb_eps_AT, hb_eps_GC are taken from the arguments. */
void oxo_dna2_params_init(oxo_dna2_params *P, double T, double salt, int dh_half_charged_ends,
		int use_mbf, double mbf_fmax, double mbf_finf);
/* interaction_type = DNA (class DNAInteraction); grooving = the major_minor_grooving key */
void oxo_dna1_params_init(oxo_dna2_params *P, double T, int grooving, int use_mbf, double mbf_fmax, double mbf_finf);
void oxo_dna2_params_seqdep(oxo_dna2_params *P, const double *stck_raw16, double stck_fact_eps, double hb_AT, double hb_GC);

/* particle = -1: every particle.  aux / iaux: see the table in include/oxdna_b200.h (same conventions) */
typedef struct { int type; int particle; int ref; int pbc; double stiff, r0, rate, stiff_rate, F0; double dir[3], pos0[3]; double aux[8]; int iaux; } oxo_ext_force;
enum { OXO_EXT_STRING = 0, OXO_EXT_TRAP = 1, OXO_EXT_MUTUAL = 2, OXO_EXT_LOWDIM = 3, OXO_EXT_REPULSION_PLANE = 4, OXO_EXT_ATTRACTION_PLANE = 5,
	OXO_EXT_SPHERE = 6, OXO_EXT_LJ_WALL = 7, OXO_EXT_TWIST = 8, OXO_EXT_SPHERE_SMOOTH = 9, OXO_EXT_ELLIPSOID = 10,
	OXO_EXT_REPULSION_PLANE_MOVING = 11, OXO_EXT_GENERIC_CENTRAL = 12, OXO_EXT_LJ_CONE = 13, OXO_EXT_COM = 14, OXO_EXT_YUKAWA_SPHERE = 15, OXO_EXT_SPHERE_MOVING = 16, OXO_EXT_META_COM_TRAP = 17 };

/* axes: N x 9 doubles = a1(3) a2(3) a3(3).  pairs: npairs x 2 ints (non-bonded candidates, each unique pair once).
 * Outputs (any may be NULL): force N x 3 (lab), torque_lab N x 3, torque_body N x 3, eterms[OXO_NTERMS] totals,
 * epart[N] per-particle energy (half of each pair energy to each partner). */
void oxo_dna2_forces(const oxo_dna2_params *P, int N, const double *pos, const double *axes, const int *btype,
		const int *n3, const int *n5, const double *box, const int *pairs, long long npairs,
		double *force, double *torque_lab, double *torque_body, double *eterms, double *epart);

/* external forces (src/Forces/{ConstantRateForce,MovingTrap,MutualTrap,LowdimMovingTrap,RepulsionPlane,AttractionPlane,
 * RepulsiveSphere,LJWall}.cpp), added to force (lab frame) */
/* index pool of the COM forces (entry: ref = offset of com_list, iaux = its length, pbc = length of the ref_list that follows) */
void oxo_set_ext_pool(const int *pool);
void oxo_set_ext_grid(const double *grid);
/* metadynamics coordination bias (meta_coordination, LTCoordination): mode 0 hb_cutoff, 1 switching_function, 2 mixed; pairs = n_pairs x 2
 * particle indices.  Adds force and lab-frame torque on every particle of the pairs; returns the (unclamped) coordination. */
typedef struct { int mode; double mixed_weight, hb_energy_cutoff, hb_transition_width, d0, r0; int n; double coord_min, coord_max; int N_grid; const double *grid;
	int n_pairs; const int *pairs; } oxo_coord;
double oxo_meta_coordination(const oxo_coord *C, int N, const double *pos, const double *axes, const int *btype, const double *box,
		double *force, double *torque_lab);
void oxo_ext_forces(int nf, const oxo_ext_force *ef, int N, const double *pos, const double *box, long long step, double *force);

/* Verlet list exactly as src/Lists/Cells.cpp:120-181 + VerletList.cpp:35-66: unique pairs (q<p), not bonded,
 * |min_image|^2 < rv^2 with rv = rcut + 2 skin.  Returns the number of pairs; writes at most max_pairs. */
long long oxo_verlet_pairs(int N, const double *pos, const int *n3, const int *n5, const double *box, double rv,
		int *pairs, long long max_pairs);

/* axes helpers */
void oxo_axes_from_a1a3(int N, const double *a1, const double *a3, double *axes);

/* One or more NVE/thermostat-free velocity-Verlet steps exactly as src/Backends/MD_CPUBackend.cpp:67-218
 * (forces must be valid on entry; they are valid on exit).  The Verlet list is kept in `pairs` and rebuilt
 * when a particle has moved more than skin from list_pos.  Returns number of list rebuilds. */
typedef struct {
	int N;
	double *pos, *axes, *vel, *L, *force, *torque_body, *list_pos;
	const int *btype, *n3, *n5;
	double box[3];
	double dt, skin;
	int *pairs; long long npairs, max_pairs;
	long long step;
	int nf; const oxo_ext_force *ef;
	double U;
} oxo_md;
int oxo_md_steps(const oxo_dna2_params *P, oxo_md *S, int nsteps);
void oxo_md_compute_forces(const oxo_dna2_params *P, oxo_md *S);

/* ------------------------------------------------------------------ oxRNA2 (src/Interactions/RNAInteraction.cpp, RNAInteraction2.cpp, rna_model.h) */
typedef struct {
	/* sites, src/Particles/RNANucleotide.h:27-52: BACK = back[0] a1 + back[1] a2 + back[2] a3; STACK = stack_a1 a1; BASE = base_a1 a1;
	 * STACK_3 / STACK_5 = (a1, a2) coefficients; BBVECTOR_3 / _5 = p3 / p5 on (a1, a2, a3) */
	double back[3], stack_a1, base_a1, stack3[2], stack5[2], p3[3], p5[3];
	double T;
	double fene_eps, fene_r0, fene_delta, fene_delta2;
	int use_mbf;
	double mbf_xmax, mbf_fmax, mbf_finf;
	oxo_excl excl[4]; /* 0: back-back, 1: base-base, 2: base(p)-back(q), 3: back(p)-base(q) */
	double excl_eps;
	oxo_f1 hb, stck;
	oxo_f2 crst, cxst;
	double crst_kfac[5][5]; /* sequence-dependent cross-stacking multiplier [type p][type q], RNAInteraction.cpp:365-372,901-905 */
	oxo_f4 stck_t5, stck_t6, stck_tb1, stck_tb2, hb_t1, hb_t2, hb_t3, hb_t4, hb_t7, hb_t8, crst_t1, crst_t2, crst_t3, crst_t7, crst_t8,
		cxst_t1, cxst_t4, cxst_t5, cxst_t6;
	oxo_f5 stck_phi1, stck_phi2, cxst_phi3, cxst_phi4;
	double dh_minus_kappa, dh_prefactor, dh_rhigh, dh_rc, dh_b;
	int dh_half_charged_ends;
	int average;            /* use_average_seq: G-U wobble pairs exist only when 0 */
	int mismatch_repulsion; /* RNAInteraction2.cpp:43-55,96-101 */
	double mis_eps, mis_shift;
	/* bit mask: reproduce two places where the CPU class's force is NOT the gradient of its energy (the reference's CUDA kernels
	 * use the gradient): bit 0 -- the phi2 stacking term lacks the thetaB1/B2 factors (RNAInteraction.cpp:620); bit 1 -- the mirrored
	 * coaxial theta1 term has the opposite sign (RNAInteraction.cpp:1046 vs :1302 and CUDA_RNA.cuh:896).  3: the CPU class; 0: gradient.
	 * bit 2 (4): the six f4 factors of the hydrogen bonding interpolated on the CPU class's cubic meshes (RNAInteraction.cpp:769-794) instead of
	 * the analytic form of the CUDA kernels: 7 reproduces the CPU class in every term */
	int cpu_quirks;
	double rcut;
} oxo_rna2_params;

void oxo_rna2_params_init(oxo_rna2_params *P, double T, double salt, int dh_half_charged_ends, int use_mbf, double mbf_fmax,
		double mbf_finf, int mismatch_repulsion, double mismatch_strength);
/* sequence-dependent strengths (RNAInteraction.cpp:345-391): stck_raw16 / cross_raw16 = STCK_X_Y / CROSS_X_Y (order A, G, C, T=U) */
void oxo_rna2_params_seqdep(oxo_rna2_params *P, const double *stck_raw16, double st_t_dep, const double *cross_raw16, double hb_AT,
		double hb_GC, double hb_GT);
void oxo_rna2_forces(const oxo_rna2_params *P, int N, const double *pos, const double *axes, const int *btype,
		const int *n3, const int *n5, const double *box, const int *pairs, long long npairs,
		double *force, double *torque_lab, double *torque_body, double *eterms, double *epart);
int oxo_rna2_md_steps(const oxo_rna2_params *P, oxo_md *S, int nsteps);
void oxo_rna2_md_compute_forces(const oxo_rna2_params *P, oxo_md *S);

/* ------------------------------------------------------------------ oxDNA3 (src/Interactions/DNA3Interaction.cpp, class DNA3Interaction_nomesh)
 * tab: OXO_DNA3_NTAB tables of OXO_DNA3_TSIZE = 6 x 5 x 5 x 6 doubles, entry ((n3_2 * 5 + n3_1) * 5 + n5_1) * 6 + n5_2 (5 = no neighbour;
 * src/Utilities/oxdna3_utils.h:18-33), in this order (the member arrays of DNA3Interaction.h:70-113, array index fastest within a group):
 *   0 fene_r0, 1 fene_delta, 2 fene_delta2, 3 mbf_xmax, 4.. excl_s[7], 11.. excl_r[7], 18.. excl_b[7], 25.. excl_rc[7],
 *   32.. F1 {EPS, A, RC, R0, BLOW, BHIGH, RLOW, RHIGH, RCLOW, RCHIGH, SHIFT}[2], 54.. F2 {K, K_SYMM, RC, R0, BLOW, RLOW, RCLOW, BHIGH, RCHIGH, RHIGH}[4],
 *   94.. F4 {A, B, T0, TS, TC}[21], 199.. F5 {A, B, XC, XS}[4] */
enum { OXO3_FENE_R0 = 0, OXO3_FENE_DELTA = 1, OXO3_FENE_DELTA2 = 2, OXO3_MBF_XMAX = 3, OXO3_EXCL_S = 4, OXO3_EXCL_R = 11, OXO3_EXCL_B = 18, OXO3_EXCL_RC = 25,
	OXO3_F1 = 32, OXO3_F2 = 54, OXO3_F4 = 94, OXO3_F5 = 199, OXO_DNA3_NTAB = 215, OXO_DNA3_TSIZE = 900 };
typedef struct {
	const double *tab;
	double fene_eps; int use_mbf; double mbf_fmax, mbf_finf, hb_multiplier;
	double dh_rc, dh_rhigh, dh_prefactor, dh_b, dh_minus_kappa; int dh_half_charged_ends;
	double rcut;
	oxo_f4 cxst_t1, cxst_t4, cxst_t5; double cxst_t1_sa, cxst_t1_sb; /* scalar angular set of the coaxial stacking (DNA2Interaction) */
	double excl_eps;
	double back_a1, back_a2, backref_a1, pos_stack[5], pos_base[5]; /* per-type stacking / base site offsets, src/model.h:15-39 */
	/* 1 (default): the stacking phi1 / phi2 force exactly as the reference writes it (CPU class and CUDA kernel), with the oxDNA2 lever
	 * gamma = 0.74 -- NOT the gradient of the energy for the oxDNA3 stacking site at 0.37; 0: the gradient (finite-difference checks) */
	int ref_form;
	/* 1: coaxial theta4 / theta5 / theta6 through the 6-interval cubic meshes of the CPU class (DNA3Interaction_nomesh inherits the meshed scalar f4 of
	 * DNA2Interaction); 0 (default): analytic f4, as the reference's CUDA kernel */
	int cxst_mesh;
} oxo_dna3_params;
/* scalars: the block written by oxref_dna3_tables (oracle/ref_harness.cpp), also stored in the fixtures */
void oxo_dna3_params_fill(oxo_dna3_params *P, const double *tab, const double *scalars);
void oxo_dna3_forces(const oxo_dna3_params *P, int N, const double *pos, const double *axes, const int *btype,
		const int *n3, const int *n5, const double *box, const int *pairs, long long npairs,
		double *force, double *torque_lab, double *torque_body, double *eterms, double *epart);
int oxo_dna3_md_steps(const oxo_dna3_params *P, oxo_md *S, int nsteps);
void oxo_dna3_md_compute_forces(const oxo_dna3_params *P, oxo_md *S);

/* thermostat parameter derivation (src/Backends/Thermostats/{Brownian,Langevin,Bussi}Thermostat.cpp) */
void oxo_brownian_params(double T, double dt, int newtonian_steps, double pt_in, double diff_coeff, double *pt, double *pr, double *rescale);
void oxo_langevin_params(double T, double dt, double gamma_trans_in, double diff_coeff_in, double *gamma_t, double *gamma_r, double *resc_t, double *resc_r);

#ifdef __cplusplus
}
#endif
#endif
