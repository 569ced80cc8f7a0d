/* oxdna_b200 -- C ABI of the B200-native oxDNA GPU MD step.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, int status returns
 * (0 = ok; oxb_last_error() gives the message), one context = one simulated system on one GPU and one
 * CUDA stream.  Every entry point names the reference interface it replaces (paths relative to the
 * reference tree, lorenzo-rovigatti/oxDNA).
 *
 * All host arrays are in ORIGINAL particle order (the order of the topology file); the device keeps
 * particles in Hilbert-sorted order internally and remaps on the way in and out.
 */
#ifndef OXDNA_B200_H
#define OXDNA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oxb_ctx oxb_ctx;

/* backend_precision (src/Backends/BackendFactory.cpp:55-81).  Both keep an FP64 state and FP32 pair arithmetic.  MIXED (the
 * reference default) additionally evaluates the stiff terms in double -- the FENE distance of strongly stretched bonds and every
 * excluded-volume site pair in range -- so that forces stay within 1e-5 of the FP64 CPU interaction at any box size; FLOAT skips
 * that (1e-4 criterion, ~5 % faster). */
enum { OXB_PRECISION_FLOAT = 0, OXB_PRECISION_MIXED = 1 };
enum { OXB_THERMOSTAT_NONE = 0, OXB_THERMOSTAT_BROWNIAN = 1, OXB_THERMOSTAT_LANGEVIN = 2, OXB_THERMOSTAT_BUSSI = 3 };
enum { OXB_EXT_STRING = 0, OXB_EXT_TRAP = 1, OXB_EXT_MUTUAL_TRAP = 2, OXB_EXT_LOWDIM_TRAP = 3, OXB_EXT_REPULSION_PLANE = 4,
	OXB_EXT_ATTRACTION_PLANE = 5, OXB_EXT_SPHERE = 6, OXB_EXT_LJ_WALL = 7, OXB_EXT_TWIST = 8, OXB_EXT_SPHERE_SMOOTH = 9, OXB_EXT_ELLIPSOID = 10,
	OXB_EXT_REPULSION_PLANE_MOVING = 11, OXB_EXT_GENERIC_CENTRAL = 12, OXB_EXT_LJ_CONE = 13, OXB_EXT_COM = 14, OXB_EXT_YUKAWA_SPHERE = 15,
	OXB_EXT_SPHERE_MOVING = 16, OXB_EXT_META_COM_TRAP = 17, OXB_EXT_META_COORDINATION = 18, OXB_EXT_NTYPES };
enum { OXB_TERM_FENE = 0, OXB_TERM_BEXC, OXB_TERM_STCK, OXB_TERM_NEXC, OXB_TERM_HB, OXB_TERM_CRST, OXB_TERM_CXST, OXB_TERM_DH, OXB_NTERMS };

/* ---- force-field parameters (device constant block).  Replaces the __constant__ upload of
 * src/CUDA/Interactions/CUDADNAInteraction.cu:61-154 (values derived from the CPU DNA2Interaction). */
typedef struct { float a, rc, r0, blow, bhigh, rlow, rhigh, rclow, rchigh; } oxb_f1;
typedef struct { float k, rc, r0, blow, rlow, rclow, bhigh, rhigh, rchigh; } oxb_f2;
typedef struct { float a, b, t0, ts, tc; } oxb_f4;
typedef struct { float a, b, xc, xs; } oxb_f5;
typedef struct { float sigma2, rstar2, b, rc, rc2; } oxb_excl;

enum { OXB_F4_STCK_T4 = 0, OXB_F4_STCK_T5, OXB_F4_HB_T1, OXB_F4_HB_T2, OXB_F4_HB_T4, OXB_F4_HB_T7, OXB_F4_CRST_T1,
	OXB_F4_CRST_T2, OXB_F4_CRST_T4, OXB_F4_CRST_T7, OXB_F4_CXST_T1, OXB_F4_CXST_T4, OXB_F4_CXST_T5, OXB_NF4 };

typedef struct {
	float back_a1, back_a2, stack_a1, base_a1, backref_a1; /* interaction-site offsets along a1 / a2 */
	float fene_eps, fene_r0, fene_delta, fene_delta2;
	int use_mbf;
	float mbf_xmax, mbf_fmax, mbf_finf, mbf_e0; /* mbf_e0 = fene(xmax) - long(xmax), the energy offset of the log tail */
	float excl_eps;
	oxb_excl excl[4]; /* back-back, base-base, base(p)-back(q), back(p)-base(q) */
	oxb_f1 hb, stck;
	float hb_eps[25], hb_shift[25], stck_eps[25], stck_shift[25]; /* [type_n3 * 5 + type_n5] */
	oxb_f2 crst, cxst;
	oxb_f4 f4[OXB_NF4];
	float f4_cmin[OXB_NF4], f4_cmax[OXB_NF4]; /* f4[k](theta) != 0 only if cos(theta) lies in (cmin, cmax): cheap pre-filter */
	float cxst_t1_sa, cxst_t1_sb;
	oxb_f5 phi1, phi2;
	float dh_minus_kappa, dh_prefactor, dh_rhigh, dh_rc, dh_b;
	int dh_half_charged_ends;
	float hb_multiplier;
	float rcut; /* global interaction cutoff on the centre-of-mass distance */
	float rcut_near; /* centre-of-mass distance beyond which only Debye-Hueckel can act */
	/* first-generation oxDNA (interaction_type = DNA / DNA_nomesh, class DNAInteraction): coaxial stacking with the mirrored
	 * f4(2 pi - theta1) and the f5(cos phi3)^2 factor (src/CUDA/Interactions/CUDA_DNA.cuh:612-719 with grooving off), no
	 * Debye-Hueckel */
	int v1;
	oxb_f5 phi3;
} oxb_dna2_params;

/* Host-side derivation of the oxDNA2 parameter block at temperature T (simulation units, K/3000) and molar
 * salt -- what DNA2Interaction::get_settings/init (src/Interactions/DNA2Interaction.cpp:61-149) followed by
 * CUDADNAInteraction::cuda_init computes.  seq-dependent tables may then be overwritten by the caller.
 * `rcut_out` receives the double-precision cutoff used for the Verlet radius. */
int oxb_dna2_params_init(oxb_dna2_params *P, double T, double salt_concentration, int dh_half_charged_ends,
		int use_max_backbone_force, double max_backbone_force, double max_backbone_force_far, double *rcut_out);
/* interaction_type = DNA (src/Interactions/DNAInteraction.cpp:12-228,295-328); grooving = the major_minor_grooving key */
int oxb_dna1_params_init(oxb_dna2_params *P, double T, int major_minor_grooving, int use_max_backbone_force, double max_backbone_force,
		double max_backbone_force_far, double *rcut_out);
/* sequence-dependent stacking / HB strengths (src/Interactions/DNAInteraction.cpp:329-375):
 * stck_raw[4][4] are the STCK_X_Y entries of the parameter file (order A, G, C, T). */
int oxb_dna2_params_seqdep(oxb_dna2_params *P, double T, const double *stck_raw16, double stck_fact_eps, double hb_AT, double hb_GC);

/* ---- oxRNA2 parameter block.  Replaces the `CUDAModel rnamodel` constant upload of
 * src/CUDA/Interactions/CUDARNAInteraction.cu:43-230,278-385 (values from src/Interactions/rna_model.h through the CPU
 * RNA2Interaction).  Same table conventions as the DNA block; every angular factor keeps its own entry because the
 * reference's `external_model` file may set theta2/theta3, theta7/theta8, theta5/theta6 independently. */
enum { OXB_RF4_STCK_T5 = 0, OXB_RF4_STCK_T6, OXB_RF4_STCK_TB1, OXB_RF4_STCK_TB2, OXB_RF4_HB_T1, OXB_RF4_HB_T2, OXB_RF4_HB_T3,
	OXB_RF4_HB_T4, OXB_RF4_HB_T7, OXB_RF4_HB_T8, OXB_RF4_CRST_T1, OXB_RF4_CRST_T2, OXB_RF4_CRST_T3, OXB_RF4_CRST_T7, OXB_RF4_CRST_T8,
	OXB_RF4_CXST_T1, OXB_RF4_CXST_T4, OXB_RF4_CXST_T5, OXB_RF4_CXST_T6, OXB_NRF4 };

typedef struct {
	/* sites (src/Particles/RNANucleotide.h:27-52): BACK = back_a1 a1 + back_a2 a2 + back_a3 a3, STACK = stack_a1 a1,
	 * BASE = base_a1 a1, STACK_3 / STACK_5 on (a1, a2), backbone-direction unit vectors p3 / p5 on (a1, a2, a3) */
	float back_a1, back_a2, back_a3, stack_a1, base_a1;
	float stack3_a1, stack3_a2, stack5_a1, stack5_a2;
	float p3[3], p5[3];
	float fene_eps, fene_r0, fene_delta, fene_delta2;
	int use_mbf;
	float mbf_xmax, mbf_fmax, mbf_finf, mbf_e0;
	float excl_eps;
	oxb_excl excl[4]; /* back-back, base-base, base(p)-back(q), back(p)-base(q) */
	oxb_f1 hb, stck;
	float hb_eps[25], hb_shift[25], stck_eps[25], stck_shift[25]; /* [type_n3 * 5 + type_n5] */
	oxb_f2 crst, cxst;
	float crst_kfac[25]; /* sequence-dependent cross-stacking multiplier [type_p * 5 + type_q] (1 for the average model) */
	oxb_f4 f4[OXB_NRF4];
	float f4_cmin[OXB_NRF4], f4_cmax[OXB_NRF4];
	oxb_f5 phi1, phi2, phi3, phi4;
	float dh_minus_kappa, dh_prefactor, dh_rhigh, dh_rc, dh_b;
	int dh_half_charged_ends;
	int average;            /* use_average_seq: G-U wobble pairs hydrogen-bond only when 0 */
	int mismatch_repulsion; /* src/Interactions/RNAInteraction2.cpp:43-55,96-101 */
	float mis_eps, mis_shift;
	float hb_multiplier;
	float rcut, rcut_near;
} oxb_rna2_params;

/* RNA2Interaction::get_settings/init (src/Interactions/RNAInteraction.cpp:60-397, RNAInteraction2.cpp:31-102) */
int oxb_rna2_params_init(oxb_rna2_params *P, double T, double salt_concentration, int dh_half_charged_ends,
		int use_max_backbone_force, double max_backbone_force, double max_backbone_force_far, int mismatch_repulsion,
		double mismatch_repulsion_strength, double *rcut_out);
/* rna_sequence_dependent_parameters.txt (RNAInteraction.cpp:345-391): STCK_X_Y, ST_T_DEP, CROSS_X_Y, HYDR_A_T, HYDR_C_G,
 * HYDR_G_T; 4 x 4 tables in the order A, G, C, U.  Clears `average`. */
int oxb_rna2_params_seqdep(oxb_rna2_params *P, double T, const double *stck_raw16, double st_t_dep, const double *cross_raw16,
		double hb_AT, double hb_GC, double hb_GT);

/* One external force acting on one particle -- or on every particle when particle = -1 (`particle = all`), in which case
 * the entry is kept once and evaluated per particle (the reference stores 15 union slots per particle).
 *   type                 reference class (src/Forces/)   fields used
 *   STRING               ConstantRateForce               F0, rate, dir; pbc = 1 is the reference's dir_as_centre: the force points from the
 *                                                        particle to the point pos0 (src/CUDA/Backends/CUDA_MD.cuh:114-130)
 *   TRAP                 MovingTrap                      stiff, rate, dir, pos0
 *   MUTUAL_TRAP          MutualTrap                      ref, pbc, stiff, r0, rate, stiff_rate
 *   LOWDIM_TRAP          LowdimMovingTrap                stiff, rate, dir, pos0, iaux = visibility mask (bit 0 x, 1 y, 2 z)
 *   REPULSION_PLANE      RepulsionPlane                  stiff, dir, aux[0] = position, aux[1] = v, aux[2] = end_position
 *   ATTRACTION_PLANE     AttractionPlane                 stiff, dir, aux[0] = position
 *   SPHERE               RepulsiveSphere                 stiff, r0, rate, pos0 = center, aux[0] = r_ext
 *   LJ_WALL              LJWall                          stiff, dir, aux[0] = position, aux[1] = sigma, aux[2] = cutoff, iaux = n
 *   TWIST                ConstantRateTorque              stiff, rate, F0 = base, dir = axis, pos0, aux[0..2] = center, aux[3..5] = mask
 *   SPHERE_SMOOTH        RepulsiveSphereSmooth           stiff, r0, pos0 = center, aux[0] = r_ext, aux[1] = smooth, aux[2] = alpha
 *   ELLIPSOID            RepulsiveEllipsoid              stiff, pos0 = center, aux[0..2] = r_2 (inner), aux[3..5] = r_1 (outer)
 *   REPULSION_PLANE_MOVING RepulsionPlaneMoving          stiff, dir, ref = lowest and iaux = highest original index of the (contiguous) ref_particle range
 *   GENERIC_CENTRAL      GenericCentralForce (gravity)   F0, pos0 = center, aux[0] = inner_cut_off^2, aux[1] = outer_cut_off^2 (0 = none); the
 *                                                        `interpolated` flavour is CPU-only in the reference as well (forces_defs.cuh:292-310)
 *   LJ_CONE              LJCone                          stiff, dir, pos0 = apex, aux[0] = sigma, aux[1] = cutoff, aux[2] = alpha, iaux = n
 *   COM                  COMForce                        ONE entry per force (not per particle): stiff, r0, rate, ref = offset of com_list in
 *                                                        the index pool (oxb_set_ext_index_pool), iaux = its length, ref_list follows it
 *                                                        directly and has pbc entries; particle is ignored
 *   YUKAWA_SPHERE        YukawaSphere                    pos0 = center, r0 = radius, stiff = WCA_epsilon, aux[0] = WCA sigma, aux[1] = WCA cutoff,
 *                                                        aux[2] = debye_length, aux[3] = debye_A, aux[4] = cutoff, iaux = WCA_n
 *   SPHERE_MOVING        RepulsiveSphereMoving           stiff, r0, rate, pos0 = origin, aux[0] = r_ext, aux[1..3] = target, aux[4] = steps
 *   META_COM_TRAP        LTCOMTrap (meta_com_trap)       ONE entry per force: ref = offset of p1a in the index pool, iaux = its length, p2a follows with
 *                                                        pbc entries; aux[0] = xmin, aux[1] = dX, aux[2] = N_grid, aux[3] = mode (1: acts on p1a,
 *                                                        2: on p2a), aux[4] = offset of potential_grid in the grid pool
 *                                                        (oxb_set_ext_grid_pool), aux[5] = PBC
 *   META_COORDINATION    LTCoordination (meta_coordination) ONE entry per force: ref = offset of the hydrogen-bond candidate pairs in the index pool
 *                                                        (p0, q0, p1, q1, ...), iaux = number of pairs; aux[0] = coord_min, aux[1] = d_coord,
 *                                                        aux[2] = N_grid, aux[3] = coordination type (0 hb_cutoff, 1 switching_function, 2 mixed),
 *                                                        aux[4] = offset of potential_grid in the grid pool, aux[5] = mixed_weight,
 *                                                        aux[6] = hb_energy_cutoff, aux[7] = hb_transition_width, r0 = d0, stiff = r0 (of the
 *                                                        switching function), pbc = n, F0 = coord_max.  Adds forces AND lab-frame torques. */
typedef struct {
	int type;      /* OXB_EXT_* */
	int particle;  /* original index, or -1 = all particles */
	int ref;       /* mutual trap partner, original index */
	int pbc;
	double stiff, r0, rate, stiff_rate, F0;
	double dir[3], pos0[3];
	double aux[8];
	int iaux;
} oxb_ext_force;

/* ---- replica batching (examples/OXPY_REMD/remd.py runs one process + one GPU context per temperature replica).
 * A context can hold R independent replicas of one system as ONE batch: particles [r * N/R, (r + 1) * N/R) of the context belong to
 * replica r, every kernel of the step is launched once for all of them (the replica is a grid offset: slots stay replica-contiguous,
 * the cell table and the Hilbert key carry the replica index, so no pair ever crosses replicas).  Replicas share box, topology and
 * every temperature-independent constant; what depends on the temperature lives in a device table of one oxb_replica_consts per
 * replica that the kernels index by replica -- a temperature swap rewrites 2 table rows and touches neither kernel arguments nor
 * captured graphs (the reference re-initialises interaction and thermostat and re-uploads constant memory, remd.py:104-147). */
typedef struct {
	float stck_eps[25], stck_shift[25];                           /* stacking strength f1(eps(T)), DNAInteraction.cpp:329-375 */
	float dh_minus_kappa, dh_prefactor, dh_rhigh, dh_rc, dh_b;    /* Debye-Hueckel, DNA2Interaction.cpp:99-149 */
	float rcut2;                                                  /* squared interaction cutoff of this replica's Hamiltonian */
	float th_a, th_b, th_c, th_d;                                 /* thermostat constants, meaning as in oxb_set_thermostat */
} oxb_replica_consts;

/* sizeof() of the ABI structures as compiled into the library (0: oxb_dna2_params, 1: oxb_rna2_params, 2: oxb_ext_force,
 * 3: oxb_replica_consts), so that foreign-language bindings can assert their mirrors */
int oxb_sizeof(int which);

/* ---- life cycle.  Replaces MD_CUDABackend / CUDAMixedBackend construction + init_cuda
 * (src/CUDA/Backends/CUDABaseBackend.cu:143-242, MD_CUDABackend.cu:674-747). */
int oxb_create(oxb_ctx **out, int device, int N, int precision);
void oxb_destroy(oxb_ctx *ctx);
const char *oxb_last_error(const oxb_ctx *ctx);
/* use an existing CUDA stream (cudaStream_t passed as void*); default: a private non-blocking stream */
int oxb_set_stream(oxb_ctx *ctx, void *cuda_stream);

int oxb_set_box(oxb_ctx *ctx, const double box[3]);                                           /* CUDABox, src/CUDA/cuda_utils/CUDABox.h */
int oxb_set_topology(oxb_ctx *ctx, const int *btype, const int *n3, const int *n5, const int *strand);
int oxb_set_model_dna2(oxb_ctx *ctx, const oxb_dna2_params *P, double rcut);                  /* CUDADNAInteraction::cuda_init */
int oxb_set_model_rna2(oxb_ctx *ctx, const oxb_rna2_params *P, double rcut);                  /* CUDARNAInteraction::cuda_init */

/* ---- oxDNA3 (interaction_type = DNA3; src/Interactions/DNA3Interaction.{h,cpp}, src/CUDA/Interactions/CUDADNA3Interaction.cu, CUDA_DNA3.cuh).
 * Every parameter of the oxDNA2 functional forms is a table over the tetramer (n3_2, n3_1, n5_1, n5_2): 6 x 5 x 5 x 6 = 900 entries,
 * entry ((n3_2 * 5 + n3_1) * 5 + n5_1) * 6 + n5_2, 5 = no such neighbour (src/Utilities/oxdna3_utils.h:18-33).  The tables are INPUT: the host
 * hands over what DNA3Interaction::init leaves in the class -- exactly the arrays CUDADNA3Interaction::cuda_init uploads
 * (CUDADNA3Interaction.cu:46-150) -- as OXB_DNA3_NTAB tables of OXB_DNA3_TSIZE doubles in this order (array index fastest within a group):
 *   0 _fene_r0_SD, 1 _fene_delta_SD, 2 _fene_delta2_SD, 3 _mbf_xmax_SD, 4.. _excl_s[7], 11.. _excl_r[7], 18.. _excl_b[7], 25.. _excl_rc[7],
 *   32.. F1_SD_{EPS, A, RC, R0, BLOW, BHIGH, RLOW, RHIGH, RCLOW, RCHIGH, SHIFT}[2],
 *   54.. F2_SD_{K, K_SYMM, RC, R0, BLOW, RLOW, RCLOW, BHIGH, RCHIGH, RHIGH}[4], 94.. F4_SD_THETA_{A, B, T0, TS, TC}[21], 199.. F5_SD_PHI_{A, B, XC, XS}[4]
 * The library repacks them into per-tetramer records (csrc/dna3_model.cuh).  use_edge = 0: one deterministic particle-centric kernel in
 * cost-split passes; use_edge = 1: the staged edge pipeline (csrc/forces.cu, k3_*).  Replica batching is not available for this interaction. */
enum { OXB_DNA3_FENE_R0 = 0, OXB_DNA3_FENE_DELTA = 1, OXB_DNA3_FENE_DELTA2 = 2, OXB_DNA3_MBF_XMAX = 3, OXB_DNA3_EXCL_S = 4, OXB_DNA3_EXCL_R = 11,
	OXB_DNA3_EXCL_B = 18, OXB_DNA3_EXCL_RC = 25, OXB_DNA3_F1 = 32, OXB_DNA3_F2 = 54, OXB_DNA3_F4 = 94, OXB_DNA3_F5 = 199, OXB_DNA3_NTAB = 215, OXB_DNA3_TSIZE = 900 };
typedef struct {
	double fene_eps;                 /* _fene_eps */
	double use_mbf, mbf_fmax, mbf_finf; /* max_backbone_force (non-zero = on), DNAInteraction.h:35-38 */
	double hb_multiplier;
	double dh_rc, dh_rhigh, dh_prefactor, dh_b, dh_minus_kappa, dh_half_charged_ends; /* DNA2Interaction.h:32-40 */
	double rcut;                     /* get_rcut() */
	double cxst_t1[5], cxst_t4[5], cxst_t5[5]; /* {A, B, T0, TS, TC} of the scalar F4_THETA_* at CXST_F4_THETA1 / 4 / 5: coaxial stacking keeps the oxDNA2 angular set */
	double cxst_t1_sa, cxst_t1_sb;   /* F4_THETA_SA / SB [CXST_F4_THETA1], DNA2Interaction.h:111-112 */
} oxb_dna3_scalars;
int oxb_set_model_dna3(oxb_ctx *ctx, const double *tables, const oxb_dna3_scalars *S);       /* CUDADNA3Interaction::cuda_init */

/* Replica batching: declare that the N particles are n_replicas copies of one N / n_replicas-particle system (topology, state and
 * external forces are given for all N particles, replica after replica; bonds must not cross replicas).  The model set with
 * oxb_set_model_* fixes the list radii and must be the one of the HOTTEST temperature any replica will visit (largest Debye-Hueckel
 * range); oxb_set_replica_consts installs the per-replica temperature-dependent constants (n = n_replicas rows; oxb_replica_consts_dna2 /
 * _rna2 fill a row from a parameter block and the thermostat constants a..d of oxb_set_thermostat at that temperature).
 * oxb_replica_energies: potential energy of every replica for the current table (one force pass if the forces are not current, then
 * a segmented reduction), what OxpyManager::system_energy() returns per process in the reference's REMD driver.
 * Not available with the Bussi thermostat (one global kinetic energy) and oxb_energy_split. */
int oxb_set_replicas(oxb_ctx *ctx, int n_replicas);
int oxb_set_replica_consts(oxb_ctx *ctx, int n, const oxb_replica_consts *rows);
int oxb_replica_energies(oxb_ctx *ctx, double *U);
void oxb_replica_consts_dna2(const oxb_dna2_params *P, double th_a, double th_b, double th_c, double th_d, oxb_replica_consts *out);
void oxb_replica_consts_rna2(const oxb_rna2_params *P, double th_a, double th_b, double th_c, double th_d, oxb_replica_consts *out);
/* CUDASimpleVerletList::get_settings/init (src/CUDA/Lists/CUDASimpleVerletList.cu:47-56,165-202) + CUDA_sort_every, use_edge */
int oxb_set_lists(oxb_ctx *ctx, double verlet_skin, int use_edge, int sort_every, double max_density_multiplier);
int oxb_set_dt(oxb_ctx *ctx, double dt);
/* CUDAThermostatFactory + CUDA{Brownian,Langevin,Bussi}Thermostat (src/CUDA/Thermostats/); already-derived parameters:
 *  brownian: a = pt, b = pr, c = rescale factor sqrt(T), every = newtonian_steps
 *  langevin: a = gamma_trans, b = gamma_rot, c = rescale_trans, d = rescale_rot, every = 1
 *  bussi:    a = T, b = exp(-newtonian_steps / tau), every = newtonian_steps */
int oxb_set_thermostat(oxb_ctx *ctx, int type, int every, double a, double b, double c, double d, unsigned long long seed);
/* MD_CUDABackend::_apply_external_forces_changes (src/CUDA/Backends/MD_CUDABackend.cu:108-229) */
int oxb_set_ext_forces(oxb_ctx *ctx, int n, const oxb_ext_force *forces);
/* particle-index lists (original indices) referenced by OXB_EXT_COM entries: COMForce::_com_list / _ref_list
 * (src/Forces/COMForce.cpp:31-44; the reference uploads them per force, src/CUDA/Forces/forces_defs.cuh:367-393).  Call before
 * oxb_set_ext_forces. */
int oxb_set_ext_index_pool(oxb_ctx *ctx, int n, const int *indices);
/* tabulated bias potentials referenced by OXB_EXT_META_COM_TRAP entries: LTCOMTrap::potential_grid (src/Forces/Metadynamics/LTCOMTrap.cpp:36-41;
 * uploaded per force by the reference, src/CUDA/Forces/metad_forces.cuh:33-66).  Call before oxb_set_ext_forces. */
int oxb_set_ext_grid_pool(oxb_ctx *ctx, int n, const double *values);

/* ---- state marshalling.  Replaces apply_changes_to_simulation_data / apply_simulation_data_changes
 * (src/CUDA/Backends/MD_CUDABackend.cu:231-394).  pos, a1, a3, vel, L: N x 3 doubles, original order. */
int oxb_set_state(oxb_ctx *ctx, const double *pos, const double *a1, const double *a3, const double *vel, const double *L);
int oxb_get_state(oxb_ctx *ctx, double *pos, double *a1, double *a3, double *vel, double *L);
/* trajectory / last_conf frame in the reference's text format (docs/source/configurations.md:14-34; Configuration::_headers and
 * ::_particle, src/Observables/Configurations/Configuration.cpp:85-140) straight from the device state, without the BaseParticle
 * round trip and CPU energy evaluation of SimBackend::print_conf: "t = step / b = box / E = Etot U K (per particle)" + one line
 * per particle (original order): pos a1 a3 [vel L].  append != 0 adds a frame to an existing trajectory file. */
int oxb_write_conf(oxb_ctx *ctx, const char *path, int append, int print_momenta);
/* the same frame in the reference's binary format (BinaryConfiguration::_headers / _configuration,
 * src/Observables/Configurations/BinaryConfiguration.cpp:20-92): step, rng state (3 unsigned shorts, zeros if NULL), box, E U K per
 * particle, then per particle pos, pos_shift (3 ints per particle, zeros if NULL), a1, a2, a3, vel, L as doubles */
int oxb_write_conf_binary(oxb_ctx *ctx, const char *path, int append, const unsigned short rng_state[3], const int *pos_shift);
int oxb_set_step(oxb_ctx *ctx, long long step);
long long oxb_get_step(const oxb_ctx *ctx);

/* ---- the individual operators of the step (also reachable one by one for parity tests) */
int oxb_sort(oxb_ctx *ctx);                 /* MD_CUDABackend::_sort_particles, src/CUDA/CUDA_sort.cu */
int oxb_update_lists(oxb_ctx *ctx);         /* CUDASimpleVerletList::update */
int oxb_compute_forces(oxb_ctx *ctx);       /* set_external_forces + CUDADNAInteraction::compute_forces */
int oxb_first_step(oxb_ctx *ctx);           /* first_step[_mixed] */
int oxb_second_step(oxb_ctx *ctx);          /* second_step[_mixed] */
int oxb_thermostat(oxb_ctx *ctx);           /* CUDABaseThermostat::apply_cuda at the current step */

/* ---- the hot loop: n calls of MD_CUDABackend::sim_step (src/CUDA/Backends/MD_CUDABackend.cu:567-619), device-resident,
 * no host synchronisation inside; forces must be valid on entry (oxb_set_state makes them so). */
int oxb_run(oxb_ctx *ctx, long long n_steps);
int oxb_synchronize(oxb_ctx *ctx);

/* ---- read-backs (original order; any pointer may be NULL).  force: lab frame; torque: body frame as the reference
 * stores it; torque_lab: lab frame; energy: per-particle sum of pair energies (system U = sum/2); hb_energy likewise. */
int oxb_get_forces(oxb_ctx *ctx, double *force, double *torque_body, double *torque_lab, double *energy, double *hb_energy);
/* potential energy U and kinetic energy K of the whole system (GpuUtils::sum_c_number4_to_double_on_GPU, CUDA_print_energy) */
int oxb_energy(oxb_ctx *ctx, double *U, double *K);

/* MC barostat: MD_CUDABackend::_apply_barostat (src/CUDA/Backends/MD_CUDABackend.cu:451-516) with its kernels compute_molecular_coms /
 * rescale_molecular_positions / rescale_positions (src/CUDA/Backends/CUDA_MD.cuh:62-95).  The host draws the new box sides
 * (L + delta_L (drand48() - 0.5), isotropic or per axis) and the acceptance number u, as the reference does.
 *   oxb_barostat_move: U(old) -> rescale (molecular: every strand is translated with its centre of mass; atomic: every position is
 *     scaled) -> rebuild lists -> U(new) -> accept iff exp(-(dE + P dV - N_objs T ln(V'/V)) / T) > u, N_objs = strands or particles;
 *     a rejected move restores positions and box EXACTLY (device snapshot; the reference applies the inverse scaling in FP32).
 *   oxb_barostat_trial / _accept / _reject: the same move in pieces, for callers that evaluate their own acceptance rule.
 * Valid between completed steps (not after oxb_first_step).  Works with use_edge = 1 as well (the reference forbids that
 * combination, MD_CUDABackend.cu:629-631). */
int oxb_barostat_move(oxb_ctx *ctx, const double new_box[3], int molecular, double P, double T, double u, int *accepted, double *dE);
int oxb_barostat_trial(oxb_ctx *ctx, const double new_box[3], int molecular);
int oxb_barostat_accept(oxb_ctx *ctx);
int oxb_barostat_reject(oxb_ctx *ctx);
/* SimBackend::fix_diffusion (src/Backends/SimBackend.cpp:786-882; CubicBox::shift_particle, src/Boxes/CubicBox.cpp:70-77) on the device:
 * every strand is translated by whole box sides so that its centre of mass lies in [0, L), quaternions are renormalised.  Minimum-image
 * separations, lists and forces are unaffected (no energy check needed, unlike the matrix re-orthonormalisation of the reference).
 * shifts (may be NULL): 3 ints per particle (original order) = floor(com / L), what the reference adds to BaseParticle::_pos_shift.
 * As in the reference's CUDA backend, forces that read absolute positions (trap, twist, planes) see the translated coordinates. */
int oxb_fix_diffusion(oxb_ctx *ctx, int *shifts);
/* How the host thread waits for the device at the end of a batch of steps inside oxb_run: 0 (default) spins in the driver (lowest
 * latency: one system per GPU), 1 sleeps on a blocking-sync event (replica ensembles that run more host threads than the box has
 * cores: 64 replicas on 8 GPUs are 64 driver threads).  No counterpart in the reference (one process per replica, examples/OXPY_REMD). */
int oxb_set_host_wait(oxb_ctx *ctx, int blocking);
/* current box sides (they change under the barostat) */
int oxb_get_box(oxb_ctx *ctx, double box[3]);
/* potential energy of the whole system split into the reference's terms, terms[OXB_NTERMS] in the order of OXB_TERM_*
 * (BaseInteraction::get_system_energy_split, src/Interactions/BaseInteraction.cpp:61-90; the `potential_energy` observable
 * with split = true, src/Observables/PotentialEnergy.cpp), evaluated on the device for the current positions */
int oxb_energy_split(oxb_ctx *ctx, double *terms);
/* unique Verlet pairs (i < j, original ids); call with pairs = NULL to get the count */
int oxb_get_pairs(oxb_ctx *ctx, int *pairs, long long max_pairs, long long *n_pairs);
/* number of list rebuilds / sorts so far; current neighbour-matrix capacity; overflow flags (0 = ok) */
int oxb_get_stats(oxb_ctx *ctx, long long *n_list_updates, long long *n_sorts, int *max_neigh, int *error_flags);
/* device pointers for zero-copy consumers (CUDABaseInteraction plugin seam): float4 positions with packed
 * (btype<<22 | index) in .w, float4 quaternions, column-major neighbour matrix, neighbour counts, and the int2 list of
 * unique pairs closer than rcut_near + 2 skin at the last rebuild (the pairs that can feel more than Debye-Hueckel; .x = from slot,
 * .y = to slot | class << 24, the class being the list builder's bit mask of the site-pair families that can come into range before the
 * next rebuild; the list comes in three segments by class group -- n_edges[0] = total, n_edges[1..3] = the segment lengths -- each grouped
 * by `from` in slot order).  Asking for matrix_neighs / number_neighs switches the context to full builds (both directions of every pair, plain
 * slots), which the edge pipeline's own half-shell builds do not produce. */
int oxb_device_views(oxb_ctx *ctx, void **poss_f4, void **orientations_f4, void **matrix_neighs, void **number_neighs,
		void **edge_list, void **n_edges);
/* ---- plugin seam: a third-party interaction replaces the built-in force field.  Replaces the call
 * CUDABaseInteraction::compute_forces(lists, d_poss, d_orientations, d_forces, d_torques, d_bonds, d_box) that the reference's backend
 * makes on whatever CUDAInteractionFactory returned -- for unknown interaction types the class that PluginManager finds in
 * CUDA<type>.so through make_CUDA<type> (src/CUDA/Interactions/CUDAInteractionFactory.cu:44-51, src/PluginManagement/PluginManager.cpp:89-180,
 * src/CUDA/Interactions/CUDABaseInteraction.h:60).  The callback is invoked on the host wherever the built-in force pass would be launched
 * (first forces, every MD step inside oxb_run, energy read-backs); it must enqueue its kernels on views->stream and return 0.  Arrays are
 * in the reference's layouts, indexed by slot (the Hilbert order of the last re-sort): poss float4 (absolute position, .w = btype << 22 |
 * original index, MD_CUDABackend.cu:243-254), orientations float4 quaternions (GPU_quat), matrix_neighs column-major
 * (matrix_neighs[k * stride + i], k < number_neighs[i], both directions of every pair, bonded neighbours excluded), bonds int2 (n3, n5
 * slots, -1 = none).  forces / torques are float4 accumulators, zero on entry: x, y, z in the LAB frame (the integrator rotates the
 * torque into the body frame; the reference's kernels do that themselves), forces .w = the particle's share of the potential energy
 * summed from both ends of every pair (U = sum / 2, as in the reference), torques .w = hydrogen-bonding energy or 0.
 * Setting a callback defines the interaction: rcut fixes the Verlet radius, no built-in model is needed; use_edge and captured graphs
 * are not available with it.  fn = NULL removes it. */
typedef struct oxb_force_views {
	int N, stride;
	const void *poss, *orientations;
	const int *matrix_neighs, *number_neighs;
	const void *bonds;
	void *forces, *torques;
	double box[3];
	long long step;
	void *stream; /* cudaStream_t */
} oxb_force_views;
typedef int (*oxb_force_callback)(void *user, const oxb_force_views *views);
int oxb_set_force_callback(oxb_ctx *ctx, oxb_force_callback fn, void *user, double rcut);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches claim) */
long long oxb_launch_count(const oxb_ctx *ctx);

/* ---- per-kernel timing hooks for the benchmark: time `reps` back-to-back launches of one operator with CUDA events on
 * the context's stream; returns milliseconds per launch.  which: 0 = non-bonded+bonded force pass, 1 = integrate (first
 * step), 2 = list rebuild, 3 = sort */
int oxb_time_kernel(oxb_ctx *ctx, int which, int reps, float *ms_per_launch);
/* Timeline of the hot loop measured INSIDE oxb_run (graph-launched batches included): while enabled, the first thread of the kernel
 * that opens each phase of the step stamps %globaltimer on the device and the time since the previous stamp is charged to the phase
 * that was open, so the phases add up to the device time of the run (launch gaps and host-synchronisation bubbles included).
 * Phases (OXB_PROF_*): 0 other, 1 force pass, 2 integrate, 3 halted launches + host wait before a rebuild, 4 Hilbert sort, 5 list build,
 * 6 gap between the start of a batch (k_batch_begin) and its first force kernel (launch latency of the batch), 7 the gather pass of the
 * re-sort (4 = keys + radix sort + inverse permutation), 8 edge-list scan + fill (5 = cell binning + neighbour scan).
 * Replaces the reference's TimingManager around sim_step (src/Utilities/Timings.cpp, MD_CUDABackend.cu:567-619), which needs a
 * cudaDeviceSynchronize per timer.  oxb_set_profile zeroes the accumulators; oxb_get_profile returns milliseconds and the number of
 * times each phase was entered (OXB_PROF_NPHASES values each). */
#define OXB_PROF_NPHASES 9
int oxb_set_profile(oxb_ctx *ctx, int enable);
int oxb_get_profile(oxb_ctx *ctx, double *ms, long long *entries);

#ifdef __cplusplus
}
#endif
#endif
