// Internal launch interface between the context (context.cu) and the kernel translation units.
#pragma once

#include "common.cuh"

// device-side control words (one int array per context)
enum : int {
	OXB_FLAG_UNUSED0 = 0,
	OXB_FLAG_ERROR = 1,      // OXB_ERR_* bits
	OXB_FLAG_STEPS_DONE = 2, // full steps completed inside the current batch
	OXB_FLAG_PENDING = 3,    // 1 = positions of step `STEPS_DONE` are already advanced, forces not yet computed
	OXB_FLAG_MAX_NEIGH_SEEN = 4,
	// Two "halt" words, [OXB_FLAG_COUNT + 0/1].  The integrator launch with running index e reads word (e & 1) and may set
	// word ((e + 1) & 1) when a particle has moved further than the skin; every kernel launched after it reads that word
	// and turns into a no-op, so a speculative batch of launches stops at the step that needs a list rebuild without any
	// host synchronisation inside the batch.  No kernel ever reads a word that it can write itself.
	OXB_FLAG_COUNT = 8,
	OXB_FLAG_PROF_ON = 12,   // 1 = the first thread of the step's kernels stamps %globaltimer into the profile area (oxb_set_profile)
	OXB_FLAG_WORDS = 16,     // words copied back by read_flags
	OXB_PROF_OFFSET = 32,    // profile area (unsigned long long words) starts at this int offset
	OXB_FLAG_ALLOC = 128,
};

// Device-side timeline of the hot loop: every kernel that opens a phase of the step has its first thread stamp %globaltimer; the time
// since the previous stamp is charged to the phase that was open.  Works inside graph-launched batches and costs one thread a few
// global accesses, so the decomposition is measured IN the timed region (bench.py roofline), launch gaps included: the phases sum to
// the device time line of the run.  Layout (unsigned long long): [0] last stamp, [1] open phase, [2 + p] ns in phase p, [2 + NPHASE + p] entries into p.
enum { OXB_PROF_OTHER = 0, OXB_PROF_FORCE = 1, OXB_PROF_INTEG = 2, OXB_PROF_WAIT = 3, OXB_PROF_SORT = 4, OXB_PROF_BUILD = 5, OXB_PROF_GAP = 6, OXB_PROF_PERMUTE = 7,
	OXB_PROF_EDGES = 8, OXB_PROF_NPHASE = 9 };
#ifdef __CUDACC__
__device__ __forceinline__ void prof_mark(int *flags, int phase, bool reset = false) {
	if(flags[OXB_FLAG_PROF_ON] == 0) return;
	unsigned long long *prof = reinterpret_cast<unsigned long long *>(flags + OXB_PROF_OFFSET);
	// the kernels of one force pass start concurrently on forked streams: whichever arrives first opens the phase (atomic exchange of the
	// open-phase word), the others find it open and leave; the bookkeeping below then has a single writer
	const unsigned long long open = atomicExch(prof + 1, (unsigned long long) phase);
	if(!reset && open == (unsigned long long) phase) return;
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	if(!reset && prof[0] != 0ull) prof[2 + open] += t - prof[0];
	prof[2 + OXB_PROF_NPHASE + phase] += 1ull;
	prof[0] = t;
}
#endif

struct DevExtForce {
	int type, particle, ref, pbc;
	float stiff, r0, rate, stiff_rate, F0;
	float dir[3];
	double pos0[3];
	float aux[8];
	int iaux;
	double daux;  // SPHERE_MOVING: number of steps origin -> target
	double r0d;   // r0 at full precision (radii of the steep walls)
};

struct ThermostatCfg {
	int type;
	int every;
	float a, b, c, d; // see oxb_set_thermostat
	unsigned long long seed;
};

// sums for the Bussi thermostat and for kinetic-energy read-back (doubles, atomically accumulated per block)
struct KinSums {
	double vx, vy, vz; // sum of velocities
	double v2;         // sum |v|^2
	double L2;         // sum |L|^2
	double factor_t, factor_r; // Bussi rescale factors computed on device
	double K_t, K_r;           // Bussi target kinetic energies (state)
};

// phases: bit 0 = second half-kick of step `step`, bit 1 = thermostat of step `step`, bit 2 = first half-kick + drift of step+1
enum { OXB_PH_SECOND = 1, OXB_PH_THERMO = 2, OXB_PH_FIRST = 4, OXB_PH_BUSSI_SUMS = 8, OXB_PH_BUSSI_APPLY = 16, OXB_PH_COUNT_STEP = 32 };

// ---- oxDNA3 (dna3_model.cuh, forces_dna3.cu): packed parameter records, see dna3_model.cuh for the layout
enum { OXB3_REC_BONDED = 48, OXB3_REC_CRST = 32, OXB3_REC_CXST = 12, OXB3_REC_HB = 32, OXB3_REC_NEXCL = 16 };

// kernel argument (by value): device pointers to the packed records + the scalars of the model
struct oxb_dna3_dev {
	const float4 *bonded; // 900 x 12 float4: fene {r0, delta2, xmax, e0} | excl 4, 5, 6 {sigma2, rstar2, b, rc} | stacking f1 (12) | f4 theta4, theta5 (12) | f5 phi1, phi2 (8)
	const float4 *crst;   // 2 x 900 x 8 float4 (3'3' diagonal, then 5'5'): f2 (12) | f4 theta1, theta2, theta4, theta7 (20)
	const float4 *cxst;   // 900 x 3 float4: f2 with K, K_SYMM in slot 9
	const float4 *hb;     // 25 x 8 float4 [type q * 5 + type p]: f1 (12) | f4 theta1, theta2, theta4, theta7 (20)
	const float4 *nexcl;  // 25 x 4 float4 [type q * 5 + type p]: excl 0..3 {sigma2, rstar2, b, rc}
	const int *tcode;     // per ORIGINAL id: type | n3 type << 3 | n5 type << 6 | (btype == 4) << 9   (5 = no neighbour)
	float fene_eps, mbf_fmax, mbf_finf, hb_multiplier, excl_eps;
	int use_mbf;
	float dh_minus_kappa, dh_prefactor, dh_rhigh, dh_rc, dh_b;
	int dh_half_charged_ends;
	float rcut2;
	float r2_excl_max, r2_base_max, r2_stack_max; // squares of: longest excluded-volume range + both levers; longest HB / cross-stacking range; longest coaxial range
	float range_bb, range_eb, range_bk; // longest non-bonded excluded-volume ranges: backbone-backbone, base-base, base-backbone (list classification)
	float r2_near_max;                            // square of the largest centre-centre distance at which any term but Debye-Hueckel can act
	oxb_f4 cxst_t1, cxst_t4, cxst_t5;
	float cxst_t1_sa, cxst_t1_sb;
	float back_a1, back_a2, backref_a1, gamma;
	float pos_stack[5], pos_base[5];
};

namespace oxb {

// the force field of a context: exactly one of the two blocks is set
struct ModelRef {
	const oxb_dna2_params *dna;
	const oxb_rna2_params *rna;
	const oxb_dna3_dev *dna3; // oxDNA3 (forces_dna3.cu: particle-centric pass; forces.cu k3_*: staged edge pipeline)
};

// ---- forces_dna3.cu
void launch_forces_dna3(cudaStream_t s, const oxb_dna3_dev &M, BoxF box, int N, const int4 *ipos, const int4 *iback, const float4 *axf,
		const double4 *posd, const double4 *quatd, const int2 *bonds, const int *nbr, const int *nnbr, int stride, float4 *F, float4 *T, int *flags, int hw);
void launch_energy_split_dna3(cudaStream_t s, const oxb_dna3_dev &M, BoxF box, int N, const int4 *ipos, const float4 *axf, const int2 *bonds,
		const int *nbr, const int *nnbr, int stride, double *out);

// ---- forces.cu
void launch_forces_particle(cudaStream_t s, const ModelRef &M, BoxF box, int N, const int4 *ipos, const int4 *iback, const float4 *axf,
		const double4 *posd, const double4 *quatd, const int2 *bonds,
		const int *nbr, const int *nnbr, int stride, float4 *F, float4 *T, const oxb_replica_consts *rep, int n_per, int *flags, int hw);
struct EdgeArgs {
	int N;
	const oxb_replica_consts *rep; // replica batching: one row per replica (null: single system), n_per particles per replica
	int n_per;
	const int4 *ipos, *iback;
	const float4 *axf;
	const double4 *posd, *quatd; // FP64 state: read only where an excluded-volume term is active (ExclRefine)
	const int2 *bonds, *edges; // near edges
	const int *n_edges;
	const int4 *edge_cnt;    // N + 1 entries, .x = first near edge of every `from` slot when the list is one segment (the tile variant of the near-edge kernel)
	int near_tile;           // 1: one block of the near-edge kernel per tile of 128 `from` slots, staged in shared memory (n_seg = tiles)
	const int *dh_nbr, *dh_nnbr; // Debye-Hueckel neighbour matrix, column-major, stride N
	float4 *F, *T, *Fb;
	// work lists segmented by producer block (n_seg blocks): hydrogen-bonding pairs, coaxial-stacking pairs, cross-stacking-only
	// pairs; seg_counts[l * n_seg + b] = length of block b's segment of list l
	int2 *hb_list, *cx_list, *cr_list;
	int *seg_counts;
	int n_seg, hb_seg, cx_seg, cr_seg;
	int dh_half;  // the Debye-Hueckel matrix holds every pair once: the kernel adds the partner's share atomically, Fb must be zero on entry
	int hb_split; // consumer blocks per segment of the hydrogen-bonding / cross-stacking list
	// pairs with an excluded-volume site pair in range, evaluated in double by k_excl_fix: near edges (p, q, mask, -) segmented by
	// producer block like the lists above (ex_counts[b] entries in block b's segment of ex_seg), bonds as one mask per particle
	int4 *ex_list;
	int *ex_counts, *ex_bonded;
	int ex_seg;
	int fold;   // 1: the coaxial-stacking pairs and the parked excluded-volume pairs are evaluated in the tails of k_edge_near / k_bonded (no stage 3 / 6 launches)
	int refine; // 1 (backend_precision = mixed): FENE and excluded volume in double; 0 (float): FP32 pair arithmetic throughout
};
void launch_edge_stage(cudaStream_t s, int which, const ModelRef &M, BoxF box, const EdgeArgs &a, int *flags, int hw);
void launch_energy_split(cudaStream_t s, const ModelRef &M, BoxF box, int N, const int4 *ipos, const float4 *axf, const int2 *bonds, const int *nbr,
		const int *nnbr, int stride, double *out);
void launch_ext_forces(cudaStream_t s, int n, const DevExtForce *ef, const int *slot_of, const int4 *ipos, const double4 *posd, BoxF box,
		long long step, const long long *cur_step, float4 *F, const int *flags, int hw);
// entries that act on every particle (particle = all): one thread per particle, no atomics
void launch_ext_forces_all(cudaStream_t s, int N, int n_all, const DevExtForce *ef_all, const int *slot_of, const int4 *ipos, const double4 *posd, BoxF box,
		long long step, const long long *cur_step, float4 *F, const int *flags, int hw);
// COM forces: one block per entry; `pool` holds the com_list / ref_list original indices
void launch_ext_com(cudaStream_t s, int n, const DevExtForce *ef_com, const int *pool, const float *grid, const int *slot_of, const double4 *posd,
		const double4 *quatd, const int4 *ipos, const double *box, long long step, const long long *cur_step, float4 *F, float4 *T, const int *flags, int hw);

// ---- integrate.cu
struct IntegrateArgs {
	int N;
	double dt;
	double box_inv[3];
	float skin2;        // (skin - quantisation error of the packed references)^2
	BoxF box;
	double4 *posd, *veld, *Ld, *quatd;
	int4 *ipos;
	float4 *axf;
	float4 *F, *T, *Fb; // lab-frame force / torque accumulators (zeroed by the first-half phase once consumed)
	int zero_Fb;        // Fb is an accumulator too (dh_half)
	int4 *iback;        // fixed-point backbone-site position, .w bit 0 = strand end
	float back_a1, back_a2, back_a3, base_a1;
	int *flags;
	KinSums *sums;
	ThermostatCfg th;
	const oxb_replica_consts *rep; // replica batching: per-replica thermostat constants (null: th applies to every particle)
	int n_per;
	long long step;      // step index of the thermostat application, or < 0: read it from cur_step (graph-launched batches)
	long long *cur_step; // two device words, see k_integrate
};
void launch_integrate_epoch(cudaStream_t s, const IntegrateArgs &a, int phases, int epoch);
void launch_bussi_update_epoch(cudaStream_t s, KinSums *sums, int N, ThermostatCfg th, long long step, const int *flags, int epoch);
void launch_clear_sums(cudaStream_t s, KinSums *sums, const int *flags, int epoch);
void launch_kinetic_sums(cudaStream_t s, int N, const double4 *veld, const double4 *Ld, KinSums *sums);
// ---- marshal.cu: state marshalling on the device (flat N x 3 double arrays in the order of the original particle ids)
struct MarshalArgs {
	int N;
	const double *pos, *a1, *a3, *vel, *L; // device staging; vel / L may be null (zero momenta)
	const int4 *topo;                      // btype, n3, n5, strand per original id
	double box_inv[3];
	float back_a1, back_a2, back_a3;
	double4 *posd, *veld, *Ld, *quatd;
	int4 *ipos, *iback;
	float4 *axf;
	int2 *bonds;
	int *slot_of;
	int *err; // smallest particle id with a null orientation vector (INT_MAX = none)
};
void launch_state_in(cudaStream_t s, const MarshalArgs &a);
void launch_state_out(cudaStream_t s, int N, const int4 *ipos, const double4 *posd, const double4 *veld, const double4 *Ld, const double4 *quatd,
		double *pos, double *a1, double *a3, double *vel, double *L);

// MC barostat: molecular centres of mass (FP64, one atomicAdd triple per particle) and position rescaling + fixed-point re-encode
struct RescaleArgs {
	int N, molecular;
	double f[3];        // molecular: shift factor (new/old - 1); atomic: ratio new/old
	double box_inv[3];  // of the NEW box
	double4 *posd;
	const double4 *quatd;
	int4 *ipos, *iback;
	const int *mol_of;  // molecule of an original particle id
	const double *coms; // 3 doubles per molecule
	double4 *backup;         // if set: positions before the move, indexed by original id
	const double4 *restore;  // if set: positions are taken from here (rejected move) instead of being rescaled
	float back_a1, back_a2, back_a3;
};
void launch_mol_coms(cudaStream_t s, int N, int n_mol, const int4 *ipos, const int *mol_of, const double *inv_size, const double4 *posd, double *coms);
void launch_rescale_positions(cudaStream_t s, const RescaleArgs &a);
// fix_diffusion: strands translated by whole box sides back into the box (coms from launch_mol_coms), quaternions renormalised
void launch_fix_diffusion(cudaStream_t s, int N, const int4 *ipos, const int *mol_of, const double *coms, const double *box, double4 *posd,
		double4 *quatd, float4 *axf, int *shifts);
void launch_energy_sum(cudaStream_t s, int N, const float4 *F, const float4 *Fb, double *out);
// per-replica potential energies: out[r] = sum over the slots of replica r (n_rep doubles, zeroed here)
void launch_energy_sum_replicas(cudaStream_t s, int N, int n_rep, int n_per, const float4 *F, const float4 *Fb, double *out);

// ---- lists.cu
struct ListArgs {
	int N;
	int n_rep, n_per; // replica batching: the cell table holds n_rep copies of the grid, cell id = replica * ncells + cell
	double box[3];
	BoxF boxf;
	int ncell[3];
	double rv;      // Verlet radius rcut + 2 skin (double: the exact predicate)
	const double4 *posd;
	const int4 *ipos;
	const int2 *bonds;
	int *cell_key, *cell_key_sorted, *cell_val, *cell_val_sorted, *cell_start; // N, N, N, N, ncells + 1
	int *nbr, *nnbr;
	int max_neigh, stride;
	int2 *edges;       // unique pairs that can come within rcut_near before the next rebuild ("near" edges), grouped by `from`
	int4 *edge_cnt;    // N + 1: near edges of every `from` slot per class group (x, y, z), exclusive-scanned in place; entry N = totals
	bool class_groups; // lay the edge list out in three segments by class group (common.cuh: cls_group); false: one segment
	ulonglong2 *near_mask; // per row: which entries are near edges to a higher slot
	bool direct;       // particle arrays are ordered by cell: cell members are the slots [start, end) themselves
	bool ranges_done;  // ... and the re-sort's gather pass already filled the cell table and the staleness references
	int *n_edges;      // device-side length of `edges`
	float rnear2;      // (rcut_near + 2 skin + margin)^2
	// Debye-Hueckel neighbour matrix (full, both directions), selected on the backbone-site distance
	const int4 *iback;
	double4 *ref_pos, *ref_vel, *ref_L; // the FP64 state arrays whose .w lanes take the staleness references (common.cuh, pack_ref)
	const float4 *axf;
	float base_a1, stack_a1;
	// squared site-site selection radii of the near-edge list (range + 2 skin + margin): backbone-backbone and base-base excluded volume,
	// base-base hydrogen bonding / cross stacking (r2_base), base-backbone excluded volume, stack-stack coaxial stacking
	float r2_bb, r2_eb, r2_base, r2_bk, r2_stack;
	int *dh_nbr, *dh_nnbr;
	int max_dh;
	bool dh_half;      // each Debye-Hueckel pair appears in one row only (see k_dh_particle)
	bool half_shell;   // edge pipeline: only partners in higher slots are scanned; nbr / nnbr / dh_nbr hold every pair once (row of the lower slot)
	float rdh2;        // (dh_rc + 2 skin + margin)^2
	long long edge_capacity;
	int *flags;
	void *cub_tmp;
	size_t cub_tmp_bytes;
	bool build_edges;
};
size_t lists_tmp_bytes(int N, int ncells);
void launch_build_lists(cudaStream_t s, const ListArgs &a);

// ---- sort.cu
struct SortArgs {
	int N;
	int n_rep, n_per; // replica batching: the replica index forms the top bits of the key (slots stay replica-contiguous)
	double box[3];
	const double4 *posd;
	int ncell[3];            // > 0: sort by the Hilbert index of the list builder's cell coordinates (binning for free)
	unsigned *keys, *keys_sorted;
	int *vals, *vals_sorted; // vals_sorted[new_slot] = old_slot
	int *inv;                // inv[old_slot] = new_slot
	void *cub_tmp;
	size_t cub_tmp_bytes;
	int *flags;
	int *small_tmp; // scratch of the one-launch ordering of small systems (sort_small_bytes), or null: radix sort
};
size_t sort_tmp_bytes(int N);
size_t sort_small_bytes(int N, const int ncell[3], int n_rep);
void launch_hilbert_order(cudaStream_t s, const SortArgs &a);
struct PermuteArgs {
	int N;
	const int *perm; // perm[new] = old
	const int *inv;  // inv[old] = new
	const double4 *posd_in, *veld_in, *Ld_in, *quatd_in;
	double4 *posd_out, *veld_out, *Ld_out, *quatd_out;
	const int4 *ipos_in, *iback_in;
	int4 *ipos_out, *iback_out;
	const float4 *axf_in, *F_in, *T_in;
	float4 *axf_out, *F_out, *T_out;
	const int2 *bonds_in;
	int2 *bonds_out;
	int *slot_of; // slot_of[original id] = new slot
	int *flags;   // device control words (phase timeline)
	int *cell_lin; // optional: linear cell id of every new slot (what the list builder's binning would have produced)
	// optional (with cell_lin): the list builder's cell table and staleness references are produced by the same pass (sort.cu: k_permute)
	const unsigned *keys_sorted; // the sorted Hilbert keys of the cells: equal keys = equal cell
	int *cell_start, *cell_end;  // ncells_total entries each, contiguous (cell_end = cell_start + ncells_total), zeroed by launch_permute
	long long ncells_total;
	float base_a1;
	BoxF boxf;
	int n_per;     // replica batching: particles per replica (cell ids are offset by replica * ncells)
	double box[3];
	int ncell[3];
};
void launch_permute(cudaStream_t s, const PermuteArgs &a);

} // namespace oxb
