// Context + C ABI of oxdna_b200 (see include/oxdna_b200.h).  Host-side orchestration of the MD step:
// replaces MD_CUDABackend::sim_step and its helpers (src/CUDA/Backends/MD_CUDABackend.cu:518-619,
// MD_CUDAMixedBackend.cu:82-148).  There is no CPU fallback: every entry point that computes launches CUDA kernels.
//
// Step structure (per step, steady state): [forces (+ext)] -> [one fused integrate kernel: second half-kick(n),
// thermostat(n), first half-kick + drift + rotation(n+1), staleness check].  Launches are issued in speculative
// batches: if the staleness check fires, a device-side "halt" word turns the rest of the batch into no-ops, the host
// synchronises once, rebuilds (sort + cells + Verlet list) and resumes at the pending force evaluation.  The reference
// synchronises >= 6 times per step (timers) and reads a pinned flag every step.
#include "kernels.h"
#include "dna3_pack.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <vector>

namespace oxb {
void launch_integrate_epoch(cudaStream_t s, const IntegrateArgs &a, int phases, int epoch);
}

struct oxb_ctx {
	int device = 0, N = 0, precision = OXB_PRECISION_MIXED, n_sm = 148;
	cudaStream_t stream = nullptr;
	bool own_stream = true;
	std::string err;
	long long launches = 0;

	// host-side topology (original order)
	std::vector<int> h_btype, h_n3, h_n5, h_strand;
	bool have_topology = false, have_box = false, have_model = false, have_state = false;
	double box[3] = { 0, 0, 0 };
	BoxF boxf;

	// double-buffered state (slot order)
	int cur = 0;
	double4 *posd[2] = { nullptr, nullptr }, *veld[2] = { nullptr, nullptr }, *Ld[2] = { nullptr, nullptr }, *quatd[2] = { nullptr, nullptr };
	int4 *ipos[2] = { nullptr, nullptr }, *iback[2] = { nullptr, nullptr };
	float4 *Fb = nullptr;
	float4 *axf[2] = { nullptr, nullptr }; // FP32 orientation records (a1, a3): 2 float4 per particle (common.cuh)
	float4 *quat_f4 = nullptr;             // optional reference-layout quaternion view, filled on demand
	float4 *F[2] = { nullptr, nullptr }, *T[2] = { nullptr, nullptr };
	int2 *bonds[2] = { nullptr, nullptr };
	int *slot_of = nullptr;
	float4 *pos_f4 = nullptr; // optional reference-layout view, filled on demand

	// control
	int *flags = nullptr;       // device, OXB_FLAG_WORDS
	int *h_flags = nullptr;     // pinned
	KinSums *sums = nullptr;
	double *d_energy = nullptr; // device scalar
	double *h_scalars = nullptr; // pinned

	// model
	// exactly one force field is active.  For oxRNA2 the fields of `model` that the context itself reads (site offsets,
	// radial ranges for the list radii, dh_rc, rcut_near) are mirrored from `rmodel`; the kernels get the RNA block.
	oxb_dna2_params model;
	oxb_rna2_params rmodel;
	bool is_rna = false;
	float back_a3 = 0.f;
	double rcut = 0.;
	// oxDNA3: `model` mirrors site offsets and radii as for oxRNA2; the kernels get the packed records (d3; one device allocation d3_buf)
	bool is_dna3 = false;
	oxb_dna3_dev d3;
	float4 *d3_buf = nullptr;
	int *d3_tcode = nullptr;
	bool d3_tcode_valid = false;
	oxb::ModelRef mref() const {
		if(is_dna3) return oxb::ModelRef{ nullptr, nullptr, &d3 };
		return is_rna ? oxb::ModelRef{ nullptr, &rmodel, nullptr } : oxb::ModelRef{ &model, nullptr, nullptr };
	}

	// replica batching (oxb_set_replicas): n_rep replicas of n_per particles, one row of temperature-dependent constants each
	int n_rep = 1, n_per = 0;
	oxb_replica_consts *rep = nullptr; // device table, n_rep rows (null while n_rep == 1)
	bool have_rep_consts = false;
	double *d_rep_energy = nullptr, *h_rep_energy = nullptr;

	// lists
	double skin = 0.05, max_density_multiplier = 3.;
	int use_edge = 0, sort_every = 0;
	int ncell[3] = { 0, 0, 0 };
	int max_neigh = 0;
	int *cell_key = nullptr, *cell_key_sorted = nullptr, *cell_val = nullptr, *cell_val_sorted = nullptr, *cell_start = nullptr;
	int *nbr = nullptr, *nnbr = nullptr;
	int2 *edges = nullptr;
	int4 *edge_cnt = nullptr;
	int *n_edges = nullptr;
	ulonglong2 *near_mask = nullptr;
	bool slots_cell_ordered = false; // the last re-sort ordered the slots by cell and left the cell ids in cell_key_sorted
	int sort_ncell[3] = { 0, 0, 0 }; // ... for this cell grid (a model change in between invalidates the shortcut)
	long long edge_capacity = 0;
	int edge_hint = 0;
	int *dh_nbr = nullptr, *dh_nnbr = nullptr;
	int max_dh = 0;
	int2 *hb_list = nullptr, *cx_list = nullptr, *cr_list = nullptr;
	int4 *ex_list = nullptr;
	int *ex_counts = nullptr, *ex_bonded = nullptr;
	int ex_seg = 1;
	int *seg_counts = nullptr;
	int n_seg = 1, hb_seg = 1, cx_seg = 1, cr_seg = 1;
	void *cub_tmp = nullptr;
	size_t cub_tmp_bytes = 0;
	unsigned *hkeys = nullptr, *hkeys_sorted = nullptr;
	int *hvals = nullptr, *hvals_sorted = nullptr, *hinv = nullptr;
	int *sort_small = nullptr; // scratch of the one-launch ordering of small systems (sort.cu: k_sort_small)
	size_t sort_small_bytes = 0;
	bool lists_allocated = false, lists_valid = false, forces_valid = false;
	bool need_full_matrix = false; // a consumer of both directions of every pair (oxb_device_views) has shown up: no half-shell builds any more
	bool half_shell_ok = true;     // OXB_HALF_SHELL=0 switches the half-shell scan off
	bool nbr_is_half = false;      // what the current lists hold
	// what the current lists were built for: they stay usable for a model with radii that are not larger (a replica-exchange
	// energy evaluation at a lower temperature), but only an exact match reproduces the reference's pair set
	double lists_rv = 0., lists_dh_rc = 0.;
	long long ncells_alloc = 0;
	long long n_list_updates = 0, n_sorts = 0;
	int error_flags = 0;

	// dynamics
	double dt = 0.003;
	long long step = 0;
	double spec_factor = 1.0;       // speculative batch length in units of the running mean rebuild interval (OXB_SPEC_FACTOR;
	                                // profiles/spec_sweep_r01.txt: 1.0 beats 0.8 by 3 % at 81,920 nt, 0.5 % at 1M)
	bool blocking_wait = false;     // oxb_set_host_wait
	cudaEvent_t ev_wait = nullptr;
	bool defer_build_checks = true; // OXB_DEFER_BUILD_CHECK=0 restores one host synchronisation per rebuild
	bool build_unchecked = false; // a list rebuild was launched without waiting for its overflow flags (oxb_run); the next batch checks
	double seg_scale = 1.;   // growth factor of the work-list segments (doubled whenever one overflows; OXB_SEG_SCALE sets the start value)
	bool dirty_acc = false;  // F / T / Fb hold the partial sums of an incomplete force pass: clear them before the next one
	bool class_groups = true; // edge list in three segments by class group (OXB_CLASS_GROUPS=0: one segment)
	bool near_tile = false;  // OXB_NEAR_TILE=1: tile-staged variant of the near-edge kernel (experiment)
	bool fold_hb = false;  // ... and hydrogen bonding / cross stacking in the tail of k_edge_near (OXB_FOLD_HB=0/1; default: systems below 300,000 particles)
	bool fold_hb_set = false;
	bool fold_tails = true; // coaxial stacking + FP64 excluded volume in the tails of the producing kernels (OXB_FOLD=0: separate launches)
	bool dh_half = true; // Debye-Hueckel matrix with every pair in one row + partner atomics (OXB_DH_HALF=0: full matrix, no atomics)
	bool fork_streams = true; // force pass on three concurrent streams (OXB_FORK=0/1 overrides the size-based default)
	bool mid_step = false; // positions already advanced for `step`, forces pending
	ThermostatCfg th;
	bool bussi_init = false;
	int n_ext = 0, n_ext_all = 0; // entries bound to one particle / entries acting on every particle
	DevExtForce *ext = nullptr, *ext_all = nullptr, *ext_com = nullptr;
	// marshalling on the device (marshal.cu): topology per original id, staging for the flat N x 3 double arrays of set/get_state
	int4 *d_topo = nullptr;
	double *d_stage = nullptr; // 15 N doubles: pos, a1, a3, vel, L
	int *d_marshal_err = nullptr;
	// MC barostat (oxb_barostat_*): molecule table, FP64 centres of mass, snapshot of the positions of an open trial
	int n_mol = 0;
	int *mol_of = nullptr;
	double *mol_inv_size = nullptr, *mol_coms = nullptr;
	double4 *pos_backup = nullptr;
	double box_backup[3] = { 0., 0., 0. };
	bool trial_open = false;
	int n_ext_com = 0;           // COM forces (one entry per force, evaluated by one block each)
	int *ext_pool = nullptr;     // com_list / ref_list original indices of the COM forces
	std::vector<int> ext_pool_h;
	float *ext_grid = nullptr;   // tabulated bias potentials of the metadynamics COM traps
	size_t ext_grid_n = 0;
	double avg_interval = 8.;

	// concurrency inside one force pass (independent kernels on forked streams) and graph-captured batches of steps
	cudaStream_t aux[2] = { nullptr, nullptr };
	cudaEvent_t ev_fork = nullptr, ev_near = nullptr, ev_join[2] = { nullptr, nullptr };
	long long *cur_step = nullptr; // device, 2 words (see k_integrate)
	struct BatchGraph { int units, cur; unsigned long long cfg; cudaGraphExec_t exec; };
	std::vector<BatchGraph> graphs;
	bool use_graphs = true;
	long long graph_launches = 0;
	bool profiling = false; // oxb_set_profile: device-side phase timeline (kernels.h, prof_mark)
	// plugin seam (oxb_set_force_callback): a third-party force pass in place of the built-in kernels
	oxb_force_callback force_cb = nullptr;
	void *force_cb_user = nullptr;
};

namespace {

int fail(oxb_ctx *c, int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if(c != nullptr) c->err = buf;
	return code;
}

#define CU(call)                                                                                             \
	do {                                                                                                     \
		cudaError_t e_ = (call);                                                                             \
		if(e_ != cudaSuccess) return fail(c, 100 + (int) e_, "%s failed: %s", #call, cudaGetErrorString(e_)); \
	} while(0)

template<typename T>
cudaError_t dalloc(T **p, size_t n) {
	return cudaMalloc((void **) p, sizeof(T) * std::max<size_t>(n, 1));
}

// everything a captured batch freezes (pointers, model constants, dt, thermostat, grids) invalidates the cache
void drop_graphs(oxb_ctx *c) {
	for(auto &g : c->graphs) cudaGraphExecDestroy(g.exec);
	c->graphs.clear();
}

void set_boxf(oxb_ctx *c) {
	c->boxf.lx = (float) c->box[0]; c->boxf.ly = (float) c->box[1]; c->boxf.lz = (float) c->box[2];
	c->boxf.sx = (float) (c->box[0] / 4294967296.0); c->boxf.sy = (float) (c->box[1] / 4294967296.0); c->boxf.sz = (float) (c->box[2] / 4294967296.0);
	c->boxf.dsx = c->box[0] / 4294967296.0; c->boxf.dsy = c->box[1] / 4294967296.0; c->boxf.dsz = c->box[2] / 4294967296.0;
}

void free_lists(oxb_ctx *c) {
	drop_graphs(c);
	cudaFree(c->cell_key); cudaFree(c->cell_key_sorted); cudaFree(c->cell_val); cudaFree(c->cell_val_sorted); cudaFree(c->cell_start);
	cudaFree(c->nbr); cudaFree(c->nnbr); cudaFree(c->edges); cudaFree(c->edge_cnt); cudaFree(c->n_edges); cudaFree(c->cub_tmp);
	cudaFree(c->dh_nbr); cudaFree(c->dh_nnbr); cudaFree(c->near_mask);
	c->near_mask = nullptr;
	c->slots_cell_ordered = false;
	cudaFree(c->hb_list); cudaFree(c->cx_list); cudaFree(c->cr_list); cudaFree(c->seg_counts);
	cudaFree(c->ex_list); cudaFree(c->ex_counts); cudaFree(c->ex_bonded);
	c->ex_list = nullptr; c->ex_counts = c->ex_bonded = nullptr;
	c->cell_key = c->cell_key_sorted = c->cell_val = c->cell_val_sorted = c->cell_start = c->nbr = c->nnbr = c->n_edges = nullptr;
	c->edge_cnt = nullptr;
	c->seg_counts = c->dh_nbr = c->dh_nnbr = nullptr;
	c->edges = c->hb_list = c->cx_list = c->cr_list = nullptr;
	c->cub_tmp = nullptr;
	c->ncells_alloc = 0;
	c->lists_allocated = false;
}

// cell grid for the current cutoff; the cell table is the only allocation that depends on it
int ensure_cells(oxb_ctx *c) {
	const int N = c->N;
	double rv = c->rcut + 2. * c->skin;
	long long ncells = 1;
	for(int k = 0; k < 3; k++) {
		// CUDASimpleVerletList::_compute_N_cells_side (CUDASimpleVerletList.cu:96-110): floor(L / r_verlet), at least 3
		int n = (int) std::floor(c->box[k] / rv + 1e-9);
		if(n < 3) n = 3;
		if(n > 1024) n = 1024;
		c->ncell[k] = n;
		ncells *= n;
	}
	// keep the cell table bounded for huge dilute boxes: coarsen uniformly (cells only need to be >= r_verlet wide)
	while(ncells > 8ll * (N / c->n_rep) + 4096) {
		ncells = 1;
		for(int k = 0; k < 3; k++) {
			c->ncell[k] = std::max(3, (c->ncell[k] * 4) / 5);
			ncells *= c->ncell[k];
		}
	}
	if(c->n_rep > 1) {
		// one copy of the grid per replica; the replica index also forms the top bits of the 32-bit Hilbert key
		int bits = 1, rep_bits = 0;
		while((1 << bits) < std::max(c->ncell[0], std::max(c->ncell[1], c->ncell[2]))) bits++;
		while((1 << rep_bits) < c->n_rep) rep_bits++;
		if(3 * bits + rep_bits > 32) return fail(c, 1, "replica batching: %d replicas x a %d x %d x %d cell grid do not fit the 32-bit sort key", c->n_rep, c->ncell[0], c->ncell[1], c->ncell[2]);
		ncells *= c->n_rep;
	}
	if(ncells > c->ncells_alloc) {
		CU(cudaStreamSynchronize(c->stream));
		cudaFree(c->cell_start); cudaFree(c->cub_tmp);
		c->cell_start = nullptr; c->cub_tmp = nullptr;
		c->ncells_alloc = ncells + ncells / 8;
		CU(dalloc(&c->cell_start, 2 * (size_t) c->ncells_alloc));
		c->cub_tmp_bytes = std::max(oxb::lists_tmp_bytes(N, (int) c->ncells_alloc), oxb::sort_tmp_bytes(N));
		CU(cudaMalloc(&c->cub_tmp, c->cub_tmp_bytes));
	}
	return 0;
}

int alloc_lists(oxb_ctx *c, int max_neigh) {
	free_lists(c);
	const int N = c->N;
	c->ncells_alloc = 0;
	{ int rc = ensure_cells(c); if(rc) return rc; }
	c->max_neigh = max_neigh;
	CU(dalloc(&c->cell_key, N)); CU(dalloc(&c->cell_key_sorted, N)); CU(dalloc(&c->cell_val, N)); CU(dalloc(&c->cell_val_sorted, N));
	CU(dalloc(&c->nbr, (size_t) max_neigh * N)); CU(dalloc(&c->nnbr, N));
	CU(cudaMemset(c->nbr, 0, sizeof(int) * (size_t) max_neigh * N)); // (oxb_get_pairs downloads whole rows: keep the unwritten tails defined)
	CU(dalloc(&c->edge_cnt, (size_t) N + 1)); CU(dalloc(&c->n_edges, 4)); CU(dalloc(&c->near_mask, (size_t) N));
	CU(cudaMemset(c->edge_cnt, 0, sizeof(int4) * ((size_t) N + 1)));
	CU(cudaMemset(c->n_edges, 0, 4 * sizeof(int)));
	c->max_dh = max_neigh;
	CU(dalloc(&c->dh_nbr, (size_t) c->max_dh * N)); CU(dalloc(&c->dh_nnbr, N));
	// segmented work lists of the edge pipeline: one segment per block of the near-edge kernel, sized ~4x the expected load
	// (hydrogen-bonding pairs ~0.5 N, cross-stacking-only pairs ~1.2 N, coaxial pairs << N) plus a floor for tiny systems
	// ~3.3 near edges per particle, one producer block per 128 edges up to 16 blocks per SM (grid-stride beyond that)
	{
		// blocks per SM of the near-edge kernel and of its work-list segments: 8 = what an SM holds of this kernel (64 registers x 128
		// threads), i.e. ONE full wave that strides over the edges, no tail wave.  Sweep on B200 (gpurun_out r2t / r2u): at 81,920 nt 8 gives
		// 1.58e9 particle-steps/s against 1.48e9 for 10, 12 and 16 and 1.44e9 for 4; neutral at 1M nt
		const char *e = getenv("OXB_PB_NEAR");
		const int pb = (e != nullptr && atoi(e) > 0) ? atoi(e) : 8;
		c->n_seg = c->use_edge ? (int) std::max<long long>(1, std::min<long long>((long long) pb * c->n_sm, (33ll * N / 10 + 127) / 128)) : 1;
		if(c->use_edge && c->near_tile) c->n_seg = (N + 127) / 128; // one block (and one work-list segment) per tile of 128 slots
	}
	c->hb_seg = c->use_edge ? (int) (c->seg_scale * (double) (6ll * N / c->n_seg + 128)) + 3 : 1;
	c->cr_seg = 1; // the cross-stacking-only list is not produced (see forces.cu)
	c->cx_seg = c->use_edge ? (int) (c->seg_scale * (double) (2ll * N / c->n_seg + 64)) + 1 : 1;
	CU(dalloc(&c->hb_list, (size_t) c->hb_seg * c->n_seg)); CU(dalloc(&c->cx_list, (size_t) c->cx_seg * c->n_seg));
	c->ex_seg = c->use_edge ? (int) (c->seg_scale * (double) (2ll * N / c->n_seg + 64)) + 1 : 1;
	CU(dalloc(&c->ex_list, (size_t) c->ex_seg * c->n_seg)); CU(dalloc(&c->ex_counts, (size_t) c->n_seg)); CU(dalloc(&c->ex_bonded, (size_t) N));
	CU(cudaMemset(c->ex_counts, 0, sizeof(int) * (size_t) c->n_seg));
	CU(cudaMemset(c->ex_bonded, 0, sizeof(int) * (size_t) N));
	CU(dalloc(&c->cr_list, (size_t) c->cr_seg * c->n_seg)); CU(dalloc(&c->seg_counts, (size_t) 3 * c->n_seg));
	CU(cudaMemset(c->seg_counts, 0, sizeof(int) * 3 * (size_t) c->n_seg));
	c->edge_capacity = c->use_edge ? ((long long) N * max_neigh) / 4 + N : 1;
	CU(dalloc(&c->edges, (size_t) c->edge_capacity));
	c->lists_allocated = true;
	return 0;
}

oxb::ListArgs list_args(oxb_ctx *c) {
	oxb::ListArgs a;
	a.N = c->N;
	a.n_rep = c->n_rep; a.n_per = c->n_per;
	for(int k = 0; k < 3; k++) { a.box[k] = c->box[k]; a.ncell[k] = c->ncell[k]; }
	a.boxf = c->boxf;
	a.rv = c->rcut + 2. * c->skin;
	a.posd = c->posd[c->cur]; a.ipos = c->ipos[c->cur]; a.bonds = c->bonds[c->cur];
	a.cell_key = c->cell_key; a.cell_key_sorted = c->cell_key_sorted; a.cell_val = c->cell_val; a.cell_val_sorted = c->cell_val_sorted;
	a.cell_start = c->cell_start;
	a.nbr = c->nbr; a.nnbr = c->nnbr; a.max_neigh = c->max_neigh; a.stride = c->N;
	a.edges = c->edges; a.edge_cnt = c->edge_cnt; a.n_edges = c->n_edges; a.edge_capacity = c->edge_capacity;
	a.iback = c->iback[c->cur]; a.axf = c->axf[c->cur];
	a.ref_pos = c->posd[c->cur]; a.ref_vel = c->veld[c->cur]; a.ref_L = c->Ld[c->cur];
	a.base_a1 = c->model.base_a1; a.stack_a1 = c->model.stack_a1;
	{
		const oxb_dna2_params &M = c->model;
		double pad = 2. * c->skin + 0.02;
		auto sq = [](double x) { return (float) (x * x); };
		a.r2_bb = sq(M.excl[0].rc + pad);
		a.r2_base = sq(std::max((double) M.hb.rchigh, (double) M.crst.rchigh) + pad);
		a.r2_eb = sq((double) M.excl[1].rc + pad);
		a.r2_bk = sq(std::max((double) M.excl[2].rc, (double) M.excl[3].rc) + pad);
		a.r2_stack = sq((double) M.cxst.rchigh + pad);
	} a.dh_nbr = c->dh_nbr; a.dh_nnbr = c->dh_nnbr; a.max_dh = c->max_dh;
	a.dh_half = c->dh_half && c->use_edge;
	// the edge pipeline consumes every pair once (near edges from the lower slot, half Debye-Hueckel matrix): scan half a shell and leave the
	// upper triangle of the Verlet matrix only.  The full matrix (plugin seam) is built when somebody asks for it (oxb_device_views)
	a.half_shell = a.dh_half && !c->need_full_matrix && c->half_shell_ok;
	{
		double rd = (double) c->model.dh_rc + 2. * c->skin + 0.02;
		a.rdh2 = (float) (rd * rd);
	}
	{
		// pairs further apart than this at build time cannot come within rcut_near before the next rebuild (each particle
		// moves at most `skin` plus one step): they only ever feel Debye-Hueckel
		double rn = (double) c->model.rcut_near + 2. * c->skin + 0.02;
		a.rnear2 = (float) (rn * rn);
	}
	a.flags = c->flags;
	a.cub_tmp = c->cub_tmp; a.cub_tmp_bytes = c->cub_tmp_bytes;
	a.build_edges = c->use_edge != 0;
	a.class_groups = c->class_groups && !c->near_tile; // (the tile variant of the near-edge kernel walks one segment by `from`)
	a.near_mask = c->near_mask;
	a.direct = c->slots_cell_ordered && c->sort_ncell[0] == c->ncell[0] && c->sort_ncell[1] == c->ncell[1] && c->sort_ncell[2] == c->ncell[2];
	a.ranges_done = a.direct;
	return a;
}

int read_flags(oxb_ctx *c) {
	CU(cudaMemcpyAsync(c->h_flags, c->flags, sizeof(int) * OXB_FLAG_WORDS, cudaMemcpyDeviceToHost, c->stream));
	if(c->blocking_wait) {
		// replica ensembles with more host threads than cores: sleep on a blocking-sync event instead of spinning in the driver
		if(c->ev_wait == nullptr) CU(cudaEventCreateWithFlags(&c->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
		CU(cudaEventRecord(c->ev_wait, c->stream));
		CU(cudaEventSynchronize(c->ev_wait));
	}
	else CU(cudaStreamSynchronize(c->stream));
	c->error_flags |= c->h_flags[OXB_FLAG_ERROR];
	return 0;
}

int do_sort(oxb_ctx *c) {
	if(!c->lists_allocated) { int rc = alloc_lists(c, c->max_neigh > 0 ? c->max_neigh : 64); if(rc) return rc; }
	{ int rc = ensure_cells(c); if(rc) return rc; }
	const int N = c->N, a = c->cur, b = 1 - c->cur;
	if(c->hkeys == nullptr) {
		CU(dalloc(&c->hkeys, N)); CU(dalloc(&c->hkeys_sorted, N)); CU(dalloc(&c->hvals, N)); CU(dalloc(&c->hvals_sorted, N)); CU(dalloc(&c->hinv, N));
	}
	oxb::SortArgs s;
	s.N = N;
	s.n_rep = c->n_rep; s.n_per = c->n_per;
	for(int k = 0; k < 3; k++) s.box[k] = c->box[k];
	s.posd = c->posd[a];
	for(int k = 0; k < 3; k++) s.ncell[k] = c->ncell[k];
	s.keys = c->hkeys; s.keys_sorted = c->hkeys_sorted; s.vals = c->hvals; s.vals_sorted = c->hvals_sorted; s.inv = c->hinv;
	{
		const size_t need = oxb::sort_small_bytes(N, c->ncell, c->n_rep);
		if(need > c->sort_small_bytes) {
			CU(cudaStreamSynchronize(c->stream));
			cudaFree(c->sort_small);
			c->sort_small = nullptr;
			CU(cudaMalloc((void **) &c->sort_small, need));
			c->sort_small_bytes = need;
		}
		s.small_tmp = need > 0 ? c->sort_small : nullptr;
	}
	s.cub_tmp = c->cub_tmp; s.cub_tmp_bytes = c->cub_tmp_bytes;
	s.flags = c->flags;
	oxb::launch_hilbert_order(c->stream, s);
	oxb::PermuteArgs p;
	p.N = N; p.perm = c->hvals_sorted; p.inv = c->hinv;
	p.posd_in = c->posd[a]; p.veld_in = c->veld[a]; p.Ld_in = c->Ld[a]; p.quatd_in = c->quatd[a];
	p.posd_out = c->posd[b]; p.veld_out = c->veld[b]; p.Ld_out = c->Ld[b]; p.quatd_out = c->quatd[b];
	p.ipos_in = c->ipos[a]; p.ipos_out = c->ipos[b];
	p.iback_in = c->iback[a]; p.iback_out = c->iback[b];
	p.axf_in = c->axf[a]; p.F_in = c->F[a]; p.T_in = c->T[a]; p.axf_out = c->axf[b]; p.F_out = c->F[b]; p.T_out = c->T[b];
	p.bonds_in = c->bonds[a]; p.bonds_out = c->bonds[b];
	p.slot_of = c->slot_of;
	p.flags = c->flags;
	p.cell_lin = c->cell_key_sorted;
	p.n_per = c->n_per;
	for(int k = 0; k < 3; k++) { p.box[k] = c->box[k]; p.ncell[k] = c->ncell[k]; }
	p.ncells_total = (long long) c->ncell[0] * c->ncell[1] * c->ncell[2] * c->n_rep;
	p.keys_sorted = c->hkeys_sorted; p.cell_start = c->cell_start; p.cell_end = c->cell_start + p.ncells_total;
	p.base_a1 = c->model.base_a1; p.boxf = c->boxf;
	oxb::launch_permute(c->stream, p);
	c->slots_cell_ordered = true; // consumed (and cleared) by the list build that follows
	for(int k = 0; k < 3; k++) c->sort_ncell[k] = c->ncell[k];
	c->launches += 5;
	c->cur = b;
	c->n_sorts++;
	c->lists_valid = false;
	c->forces_valid = false; // Fb (backbone-site force sums) is not permuted
	CU(cudaGetLastError());
	return 0;
}

// deferred: launch the rebuild and return WITHOUT reading its overflow flags (one host synchronisation less per rebuild inside
// oxb_run).  The batch that follows starts with k_batch_begin(check = 1), which turns an overflow into a halt: every kernel of the
// batch is then a no-op, oxb_run sees "halted with zero steps done + overflow bits" and repeats the rebuild on the checked path.
int do_build(oxb_ctx *c, bool deferred = false) {
	if(!c->lists_allocated) { int rc = alloc_lists(c, 64); if(rc) return rc; deferred = false; }
	{ int rc = ensure_cells(c); if(rc) return rc; }
	for(int attempt = 0; attempt < 6; attempt++) {
		CU(cudaMemsetAsync(c->flags + OXB_FLAG_ERROR, 0, sizeof(int), c->stream));
		{
			const oxb::ListArgs la = list_args(c);
			oxb::launch_build_lists(c->stream, la);
			c->nbr_is_half = la.half_shell;
			// neighbour scan (+ scan + edge fill); without a preceding re-sort also cell keys, radix sort (>= 3 kernels), cell ranges
			c->launches += (c->use_edge ? 4 : 1) + (la.direct ? 0 : 5);
		}
		CU(cudaGetLastError());
		if(deferred) {
			c->lists_valid = true;
			c->lists_rv = c->rcut + 2. * c->skin;
			c->lists_dh_rc = (double) c->model.dh_rc;
			c->slots_cell_ordered = false;
			c->n_list_updates++;
			c->build_unchecked = true;
			return 0;
		}
		int rc = read_flags(c);
		if(rc) return rc;
		int seen = c->h_flags[OXB_FLAG_MAX_NEIGH_SEEN];
		if((c->h_flags[OXB_FLAG_ERROR] & (OXB_ERR_NEIGH_OVERFLOW | OXB_ERR_EDGE_OVERFLOW)) == 0) {
			c->error_flags &= ~(OXB_ERR_NEIGH_OVERFLOW | OXB_ERR_EDGE_OVERFLOW);
			c->lists_valid = true;
			c->lists_rv = c->rcut + 2. * c->skin;
			c->lists_dh_rc = (double) c->model.dh_rc;
			c->slots_cell_ordered = false; // particles move on: the next build bins them itself unless a re-sort precedes it
			c->n_list_updates++;
			return 0;
		}
		// a row overflowed: grow the matrix (the reference silently corrupts the next column here) and rebuild
		int want = ((int) (seen * 1.5) + 15) / 8 * 8;
		c->error_flags &= ~(OXB_ERR_NEIGH_OVERFLOW | OXB_ERR_EDGE_OVERFLOW);
		rc = alloc_lists(c, std::max(want, c->max_neigh * 2));
		if(rc) return rc;
	}
	return fail(c, 3, "neighbour list does not fit after repeated growth (max_neigh = %d)", c->max_neigh);
}

int ensure_lists(oxb_ctx *c, bool deferred = false) {
	if(c->lists_valid) return 0;
	if(!c->have_state || !c->have_model || !c->have_box) return fail(c, 2, "box, model and state must be set before building lists");
	if(c->sort_every > 0 && (c->n_list_updates % c->sort_every) == 0) {
		int rc = do_sort(c);
		if(rc) return rc;
	}
	return do_build(c, deferred);
}

__global__ void k_pos_view(int N, const double4 *__restrict__ posd, const int4 *__restrict__ ipos, float4 *__restrict__ out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	double4 p = posd[i];
	out[i] = make_float4((float) p.x, (float) p.y, (float) p.z, __int_as_float(ipos[i].w));
}
__global__ void k_quat_view(int N, const double4 *__restrict__ quatd, float4 *__restrict__ out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	double4 q = quatd[i];
	out[i] = make_float4((float) q.x, (float) q.y, (float) q.z, (float) q.w);
}

// plugin seam: reference-layout views of the current state, then the third-party force pass on the context's stream.  The accumulators
// are cleared first in every case: launches behind a halt are no-ops for the built-in kernels but not for a plugin's.
int launch_forces_callback(oxb_ctx *c, long long step) {
	const int a = c->cur, N = c->N;
	cudaStream_t m = c->stream;
	if(c->pos_f4 == nullptr) CU(dalloc(&c->pos_f4, N));
	if(c->quat_f4 == nullptr) CU(dalloc(&c->quat_f4, N));
	CU(cudaMemsetAsync(c->F[a], 0, sizeof(float4) * (size_t) N, m));
	CU(cudaMemsetAsync(c->T[a], 0, sizeof(float4) * (size_t) N, m));
	k_pos_view<<<(N + 255) / 256, 256, 0, m>>>(N, c->posd[a], c->ipos[a], c->pos_f4);
	k_quat_view<<<(N + 255) / 256, 256, 0, m>>>(N, c->quatd[a], c->quat_f4);
	c->launches += 2;
	oxb_force_views v;
	v.N = N; v.stride = N;
	v.poss = c->pos_f4; v.orientations = c->quat_f4;
	v.matrix_neighs = c->nbr; v.number_neighs = c->nnbr;
	v.bonds = c->bonds[a];
	v.forces = c->F[a]; v.torques = c->T[a];
	for(int k = 0; k < 3; k++) v.box[k] = c->box[k];
	v.step = step < 0 ? c->step : step;
	v.stream = (void *) m;
	const int rc = c->force_cb(c->force_cb_user, &v);
	if(rc != 0) return fail(c, 10, "the force callback of the plugged-in interaction failed with code %d", rc);
	c->launches += 1;
	CU(cudaGetLastError());
	return 0;
}

// hw: index of the halt word the launched kernels must honour; clear: F/T are not known to be zero; step < 0: kernels read
// the step index from the device counter.  Edge pipeline = 5 kernels (6 in mixed precision: + k_excl_fix on aux1) on 3 streams:
//   main: near edges -> hydrogen bonding / cross stacking      aux0: Debye-Hueckel      aux1: bonds, external forces, coaxial stacking
// joined back into main before the integrator.  Under stream capture the same calls become the fork/join edges of a graph.
int launch_forces(oxb_ctx *c, int hw, bool clear, long long step) {
	const int a = c->cur;
	cudaStream_t m = c->stream;
	if(c->use_edge) {
		if(clear) {
			CU(cudaMemsetAsync(c->F[a], 0, sizeof(float4) * (size_t) c->N, m));
			CU(cudaMemsetAsync(c->T[a], 0, sizeof(float4) * (size_t) c->N, m));
			if(c->dh_half) CU(cudaMemsetAsync(c->Fb, 0, sizeof(float4) * (size_t) c->N, m));
		}
		oxb::EdgeArgs e;
		e.rep = c->rep; e.n_per = c->n_per;
		e.N = c->N; e.ipos = c->ipos[a]; e.iback = c->iback[a]; e.axf = c->axf[a]; e.posd = c->posd[a]; e.quatd = c->quatd[a]; e.bonds = c->bonds[a]; e.edges = c->edges;
		e.n_edges = c->n_edges; e.dh_nbr = c->dh_nbr; e.dh_nnbr = c->dh_nnbr;
		e.edge_cnt = c->edge_cnt; e.near_tile = c->near_tile ? 1 : 0;
		e.F = c->F[a]; e.T = c->T[a]; e.Fb = c->Fb; e.hb_list = c->hb_list; e.cx_list = c->cx_list; e.cr_list = c->cr_list; e.seg_counts = c->seg_counts;
		e.ex_list = c->ex_list; e.ex_counts = c->ex_counts; e.ex_bonded = c->ex_bonded; e.ex_seg = c->ex_seg;
		e.refine = (c->precision == OXB_PRECISION_MIXED) ? 1 : 0;
		e.dh_half = c->dh_half ? 1 : 0;
		// bit 0: coaxial stacking + FP64 excluded volume in the tails of the producers; bit 1: hydrogen bonding / cross stacking too (no stage 2 launch)
		e.fold = c->is_dna3 ? 0 : (c->fold_tails ? (c->fold_hb ? 3 : 1) : 0); // (oxDNA3: every stage is its own launch, forces.cu)
		e.n_seg = c->n_seg; e.hb_seg = c->hb_seg; e.cx_seg = c->cx_seg; e.cr_seg = c->cr_seg;
		// ~1.7 items per particle in that list; aim at ~2 items per consumer thread
		e.hb_split = (int) std::max<long long>(1, std::min<long long>(8, (17ll * c->N / 10 / c->n_seg + 64) / 128));
		// fork = 1: three concurrent streams (pays while one kernel cannot fill the GPU); fork = 0: the same launches in line on the
		// main stream (large systems: every kernel fills the machine by itself and concurrent kernels only evict each other's
		// gathers from L2 -- measured at 1M nucleotides, profiles/fork_sweep_r01.txt)
		const bool fork = c->fork_streams;
		cudaStream_t s0 = fork ? c->aux[0] : m, s1 = fork ? c->aux[1] : m;
		if(fork) {
			CU(cudaEventRecord(c->ev_fork, m));
			CU(cudaStreamWaitEvent(c->aux[0], c->ev_fork, 0));
			CU(cudaStreamWaitEvent(c->aux[1], c->ev_fork, 0));
		}
		oxb::launch_edge_stage(s0, 0, c->mref(), c->boxf, e, c->flags, hw);
		oxb::launch_edge_stage(s1, 4, c->mref(), c->boxf, e, c->flags, hw);
		oxb::launch_edge_stage(m, 1, c->mref(), c->boxf, e, c->flags, hw);
		if(fork) CU(cudaEventRecord(c->ev_near, m));
		if(!(e.fold & 2)) oxb::launch_edge_stage(m, 2, c->mref(), c->boxf, e, c->flags, hw);
		if(c->n_ext > 0) {
			oxb::launch_ext_forces(s1, c->n_ext, c->ext, c->slot_of, c->ipos[a], c->posd[a], c->boxf, step, c->cur_step, c->F[a], c->flags, hw);
			c->launches += 1;
		}
		if(c->n_ext_all > 0) {
			oxb::launch_ext_forces_all(s1, c->N, c->n_ext_all, c->ext_all, c->slot_of, c->ipos[a], c->posd[a], c->boxf, step, c->cur_step, c->F[a], c->flags, hw);
			c->launches += 1;
		}
		if(c->n_ext_com > 0) {
			oxb::launch_ext_com(s1, c->n_ext_com, c->ext_com, c->ext_pool, c->ext_grid, c->slot_of, c->posd[a], c->quatd[a], c->ipos[a], c->box, step, c->cur_step, c->F[a], c->T[a], c->flags, hw);
			c->launches += 1;
		}
		if(fork) CU(cudaStreamWaitEvent(c->aux[1], c->ev_near, 0));
		if(!e.fold) {
			oxb::launch_edge_stage(s1, 3, c->mref(), c->boxf, e, c->flags, hw);
			if(e.refine && !c->is_dna3) oxb::launch_edge_stage(s1, 6, c->mref(), c->boxf, e, c->flags, hw); // excluded volume in double for the parked pairs (after near + bonded)
		}
		if(fork) {
			CU(cudaEventRecord(c->ev_join[0], c->aux[0]));
			CU(cudaEventRecord(c->ev_join[1], c->aux[1]));
			CU(cudaStreamWaitEvent(m, c->ev_join[0], 0));
			CU(cudaStreamWaitEvent(m, c->ev_join[1], 0));
		}
		c->launches += c->is_dna3 ? 5 : ((e.fold & 2) ? 3 : (e.fold ? 4 : (e.refine ? 6 : 5)));
	}
	else {
		if(c->force_cb != nullptr) { int rc = launch_forces_callback(c, step); if(rc) return rc; }
		else oxb::launch_forces_particle(m, c->mref(), c->boxf, c->N, c->ipos[a], c->iback[a], c->axf[a],
				c->precision == OXB_PRECISION_MIXED ? c->posd[a] : nullptr, c->quatd[a], c->bonds[a], c->nbr, c->nnbr, c->N, c->F[a], c->T[a],
				c->rep, c->n_per, c->flags, hw);
		if(c->force_cb == nullptr) c->launches += 1;
		if(c->n_ext > 0) {
			oxb::launch_ext_forces(m, c->n_ext, c->ext, c->slot_of, c->ipos[a], c->posd[a], c->boxf, step, c->cur_step, c->F[a], c->flags, hw);
			c->launches += 1;
		}
		if(c->n_ext_all > 0) {
			oxb::launch_ext_forces_all(m, c->N, c->n_ext_all, c->ext_all, c->slot_of, c->ipos[a], c->posd[a], c->boxf, step, c->cur_step, c->F[a], c->flags, hw);
			c->launches += 1;
		}
		if(c->n_ext_com > 0) {
			oxb::launch_ext_com(m, c->n_ext_com, c->ext_com, c->ext_pool, c->ext_grid, c->slot_of, c->posd[a], c->quatd[a], c->ipos[a], c->box, step, c->cur_step, c->F[a], c->T[a], c->flags, hw);
			c->launches += 1;
		}
	}
	return 0;
}

oxb::IntegrateArgs integ_args(oxb_ctx *c, long long step) {
	oxb::IntegrateArgs a;
	const int k = c->cur;
	a.N = c->N; a.dt = c->dt;
	for(int d = 0; d < 3; d++) a.box_inv[d] = 1. / c->box[d];
	{
		// the staleness references are 21-bit fixed point (common.cuh, pack_ref): tighten the threshold by their worst-case error
		const double lmax = std::max(c->box[0], std::max(c->box[1], c->box[2]));
		const double s_eff = std::max(c->skin - 1.7321 * lmax / 4194304., 0.5 * c->skin);
		a.skin2 = (float) (s_eff * s_eff);
	}
	a.box = c->boxf;
	a.posd = c->posd[k]; a.veld = c->veld[k]; a.Ld = c->Ld[k]; a.quatd = c->quatd[k];
	a.ipos = c->ipos[k]; a.axf = c->axf[k];
	a.F = c->F[k]; a.T = c->T[k]; a.Fb = c->use_edge ? c->Fb : nullptr; a.iback = c->iback[k];
	a.back_a1 = c->model.back_a1; a.back_a2 = c->model.back_a2; a.back_a3 = c->back_a3; a.base_a1 = c->model.base_a1;
	a.flags = c->flags; a.sums = c->sums; a.th = c->th; a.step = step;
	a.rep = c->rep; a.n_per = c->n_per;
	a.zero_Fb = (c->use_edge && c->dh_half) ? 1 : 0;
	a.cur_step = c->cur_step;
	return a;
}

__global__ void k_batch_begin(int *flags, long long *cur_step, long long step, int check_build) {
	prof_mark(flags, OXB_PROF_GAP); // from here to the first kernel of the force pass: launch latency of the batch (graph launch included)
	// check_build: the list rebuild in front of this batch was not checked by the host; an overflow halts the whole batch
	const int halt = (check_build && (flags[OXB_FLAG_ERROR] & (OXB_ERR_NEIGH_OVERFLOW | OXB_ERR_EDGE_OVERFLOW))) ? 1 : 0;
	flags[OXB_FLAG_STEPS_DONE] = 0;
	flags[OXB_FLAG_COUNT] = halt;
	flags[OXB_FLAG_COUNT + 1] = halt;
	cur_step[0] = step;
	cur_step[1] = step;
}

__global__ void k_prof_mark(int *flags, int phase, int reset) { prof_mark(flags, phase, reset != 0); }

// start of a batch of launches: clears the halt words and the completed-step counter, seeds the device-side step index
int reset_batch_flags(oxb_ctx *c) {
	k_batch_begin<<<1, 1, 0, c->stream>>>(c->flags, c->cur_step, c->step, c->build_unchecked ? 1 : 0);
	c->launches++;
	CU(cudaGetLastError());
	return 0;
}

bool thermostat_active(const oxb_ctx *c, long long s) {
	if(c->th.type == OXB_THERMOSTAT_NONE) return false;
	if(c->th.type == OXB_THERMOSTAT_LANGEVIN) return true;
	return (s % c->th.every) == 0;
}

int init_bussi(oxb_ctx *c) {
	KinSums h;
	std::memset(&h, 0, sizeof(h));
	double T = c->th.a;
	h.K_t = 0.5 * (3. * (c->N - 1)) * T; // BussiThermostat.cpp:36-43,152-158
	h.K_r = 0.5 * (3. * c->N) * T;
	h.factor_t = h.factor_r = 1.;
	CU(cudaMemcpyAsync(c->sums, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
	c->bussi_init = true;
	return 0;
}

// the per-particle type word of the oxDNA3 kernels (by ORIGINAL id): type | n3 type << 3 | n5 type << 6 | (btype == 4) << 9, 5 = no neighbour
// (neigh_types of the reference, CUDA_DNA3.cuh:96-108; DNANucleotide::set_positions for the dummy base, DNANucleotide.cpp:71-79)
int dna3_upload_codes(oxb_ctx *c) {
	if(c->d3_tcode_valid) return 0;
	if(!c->have_topology) return fail(c, 2, "the topology must be set before oxDNA3 forces are evaluated");
	const int N = c->N;
	std::vector<int> code(N);
	for(int i = 0; i < N; i++) {
		// the dummy base 'D' has btype = type = 4 (TopologyParser.cpp:96-99, Utils.cpp:33-34)
		auto ty = [](int b) { return b == 4 ? 4 : btype_to_type(b); };
		const int t = ty(c->h_btype[i]);
		const int t3 = c->h_n3[i] >= 0 ? ty(c->h_btype[c->h_n3[i]]) : 5;
		const int t5 = c->h_n5[i] >= 0 ? ty(c->h_btype[c->h_n5[i]]) : 5;
		code[i] = t | (t3 << 3) | (t5 << 6) | ((c->h_btype[i] == 4) ? (1 << 9) : 0);
	}
	if(c->d3_tcode == nullptr) CU(dalloc(&c->d3_tcode, (size_t) N));
	CU(cudaMemcpy(c->d3_tcode, code.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
	c->d3.tcode = c->d3_tcode;
	c->d3_tcode_valid = true;
	return 0;
}

int check_ready(oxb_ctx *c) {
	// any host thread may drive a context (replica ensembles run one thread per local replica): bind the calling thread
	cudaSetDevice(c->device);
	if(!c->have_topology) return fail(c, 2, "topology not set");
	if(!c->have_box) return fail(c, 2, "box not set");
	if(!c->have_model) return fail(c, 2, "interaction model not set");
	if(!c->have_state) return fail(c, 2, "state not set");
	if(c->n_rep > 1) {
		if(!c->have_rep_consts) return fail(c, 2, "replica batching: oxb_set_replica_consts has not been called");
		if(c->th.type == OXB_THERMOSTAT_BUSSI) return fail(c, 1, "replica batching is not available with the Bussi thermostat");
		if(c->is_dna3) return fail(c, 1, "replica batching is not available for oxDNA3");
	}
	if(c->is_dna3) return dna3_upload_codes(c);
	return 0;
}

int ensure_forces(oxb_ctx *c) {
	int rc = check_ready(c);
	if(rc) return rc;
	rc = ensure_lists(c);
	if(rc) return rc;
	if(!c->forces_valid) {
		rc = reset_batch_flags(c);
		if(rc) return rc;
		rc = launch_forces(c, OXB_FLAG_COUNT, true, c->step);
		if(rc) return rc;
		CU(cudaGetLastError());
		c->forces_valid = true;
		c->dirty_acc = false; // (the pass started from cleared accumulators)
	}
	return 0;
}

// A force pass that overflowed a work-list segment has dropped pairs.  The segments are sized from N-averaged heuristics; a dense aggregate
// (compact origami, high-pressure NPT) can exceed them.  Recovery, like the neighbour matrix and the edge list: double the segments
// (the lists are rebuilt: every array of the edge pipeline is reallocated) and repeat the pass.
int grow_segments(oxb_ctx *c, const char *what) {
	c->error_flags &= ~OXB_ERR_SEG_OVERFLOW;
	if(c->seg_scale > 256.) return fail(c, 8, "%s: a work-list segment of the edge pipeline overflowed even at %g x its default size", what, c->seg_scale);
	c->seg_scale *= 2.;
	int rc = alloc_lists(c, c->max_neigh > 0 ? c->max_neigh : 64);
	if(rc) return rc;
	c->lists_valid = false; c->forces_valid = false;
	c->dirty_acc = true;
	return 0;
}

// Force passes launched outside oxb_run (energies, force read-backs, barostat trials) must not swallow such an overflow either (a barostat
// trial could accept on an energy with dropped pairs).  One 64-byte read-back behind the synchronisation these callers do anyway.  A FENE
// bond out of range is NOT an error here: the reference's CPU energy of such a state is a huge number that a barostat trial simply
// rejects; oxb_run reports it for states it would integrate.  Returns -1 if the pass has to be repeated.
int check_force_flags(oxb_ctx *c, const char *what) {
	CU(cudaMemcpyAsync(c->h_flags, c->flags, sizeof(int) * OXB_FLAG_WORDS, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	if(c->h_flags[OXB_FLAG_ERROR] & OXB_ERR_SEG_OVERFLOW) {
		int rc = grow_segments(c, what);
		return rc ? rc : -1;
	}
	return 0;
}

int ensure_forces_checked(oxb_ctx *c, const char *what) {
	for(int attempt = 0; attempt < 12; attempt++) {
		int rc = ensure_forces(c);
		if(rc) return rc;
		rc = check_force_flags(c, what);
		if(rc != -1) return rc;
	}
	return fail(c, 8, "%s: work-list segments keep overflowing", what);
}

// One unit of the hot loop with launch index `epoch`: force pass for the pending positions, then ONE integrate launch doing
// the second half-kick + thermostat of step s and (with_first) the first half-kick + drift + rotation of step s + 1.
// step < 0: graph capture, the kernels take the step index from the device counter.
int launch_unit(oxb_ctx *c, int epoch, long long step, bool with_first) {
	int rc = launch_forces(c, OXB_FLAG_COUNT + (epoch & 1), false, step);
	if(rc) return rc;
	oxb::IntegrateArgs a = integ_args(c, step);
	const bool th_cfg = (c->th.type == OXB_THERMOSTAT_BROWNIAN || c->th.type == OXB_THERMOSTAT_LANGEVIN);
	if(c->th.type == OXB_THERMOSTAT_BUSSI && step >= 0 && thermostat_active(c, step)) {
		oxb::launch_clear_sums(c->stream, c->sums, c->flags, epoch);
		oxb::launch_integrate_epoch(c->stream, a, OXB_PH_SECOND | OXB_PH_BUSSI_SUMS | OXB_PH_COUNT_STEP, epoch);
		// the two extra launches of a Bussi step keep the halt-word parity: update reads the word integrate wrote, apply
		// reads it too and writes the one the next unit reads
		oxb::launch_bussi_update_epoch(c->stream, c->sums, c->N, c->th, step, c->flags, epoch + 1);
		a.step = step + 1; // the apply launch carries the step counter forward without counting a step
		oxb::launch_integrate_epoch(c->stream, a, with_first ? (OXB_PH_BUSSI_APPLY | OXB_PH_FIRST) : OXB_PH_BUSSI_APPLY, epoch + 1);
		c->launches += 4;
		return 2; // consumed two launch indices
	}
	int ph = OXB_PH_SECOND | OXB_PH_COUNT_STEP | (th_cfg ? OXB_PH_THERMO : 0) | (with_first ? OXB_PH_FIRST : 0);
	oxb::launch_integrate_epoch(c->stream, a, ph, epoch);
	c->launches++;
	return 0;
}

// cached graph of `units` consecutive full units starting at an even launch index, for the current state buffers
unsigned long long config_hash(const oxb_ctx *c) {
	// FNV-1a over the by-value kernel arguments a captured batch freezes: model constants, thermostat, time step, skin
	unsigned long long h = 1469598103934665603ull;
	auto mix = [&](const void *p, size_t n) {
		const unsigned char *b = (const unsigned char *) p;
		for(size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
	};
	mix(&c->model, sizeof(c->model));
	// oxRNA: `model` only mirrors what the context reads; the kernels freeze the full RNA block (stacking strengths, sequence-dependent
	// tables, mismatch repulsion ...) by value
	if(c->is_rna) { mix(&c->rmodel, sizeof(c->rmodel)); mix(&c->back_a3, sizeof(float)); }
	if(c->is_dna3) mix(&c->d3, sizeof(c->d3)); // scalars by value, tables behind stable pointers (re-uploaded in place)
	mix(&c->precision, sizeof(int));
	mix(&c->th.type, sizeof(int)); mix(&c->th.every, sizeof(int)); mix(&c->th.a, 4 * sizeof(float)); mix(&c->th.seed, sizeof(c->th.seed));
	mix(&c->dt, sizeof(double)); mix(&c->skin, sizeof(double));
	return h;
}

int batch_graph(oxb_ctx *c, int units, cudaGraphExec_t *out) {
	// temperature changes (replica exchange) alternate between a few configurations: graphs are kept per configuration
	const unsigned long long cfg = config_hash(c);
	for(auto &g : c->graphs) if(g.units == units && g.cur == c->cur && g.cfg == cfg) { *out = g.exec; return 0; }
	if(c->graphs.size() >= 48) drop_graphs(c);
	cudaGraph_t graph = nullptr;
	const long long launches0 = c->launches;
	CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
	int rc = 0;
	for(int k = 0; k < units && rc == 0; k++) rc = launch_unit(c, k, -1, true);
	cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
	c->launches = launches0; // capturing launches nothing
	if(rc) { if(graph) cudaGraphDestroy(graph); return rc; }
	if(e != cudaSuccess) return fail(c, 100 + (int) e, "graph capture failed: %s", cudaGetErrorString(e));
	cudaGraphExec_t exec = nullptr;
	e = cudaGraphInstantiate(&exec, graph, 0);
	cudaGraphDestroy(graph);
	if(e != cudaSuccess) return fail(c, 100 + (int) e, "graph instantiation failed: %s", cudaGetErrorString(e));
	c->graphs.push_back({ units, c->cur, cfg, exec });
	*out = exec;
	return 0;
}

// `n` full units for steps step0, step0 + 1, ...; `epoch` is advanced by the number of launch indices used.  Graph-launched
// in chunks of 8/4/2 units (even sizes keep the frozen parity valid) while the index is even; the rest goes out as plain
// stream launches.
int launch_full_units(oxb_ctx *c, long long n, long long step0, int &epoch) {
	const bool graphable = c->use_graphs && c->th.type != OXB_THERMOSTAT_BUSSI && c->force_cb == nullptr;
	const int per_unit = (c->use_edge ? (c->is_dna3 ? 5 : (c->fold_tails ? (c->fold_hb ? 3 : 4) : (c->precision == OXB_PRECISION_MIXED ? 6 : 5))) : 1) + (c->n_ext > 0 ? 1 : 0) + (c->n_ext_all > 0 ? 1 : 0) + (c->n_ext_com > 0 ? 1 : 0) + 1;
	long long k = 0;
	while(k < n) {
		int chunk = 0;
		if(graphable && (epoch & 1) == 0) {
			if(n - k >= 8) chunk = 8;
			else if(n - k >= 4) chunk = 4;
			else if(n - k >= 2) chunk = 2;
		}
		if(chunk > 0) {
			cudaGraphExec_t g = nullptr;
			int rc = batch_graph(c, chunk, &g);
			if(rc) return rc;
			CU(cudaGraphLaunch(g, c->stream));
			c->launches += (long long) per_unit * chunk;
			c->graph_launches++;
			epoch += chunk;
			k += chunk;
		}
		else {
			int rc = launch_unit(c, epoch, step0 + k, true);
			if(rc != 0 && rc != 2) return rc;
			epoch += (rc == 2) ? 2 : 1;
			k += 1;
		}
	}
	return 0;
}

} // namespace

// ------------------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int oxb_create(oxb_ctx **out, int device, int N, int precision) {
	if(out == nullptr || N <= 0) return 1;
	if(N >= (1 << 22)) return 1; // packed index is 22 bits, as in the reference (MD_CUDABackend.cu:243-254)
	oxb_ctx *c = new oxb_ctx();
	*out = c;
	c->device = device; c->N = N; c->precision = precision;
	c->n_per = N;
	std::memset(&c->th, 0, sizeof(c->th));
	c->th.every = 1;
	// backend_precision: both keep the FP64 state and FP32 pair arithmetic.  mixed (the reference default) additionally takes the two
	// stiff pieces of the model -- the FENE distance and every excluded-volume term that is in range -- in double (1e-5 force
	// criterion at any box size); float skips that refinement (pure FP32 pair arithmetic, 1e-4 criterion, ~5 % faster).
	if(precision != OXB_PRECISION_MIXED && precision != OXB_PRECISION_FLOAT) return fail(c, 4, "backend_precision must be mixed or float (double is not available)");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if(e != cudaSuccess || ndev == 0) return fail(c, 5, "no CUDA device available (%s): oxdna_b200 has no CPU fallback", cudaGetErrorString(e));
	CU(cudaSetDevice(device));
	CU(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device));
	CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	for(int k = 0; k < 2; k++) {
		CU(cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking));
		CU(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
	}
	CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&c->ev_near, cudaEventDisableTiming));
	CU(dalloc(&c->cur_step, 2));
	CU(cudaMemset(c->cur_step, 0, 2 * sizeof(long long)));
	{
		const char *g = getenv("OXB_NO_GRAPHS");
		c->use_graphs = !(g != nullptr && g[0] == '1');
		const char *sf = getenv("OXB_SPEC_FACTOR");
		if(sf != nullptr && atof(sf) > 0.) c->spec_factor = atof(sf);
		const char *db = getenv("OXB_DEFER_BUILD_CHECK");
		if(db != nullptr) c->defer_build_checks = (db[0] != '0');
		const char *fo = getenv("OXB_FOLD");
		if(fo != nullptr) c->fold_tails = (fo[0] != '0');
		const char *dhh = getenv("OXB_DH_HALF");
		if(dhh != nullptr) c->dh_half = (dhh[0] != '0');
		const char *cg = getenv("OXB_CLASS_GROUPS");
		if(cg != nullptr) c->class_groups = (cg[0] != '0');
		const char *nt = getenv("OXB_NEAR_TILE");
		if(nt != nullptr) c->near_tile = (nt[0] != '0');
		const char *ss = getenv("OXB_SEG_SCALE");
		if(ss != nullptr && atof(ss) > 0.) c->seg_scale = atof(ss);
		const char *fh = getenv("OXB_FOLD_HB");
		if(fh != nullptr) { c->fold_hb = (fh[0] != '0'); c->fold_hb_set = true; }
		if(!c->fold_hb_set) c->fold_hb = N < 300000;
		const char *hs = getenv("OXB_HALF_SHELL");
		if(hs != nullptr) c->half_shell_ok = (hs[0] != '0');
		const char *f = getenv("OXB_FORK");
		if(f != nullptr) c->fork_streams = (f[0] != '0');
	}
	for(int k = 0; k < 2; k++) {
		CU(dalloc(&c->posd[k], N)); CU(dalloc(&c->veld[k], N)); CU(dalloc(&c->Ld[k], N)); CU(dalloc(&c->quatd[k], N));
		CU(dalloc(&c->ipos[k], N)); CU(dalloc(&c->iback[k], N));
		CU(dalloc(&c->axf[k], 2 * (size_t) N)); CU(dalloc(&c->F[k], N)); CU(dalloc(&c->T[k], N));
		CU(dalloc(&c->bonds[k], N));
		CU(cudaMemset(c->F[k], 0, sizeof(float4) * N)); CU(cudaMemset(c->T[k], 0, sizeof(float4) * N));
	}
	CU(dalloc(&c->slot_of, N));
	CU(dalloc(&c->Fb, N));
	CU(cudaMemset(c->Fb, 0, sizeof(float4) * N));
	CU(dalloc(&c->flags, OXB_FLAG_ALLOC));
	CU(cudaMemset(c->flags, 0, sizeof(int) * OXB_FLAG_ALLOC));
	CU(cudaMallocHost((void **) &c->h_flags, sizeof(int) * OXB_FLAG_WORDS));
	CU(dalloc(&c->sums, 1));
	CU(cudaMemset(c->sums, 0, sizeof(KinSums)));
	CU(dalloc(&c->d_energy, 16));
	CU(cudaMallocHost((void **) &c->h_scalars, sizeof(double) * 16));
	return 0;
}

void oxb_destroy(oxb_ctx *c) {
	if(c == nullptr) return;
	cudaSetDevice(c->device);
	if(c->stream) cudaStreamSynchronize(c->stream);
	drop_graphs(c);
	for(int k = 0; k < 2; k++) {
		if(c->aux[k]) { cudaStreamSynchronize(c->aux[k]); cudaStreamDestroy(c->aux[k]); }
		if(c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
	}
	if(c->ev_fork) cudaEventDestroy(c->ev_fork);
	if(c->ev_near) cudaEventDestroy(c->ev_near);
	if(c->ev_wait) cudaEventDestroy(c->ev_wait);
	cudaFree(c->cur_step);
	for(int k = 0; k < 2; k++) {
		cudaFree(c->posd[k]); cudaFree(c->veld[k]); cudaFree(c->Ld[k]); cudaFree(c->quatd[k]); cudaFree(c->ipos[k]);
		cudaFree(c->axf[k]); cudaFree(c->F[k]); cudaFree(c->T[k]); cudaFree(c->bonds[k]); cudaFree(c->iback[k]);
	}
	cudaFree(c->Fb);
	cudaFree(c->rep); cudaFree(c->d_rep_energy);
	cudaFree(c->d3_buf); cudaFree(c->d3_tcode);
	if(c->h_rep_energy) cudaFreeHost(c->h_rep_energy);
	cudaFree(c->slot_of); cudaFree(c->flags); cudaFree(c->sums); cudaFree(c->d_energy); cudaFree(c->ext); cudaFree(c->ext_all); cudaFree(c->ext_com); cudaFree(c->ext_pool); cudaFree(c->ext_grid);
	cudaFree(c->d_topo); cudaFree(c->d_stage); cudaFree(c->d_marshal_err);
	cudaFree(c->mol_of); cudaFree(c->mol_inv_size); cudaFree(c->mol_coms); cudaFree(c->pos_backup); cudaFree(c->pos_f4); cudaFree(c->quat_f4);
	cudaFree(c->hkeys); cudaFree(c->hkeys_sorted); cudaFree(c->hvals); cudaFree(c->hvals_sorted); cudaFree(c->hinv); cudaFree(c->sort_small);
	free_lists(c);
	if(c->h_flags) cudaFreeHost(c->h_flags);
	if(c->h_scalars) cudaFreeHost(c->h_scalars);
	if(c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

const char *oxb_last_error(const oxb_ctx *c) { return c ? c->err.c_str() : "null context"; }

int oxb_set_stream(oxb_ctx *c, void *s) {
	if(c == nullptr) return 1;
	if(c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
	c->stream = (cudaStream_t) s;
	c->own_stream = false;
	drop_graphs(c);
	return 0;
}

int oxb_set_box(oxb_ctx *c, const double box[3]) {
	if(c == nullptr || box == nullptr) return 1;
	for(int k = 0; k < 3; k++) {
		if(!(box[k] > 0)) return fail(c, 1, "box sides must be positive");
		c->box[k] = box[k];
	}
	set_boxf(c);
	drop_graphs(c);
	c->have_box = true;
	c->lists_valid = false; c->forces_valid = false;
	if(c->lists_allocated) free_lists(c);
	return 0;
}

int oxb_set_topology(oxb_ctx *c, const int *btype, const int *n3, const int *n5, const int *strand) {
	if(c == nullptr || btype == nullptr || n3 == nullptr || n5 == nullptr) return 1;
	const int N = c->N;
	c->h_btype.assign(btype, btype + N); c->h_n3.assign(n3, n3 + N); c->h_n5.assign(n5, n5 + N);
	if(strand) c->h_strand.assign(strand, strand + N); else c->h_strand.assign(N, 0);
	for(int i = 0; i < N; i++) {
		if(n3[i] >= N || n5[i] >= N) return fail(c, 1, "wrong topology for particle %d (neighbour index out of range)", i);
		if(btype[i] > 511 || btype[i] < -511) return fail(c, 1, "base type of particle %d does not fit the packed word (|btype| <= 511)", i);
		if(c->n_rep > 1 && ((n3[i] >= 0 && n3[i] / c->n_per != i / c->n_per) || (n5[i] >= 0 && n5[i] / c->n_per != i / c->n_per)))
			return fail(c, 1, "replica batching: particle %d is bonded across replicas", i);
	}
	{
		std::vector<int4> ht(N);
		for(int i = 0; i < N; i++) ht[i] = make_int4(btype[i], n3[i], n5[i], c->h_strand[i]);
		if(c->d_topo == nullptr) CU(dalloc(&c->d_topo, (size_t) N));
		CU(cudaMemcpy(c->d_topo, ht.data(), sizeof(int4) * N, cudaMemcpyHostToDevice));
	}
	c->have_topology = true;
	c->d3_tcode_valid = false;
	c->have_state = false;
	cudaFree(c->mol_of); cudaFree(c->mol_inv_size); cudaFree(c->mol_coms); cudaFree(c->pos_backup);
	c->mol_of = nullptr; c->mol_inv_size = nullptr; c->mol_coms = nullptr; c->pos_backup = nullptr;
	c->trial_open = false;
	return 0;
}

int oxb_set_model_dna2(oxb_ctx *c, const oxb_dna2_params *P, double rcut) {
	if(c == nullptr || P == nullptr) return 1;
	if(c->is_rna || c->is_dna3) { drop_graphs(c); c->lists_valid = false; }
	c->is_dna3 = false;
	c->model = *P;
	c->is_rna = false;
	c->back_a3 = 0.f;
	c->rcut = rcut;
	c->have_model = true;
	c->forces_valid = false;
	// Lists built for radii that are not smaller than the new model's stay valid for forces and energies (every pair inside
	// the new cutoffs is listed, the kernels re-test the distances): a replica-exchange energy evaluation under the partner's
	// colder Hamiltonian costs one force pass, no rebuild and no reallocation.  oxb_update_lists / oxb_get_pairs rebuild for
	// the exact radius when it differs.
	if(c->lists_valid && (rcut + 2. * c->skin > c->lists_rv || (double) P->dh_rc > c->lists_dh_rc)) c->lists_valid = false;
	return 0;
}

int oxb_set_model_rna2(oxb_ctx *c, const oxb_rna2_params *P, double rcut) {
	if(c == nullptr || P == nullptr) return 1;
	if(!c->is_rna && c->have_model) { drop_graphs(c); c->lists_valid = false; }
	c->is_dna3 = false;
	c->rmodel = *P;
	c->is_rna = true;
	// the subset the context reads (list radii, site offsets)
	oxb_dna2_params &M = c->model;
	std::memset(&M, 0, sizeof(M));
	M.back_a1 = P->back_a1; M.back_a2 = P->back_a2; c->back_a3 = P->back_a3;
	M.base_a1 = P->base_a1; M.stack_a1 = P->stack_a1;
	for(int k = 0; k < 4; k++) M.excl[k] = P->excl[k];
	M.hb = P->hb; M.crst = P->crst; M.cxst = P->cxst;
	M.dh_rc = P->dh_rc; M.rcut = P->rcut; M.rcut_near = P->rcut_near;
	c->rcut = rcut;
	c->have_model = true;
	c->forces_valid = false;
	if(c->lists_valid && (rcut + 2. * c->skin > c->lists_rv || (double) P->dh_rc > c->lists_dh_rc)) c->lists_valid = false;
	return 0;
}

int oxb_set_model_dna3(oxb_ctx *c, const double *tab, const oxb_dna3_scalars *S) {
	if(c == nullptr || tab == nullptr || S == nullptr) return 1;
	if(c->n_rep > 1) return fail(c, 1, "replica batching is not available for oxDNA3");
	cudaSetDevice(c->device);
	CU(cudaStreamSynchronize(c->stream));
	const bool entering = !c->is_dna3;
	if(entering && c->have_model) { drop_graphs(c); c->lists_valid = false; }
	std::vector<float> h;
	oxb_dna3_dev &D = c->d3;
	const int *keep_tcode = c->d3_tcode_valid ? c->d3_tcode : nullptr;
	size_t off[5];
	dna3_pack(tab, S, h, D, off);
	if(c->d3_buf == nullptr) CU(dalloc(&c->d3_buf, h.size() / 4));
	CU(cudaMemcpy(c->d3_buf, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice));
	D.bonded = c->d3_buf + off[0]; D.crst = c->d3_buf + off[1]; D.cxst = c->d3_buf + off[2]; D.hb = c->d3_buf + off[3]; D.nexcl = c->d3_buf + off[4];
	D.tcode = keep_tcode;
	// the subset the context itself reads (list radii, site offsets for the integrator's fixed-point sites)
	oxb_dna2_params &M = c->model;
	std::memset(&M, 0, sizeof(M));
	M.back_a1 = D.back_a1; M.back_a2 = D.back_a2; c->back_a3 = 0.f;
	M.base_a1 = 0.43f; M.stack_a1 = 0.37f; M.backref_a1 = D.backref_a1;
	M.dh_rc = D.dh_rc; M.rcut = (float) S->rcut;
	// thresholds of the list builder's near-edge selection (list_args): the longest range of any tetramer per family of site pairs.  The
	// builder places every base site at 0.43 a1 and every stacking site at 0.37 a1; the true offsets are 0.37 ... 0.43 (0.34 for a dummy
	// base): a base-base distance is off by at most 0.12, a base-backbone or stack-stack distance by at most 0.06
	M.rcut_near = std::min((float) S->rcut, std::sqrt(D.r2_near_max));
	M.excl[0].rc = D.range_bb; M.excl[1].rc = D.range_eb + 0.12f; M.excl[2].rc = M.excl[3].rc = D.range_bk + 0.06f;
	M.hb.rchigh = M.crst.rchigh = std::sqrt(D.r2_base_max) + 0.12f;
	M.cxst.rchigh = std::sqrt(D.r2_stack_max) + 0.06f;
	c->is_rna = false;
	c->is_dna3 = true;
	c->lists_valid = false;
	c->rcut = S->rcut;
	c->have_model = true;
	c->forces_valid = false;
	if(c->lists_valid && (S->rcut + 2. * c->skin > c->lists_rv || (double) D.dh_rc > c->lists_dh_rc)) c->lists_valid = false;
	return 0;
}

int oxb_set_force_callback(oxb_ctx *c, oxb_force_callback fn, void *user, double rcut) {
	if(c == nullptr) return 1;
	if(fn != nullptr && !(rcut > 0.)) return fail(c, 1, "oxb_set_force_callback: the interaction cutoff must be positive");
	if(fn != nullptr && c->n_rep > 1) return fail(c, 1, "a plugged-in interaction cannot be combined with replica batching");
	cudaSetDevice(c->device);
	CU(cudaStreamSynchronize(c->stream));
	drop_graphs(c);
	c->force_cb = fn; c->force_cb_user = user;
	c->forces_valid = false; c->lists_valid = false;
	if(fn != nullptr) {
		// the context only needs the Verlet radius: an all-zero built-in block with this cutoff (no Debye-Hueckel rows, no site geometry)
		std::memset(&c->model, 0, sizeof(c->model));
		c->model.rcut = (float) rcut; c->model.rcut_near = (float) rcut;
		c->is_rna = false; c->back_a3 = 0.f;
		c->rcut = rcut;
		c->have_model = true;
	}
	else c->have_model = false;
	return 0;
}

int oxb_set_replicas(oxb_ctx *c, int n_replicas) {
	if(c == nullptr) return 1;
	if(n_replicas < 1 || c->N % n_replicas != 0) return fail(c, 1, "the number of particles (%d) is not a multiple of the number of replicas (%d)", c->N, n_replicas);
	cudaSetDevice(c->device);
	const int n_per = c->N / n_replicas;
	if(c->have_topology) {
		for(int i = 0; i < c->N; i++) {
			if((c->h_n3[i] >= 0 && c->h_n3[i] / n_per != i / n_per) || (c->h_n5[i] >= 0 && c->h_n5[i] / n_per != i / n_per))
				return fail(c, 1, "replica batching: particle %d is bonded across replicas", i);
		}
	}
	CU(cudaStreamSynchronize(c->stream));
	drop_graphs(c);
	cudaFree(c->rep); cudaFree(c->d_rep_energy);
	if(c->h_rep_energy) cudaFreeHost(c->h_rep_energy);
	c->rep = nullptr; c->d_rep_energy = nullptr; c->h_rep_energy = nullptr;
	c->n_rep = n_replicas; c->n_per = n_per;
	c->have_rep_consts = false;
	if(n_replicas > 1) {
		CU(dalloc(&c->rep, (size_t) n_replicas));
		CU(dalloc(&c->d_rep_energy, (size_t) n_replicas));
		CU(cudaMallocHost((void **) &c->h_rep_energy, sizeof(double) * (size_t) n_replicas));
	}
	c->lists_valid = false; c->forces_valid = false;
	if(c->lists_allocated) free_lists(c);
	return 0;
}

int oxb_set_replica_consts(oxb_ctx *c, int n, const oxb_replica_consts *rows) {
	if(c == nullptr || rows == nullptr) return 1;
	if(c->n_rep < 2 || n != c->n_rep) return fail(c, 1, "oxb_set_replica_consts: %d rows for %d replicas (call oxb_set_replicas first)", n, c->n_rep);
	if(!c->have_model) return fail(c, 2, "the interaction model (hottest replica: list radii) must be set before the replica constants");
	cudaSetDevice(c->device);
	for(int r = 0; r < n; r++) {
		// the lists are built for the radii of the model block: no replica may reach further
		if(rows[r].dh_rc > c->model.dh_rc * (1.f + 1e-6f) || rows[r].rcut2 > (float) (c->rcut * c->rcut) * (1.f + 1e-6f))
			return fail(c, 1, "replica %d: Debye-Hueckel range %g / cutoff %g exceed those of the model block (%g / %g): set the model of the hottest temperature", r,
					(double) rows[r].dh_rc, std::sqrt((double) rows[r].rcut2), (double) c->model.dh_rc, c->rcut);
	}
	// stream-ordered: kernels already queued keep the old table
	CU(cudaMemcpyAsync(c->rep, rows, sizeof(oxb_replica_consts) * (size_t) n, cudaMemcpyHostToDevice, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	c->have_rep_consts = true;
	c->forces_valid = false;
	return 0;
}

int oxb_replica_energies(oxb_ctx *c, double *U) {
	if(c == nullptr || U == nullptr) return 1;
	if(c->n_rep < 2) return oxb_energy(c, U, nullptr);
	int rc = ensure_forces_checked(c, "replica_energies");
	if(rc) return rc;
	const int k = c->cur;
	oxb::launch_energy_sum_replicas(c->stream, c->N, c->n_rep, c->n_per, c->F[k], c->use_edge ? c->Fb : nullptr, c->d_rep_energy);
	c->launches += 1;
	CU(cudaMemcpyAsync(c->h_rep_energy, c->d_rep_energy, sizeof(double) * (size_t) c->n_rep, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	for(int r = 0; r < c->n_rep; r++) U[r] = 0.5 * c->h_rep_energy[r];
	return 0;
}

int oxb_set_lists(oxb_ctx *c, double verlet_skin, int use_edge, int sort_every, double max_density_multiplier) {
	if(c == nullptr) return 1;
	if(!(verlet_skin > 0)) return fail(c, 1, "verlet_skin must be > 0");
	c->skin = verlet_skin; c->use_edge = use_edge ? 1 : 0; c->sort_every = sort_every < 0 ? 0 : sort_every;
	c->max_density_multiplier = max_density_multiplier;
	c->lists_valid = false; c->forces_valid = false;
	if(c->lists_allocated) free_lists(c);
	return 0;
}

int oxb_set_dt(oxb_ctx *c, double dt) {
	if(c == nullptr) return 1;
	c->dt = dt;
	return 0;
}

int oxb_set_thermostat(oxb_ctx *c, int type, int every, double a, double b, double cc, double d, unsigned long long seed) {
	if(c == nullptr) return 1;
	if(type < OXB_THERMOSTAT_NONE || type > OXB_THERMOSTAT_BUSSI) return fail(c, 1, "unknown thermostat %d", type);
	if(type != OXB_THERMOSTAT_NONE && every < 1) return fail(c, 1, "'newtonian_steps' must be > 0");
	c->th.type = type; c->th.every = every < 1 ? 1 : every;
	c->th.a = (float) a; c->th.b = (float) b; c->th.c = (float) cc; c->th.d = (float) d; c->th.seed = seed;
	c->bussi_init = false;
	return 0;
}

int oxb_set_ext_forces(oxb_ctx *c, int n, const oxb_ext_force *f) {
	if(c == nullptr || n < 0 || (n > 0 && f == nullptr)) return 1;
	std::vector<DevExtForce> h, hall, hcom;
	for(int k = 0; k < n; k++) {
		if(f[k].type < OXB_EXT_STRING || f[k].type >= OXB_EXT_NTYPES) return fail(c, 1, "external force %d: unsupported type %d", k, f[k].type);
		if(f[k].particle < -1 || f[k].particle >= c->N) return fail(c, 1, "external force %d: invalid particle %d", k, f[k].particle);
		if(f[k].type == OXB_EXT_MUTUAL_TRAP && (f[k].ref < 0 || f[k].ref >= c->N)) return fail(c, 1, "Invalid reference particle %d for Mutual Trap", f[k].ref);
		if(f[k].type == OXB_EXT_MUTUAL_TRAP && f[k].particle < 0) return fail(c, 1, "external force %d: a mutual trap needs one particle", k);
		if(f[k].type == OXB_EXT_LJ_WALL && (f[k].iaux % 2 != 0 || f[k].iaux <= 0)) return fail(c, 1, "LJWall: n (%d) should be an even integer. Aborting", f[k].iaux);
		if(f[k].type == OXB_EXT_LJ_CONE && (f[k].iaux % 2 != 0 || f[k].iaux <= 0)) return fail(c, 1, "RepulsiveCone: n (%d) should be an even integer. Aborting", f[k].iaux);
		if(f[k].type == OXB_EXT_REPULSION_PLANE_MOVING && (f[k].ref < 0 || f[k].iaux < f[k].ref || f[k].iaux >= c->N))
			return fail(c, 1, "RepulsionPlaneMoving requires the list of ref_particle indices to be contiguous (got %d..%d)", f[k].ref, f[k].iaux);
		if(f[k].type == OXB_EXT_META_COM_TRAP) {
			const long long ng = (long long) f[k].aux[2], go = (long long) f[k].aux[4];
			const int mode = (int) f[k].aux[3];
			if(mode != 1 && mode != 2) return fail(c, 1, "LTCOMTrap: unsupported mode '%d' (should be '1' or '2')", mode);
			if(ng < 2 || go < 0 || go + ng > (long long) c->ext_grid_n || !(f[k].aux[1] > 0.))
				return fail(c, 1, "external force %d: potential grid [%lld, +%lld) outside the grid pool (%zu values; call oxb_set_ext_grid_pool first)", k, go, ng, c->ext_grid_n);
		}
		if(f[k].type == OXB_EXT_META_COORDINATION) {
			const long long ng = (long long) f[k].aux[2], go = (long long) f[k].aux[4], lo = f[k].ref, np_ = f[k].iaux;
			const int mode = (int) f[k].aux[3];
			if(mode < 0 || mode > 2) return fail(c, 1, "Coordination: unknown coordination_type %d (0 hb_cutoff, 1 switching_function, 2 mixed)", mode);
			if(mode != 0 && (f[k].pbc % 2) != 0) return fail(c, 1, "LTCoordination: exponent n must be an even integer");
			if(ng < 2 || go < 0 || go + ng > (long long) c->ext_grid_n || !(f[k].aux[1] > 0.))
				return fail(c, 1, "external force %d: potential grid [%lld, +%lld) outside the grid pool (%zu values; call oxb_set_ext_grid_pool first)", k, go, ng, c->ext_grid_n);
			if(lo < 0 || np_ < 1 || lo + 2 * np_ > (long long) c->ext_pool_h.size())
				return fail(c, 1, "external force %d: pair list [%lld, +2 x %lld) outside the index pool (%zu entries)", k, lo, np_, c->ext_pool_h.size());
		}
		if(f[k].type == OXB_EXT_COM || f[k].type == OXB_EXT_META_COM_TRAP) {
			const long long lo = f[k].ref, n_com = f[k].iaux, n_ref = f[k].pbc;
			if(lo < 0 || n_com < 1 || n_ref < 1 || lo + n_com + n_ref > (long long) c->ext_pool_h.size())
				return fail(c, 1, "external force %d: COM force index lists [%lld, +%lld, +%lld) outside the index pool (%zu entries; call oxb_set_ext_index_pool first)",
						k, lo, n_com, n_ref, c->ext_pool_h.size());
		}
		DevExtForce d;
		std::memset(&d, 0, sizeof(d));
		d.type = f[k].type; d.particle = f[k].particle; d.ref = f[k].ref < 0 ? 0 : f[k].ref; d.pbc = f[k].pbc;
		d.stiff = (float) f[k].stiff; d.r0 = (float) f[k].r0; d.rate = (float) f[k].rate; d.stiff_rate = (float) f[k].stiff_rate; d.F0 = (float) f[k].F0;
		double nrm = std::sqrt(f[k].dir[0] * f[k].dir[0] + f[k].dir[1] * f[k].dir[1] + f[k].dir[2] * f[k].dir[2]);
		for(int x = 0; x < 3; x++) {
			d.dir[x] = (float) ((f[k].type != OXB_EXT_MUTUAL_TRAP && nrm > 0) ? f[k].dir[x] / nrm : f[k].dir[x]); // unit direction / axis
			d.pos0[x] = f[k].pos0[x];
		}
		for(int x = 0; x < 8; x++) d.aux[x] = (float) f[k].aux[x];
		d.iaux = f[k].iaux;
		d.r0d = f[k].r0;
		if(f[k].type == OXB_EXT_LJ_CONE) { d.aux[3] = (float) std::sin(f[k].aux[2]); d.aux[4] = (float) std::cos(f[k].aux[2]); d.aux[5] = (float) std::tan(f[k].aux[2]); }
		if(f[k].type == OXB_EXT_SPHERE_MOVING) d.daux = f[k].aux[4];
		if(f[k].type == OXB_EXT_COM || f[k].type == OXB_EXT_META_COM_TRAP || f[k].type == OXB_EXT_META_COORDINATION) { d.ref = f[k].ref; hcom.push_back(d); continue; }
		(f[k].particle < 0 ? hall : h).push_back(d);
	}
	cudaFree(c->ext);
	cudaFree(c->ext_all);
	cudaFree(c->ext_com);
	c->ext = c->ext_all = c->ext_com = nullptr;
	CU(dalloc(&c->ext_com, std::max<size_t>(hcom.size(), 1)));
	if(!hcom.empty()) CU(cudaMemcpy(c->ext_com, hcom.data(), sizeof(DevExtForce) * hcom.size(), cudaMemcpyHostToDevice));
	c->n_ext_com = (int) hcom.size();
	CU(dalloc(&c->ext, std::max<size_t>(h.size(), 1)));
	CU(dalloc(&c->ext_all, std::max<size_t>(hall.size(), 1)));
	if(!h.empty()) CU(cudaMemcpy(c->ext, h.data(), sizeof(DevExtForce) * h.size(), cudaMemcpyHostToDevice));
	if(!hall.empty()) CU(cudaMemcpy(c->ext_all, hall.data(), sizeof(DevExtForce) * hall.size(), cudaMemcpyHostToDevice));
	c->n_ext = (int) h.size();
	c->n_ext_all = (int) hall.size();
	c->forces_valid = false;
	drop_graphs(c);
	return 0;
}

int oxb_set_ext_grid_pool(oxb_ctx *c, int n, const double *values) {
	if(c == nullptr || n < 0 || (n > 0 && values == nullptr)) return 1;
	if(c->n_ext_com > 0) return fail(c, 1, "the grid pool cannot change while forces that refer to it are set");
	std::vector<float> h(values, values + n);
	cudaFree(c->ext_grid);
	c->ext_grid = nullptr;
	CU(dalloc(&c->ext_grid, std::max<size_t>((size_t) n, 1)));
	if(n > 0) CU(cudaMemcpy(c->ext_grid, h.data(), sizeof(float) * (size_t) n, cudaMemcpyHostToDevice));
	c->ext_grid_n = (size_t) n;
	return 0;
}

int oxb_set_ext_index_pool(oxb_ctx *c, int n, const int *idx) {
	if(c == nullptr || n < 0 || (n > 0 && idx == nullptr)) return 1;
	for(int k = 0; k < n; k++) if(idx[k] < 0 || idx[k] >= c->N) return fail(c, 1, "external-force index pool: invalid particle %d", idx[k]);
	if(c->n_ext_com > 0) return fail(c, 1, "the index pool cannot change while COM forces that refer to it are set");
	c->ext_pool_h.assign(idx, idx + n);
	cudaFree(c->ext_pool);
	c->ext_pool = nullptr;
	CU(dalloc(&c->ext_pool, std::max<size_t>((size_t) n, 1)));
	if(n > 0) CU(cudaMemcpy(c->ext_pool, idx, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice));
	return 0;
}

int oxb_set_state(oxb_ctx *c, const double *pos, const double *a1, const double *a3, const double *vel, const double *L) {
	if(c == nullptr || pos == nullptr || a1 == nullptr || a3 == nullptr) return 1;
	if(!c->have_topology) return fail(c, 2, "topology must be set before the state");
	if(!c->have_box) return fail(c, 2, "box must be set before the state");
	if(!c->have_model) return fail(c, 2, "interaction model must be set before the state (it fixes the backbone-site geometry)");
	// flat arrays -> device staging (straight from the caller's buffers: full PCIe rate when they are pinned), conversion on the device
	const size_t N = (size_t) c->N, n3d = 3 * N;
	if(c->d_stage == nullptr) CU(dalloc(&c->d_stage, 15 * N));
	if(c->d_marshal_err == nullptr) CU(dalloc(&c->d_marshal_err, (size_t) 1));
	double *dp = c->d_stage, *d1 = dp + n3d, *d3 = d1 + n3d, *dv = d3 + n3d, *dL = dv + n3d;
	CU(cudaMemcpyAsync(dp, pos, sizeof(double) * n3d, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(d1, a1, sizeof(double) * n3d, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(d3, a3, sizeof(double) * n3d, cudaMemcpyHostToDevice, c->stream));
	if(vel) CU(cudaMemcpyAsync(dv, vel, sizeof(double) * n3d, cudaMemcpyHostToDevice, c->stream));
	if(L) CU(cudaMemcpyAsync(dL, L, sizeof(double) * n3d, cudaMemcpyHostToDevice, c->stream));
	int no_err = 0x7fffffff, first_bad = 0;
	CU(cudaMemcpyAsync(c->d_marshal_err, &no_err, sizeof(int), cudaMemcpyHostToDevice, c->stream));
	const int k = c->cur;
	oxb::MarshalArgs a;
	a.N = c->N; a.pos = dp; a.a1 = d1; a.a3 = d3; a.vel = vel ? dv : nullptr; a.L = L ? dL : nullptr;
	a.topo = c->d_topo;
	for(int d = 0; d < 3; d++) a.box_inv[d] = 1. / c->box[d];
	a.back_a1 = c->model.back_a1; a.back_a2 = c->model.back_a2; a.back_a3 = c->back_a3;
	a.posd = c->posd[k]; a.veld = c->veld[k]; a.Ld = c->Ld[k]; a.quatd = c->quatd[k];
	a.ipos = c->ipos[k]; a.iback = c->iback[k]; a.axf = c->axf[k]; a.bonds = c->bonds[k]; a.slot_of = c->slot_of;
	a.err = c->d_marshal_err;
	oxb::launch_state_in(c->stream, a);
	c->launches++;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(&first_bad, c->d_marshal_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	if(first_bad != no_err) {
		c->have_state = false;
		return fail(c, 1, "Invalid orientation for particle %d: at least one of the vectors is a null vector", first_bad);
	}
	c->have_state = true;
	c->lists_valid = false; c->forces_valid = false; c->mid_step = false;
	c->trial_open = false;
	c->slots_cell_ordered = false; // the slots were rewritten in identity order
	c->error_flags = 0;            // errors (a broken FENE bond ...) belonged to the previous state
	return 0;
}

int oxb_get_state(oxb_ctx *c, double *pos, double *a1, double *a3, double *vel, double *L) {
	if(c == nullptr) return 1;
	if(!c->have_state) return fail(c, 2, "state not set");
	// slot order -> original order and quaternion -> a1 / a3 on the device, then flat copies straight into the caller's buffers
	const size_t N = (size_t) c->N, n3d = 3 * N;
	const int k = c->cur;
	if(c->d_stage == nullptr) CU(dalloc(&c->d_stage, 15 * N));
	double *dp = c->d_stage, *d1 = dp + n3d, *d3 = d1 + n3d, *dv = d3 + n3d, *dL = dv + n3d;
	oxb::launch_state_out(c->stream, c->N, c->ipos[k], c->posd[k], c->veld[k], c->Ld[k], c->quatd[k], pos ? dp : nullptr, a1 ? d1 : nullptr,
			a3 ? d3 : nullptr, vel ? dv : nullptr, L ? dL : nullptr);
	c->launches++;
	CU(cudaGetLastError());
	if(pos) CU(cudaMemcpyAsync(pos, dp, sizeof(double) * n3d, cudaMemcpyDeviceToHost, c->stream));
	if(a1) CU(cudaMemcpyAsync(a1, d1, sizeof(double) * n3d, cudaMemcpyDeviceToHost, c->stream));
	if(a3) CU(cudaMemcpyAsync(a3, d3, sizeof(double) * n3d, cudaMemcpyDeviceToHost, c->stream));
	if(vel) CU(cudaMemcpyAsync(vel, dv, sizeof(double) * n3d, cudaMemcpyDeviceToHost, c->stream));
	if(L) CU(cudaMemcpyAsync(L, dL, sizeof(double) * n3d, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

int oxb_write_conf(oxb_ctx *c, const char *path, int append, int print_momenta) {
	if(c == nullptr || path == nullptr) return 1;
	double U = 0., K = 0.;
	int rc = oxb_energy(c, &U, &K);
	if(rc) return rc;
	const int N = c->N;
	std::vector<double> pos(3 * (size_t) N), a1(3 * (size_t) N), a3(3 * (size_t) N), vel(3 * (size_t) N), L(3 * (size_t) N);
	rc = oxb_get_state(c, pos.data(), a1.data(), a3.data(), vel.data(), L.data());
	if(rc) return rc;
	FILE *f = std::fopen(path, append ? "a" : "w");
	if(f == nullptr) return fail(c, 9, "cannot open '%s' for writing", path);
	std::vector<char> buf(1 << 22);
	std::setvbuf(f, buf.data(), _IOFBF, buf.size());
	// header and particle lines exactly as Configuration::_headers / _particle print them (stream precision 15 = %.15g)
	std::fprintf(f, "t = %lld\nb = %.15g %.15g %.15g\nE = %.15g %.15g %.15g\n", c->step, c->box[0], c->box[1], c->box[2], (U + K) / N, U / N, K / N);
	for(int i = 0; i < N; i++) {
		const double *p = &pos[3 * (size_t) i], *x = &a1[3 * (size_t) i], *z = &a3[3 * (size_t) i], *v = &vel[3 * (size_t) i], *l = &L[3 * (size_t) i];
		if(print_momenta) {
			std::fprintf(f, "%.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g\n", p[0], p[1], p[2], x[0], x[1], x[2], z[0],
					z[1], z[2], v[0], v[1], v[2], l[0], l[1], l[2]);
		}
		else std::fprintf(f, "%.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g %.15g\n", p[0], p[1], p[2], x[0], x[1], x[2], z[0], z[1], z[2]);
	}
	std::fclose(f);
	return 0;
}

int oxb_write_conf_binary(oxb_ctx *c, const char *path, int append, const unsigned short rng_state[3], const int *pos_shift) {
	if(c == nullptr || path == nullptr) return 1;
	double U = 0., K = 0.;
	int rc = oxb_energy(c, &U, &K);
	if(rc) return rc;
	const int N = c->N;
	std::vector<double> pos(3 * (size_t) N), a1(3 * (size_t) N), a3(3 * (size_t) N), vel(3 * (size_t) N), L(3 * (size_t) N);
	rc = oxb_get_state(c, pos.data(), a1.data(), a3.data(), vel.data(), L.data());
	if(rc) return rc;
	FILE *f = std::fopen(path, append ? "ab" : "wb");
	if(f == nullptr) return fail(c, 9, "cannot open '%s' for writing", path);
	std::vector<char> buf(1 << 22);
	std::setvbuf(f, buf.data(), _IOFBF, buf.size());
	// BinaryConfiguration::_headers / _configuration (src/Observables/Configurations/BinaryConfiguration.cpp:20-92): step, rng state,
	// box, E U K per particle; then per particle pos, pos_shift, the three axes as rows, vel, L -- all native-endian, doubles and ints
	const long long step = c->step;
	const unsigned short zero_seed[3] = { 0, 0, 0 };
	const double e[3] = { (U + K) / N, U / N, K / N };
	std::fwrite(&step, sizeof(long long), 1, f);
	std::fwrite(rng_state ? rng_state : zero_seed, sizeof(unsigned short), 3, f);
	std::fwrite(c->box, sizeof(double), 3, f);
	std::fwrite(e, sizeof(double), 3, f);
	for(int i = 0; i < N; i++) {
		const double *x = &a1[3 * (size_t) i], *z = &a3[3 * (size_t) i];
		const double y[3] = { z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0] }; // a2 = a3 x a1
		const int none[3] = { 0, 0, 0 };
		std::fwrite(&pos[3 * (size_t) i], sizeof(double), 3, f);
		std::fwrite(pos_shift ? pos_shift + 3 * (size_t) i : none, sizeof(int), 3, f);
		std::fwrite(x, sizeof(double), 3, f);
		std::fwrite(y, sizeof(double), 3, f);
		std::fwrite(z, sizeof(double), 3, f);
		std::fwrite(&vel[3 * (size_t) i], sizeof(double), 3, f);
		std::fwrite(&L[3 * (size_t) i], sizeof(double), 3, f);
	}
	std::fclose(f);
	return 0;
}

int oxb_set_step(oxb_ctx *c, long long step) {
	if(c == nullptr) return 1;
	c->step = step;
	c->forces_valid = c->forces_valid && (c->n_ext == 0) && (c->n_ext_all == 0) && (c->n_ext_com == 0);
	return 0;
}

long long oxb_get_step(const oxb_ctx *c) { return c ? c->step : -1; }

int oxb_sort(oxb_ctx *c) {
	if(c == nullptr) return 1;
	int rc = check_ready(c);
	if(rc) return rc;
	return do_sort(c);
}

int oxb_update_lists(oxb_ctx *c) {
	if(c == nullptr) return 1;
	int rc = check_ready(c);
	if(rc) return rc;
	c->lists_valid = false;
	return ensure_lists(c);
}

int oxb_compute_forces(oxb_ctx *c) {
	if(c == nullptr) return 1;
	c->forces_valid = false;
	return ensure_forces_checked(c, "compute_forces");
}

int oxb_first_step(oxb_ctx *c) {
	if(c == nullptr) return 1;
	int rc = ensure_forces_checked(c, "first_step");
	if(rc) return rc;
	rc = reset_batch_flags(c);
	if(rc) return rc;
	oxb::launch_integrate_epoch(c->stream, integ_args(c, c->step), OXB_PH_FIRST, 0);
	c->launches++;
	rc = read_flags(c);
	if(rc) return rc;
	if(c->h_flags[OXB_FLAG_COUNT] || c->h_flags[OXB_FLAG_COUNT + 1]) c->lists_valid = false;
	c->forces_valid = false;
	c->mid_step = true;
	return 0;
}

int oxb_second_step(oxb_ctx *c) {
	if(c == nullptr) return 1;
	int rc = ensure_forces_checked(c, "second_step");
	if(rc) return rc;
	rc = reset_batch_flags(c);
	if(rc) return rc;
	oxb::launch_integrate_epoch(c->stream, integ_args(c, c->step), OXB_PH_SECOND | OXB_PH_COUNT_STEP, 0);
	c->launches++;
	c->mid_step = false;
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

int oxb_thermostat(oxb_ctx *c) {
	if(c == nullptr) return 1;
	int rc = check_ready(c);
	if(rc) return rc;
	if(!thermostat_active(c, c->step)) return 0;
	rc = reset_batch_flags(c);
	if(rc) return rc;
	oxb::IntegrateArgs a = integ_args(c, c->step);
	if(c->th.type == OXB_THERMOSTAT_BUSSI) {
		if(!c->bussi_init) { rc = init_bussi(c); if(rc) return rc; }
		oxb::launch_kinetic_sums(c->stream, c->N, c->veld[c->cur], c->Ld[c->cur], c->sums);
		oxb::launch_bussi_update_epoch(c->stream, c->sums, c->N, c->th, c->step, c->flags, 0);
		oxb::launch_integrate_epoch(c->stream, a, OXB_PH_BUSSI_APPLY, 0);
		c->launches += 4;
	}
	else {
		oxb::launch_integrate_epoch(c->stream, a, OXB_PH_THERMO, 0);
		c->launches++;
	}
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

int oxb_run(oxb_ctx *c, long long n_steps) {
	if(c == nullptr || n_steps < 0) return 1;
	if(n_steps == 0) return 0;
	int rc = check_ready(c);
	if(rc) return rc;
	if(c->th.type == OXB_THERMOSTAT_BUSSI && !c->bussi_init) { rc = init_bussi(c); if(rc) return rc; }
	if(c->build_unchecked) {
		// a previous run ended (on an error) between an unchecked rebuild and the batch that would have checked it: rebuild on the checked path
		c->build_unchecked = false;
		c->lists_valid = false;
	}
	if(c->profiling) { k_prof_mark<<<1, 1, 0, c->stream>>>(c->flags, OXB_PROF_OTHER, 1); c->launches++; } // time between runs is nobody's
	long long remaining = n_steps;
	long long since_rebuild = 0;
	bool since_rebuild_valid = false;
	while(remaining > 0) {
		// rebuilds in the middle of a run are launched unchecked (see do_build); the first one of a run is checked
		rc = ensure_lists(c, c->defer_build_checks && since_rebuild_valid);
		if(rc) return rc;
		since_rebuild_valid = true;
		rc = reset_batch_flags(c);
		if(rc) return rc;
		const long long step0 = c->step;
		const bool started_mid = c->mid_step;
		if(c->dirty_acc) {
			// an incomplete force pass left partial sums behind
			const int k = c->cur;
			CU(cudaMemsetAsync(c->F[k], 0, sizeof(float4) * (size_t) c->N, c->stream));
			CU(cudaMemsetAsync(c->T[k], 0, sizeof(float4) * (size_t) c->N, c->stream));
			CU(cudaMemsetAsync(c->Fb, 0, sizeof(float4) * (size_t) c->N, c->stream));
			c->dirty_acc = false;
		}
		if(!c->mid_step) {
			// start of a run: forces for the current positions, then the first half-kick + drift.  Launch index -1: reads halt
			// word 1, writes word 0, so that the first full unit below always has index 0 (captured graphs freeze the parity)
			if(!c->forces_valid) { rc = launch_forces(c, OXB_FLAG_COUNT + 1, true, step0); if(rc) return rc; c->forces_valid = true; }
			oxb::launch_integrate_epoch(c->stream, integ_args(c, step0), OXB_PH_FIRST, -1);
			c->launches++;
		}
		// the batch: `full` units that end with the first half of the following step, then (only at the very end of the run)
		// one closing unit without it
		// speculate up to ~1.0 x the running mean rebuild interval past the last rebuild, then in pairs (launches behind a halt are
		// no-ops but still cost their launch latency); even sizes keep the graph path usable
		long long budget = (long long) (c->spec_factor * c->avg_interval + 0.5) - since_rebuild;
		budget = std::max<long long>(2, std::min<long long>(64, budget)) & ~1ll;
		const bool closes = (remaining <= budget);
		long long full = closes ? remaining - 1 : budget;
		int epoch = 0;
		rc = launch_full_units(c, full, step0, epoch);
		if(rc) return rc;
		if(closes) {
			rc = launch_unit(c, epoch++, step0 + full, false);
			if(rc != 0 && rc != 2) return rc;
		}
		const long long batch = full + (closes ? 1 : 0);
		CU(cudaGetLastError());
		rc = read_flags(c);
		if(rc) return rc;
		const int done = c->h_flags[OXB_FLAG_STEPS_DONE];
		const bool halted = c->h_flags[OXB_FLAG_COUNT] || c->h_flags[OXB_FLAG_COUNT + 1];
		if(c->build_unchecked) {
			c->build_unchecked = false;
			if(halted && done == 0 && (c->h_flags[OXB_FLAG_ERROR] & (OXB_ERR_NEIGH_OVERFLOW | OXB_ERR_EDGE_OVERFLOW))) {
				// the unchecked rebuild overflowed and k_batch_begin halted the batch: nothing ran.  Redo it on the checked path
				// (which grows the arrays); it is the same list update, not a new one
				c->error_flags &= ~(OXB_ERR_NEIGH_OVERFLOW | OXB_ERR_EDGE_OVERFLOW);
				c->lists_valid = false;
				c->n_list_updates--;
				rc = do_build(c, false);
				if(rc) return rc;
				continue;
			}
		}
		c->step += done;
		remaining -= done;
		since_rebuild += done;
		if(c->error_flags & OXB_ERR_SEG_OVERFLOW) {
			// a force pass of this batch dropped pairs: the integrator launch behind it halted the batch without using the forces.  Enlarge
			// the segments and repeat that pass
			rc = grow_segments(c, "run");
			if(rc) return rc;
			c->mid_step = started_mid || done > 0;
			since_rebuild = 0;
			continue;
		}
		if(c->error_flags & OXB_ERR_EDGE_OVERFLOW) {
			return fail(c, 8, "the edge list overflowed around step %lld after repeated growth", c->step);
		}
		if(c->error_flags & OXB_ERR_FENE_BROKEN) {
			return fail(c, 6, "the distance between bonded neighbors exceeded acceptable values (FENE range) around step %lld", c->step);
		}
		if(halted) {
			c->lists_valid = false;
			c->mid_step = true;
			c->forces_valid = false;
			c->avg_interval = 0.7 * c->avg_interval + 0.3 * (double) std::max<long long>(since_rebuild, 1);
			since_rebuild = 0;
		}
		else {
			c->mid_step = (remaining > 0);
			c->forces_valid = false;
			if(done != batch) return fail(c, 7, "internal error: batch of %lld steps completed %d without a halt", batch, done);
		}
	}
	if(c->profiling) { k_prof_mark<<<1, 1, 0, c->stream>>>(c->flags, OXB_PROF_OTHER, 0); c->launches++; } // closes the last integrate launch
	// at the end of a run the forces in memory belong to the last completed step's positions: still valid
	c->forces_valid = true;
	c->mid_step = false;
	return 0;
}

int oxb_synchronize(oxb_ctx *c) {
	if(c == nullptr) return 1;
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

int oxb_get_forces(oxb_ctx *c, double *force, double *torque_body, double *torque_lab, double *energy, double *hb_energy) {
	if(c == nullptr) return 1;
	int rc = ensure_forces_checked(c, "get_forces");
	if(rc) return rc;
	const int N = c->N, k = c->cur;
	std::vector<float4> hF(N), hT(N), hB(N, make_float4(0.f, 0.f, 0.f, 0.f));
	std::vector<double4> hq(N);
	std::vector<int4> hi(N);
	if(c->use_edge) CU(cudaMemcpyAsync(hB.data(), c->Fb, sizeof(float4) * N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(hF.data(), c->F[k], sizeof(float4) * N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(hT.data(), c->T[k], sizeof(float4) * N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(hq.data(), c->quatd[k], sizeof(double4) * N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(hi.data(), c->ipos[k], sizeof(int4) * N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	for(int s = 0; s < N; s++) {
		int i = word_index(hi[s].w);
		if(c->use_edge) {
			// edge pipeline: the Debye-Hueckel force sum acts at the backbone site and is kept apart on the device (the integrator
			// folds it in); fold it here the same way: F += Fb, tau += back x Fb, energy += Fb.w
			quatd q = { hq[s].x, hq[s].y, hq[s].z, hq[s].w };
			double x1[3], x2[3], x3[3];
			axes_from_quatd(q, x1, x2, x3);
			double bk[3], g[3] = { hB[s].x, hB[s].y, hB[s].z };
			for(int d = 0; d < 3; d++) bk[d] = (double) c->model.back_a1 * x1[d] + (double) c->model.back_a2 * x2[d] + (double) c->back_a3 * x3[d];
			hF[s].x += hB[s].x; hF[s].y += hB[s].y; hF[s].z += hB[s].z; hF[s].w += hB[s].w;
			hT[s].x += (float) (bk[1] * g[2] - bk[2] * g[1]); hT[s].y += (float) (bk[2] * g[0] - bk[0] * g[2]); hT[s].z += (float) (bk[0] * g[1] - bk[1] * g[0]);
		}
		if(force) { force[3 * i] = hF[s].x; force[3 * i + 1] = hF[s].y; force[3 * i + 2] = hF[s].z; }
		// the device keeps the torque in the lab frame; the reference stores it in the body frame (R^T tau)
		if(torque_lab) { torque_lab[3 * i] = hT[s].x; torque_lab[3 * i + 1] = hT[s].y; torque_lab[3 * i + 2] = hT[s].z; }
		if(torque_body) {
			quatd q = { hq[s].x, hq[s].y, hq[s].z, hq[s].w };
			double x1[3], x2[3], x3[3];
			axes_from_quatd(q, x1, x2, x3);
			torque_body[3 * i] = x1[0] * hT[s].x + x1[1] * hT[s].y + x1[2] * hT[s].z;
			torque_body[3 * i + 1] = x2[0] * hT[s].x + x2[1] * hT[s].y + x2[2] * hT[s].z;
			torque_body[3 * i + 2] = x3[0] * hT[s].x + x3[1] * hT[s].y + x3[2] * hT[s].z;
		}
		if(energy) energy[i] = hF[s].w;
		if(hb_energy) hb_energy[i] = hT[s].w;
	}
	return 0;
}

int oxb_energy(oxb_ctx *c, double *U, double *K) {
	if(c == nullptr) return 1;
	int rc = ensure_forces_checked(c, "energy");
	if(rc) return rc;
	const int k = c->cur;
	oxb::launch_energy_sum(c->stream, c->N, c->F[k], c->use_edge ? c->Fb : nullptr, c->d_energy);
	KinSums hs;
	// keep the Bussi state words intact: only the five running sums are cleared/refilled
	oxb::launch_kinetic_sums(c->stream, c->N, c->veld[k], c->Ld[k], c->sums);
	c->launches += 3;
	CU(cudaMemcpyAsync(c->h_scalars, c->d_energy, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(&hs, c->sums, sizeof(KinSums), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	if(U) *U = 0.5 * c->h_scalars[0];
	if(K) *K = 0.5 * (hs.v2 + hs.L2);
	return 0;
}

// ---- MC barostat (MD_CUDABackend::_apply_barostat, src/CUDA/Backends/MD_CUDABackend.cu:451-516)
static int barostat_tables(oxb_ctx *c) {
	if(c->mol_of != nullptr) return 0;
	const int N = c->N;
	// molecules = strands (ConfigInfo::molecules() of a nucleic-acid topology); ids compacted in order of first appearance
	std::map<int, int> ids;
	std::vector<int> mol(N);
	for(int i = 0; i < N; i++) {
		auto it = ids.find(c->h_strand[i]);
		if(it == ids.end()) it = ids.insert(std::make_pair(c->h_strand[i], (int) ids.size())).first;
		mol[i] = it->second;
	}
	c->n_mol = (int) ids.size();
	std::vector<double> inv(c->n_mol, 0.);
	for(int i = 0; i < N; i++) inv[mol[i]] += 1.;
	for(auto &x : inv) x = 1. / x;
	CU(dalloc(&c->mol_of, (size_t) N));
	CU(dalloc(&c->mol_inv_size, (size_t) c->n_mol));
	CU(dalloc(&c->mol_coms, 3 * (size_t) c->n_mol));
	CU(dalloc(&c->pos_backup, (size_t) N));
	CU(cudaMemcpy(c->mol_of, mol.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(c->mol_inv_size, inv.data(), sizeof(double) * c->n_mol, cudaMemcpyHostToDevice));
	return 0;
}

int oxb_fix_diffusion(oxb_ctx *c, int *shifts) {
	if(c == nullptr) return 1;
	int rc = check_ready(c);
	if(rc) return rc;
	if(c->mid_step) return fail(c, 2, "fix_diffusion needs a completed step (call between runs)");
	rc = barostat_tables(c);
	if(rc) return rc;
	const int N = c->N, k = c->cur;
	int *d_shifts = nullptr;
	if(shifts) {
		// staging (15 N doubles) is free between set/get_state calls
		if(c->d_stage == nullptr) CU(dalloc(&c->d_stage, 15 * (size_t) N));
		d_shifts = reinterpret_cast<int *>(c->d_stage);
	}
	oxb::launch_mol_coms(c->stream, N, c->n_mol, c->ipos[k], c->mol_of, c->mol_inv_size, c->posd[k], c->mol_coms);
	oxb::launch_fix_diffusion(c->stream, N, c->ipos[k], c->mol_of, c->mol_coms, c->box, c->posd[k], c->quatd[k], c->axf[k], d_shifts);
	c->launches += 2;
	CU(cudaGetLastError());
	if(shifts) CU(cudaMemcpyAsync(shifts, d_shifts, sizeof(int) * 3 * (size_t) N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	// positions moved by whole box sides: fixed-point images, lists and forces stay valid
	return 0;
}

int oxb_set_host_wait(oxb_ctx *c, int blocking) {
	if(c == nullptr) return 1;
	c->blocking_wait = blocking != 0;
	return 0;
}

int oxb_get_box(oxb_ctx *c, double box[3]) {
	if(c == nullptr || box == nullptr) return 1;
	if(!c->have_box) return fail(c, 2, "box not set");
	for(int k = 0; k < 3; k++) box[k] = c->box[k];
	return 0;
}

int oxb_barostat_trial(oxb_ctx *c, const double new_box[3], int molecular) {
	if(c == nullptr || new_box == nullptr) return 1;
	int rc = check_ready(c);
	if(rc) return rc;
	if(c->mid_step) return fail(c, 2, "a barostat move needs a completed step (call between runs)");
	if(c->trial_open) return fail(c, 2, "a barostat trial is already open");
	for(int k = 0; k < 3; k++) if(!(new_box[k] > 0)) return fail(c, 1, "box sides must be positive");
	rc = barostat_tables(c);
	if(rc) return rc;
	const int N = c->N, k = c->cur;
	oxb::RescaleArgs a;
	a.backup = c->pos_backup; a.restore = nullptr;
	a.N = N; a.molecular = molecular ? 1 : 0;
	for(int d = 0; d < 3; d++) {
		c->box_backup[d] = c->box[d];
		a.f[d] = molecular ? new_box[d] / c->box[d] - 1. : new_box[d] / c->box[d];
		a.box_inv[d] = 1. / new_box[d];
	}
	a.posd = c->posd[k]; a.quatd = c->quatd[k]; a.ipos = c->ipos[k]; a.iback = c->iback[k];
	a.mol_of = c->mol_of; a.coms = c->mol_coms;
	a.back_a1 = c->model.back_a1; a.back_a2 = c->model.back_a2; a.back_a3 = c->back_a3;
	if(molecular) { oxb::launch_mol_coms(c->stream, N, c->n_mol, c->ipos[k], c->mol_of, c->mol_inv_size, c->posd[k], c->mol_coms); c->launches++; }
	oxb::launch_rescale_positions(c->stream, a);
	c->launches++;
	CU(cudaGetLastError());
	c->trial_open = true;
	return oxb_set_box(c, new_box);
}

int oxb_barostat_reject(oxb_ctx *c) {
	if(c == nullptr) return 1;
	if(!c->trial_open) return fail(c, 2, "no barostat trial is open");
	// the slots may have been re-sorted since the trial (list rebuild in the new box): restore by original id, then re-encode the
	// fixed-point centre and backbone site for the old box
	const int N = c->N, k = c->cur;
	oxb::RescaleArgs a;
	a.N = N; a.molecular = 0;
	for(int d = 0; d < 3; d++) { a.f[d] = 1.; a.box_inv[d] = 1. / c->box_backup[d]; }
	a.posd = c->posd[k]; a.quatd = c->quatd[k]; a.ipos = c->ipos[k]; a.iback = c->iback[k];
	a.mol_of = c->mol_of; a.coms = c->mol_coms; a.backup = nullptr; a.restore = c->pos_backup;
	a.back_a1 = c->model.back_a1; a.back_a2 = c->model.back_a2; a.back_a3 = c->back_a3;
	oxb::launch_rescale_positions(c->stream, a);
	c->launches++;
	CU(cudaGetLastError());
	c->trial_open = false;
	return oxb_set_box(c, c->box_backup);
}

int oxb_barostat_accept(oxb_ctx *c) {
	if(c == nullptr) return 1;
	if(!c->trial_open) return fail(c, 2, "no barostat trial is open");
	c->trial_open = false;
	return 0;
}

int oxb_barostat_move(oxb_ctx *c, const double new_box[3], int molecular, double P, double T, double u, int *accepted, double *dE_out) {
	if(c == nullptr || new_box == nullptr) return 1;
	double U0 = 0., U1 = 0.;
	int rc = oxb_energy(c, &U0, nullptr);
	if(rc) return rc;
	const double V0 = c->box[0] * c->box[1] * c->box[2], V1 = new_box[0] * new_box[1] * new_box[2];
	rc = oxb_barostat_trial(c, new_box, molecular);
	if(rc) return rc;
	rc = oxb_energy(c, &U1, nullptr);
	if(rc) { oxb_barostat_reject(c); return rc; }
	const double n_objs = molecular ? (double) c->n_mol : (double) c->N;
	const double dE = U1 - U0;
	const double acc = std::exp(-(dE + P * (V1 - V0) - n_objs * T * std::log(V1 / V0)) / T);
	const bool ok = acc > u;
	if(dE_out) *dE_out = dE;
	if(accepted) *accepted = ok ? 1 : 0;
	return ok ? oxb_barostat_accept(c) : oxb_barostat_reject(c);
}

int oxb_energy_split(oxb_ctx *c, double *terms) {
	if(c == nullptr || terms == nullptr) return 1;
	if(c->n_rep > 1) return fail(c, 1, "oxb_energy_split is not available for a replica batch");
	int rc = check_ready(c);
	if(rc) return rc;
	rc = ensure_lists(c);
	if(rc) return rc;
	const int k = c->cur;
	oxb::launch_energy_split(c->stream, c->mref(), c->boxf, c->N, c->ipos[k], c->axf[k], c->bonds[k], c->nbr, c->nnbr, c->N, c->d_energy + 2);
	c->launches += 1;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(c->h_scalars + 2, c->d_energy + 2, sizeof(double) * OXB_NTERMS, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	for(int t = 0; t < OXB_NTERMS; t++) terms[t] = c->h_scalars[2 + t];
	return 0;
}

int oxb_get_pairs(oxb_ctx *c, int *pairs, long long max_pairs, long long *n_pairs) {
	if(c == nullptr || n_pairs == nullptr) return 1;
	int rc = check_ready(c);
	if(rc) return rc;
	if(c->lists_valid && c->lists_rv != c->rcut + 2. * c->skin) c->lists_valid = false; // the pair set is defined by the exact radius
	rc = ensure_lists(c);
	if(rc) return rc;
	const int N = c->N;
	std::vector<int> hn(N), hm((size_t) c->max_neigh * N);
	std::vector<int4> hi(N);
	CU(cudaMemcpyAsync(hn.data(), c->nnbr, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(hm.data(), c->nbr, sizeof(int) * (size_t) c->max_neigh * N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(hi.data(), c->ipos[c->cur], sizeof(int4) * N, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	long long n = 0;
	for(int s = 0; s < N; s++) {
		int i = word_index(hi[s].w);
		for(int k = 0; k < hn[s]; k++) {
			int j = word_index(hi[hm[(size_t) k * N + s] & OXB_SLOT_MASK].w); // (half-shell entries carry the near-edge class above the slot)
			// full matrix: every pair shows up from both ends, keep it once; half matrix: every entry is a pair
			if(i < j || c->nbr_is_half) {
				if(pairs != nullptr && n < max_pairs) { pairs[2 * n] = std::min(i, j); pairs[2 * n + 1] = std::max(i, j); }
				n++;
			}
		}
	}
	*n_pairs = n;
	return 0;
}

int oxb_get_stats(oxb_ctx *c, long long *n_list_updates, long long *n_sorts, int *max_neigh, int *error_flags) {
	if(c == nullptr) return 1;
	if(n_list_updates) *n_list_updates = c->n_list_updates;
	if(n_sorts) *n_sorts = c->n_sorts;
	if(max_neigh) *max_neigh = c->max_neigh;
	if(error_flags) *error_flags = c->error_flags;
	return 0;
}

int oxb_device_views(oxb_ctx *c, void **poss_f4, void **orientations_f4, void **matrix_neighs, void **number_neighs, void **edge_list, void **n_edges) {
	if(c == nullptr) return 1;
	int rc = check_ready(c);
	if(rc) return rc;
	if((matrix_neighs || number_neighs) && !c->need_full_matrix) {
		// the reference's matrix holds both directions of every pair: from now on the builds produce it
		c->need_full_matrix = true;
		if(c->nbr_is_half) c->lists_valid = false;
	}
	rc = ensure_lists(c);
	if(rc) return rc;
	if(poss_f4) {
		if(c->pos_f4 == nullptr) CU(dalloc(&c->pos_f4, c->N));
		k_pos_view<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->N, c->posd[c->cur], c->ipos[c->cur], c->pos_f4);
		c->launches++;
		CU(cudaStreamSynchronize(c->stream));
		*poss_f4 = c->pos_f4;
	}
	if(orientations_f4) {
		// GPU_quat view (src/CUDA/cuda_defs.h:58-97) of the FP64 quaternions
		if(c->quat_f4 == nullptr) CU(dalloc(&c->quat_f4, c->N));
		k_quat_view<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->N, c->quatd[c->cur], c->quat_f4);
		c->launches++;
		CU(cudaStreamSynchronize(c->stream));
		*orientations_f4 = c->quat_f4;
	}
	if(matrix_neighs) *matrix_neighs = c->nbr;
	if(number_neighs) *number_neighs = c->nnbr;
	if(edge_list) *edge_list = c->edges;
	if(n_edges) *n_edges = c->n_edges;
	return 0;
}

long long oxb_launch_count(const oxb_ctx *c) { return c ? c->launches : 0; }

int oxb_set_profile(oxb_ctx *c, int enable) {
	if(c == nullptr) return 1;
	cudaSetDevice(c->device);
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaMemset(c->flags + OXB_PROF_OFFSET, 0, sizeof(unsigned long long) * (2 + 2 * OXB_PROF_NPHASE)));
	const int on = enable ? 1 : 0;
	CU(cudaMemcpy(c->flags + OXB_FLAG_PROF_ON, &on, sizeof(int), cudaMemcpyHostToDevice));
	c->profiling = on != 0;
	return 0;
}

int oxb_get_profile(oxb_ctx *c, double *ms, long long *entries) {
	if(c == nullptr) return 1;
	cudaSetDevice(c->device);
	unsigned long long h[2 + 2 * OXB_PROF_NPHASE];
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaMemcpy(h, c->flags + OXB_PROF_OFFSET, sizeof(h), cudaMemcpyDeviceToHost));
	for(int p = 0; p < OXB_PROF_NPHASE; p++) {
		if(ms) ms[p] = 1e-6 * (double) h[2 + p];
		if(entries) entries[p] = (long long) h[2 + OXB_PROF_NPHASE + p];
	}
	return 0;
}

int oxb_time_kernel(oxb_ctx *c, int which, int reps, float *ms) {
	if(c == nullptr || ms == nullptr || reps < 1) return 1;
	int rc = ensure_forces(c);
	if(rc) return rc;
	cudaEvent_t e0, e1;
	CU(cudaEventCreate(&e0));
	CU(cudaEventCreate(&e1));
	rc = reset_batch_flags(c);
	if(rc) return rc;
	if(which == 1) {
		// untimed first launch: with lazy module loading the first use of a kernel variant costs milliseconds
		oxb::IntegrateArgs a = integ_args(c, c->step);
		a.dt = 0.;
		oxb::launch_integrate_epoch(c->stream, a, OXB_PH_SECOND | OXB_PH_FIRST | OXB_PH_COUNT_STEP, 0);
		c->launches++;
	}
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaEventRecord(e0, c->stream));
	for(int r = 0; r < reps; r++) {
		if(which == 0) { rc = launch_forces(c, OXB_FLAG_COUNT, true, c->step); if(rc) return rc; }
		else if(which == 1) {
			// dt = 0 leaves the state bit-identical while moving exactly the same bytes
			oxb::IntegrateArgs a = integ_args(c, c->step);
			a.dt = 0.;
			oxb::launch_integrate_epoch(c->stream, a, OXB_PH_SECOND | OXB_PH_FIRST | OXB_PH_COUNT_STEP, 0);
			c->launches++;
		}
		else if(which == 2) {
			// a rebuild as the hot loop does it: re-sort (if enabled) + list build
			if(c->sort_every > 0) { rc = do_sort(c); if(rc) return rc; }
			oxb::launch_build_lists(c->stream, list_args(c));
			c->slots_cell_ordered = false;
			c->launches += c->use_edge ? 7 : 4;
		}
		else if(which == 3) { rc = do_sort(c); if(rc) return rc; }
		else return fail(c, 1, "unknown kernel selector %d", which);
	}
	CU(cudaEventRecord(e1, c->stream));
	CU(cudaEventSynchronize(e1));
	float t = 0.f;
	CU(cudaEventElapsedTime(&t, e0, e1));
	*ms = t / reps;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	if(which >= 2) c->lists_valid = false;
	c->forces_valid = false;
	rc = ensure_forces(c);
	if(rc) return rc;
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

} // extern "C"
