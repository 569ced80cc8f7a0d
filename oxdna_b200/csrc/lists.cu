// Cell-binned Verlet list construction (sm_100a, HBM/L2-bound integer work).
//
// Replaces CUDASimpleVerletList::update and its kernels simple_fill_cells, simple_update_neigh_list,
// edge_update_neigh_list, compress_matrix_neighs (src/CUDA/Lists/CUDASimpleVerletList.cu:231-289,
// src/CUDA/Lists/CUDA_simple_verlet.cuh:23-175).  Differences by design:
//  * binning is a stable radix sort of (cell, slot) -> contiguous cell ranges, no fixed-capacity cells, no atomics,
//    no overflow, deterministic neighbour order;
//  * the distance predicate is decided in FP32 on fixed-point coordinates and re-decided in FP64 with the CPU's exact
//    expression (src/Boxes/CubicBox.cpp:51-61, src/Lists/Cells.cpp:170) whenever it falls within a guard band of the
//    cutoff, so the pair SET is bit-identical to the reference's CPU Verlet list built from the same FP64 positions;
//  * neighbour rows are bounds-checked (the reference is not) and overflow raises a flag;
//  * the unique-pair (edge) list is produced by an exclusive scan on the device, its length stays on the device.
#include "kernels.h"

#include <cub/cub.cuh>

#include <cstdlib>

namespace {

__device__ __forceinline__ int cell_coord(double x, double L, int n) {
	// src/Lists/Cells.h:60-65
	double f = x / L - floor(x / L);
	int c = (int) (f * (1. - 2.220446049250313e-16) * n);
	return min(c, n - 1);
}

__global__ void __launch_bounds__(256) k_cell_keys(int N, int n_per, const double4 *__restrict__ posd, double Lx, double Ly, double Lz, int nx, int ny, int nz,
		int *__restrict__ key, int *__restrict__ val, int *__restrict__ flags) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i == 0) prof_mark(flags, OXB_PROF_BUILD);
	if(i >= N) return;
	double4 p = posd[i];
	int cx = cell_coord(p.x, Lx, nx), cy = cell_coord(p.y, Ly, ny), cz = cell_coord(p.z, Lz, nz);
	key[i] = cx + nx * (cy + ny * cz) + (i / n_per) * (nx * ny * nz); // replica batching: one copy of the grid per replica
	val[i] = i;
}

// (also records the staleness references of slot j -- where its centre, backbone site and base site are at this rebuild -- in the .w lanes
// of the FP64 state; done here and not in k_build_neigh, which reads other particles' positions while it runs)
__global__ void __launch_bounds__(256) k_cell_ranges(int N, const int *__restrict__ key_sorted, int *__restrict__ cell_start, int *__restrict__ cell_end,
		const int4 *__restrict__ ipos, const int4 *__restrict__ iback, const float4 *__restrict__ axf, float base_a1, BoxF boxf,
		double4 *__restrict__ ref_pos, double4 *__restrict__ ref_vel, double4 *__restrict__ ref_L, int *__restrict__ flags) {
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if(j == 0) prof_mark(flags, OXB_PROF_BUILD);
	if(j >= N) return;
	{
		const int4 ip = ipos[j];
		const v3 a1p = load_a1(axf, j);
		int4 bs = ip;
		bs.x = (int) ((unsigned) ip.x + (unsigned) (int) rintf(a1p.x * base_a1 / boxf.sx));
		bs.y = (int) ((unsigned) ip.y + (unsigned) (int) rintf(a1p.y * base_a1 / boxf.sy));
		bs.z = (int) ((unsigned) ip.z + (unsigned) (int) rintf(a1p.z * base_a1 / boxf.sz));
		ref_pos[j].w = pack_ref(ip);
		ref_vel[j].w = pack_ref(iback[j]);
		ref_L[j].w = pack_ref(bs);
	}
	int k = key_sorted[j];
	if(j == 0 || key_sorted[j - 1] != k) cell_start[k] = j;
	if(j == N - 1 || key_sorted[j + 1] != k) cell_end[k] = j + 1;
}

__device__ __forceinline__ bool within_exact(double4 p, double4 q, double Lx, double Ly, double Lz, double rv2) {
	double dx = q.x - p.x - rint((q.x - p.x) / Lx) * Lx;
	double dy = q.y - p.y - rint((q.y - p.y) / Ly) * Ly;
	double dz = q.z - p.z - rint((q.z - p.z) / Lz) * Lz;
	// same association as LR_vector::norm(): x*x + y*y + z*z, no FMA contraction
	double n = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
	return n < rv2;
}

// Site-based selection of the "near" edges: a pair can feel something other than Debye-Hueckel before the next rebuild
// only if one of its site-site distances is within the range of the corresponding term plus twice the skin (every site
// -- centre, backbone, base and, being between centre and base, stack -- moves less than `skin` between rebuilds).
// Returns the CLASS of the pair: one bit per family of site pairs that can come into range (OXB_CLS_*, common.cuh); 0 = not a near edge.
__device__ __forceinline__ int near_pair(const oxb::ListArgs &a, v3 r, v3 a1p, v3 a1q, v3 bkp, v3 bkq) {
	int cls = 0;
	v3 rbb = r + bkq - bkp;
	if(dot(rbb, rbb) < a.r2_bb) cls |= OXB_CLS_BB;
	v3 da = a1q - a1p;
	v3 rb = r + da * a.base_a1;
	const float rb2 = dot(rb, rb);
	if(rb2 < a.r2_eb) cls |= OXB_CLS_EB;
	if(rb2 < a.r2_base) cls |= OXB_CLS_HBCR;
	v3 rs = r + da * a.stack_a1;
	if(dot(rs, rs) < a.r2_stack) cls |= OXB_CLS_ST;
	v3 d1 = r + bkq - a1p * a.base_a1; // base(p) - back(q)
	v3 d2 = r + a1q * a.base_a1 - bkp; // back(p) - base(q)
	if(dot(d1, d1) < a.r2_bk || dot(d2, d2) < a.r2_bk) cls |= OXB_CLS_BK;
	return cls;
}

// One thread per particle: visit the 27 surrounding cells, keep non-bonded particles closer than rv.
//  * the 27 (start, end) cell ranges are fetched up front as independent loads and parked in shared memory, so the scan
//    pays one memory latency for them instead of 27 dependent ones;
//  * DIRECT = the particle arrays are already ordered by cell (this rebuild re-sorted them): a cell's members are the
//    slots [start, end) themselves, no index indirection and the candidate loads are contiguous;
//  * the candidate's position is prefetched one iteration ahead;
//  * neighbours m > i that can feel more than Debye-Hueckel before the next rebuild are flagged in a 128-bit row mask
//    (bit k = row entry k), from which k_fill_edges emits the edge list without touching any geometry again.
template<bool DIRECT>
__global__ void __launch_bounds__(128) k_build_neigh(oxb::ListArgs a, const int *__restrict__ cell_start, const int *__restrict__ cell_end) {
	// (staging the rows in shared-memory columns and flushing them row by row -- whole sectors instead of 4-byte pieces scattered over as
	// many rows as the lanes have different counts -- was measured: it halves the 0.45 ms that the 35-us edge fill takes behind a
	// FULL-shell build on the device timeline, but costs occupancy: 0.41 against 0.39 ms for the half-shell build, gpurun_out r2i-r2k)
	__shared__ int2 s_range[27][128];
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(a.flags, OXB_PROF_BUILD);
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool active = i < a.N;
	if(!active) i = a.N - 1;
	const int4 ip = a.ipos[i];
	const int2 b = a.bonds[i];
	const int nx = a.ncell[0], ny = a.ncell[1], nz = a.ncell[2];
	const double4 pd = a.posd[i];
	const int cx = cell_coord(pd.x, a.box[0], nx), cy = cell_coord(pd.y, a.box[1], ny), cz = cell_coord(pd.z, a.box[2], nz);
	const int coff = (a.n_rep > 1) ? (i / a.n_per) * (nx * ny * nz) : 0; // this replica's copy of the grid
	// phase 1: the 27 ranges as independent loads; only the non-empty ones are kept (in scan order, which keeps the neighbour
	// order deterministic), so that all lanes of a warp walk through candidates instead of through mostly empty cells
	int n_rg = 0;
	{
		int2 rg[27];
		int q = 0;
#pragma unroll
		for(int dz = -1; dz <= 1; dz++) {
			int zc = cz + dz; zc += (zc < 0) ? nz : 0; zc -= (zc >= nz) ? nz : 0;
#pragma unroll
			for(int dy = -1; dy <= 1; dy++) {
				int yc = cy + dy; yc += (yc < 0) ? ny : 0; yc -= (yc >= ny) ? ny : 0;
#pragma unroll
				for(int dx = -1; dx <= 1; dx++) {
					int xc = cx + dx; xc += (xc < 0) ? nx : 0; xc -= (xc >= nx) ? nx : 0;
					int c = coff + xc + nx * (yc + ny * zc);
					rg[q] = make_int2(__ldg(cell_start + c), __ldg(cell_end + c));
					// half shell: only partners in higher slots are wanted; with cell-ordered slots whole cells (about half of the 27) drop out
					if(DIRECT && a.half_shell) rg[q].x = max(rg[q].x, i + 1);
					q++;
				}
			}
		}
#pragma unroll
		for(q = 0; q < 27; q++) {
			if(rg[q].x < rg[q].y) {
				s_range[n_rg][threadIdx.x] = rg[q];
				n_rg++;
			}
		}
	}
	if(!active || n_rg == 0) n_rg = 0;
	const float rv2f = (float) (a.rv * a.rv);
	const float band = 1e-4f * rv2f;
	const double rv2 = a.rv * a.rv;
	const int4 ib = a.iback[i];
	const v3 a1p = load_a1(a.axf, i);
	const v3 bkp = min_image_fixed(a.boxf, ip, ib);
	int count = 0, higher_near = 0, ndh = 0;
	int ng0 = 0, ng1 = 0, ng2 = 0; // near edges of this row per class group
	unsigned long long mask0 = 0ull, mask1 = 0ull;
	bool mask_overflow = false, dh_overflow = false;
	// phase 2: one flat loop over this particle's candidates (cursor = range r, slot j), next candidate prefetched
	int r = 0, j = 0, jend = 0;
	int m_next = 0;
	int4 ip_next = ip;
	if(n_rg > 0) {
		int2 g = s_range[0][threadIdx.x];
		j = g.x; jend = g.y;
		m_next = DIRECT ? j : __ldg(a.cell_val_sorted + j);
		ip_next = __ldg(a.ipos + m_next);
	}
	while(r < n_rg) {
		const int m = m_next;
		const int4 ipm = ip_next;
		// advance the cursor and prefetch
		j++;
		if(j >= jend) {
			r++;
			if(r < n_rg) {
				int2 g = s_range[r][threadIdx.x];
				j = g.x; jend = g.y;
			}
		}
		if(r < n_rg) {
			m_next = DIRECT ? j : __ldg(a.cell_val_sorted + j);
			ip_next = __ldg(a.ipos + m_next);
		}
		if(m == i || m == b.x || m == b.y || (a.half_shell && m < i)) continue;
		v3 d = min_image_fixed(a.boxf, ip, ipm);
		float d2 = dot(d, d);
		bool in = d2 < rv2f;
		if(fabsf(d2 - rv2f) < band) in = within_exact(pd, a.posd[m], a.box[0], a.box[1], a.box[2], rv2);
		if(in) {
			if(count < a.max_neigh) a.nbr[(size_t) count * a.stride + i] = m;
			count++;
		}
	}
	// phase 3: classification of the row just written (near edge? Debye-Hueckel row?).  Kept out of the candidate loop: there nearly
	// every warp iteration has a hit on SOME lane, so the ~100 instructions of this block ran for every candidate (57 per particle)
	// with a few lanes active; here the lanes walk rows of similar length together (ncu r02f: 6,900 warp instructions per warp before)
	// (sharing the rows between lanes l and 31 - l as k_dh_particle does was measured here too: 0.44 against 0.39 ms at 1M nt -- the extra
	// selects and registers cost more than the balance gains, gpurun_out r2r)
	const int nrow = min(count, a.max_neigh);
	for(int k = 0; k < nrow; k++) {
		const int m = a.nbr[(size_t) k * a.stride + i]; // this thread's own store (plain load, program order)
		const int4 ipm = __ldg(a.ipos + m);
		const int4 ibm = __ldg(a.iback + m);
		const v3 d = min_image_fixed(a.boxf, ip, ipm);
		const float d2 = dot(d, d);
		// Debye-Hueckel acts between backbone sites: keep m if the sites can come within dh_rc before the
		// next rebuild (both the centre and the backbone site of every particle move less than `skin`)
		const v3 db = min_image_fixed(a.boxf, ib, ibm);
		const int cls = (m > i && d2 < a.rnear2) ? near_pair(a, d, a1p, load_a1(a.axf, m), bkp, min_image_fixed(a.boxf, ipm, ibm)) : 0;
		if(cls != 0) {
			higher_near++;
			const int grp = a.class_groups ? (a.half_shell ? cls_group(cls) : 2) : 0; // (full builds mark their edges with every class)
			ng0 += grp == 0; ng1 += grp == 1; ng2 += grp == 2;
			if(k < 64) mask0 |= 1ull << k;
			else if(k < 128) mask1 |= 1ull << (k - 64);
			else mask_overflow = true;
			// half shell: the matrix is private to the edge pipeline, the entry carries the class to k_fill_edges
			if(a.half_shell) a.nbr[(size_t) k * a.stride + i] = m | (cls << OXB_CLS_SHIFT);
		}
		if(dot(db, db) < a.rdh2) {
			// dh_half: every Debye-Hueckel pair is kept by ONE of its particles (by the parity of i + m, so that rows stay balanced); the
			// kernel adds the partner's share with one vector atomic.  Full rows (both directions, no atomics) otherwise.
			// half shell: this scan sees every pair once (m > i) and the lower slot keeps it.  Rows are then as long as a particle has
			// partners AFTER it on the Hilbert curve (0 to twice the mean inside one warp of k_dh_particle: + 10 us per force pass at
			// 1M nt); balancing them by appending to the partner's row through an atomic counter was measured and costs 0.3 ms per
			// rebuild (gpurun_out r2h) -- three times what it gains
			if(a.half_shell || !a.dh_half || ((((i + m) & 1) == 0) == (i < m))) {
				if(ndh < a.max_dh) a.dh_nbr[(size_t) ndh * a.stride + i] = m;
				ndh++;
			}
		}
	}
	// rows that overflow max_neigh are rebuilt after the matrix has grown (the entries beyond it were never written: not classified)
	if(count > a.max_neigh) mask_overflow = true;
	if(!active) return;
	{
		// one same-address atomic per warp, not per thread
		const unsigned am = __activemask();
		const int wmax = __reduce_max_sync(am, count);
		if((int) (threadIdx.x & 31) == (__ffs(am) - 1) && wmax > a.max_neigh) atomicMax(a.flags + OXB_FLAG_MAX_NEIGH_SEEN, wmax);
	}
	if(count > a.max_neigh) {
		atomicOr(a.flags + OXB_FLAG_ERROR, OXB_ERR_NEIGH_OVERFLOW);
		count = a.max_neigh;
	}
	if(ndh > a.max_dh || dh_overflow) {
		atomicOr(a.flags + OXB_FLAG_ERROR, OXB_ERR_NEIGH_OVERFLOW);
		ndh = a.max_dh;
	}
	a.nnbr[i] = count;
	a.dh_nnbr[i] = ndh;
	if(a.build_edges) {
		// (a row that takes the geometric fallback of k_fill_edges is counted there again: same classes, same counts)
		a.edge_cnt[i] = make_int4(ng0, ng1, ng2, higher_near);
		// rows longer than the mask fall back to the geometric test in k_fill_edges (top bit of word 1 doubles as the marker:
		// entry 127 can only be flagged together with an overflow, which takes the fallback anyway)
		if(mask_overflow) mask1 |= 1ull << 63;
		a.near_mask[i] = make_ulonglong2(mask0, mask1);
	}
}

// G lanes per particle (G = 4, 8, 16).  The thread-per-particle kernel above pays one exposed L2 latency per candidate (iback / axf of a
// candidate can only be requested after its distance test) and walks ~60 candidates serially: 434 us at 1M nucleotides with the
// issue slots nearly idle.  Here the G lanes of a group split the candidates of ONE particle:
//  * the 27 cell ranges are fetched by the group in one round (lane g takes cells g, g + G, ...), prefix-summed in scan order with
//    width-G shuffles and parked in shared memory: the candidates of the particle form one flat index space [0, T);
//  * round r tests candidates r G ... r G + G - 1, one per lane (independent loads: G times the memory parallelism, a G-th of the
//    serial depth); hits are compacted in candidate order with a ballot, so every row of the neighbour matrix, of the Debye-Hueckel
//    matrix and every near-edge mask is bit-identical to what the serial kernel writes.
template<bool DIRECT, int G>
__global__ void __launch_bounds__(256, 4) k_build_neigh_g(oxb::ListArgs a, const int *__restrict__ cell_start, const int *__restrict__ cell_end) {
	constexpr int PPB = 256 / G;
	__shared__ int s_start[PPB][28], s_pref[PPB][28];
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(a.flags, OXB_PROF_BUILD);
	const int grp = threadIdx.x / G, g = threadIdx.x % G;
	const unsigned lane = threadIdx.x & 31;
	const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane / G * G);
	const unsigned lt = gmask & ((1u << lane) - 1u); // lanes of my group below me
	int i = blockIdx.x * PPB + grp;
	const bool active = i < a.N;
	if(!active) i = a.N - 1;
	const int4 ip = __ldg(a.ipos + i);
	const int2 b = __ldg(a.bonds + i);
	const int nx = a.ncell[0], ny = a.ncell[1], nz = a.ncell[2];
	const double4 pd = a.posd[i];
	const int cx = cell_coord(pd.x, a.box[0], nx), cy = cell_coord(pd.y, a.box[1], ny), cz = cell_coord(pd.z, a.box[2], nz);
	const int coff = (a.n_rep > 1) ? (i / a.n_per) * (nx * ny * nz) : 0; // this replica's copy of the grid
	int T = 0;
	{
		int carry = 0;
#pragma unroll
		for(int t = 0; t < (27 + G - 1) / G; t++) {
			const int q = g + G * t;
			int st = 0, len = 0;
			if(q < 27) {
				int zc = cz + q / 9 - 1; zc += (zc < 0) ? nz : 0; zc -= (zc >= nz) ? nz : 0;
				int yc = cy + (q / 3) % 3 - 1; yc += (yc < 0) ? ny : 0; yc -= (yc >= ny) ? ny : 0;
				int xc = cx + q % 3 - 1; xc += (xc < 0) ? nx : 0; xc -= (xc >= nx) ? nx : 0;
				const int c = coff + xc + nx * (yc + ny * zc);
				st = __ldg(cell_start + c);
				const int en = __ldg(cell_end + c);
				if(DIRECT && a.half_shell) st = max(st, i + 1); // half shell, see k_build_neigh
				len = max(en - st, 0);
			}
			int incl = len;
#pragma unroll
			for(int d = 1; d < G; d <<= 1) {
				const int v = __shfl_up_sync(0xffffffffu, incl, d, G);
				if(g >= d) incl += v;
			}
			if(q < 27) { s_start[grp][q] = st; s_pref[grp][q] = carry + incl - len; }
			carry += __shfl_sync(0xffffffffu, incl, G - 1, G);
		}
		if(g == 0) s_pref[grp][27] = carry;
		T = active ? carry : 0;
	}
	__syncwarp();
	const float rv2f = (float) (a.rv * a.rv);
	const float band = 1e-4f * rv2f;
	const double rv2 = a.rv * a.rv;
	const int4 ib = __ldg(a.iback + i);
	const v3 a1p = load_a1(a.axf, i);
	const v3 bkp = min_image_fixed(a.boxf, ip, ib);
	int count = 0, ndh = 0, higher_near = 0;
	int ng0 = 0, ng1 = 0, ng2 = 0; // near edges of this row per class group
	unsigned long long mask0 = 0ull, mask1 = 0ull;
	bool mask_overflow = false, dh_overflow = false;
	const int Tw = __reduce_max_sync(0xffffffffu, T);
	int c = 0;
	for(int k0 = 0; k0 < Tw; k0 += G) {
		const int k = k0 + g;
		bool in = false;
		int m = -1;
		int4 ipm = ip;
		float d2 = 0.f;
		v3 d = mk3(0.f, 0.f, 0.f);
		if(k < T) {
			while(k >= s_pref[grp][c + 1]) c++;
			const int j = s_start[grp][c] + (k - s_pref[grp][c]);
			m = DIRECT ? j : __ldg(a.cell_val_sorted + j);
			ipm = __ldg(a.ipos + m);
			if(m != i && m != b.x && m != b.y && !(a.half_shell && m < i)) {
				d = min_image_fixed(a.boxf, ip, ipm);
				d2 = dot(d, d);
				in = d2 < rv2f;
				if(fabsf(d2 - rv2f) < band) in = within_exact(pd, a.posd[m], a.box[0], a.box[1], a.box[2], rv2);
			}
		}
		const unsigned bal = __ballot_sync(0xffffffffu, in) & gmask;
		bool is_dh = false;
		if(in) {
			const int row = count + __popc(bal & lt);
			const int4 ibm = __ldg(a.iback + m);
			const v3 db = min_image_fixed(a.boxf, ib, ibm);
			const int cls = (m > i && d2 < a.rnear2) ? near_pair(a, d, a1p, load_a1(a.axf, m), bkp, min_image_fixed(a.boxf, ipm, ibm)) : 0;
			// (half shell: the matrix is private to the edge pipeline, the entry carries the near-edge class to k_fill_edges)
			if(row < a.max_neigh) a.nbr[(size_t) row * a.stride + i] = a.half_shell ? (m | (cls << OXB_CLS_SHIFT)) : m;
			if(cls != 0) {
				higher_near++;
				const int grp = a.class_groups ? (a.half_shell ? cls_group(cls) : 2) : 0; // (full builds mark their edges with every class)
				ng0 += grp == 0; ng1 += grp == 1; ng2 += grp == 2;
				// rows that overflow max_neigh are rebuilt after the matrix has grown: never flag an entry that was not written
				if(row >= a.max_neigh) mask_overflow = true;
				else if(row < 64) mask0 |= 1ull << row;
				else if(row < 128) mask1 |= 1ull << (row - 64);
				else mask_overflow = true;
			}
			// dh_half: every Debye-Hueckel pair is kept by ONE of its particles (by the parity of i + m, so that rows stay balanced)
			is_dh = dot(db, db) < a.rdh2 && (!a.dh_half || a.half_shell || ((((i + m) & 1) == 0) == (i < m)));

		}
		count += __popc(bal);
		const unsigned bdh = __ballot_sync(0xffffffffu, is_dh) & gmask;
		if(is_dh) {
			const int row = ndh + __popc(bdh & lt);
			if(row < a.max_dh) a.dh_nbr[(size_t) row * a.stride + i] = m;
		}
		ndh += __popc(bdh);
	}
	// fold the per-lane near-edge bookkeeping over the group
#pragma unroll
	for(int o = G >> 1; o > 0; o >>= 1) {
		mask0 |= __shfl_xor_sync(0xffffffffu, mask0, o);
		mask1 |= __shfl_xor_sync(0xffffffffu, mask1, o);
		higher_near += __shfl_xor_sync(0xffffffffu, higher_near, o);
		ng0 += __shfl_xor_sync(0xffffffffu, ng0, o); ng1 += __shfl_xor_sync(0xffffffffu, ng1, o); ng2 += __shfl_xor_sync(0xffffffffu, ng2, o);
		mask_overflow = (__shfl_xor_sync(0xffffffffu, (int) mask_overflow, o) != 0) || mask_overflow;
		dh_overflow = (__shfl_xor_sync(0xffffffffu, (int) dh_overflow, o) != 0) || dh_overflow;
	}
	{
		// the longest row only matters when one overflowed (it sizes the regrown matrix): no same-address atomic otherwise -- one per warp
		// is 250,000 serialised atomics at 1M particles and G = 8, which tripled the time of this kernel
		const int wmax = __reduce_max_sync(0xffffffffu, active ? count : 0);
		if(lane == 0 && wmax > a.max_neigh) atomicMax(a.flags + OXB_FLAG_MAX_NEIGH_SEEN, wmax);
	}
	if(!active || g != 0) return;
	if(count > a.max_neigh) {
		atomicOr(a.flags + OXB_FLAG_ERROR, OXB_ERR_NEIGH_OVERFLOW);
		count = a.max_neigh;
	}
	if(ndh > a.max_dh || dh_overflow) {
		atomicOr(a.flags + OXB_FLAG_ERROR, OXB_ERR_NEIGH_OVERFLOW);
		ndh = a.max_dh;
	}
	a.nnbr[i] = count;
	a.dh_nnbr[i] = ndh;
	if(a.build_edges) {
		a.edge_cnt[i] = make_int4(ng0, ng1, ng2, higher_near);
		if(mask_overflow) mask1 |= 1ull << 63;
		a.near_mask[i] = make_ulonglong2(mask0, mask1);
	}
}

// near edge (i, m) for every flagged row entry.  The list has three segments by class group (common.cuh: cls_group), each grouped by `from`
// in slot order; edge_cnt holds the exclusive scan of the per-row counts (entry N = totals).
struct Cnt4Sum {
	__host__ __device__ __forceinline__ int4 operator()(const int4 &x, const int4 &y) const { return make_int4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w); }
};

__global__ void __launch_bounds__(128) k_fill_edges(oxb::ListArgs a) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i == 0) prof_mark(a.flags, OXB_PROF_EDGES);
	if(i >= a.N) return;
	const int4 tot = a.edge_cnt[a.N];
	const int4 o = a.edge_cnt[i];
	int off[3] = { o.x, tot.x + o.y, tot.x + tot.y + o.z }; // cursor of this row in each segment
	const int nn = a.nnbr[i];
	ulonglong2 mk = a.near_mask[i];
	// edge = (from, to | class << 24).  Half-shell builds left the class in the matrix entry; a full (public, reference-layout) matrix holds
	// plain slots and its edges are marked with every class
	const int cls_all = a.half_shell ? 0 : (OXB_CLS_ALL << OXB_CLS_SHIFT);
	auto emit = [&](int entry) {
		const int e = entry | cls_all;
		const int g = a.class_groups ? cls_group((int) ((unsigned) e >> OXB_CLS_SHIFT)) : 0;
		const int pos = (g == 0) ? off[0]++ : (g == 1 ? off[1]++ : off[2]++);
		if(pos < a.edge_capacity) a.edges[pos] = make_int2(i, e);
	};
	if(!(mk.y >> 63)) {
		unsigned long long w = mk.x;
		while(w) {
			int k = __ffsll((long long) w) - 1;
			w &= w - 1;
			emit(a.nbr[(size_t) k * a.stride + i]);
		}
		w = mk.y;
		while(w) {
			int k = 64 + __ffsll((long long) w) - 1;
			w &= w - 1;
			emit(a.nbr[(size_t) k * a.stride + i]);
		}
	}
	else {
		// very long row: redo the geometric selection (same predicate as in k_build_neigh)
		const int4 ip = a.ipos[i];
		const int4 ib = a.iback[i];
		const v3 a1p = load_a1(a.axf, i);
		const v3 bkp = min_image_fixed(a.boxf, ip, ib);
		for(int k = 0; k < nn; k++) {
			int m = a.nbr[(size_t) k * a.stride + i] & OXB_SLOT_MASK;
			if(m > i) {
				const int4 ipm = __ldg(a.ipos + m);
				v3 d = min_image_fixed(a.boxf, ip, ipm);
				const int cls = (dot(d, d) < a.rnear2) ? near_pair(a, d, a1p, load_a1(a.axf, m), bkp, min_image_fixed(a.boxf, ipm, __ldg(a.iback + m))) : 0;
				if(cls != 0) emit(a.half_shell ? (m | (cls << OXB_CLS_SHIFT)) : m);
			}
		}
	}
	if(i == 0) {
		const long long total = (long long) tot.x + tot.y + tot.z;
		a.n_edges[0] = (total <= a.edge_capacity) ? (int) total : (int) a.edge_capacity;
		a.n_edges[1] = tot.x; a.n_edges[2] = tot.y; a.n_edges[3] = tot.z;
		if(total > a.edge_capacity) atomicOr(a.flags + OXB_FLAG_ERROR, OXB_ERR_EDGE_OVERFLOW);
	}
}

int bits_for(int n) {
	int b = 1;
	while((1ll << b) < n) b++;
	return b;
}

} // namespace

namespace oxb {

size_t lists_tmp_bytes(int N, int ncells) {
	size_t a = 0, b = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, a, (int *) nullptr, (int *) nullptr, (int *) nullptr, (int *) nullptr, N, 0, bits_for(ncells));
	cub::DeviceScan::ExclusiveScan(nullptr, b, (int4 *) nullptr, (int4 *) nullptr, Cnt4Sum(), make_int4(0, 0, 0, 0), N + 1);
	return (a > b ? a : b) + 256;
}

// cell_start has room for 2 * ncells ints: starts then ends
void launch_build_lists(cudaStream_t s, const ListArgs &a) {
	const int N = a.N;
	const int ncells = a.ncell[0] * a.ncell[1] * a.ncell[2] * a.n_rep;
	int *cell_start = a.cell_start, *cell_end = a.cell_start + ncells;
	int tpb = 256;
	size_t tmp = a.cub_tmp_bytes;
	if(!a.direct) {
		// binning = stable sort of (cell, slot); with a.direct the re-sort that has just run left the particles ordered by
		// cell and wrote each slot's cell id into cell_key_sorted (sort.cu: k_permute)
		k_cell_keys<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, a.n_per, a.posd, a.box[0], a.box[1], a.box[2], a.ncell[0], a.ncell[1], a.ncell[2], a.cell_key, a.cell_val, a.flags);
		cub::DeviceRadixSort::SortPairs(a.cub_tmp, tmp, a.cell_key, a.cell_key_sorted, a.cell_val, a.cell_val_sorted, N, 0, bits_for(ncells), s);
	}
	if(!(a.direct && a.ranges_done)) {
		cudaMemsetAsync(cell_start, 0, sizeof(int) * 2 * (size_t) ncells, s);
		k_cell_ranges<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, a.cell_key_sorted, cell_start, cell_end, a.ipos, a.iback, a.axf, a.base_a1, a.boxf, a.ref_pos, a.ref_vel,
				a.ref_L, a.flags);
	}
	cudaMemsetAsync(a.flags + OXB_FLAG_MAX_NEIGH_SEEN, 0, sizeof(int), s);

	// lanes per particle of the neighbour scan (OXB_BUILD_G = 1: the thread-per-particle kernel)
	// (measured on B200, gpurun_out r2e: 8 lanes win while the system cannot fill the machine with one thread per particle -- 82 against 98 us
	// at 81,920 nt --, the serial kernel wins at 1M nt -- it executes 2.3 x fewer instructions, ncu r02f -- and both are issue-bound there)
	static const int G_env = [] { const char *e = getenv("OXB_BUILD_G"); int g = e ? atoi(e) : 0; return (g == 1 || g == 4 || g == 8 || g == 16) ? g : 0; }();
	const int G = G_env ? G_env : (N < 300000 ? 8 : 1);
#define OXB_BUILD_LAUNCH(D)                                                                                              \
	do {                                                                                                                 \
		if(G == 4) k_build_neigh_g<D, 4><<<(N + 63) / 64, 256, 0, s>>>(a, cell_start, cell_end);                         \
		else if(G == 8) k_build_neigh_g<D, 8><<<(N + 31) / 32, 256, 0, s>>>(a, cell_start, cell_end);                    \
		else if(G == 16) k_build_neigh_g<D, 16><<<(N + 15) / 16, 256, 0, s>>>(a, cell_start, cell_end);                  \
		else k_build_neigh<D><<<(N + 127) / 128, 128, 0, s>>>(a, cell_start, cell_end);                                  \
	} while(0)
	if(a.direct) OXB_BUILD_LAUNCH(true);
	else OXB_BUILD_LAUNCH(false);
#undef OXB_BUILD_LAUNCH
	if(a.build_edges) {
		tmp = a.cub_tmp_bytes;
		// in-place exclusive scan over N + 1 entries (the last input entry is ignored: its output is the total)
		cub::DeviceScan::ExclusiveScan(a.cub_tmp, tmp, a.edge_cnt, a.edge_cnt, Cnt4Sum(), make_int4(0, 0, 0, 0), N + 1, s);
		k_fill_edges<<<(N + 127) / 128, 128, 0, s>>>(a);
	}
}

} // namespace oxb
