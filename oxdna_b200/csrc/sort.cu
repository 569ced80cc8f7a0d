// Hilbert-curve re-sorting of the particle arrays (sm_100a, HBM-bound).
//
// Replaces hilbert_curve / get_inverted_sorted_hindex / permute_particles and the six device-to-device copy-backs of
// the reference (src/CUDA/CUDA_sort.cu:37-193, src/CUDA/Backends/MD_CUDABackend.cu:535-550, CUDABaseBackend.cu:263-281).
// Differences by design: 30-bit keys (10 levels instead of 8) from the FP64 positions; the state is double-buffered so
// the gather writes straight into the arrays that become current (pointer swap, no copy-back and no precision
// round trip); everything indexed by particle is remapped -- bonds, the original-id -> slot table used by external
// forces and by host read-backs -- so sorting is legal with external forces and strand-end detection stays correct
// (SURVEY appendix B.6, B.7).
#include "kernels.h"

#include <cub/cub.cuh>
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

namespace {

// Skilling's transform (AIP Conf. Proc. 707, 381 (2004)): axes -> transposed Hilbert index, then bit interleave
__device__ __forceinline__ unsigned hilbert_key(unsigned x, unsigned y, unsigned z, int bits) {
	unsigned X[3] = { x, y, z };
	unsigned M = 1u << (bits - 1);
	for(unsigned Q = M; Q > 1; Q >>= 1) {
		unsigned P = Q - 1;
#pragma unroll
		for(int i = 0; i < 3; i++) {
			if(X[i] & Q) X[0] ^= P;
			else {
				unsigned t = (X[0] ^ X[i]) & P;
				X[0] ^= t;
				X[i] ^= t;
			}
		}
	}
	X[1] ^= X[0];
	X[2] ^= X[1];
	unsigned t = 0;
	for(unsigned Q = M; Q > 1; Q >>= 1) if(X[2] & Q) t ^= Q - 1;
	X[0] ^= t; X[1] ^= t; X[2] ^= t;
	unsigned key = 0;
	for(int b = bits - 1; b >= 0; b--) {
		key = (key << 3) | (((X[0] >> b) & 1u) << 2) | (((X[1] >> b) & 1u) << 1) | ((X[2] >> b) & 1u);
	}
	return key;
}

__global__ void __launch_bounds__(256) k_hilbert_keys(int N, int n_per, const double4 *__restrict__ posd, double Lx, double Ly, double Lz, unsigned *__restrict__ keys,
		int *__restrict__ vals, int *__restrict__ flags) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i == 0) prof_mark(flags, OXB_PROF_SORT);
	if(i >= N) return;
	double4 p = posd[i];
	const int bits = (n_per < N) ? 8 : 10; // replica batching: the replica index takes the top bits of the 32-bit key
	unsigned x = to_fixed(p.x, 1. / Lx) >> (32 - bits), y = to_fixed(p.y, 1. / Ly) >> (32 - bits), z = to_fixed(p.z, 1. / Lz) >> (32 - bits);
	keys[i] = hilbert_key(x, y, z, bits) | ((unsigned) (i / n_per) << (3 * bits));
	vals[i] = i;
}

// Same cell rule as the list builder (lists.cu cell_coord, src/Lists/Cells.h:60-65).  Sorting by the Hilbert index of the
// CELL coordinates makes one radix sort do both jobs: spatial locality for the gathers of the force kernels, and binning
// (a cell's members end up in consecutive slots, so the list builder needs no second sort and no index indirection).
__device__ __forceinline__ int cell_coord_s(double x, double L, int n) {
	double f = x / L - floor(x / L);
	int c = (int) (f * (1. - 2.220446049250313e-16) * n);
	return min(c, n - 1);
}

__global__ void __launch_bounds__(256) k_cell_hilbert_keys(int N, int n_per, const double4 *__restrict__ posd, double Lx, double Ly, double Lz, int nx, int ny, int nz,
		int bits, unsigned *__restrict__ keys, int *__restrict__ vals, int *__restrict__ flags) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i == 0) prof_mark(flags, OXB_PROF_SORT);
	if(i >= N) return;
	double4 p = posd[i];
	keys[i] = hilbert_key((unsigned) cell_coord_s(p.x, Lx, nx), (unsigned) cell_coord_s(p.y, Ly, ny), (unsigned) cell_coord_s(p.z, Lz, nz), bits)
			| ((unsigned) (i / n_per) << (3 * bits));
	vals[i] = i;
}

// Small systems: the whole ordering step -- cell keys, stable sort by key, inverse permutation -- in ONE cooperative launch.  At 81,920
// particles the seven launches of the key kernel + cub::DeviceRadixSort + inversion take 73 us on the device timeline for ~20 us of work
// (each of the tiny kernels has a floor of 3-4 us plus launch latency behind a host synchronisation).  Counting sort over the Hilbert
// keys of the cells (2^(3 bits) bins): histogram with atomics, three-stage exclusive scan, scatter; ties are then put in the order of
// the old slots (a bin holds the few members of one cell), so the permutation is the same STABLE sort the radix path produces and
// everything downstream stays bit-identical.  Grid = resident blocks only (cooperative launch), grid-stride loops throughout.
__global__ void __launch_bounds__(256) k_sort_small(oxb::SortArgs a, int bits, int nbins, int *__restrict__ hist, int *__restrict__ block_sums, int *__restrict__ tmp) {
	namespace cg = cooperative_groups;
	cg::grid_group grid = cg::this_grid();
	const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
	if(tid == 0) prof_mark(a.flags, OXB_PROF_SORT);
	for(int b = tid; b < nbins; b += nthr) hist[b] = 0;
	grid.sync();
	// keys + occupancy of the bins (the atomic's return value is NOT used as the rank: it depends on the scheduling)
	for(int i = tid; i < a.N; i += nthr) {
		const double4 p = a.posd[i];
		const unsigned key = hilbert_key((unsigned) cell_coord_s(p.x, a.box[0], a.ncell[0]), (unsigned) cell_coord_s(p.y, a.box[1], a.ncell[1]),
				(unsigned) cell_coord_s(p.z, a.box[2], a.ncell[2]), bits);
		a.keys[i] = key;
		atomicAdd(hist + key, 1);
	}
	grid.sync();
	// exclusive scan of the bins: every block scans one contiguous chunk, block 0 scans the chunk totals, every block adds its offset
	__shared__ int s_warp[8];
	__shared__ int s_carry;
	const int chunk = (nbins + gridDim.x - 1) / gridDim.x;
	const int c0 = blockIdx.x * chunk, c1 = min(c0 + chunk, nbins);
	if(threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for(int base = c0; base < c1; base += blockDim.x) {
		const int b = base + threadIdx.x;
		const int v = (b < c1) ? hist[b] : 0;
		int x = v;
		for(int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if((threadIdx.x & 31) >= o) x += y; }
		if((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
		__syncthreads();
		int woff = 0;
		for(int w = 0; w < (int) (threadIdx.x >> 5); w++) woff += s_warp[w];
		const int carry = s_carry;
		if(b < c1) hist[b] = carry + woff + x - v; // exclusive, relative to the chunk
		__syncthreads();
		if(threadIdx.x == blockDim.x - 1) s_carry = carry + woff + x;
		__syncthreads();
	}
	if(threadIdx.x == 0) block_sums[blockIdx.x] = s_carry;
	grid.sync();
	if(blockIdx.x == 0) {
		// gridDim.x totals (a few hundred): one block, same warp-scan scheme
		if(threadIdx.x == 0) s_carry = 0;
		__syncthreads();
		for(int base = 0; base < (int) gridDim.x; base += blockDim.x) {
			const int b = base + threadIdx.x;
			const int v = (b < (int) gridDim.x) ? block_sums[b] : 0;
			int x = v;
			for(int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if((threadIdx.x & 31) >= o) x += y; }
			if((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
			__syncthreads();
			int woff = 0;
			for(int w = 0; w < (int) (threadIdx.x >> 5); w++) woff += s_warp[w];
			const int carry = s_carry;
			if(b < (int) gridDim.x) block_sums[b] = carry + woff + x - v;
			__syncthreads();
			if(threadIdx.x == blockDim.x - 1) s_carry = carry + woff + x;
			__syncthreads();
		}
	}
	grid.sync();
	{
		const int off = block_sums[blockIdx.x];
		for(int b = c0 + threadIdx.x; b < c1; b += blockDim.x) hist[b] += off;
	}
	grid.sync();
	// scatter in arbitrary order inside each bin (second counter array: tmp[N ...] would do, the bins' cursors live behind the N slots)
	int *cursor = tmp + a.N;
	for(int b = tid; b < nbins; b += nthr) cursor[b] = 0;
	grid.sync();
	for(int i = tid; i < a.N; i += nthr) {
		const unsigned key = a.keys[i];
		tmp[hist[key] + atomicAdd(cursor + key, 1)] = i;
	}
	grid.sync();
	// stable order inside the bins: rank = members with a smaller old slot
	for(int i = tid; i < a.N; i += nthr) {
		const unsigned key = a.keys[i];
		const int s0 = hist[key], n = cursor[key];
		int rank = 0;
		for(int k = 0; k < n; k++) rank += (tmp[s0 + k] < i) ? 1 : 0;
		const int pos = s0 + rank;
		a.vals_sorted[pos] = i;
		a.inv[i] = pos;
		a.keys_sorted[pos] = key;
	}
}

__global__ void __launch_bounds__(256) k_invert(int N, const int *__restrict__ perm, int *__restrict__ inv) {
	int n = blockIdx.x * blockDim.x + threadIdx.x;
	if(n < N) inv[perm[n]] = n;
}

// One gather pass into the second state buffer.  When the sort key is the Hilbert index of the list builder's cells (a.cell_lin set), the
// pass also does the list builder's bookkeeping for the new slots, which would otherwise cost a second pass over the particles
// (lists.cu: k_cell_ranges): linear cell id, the (start, end) slot range of every cell (equal keys = equal cell, the table is zeroed by
// the caller) and the staleness references in the .w lanes of the FP64 state.  F and T come out zero: a re-sort invalidates the forces
// (mid-step they have been consumed and zeroed by the integrator), so gathering them would move 64 B per particle for nothing.
__global__ void __launch_bounds__(256) k_permute(oxb::PermuteArgs a) {
	int n = blockIdx.x * blockDim.x + threadIdx.x;
	if(n == 0 && a.flags != nullptr) prof_mark(a.flags, OXB_PROF_PERMUTE);
	if(n >= a.N) return;
	int o = a.perm[n];
	double4 pd = a.posd_in[o], vd = a.veld_in[o], ld = a.Ld_in[o];
	int4 ip = a.ipos_in[o];
	const int4 ib = a.iback_in[o];
	const float4 ax0 = a.axf_in[2 * (size_t) o], ax1 = a.axf_in[2 * (size_t) o + 1];
	if(a.cell_lin != nullptr) {
		const int cl = cell_coord_s(pd.x, a.box[0], a.ncell[0]) + a.ncell[0] * (cell_coord_s(pd.y, a.box[1], a.ncell[1]) + a.ncell[1] * cell_coord_s(pd.z, a.box[2], a.ncell[2]))
				+ (n / a.n_per) * (a.ncell[0] * a.ncell[1] * a.ncell[2]);
		a.cell_lin[n] = cl;
		if(a.cell_start != nullptr) {
			const unsigned k = a.keys_sorted[n];
			if(n == 0 || a.keys_sorted[n - 1] != k) a.cell_start[cl] = n;
			if(n == a.N - 1 || a.keys_sorted[n + 1] != k) a.cell_end[cl] = n + 1;
			int4 bs = ip;
			bs.x = (int) ((unsigned) ip.x + (unsigned) (int) rintf(ax0.x * a.base_a1 / a.boxf.sx));
			bs.y = (int) ((unsigned) ip.y + (unsigned) (int) rintf(ax0.y * a.base_a1 / a.boxf.sy));
			bs.z = (int) ((unsigned) ip.z + (unsigned) (int) rintf(ax0.z * a.base_a1 / a.boxf.sz));
			pd.w = pack_ref(ip);
			vd.w = pack_ref(ib);
			ld.w = pack_ref(bs);
		}
	}
	a.posd_out[n] = pd;
	a.veld_out[n] = vd;
	a.Ld_out[n] = ld;
	a.quatd_out[n] = a.quatd_in[o];
	a.ipos_out[n] = ip;
	a.iback_out[n] = ib;
	a.axf_out[2 * (size_t) n] = ax0;
	a.axf_out[2 * (size_t) n + 1] = ax1;
	a.F_out[n] = make_float4(0.f, 0.f, 0.f, 0.f);
	a.T_out[n] = make_float4(0.f, 0.f, 0.f, 0.f);
	int2 b = a.bonds_in[o];
	b.x = (b.x >= 0) ? a.inv[b.x] : b.x;
	b.y = (b.y >= 0) ? a.inv[b.y] : b.y;
	a.bonds_out[n] = b;
	a.slot_of[word_index(ip.w)] = n;
}

} // namespace

namespace oxb {

size_t sort_tmp_bytes(int N) {
	size_t a = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned *) nullptr, (unsigned *) nullptr, (int *) nullptr, (int *) nullptr, N, 0, 32);
	return a + 256;
}

// scratch of the one-launch ordering of small systems: bins + chunk totals + (N + bins) ints; 0 if the system does not qualify
size_t sort_small_bytes(int N, const int ncell[3], int n_rep) {
	static const bool off = [] { const char *e = getenv("OXB_SORT_SMALL"); return e != nullptr && e[0] == '0'; }();
	if(off || n_rep > 1 || ncell[0] <= 0 || N > 400000) return 0;
	int bits = 1;
	while((1 << bits) < std::max(ncell[0], std::max(ncell[1], ncell[2]))) bits++;
	const long long nbins = 1ll << (3 * bits);
	if(nbins > (1ll << 21) || nbins > 16ll * N + 4096) return 0; // very dilute boxes: the radix sort does not care about empty bins
	return sizeof(int) * (size_t) (2 * nbins + 4096 + N);
}

void launch_hilbert_order(cudaStream_t s, const SortArgs &a) {
	int tpb = 256, nb = (a.N + tpb - 1) / tpb;
	if(a.small_tmp != nullptr) {
		int bits = 1;
		while((1 << bits) < std::max(a.ncell[0], std::max(a.ncell[1], a.ncell[2]))) bits++;
		int nbins = 1 << (3 * bits);
		static int max_blocks = 0;
		if(max_blocks == 0) {
			int per_sm = 0, dev = 0, sms = 0;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sort_small, 256, 0);
			max_blocks = std::max(1, per_sm * sms);
		}
		int grid = std::min(max_blocks, std::min(4096, std::max(nb, (nbins + 255) / 256)));
		int *hist = a.small_tmp, *block_sums = hist + nbins, *tmp = block_sums + 4096;
		SortArgs aa = a;
		void *args[] = { (void *) &aa, (void *) &bits, (void *) &nbins, (void *) &hist, (void *) &block_sums, (void *) &tmp };
		cudaLaunchCooperativeKernel((const void *) k_sort_small, dim3(grid), dim3(256), args, 0, s);
		return;
	}
	size_t tmp = a.cub_tmp_bytes;
	int rep_bits = 0;
	while((1 << rep_bits) < a.n_rep) rep_bits++;
	if(a.ncell[0] > 0) {
		int bits = 1;
		while((1 << bits) < std::max(a.ncell[0], std::max(a.ncell[1], a.ncell[2]))) bits++;
		k_cell_hilbert_keys<<<nb, tpb, 0, s>>>(a.N, a.n_per, a.posd, a.box[0], a.box[1], a.box[2], a.ncell[0], a.ncell[1], a.ncell[2], bits, a.keys, a.vals, a.flags);
		cub::DeviceRadixSort::SortPairs(a.cub_tmp, tmp, a.keys, a.keys_sorted, a.vals, a.vals_sorted, a.N, 0, 3 * bits + rep_bits, s);
	}
	else {
		k_hilbert_keys<<<nb, tpb, 0, s>>>(a.N, a.n_per, a.posd, a.box[0], a.box[1], a.box[2], a.keys, a.vals, a.flags);
		cub::DeviceRadixSort::SortPairs(a.cub_tmp, tmp, a.keys, a.keys_sorted, a.vals, a.vals_sorted, a.N, 0, (a.n_rep > 1 ? 24 : 30) + rep_bits, s);
	}
	k_invert<<<nb, tpb, 0, s>>>(a.N, a.vals_sorted, a.inv);
}

void launch_permute(cudaStream_t s, const PermuteArgs &a) {
	int tpb = 256;
	if(a.cell_lin != nullptr && a.cell_start != nullptr) cudaMemsetAsync(a.cell_start, 0, sizeof(int) * 2 * (size_t) a.ncells_total, s);
	k_permute<<<(a.N + tpb - 1) / tpb, tpb, 0, s>>>(a);
}

} // namespace oxb
