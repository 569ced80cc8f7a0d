// oxDNA3 force pass, particle-centric (sm_100a; `use_edge = 0`).  Replaces DNA3_forces of the reference
// (src/CUDA/Interactions/CUDA_DNA3.cuh:880-1085, CUDADNA3Interaction.cu:171-212).  One thread per particle over the full Verlet matrix,
// every listed pair evaluated from both ends: no atomics, deterministic, the same launch shape as k_forces_particle (forces.cu).
// (`use_edge = 1`, DNA3_forces_edge_nonbonded / _bonded: the staged edge pipeline, forces.cu k3_*.)
//
// Data: as forces.cu, plus the packed parameter records and the per-particle type word of dna3_model.cuh.
#include "dna3_model.cuh"
#include "kernels.h"

#include <cstdlib>

namespace {

struct P3 {
	int4 ip;
	Axes ax;
	v3 back;
	int btype;
	Nuc3 n;
};

__device__ __forceinline__ P3 load_p3(const oxb_dna3_dev &M, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, int i) {
	P3 P;
	P.ip = __ldg(ipos + i);
	P.ax = load_axes(axf, i);
	P.back = P.ax.a1 * M.back_a1 + P.ax.a2 * M.back_a2;
	P.btype = word_btype(P.ip.w);
	P.n = nuc3_from_code(__ldg(M.tcode + word_index(P.ip.w)));
	return P;
}

__device__ __forceinline__ ExclRefine refine3(const oxb_dna3_dev &M, const BoxF &box, const double4 *posd, const double4 *quatd) {
	ExclRefine R;
	R.posd = posd; R.quatd = quatd; R.sp = R.sq = 0;
	R.L[0] = box.dsx * 4294967296.; R.L[1] = box.dsy * 4294967296.; R.L[2] = box.dsz * 4294967296.;
	R.b1 = (double) M.back_a1; R.b2 = (double) M.back_a2; R.b3 = 0.;
	return R;
}

// the bond p -> q = n3(p): record of the tetramer (n3(q), q, p, n5(p)); FENE in double from the fixed-point backbone sites (mixed precision)
__device__ __forceinline__ float bond3(const oxb_dna3_dev &M, const BoxF &box, const P3 &P, const P3 &Q, const int4 *__restrict__ iback, int sp, int sq,
		bool refine, PairAcc &acc, bool &broken, float *esplit) {
	const float4 *rec = M.bonded + ix4(Q.n.n3t, Q.n.type, P.n.type, P.n.n5t) * (OXB3_REC_BONDED / 4);
	FeneSite fs;
	if(refine) fs = fene_from_sites(fene3_of(M, rec), box, __ldg(iback + sp), __ldg(iback + sq), broken);
	return dna3_bonded(M, rec, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.n, Q.n, P.back, Q.back, acc, broken, esplit, refine ? &fs : nullptr);
}

// The neighbour loop is split by COST into uniform passes (the single loop ran with 8.9 of 32 lanes active: one lane inside the six-angle
// hydrogen-bonding code held the 31 others; ncu r02af): per chunk of 64 neighbours
//   pass 1  every listed pair: fixed-point centre + backbone site of the neighbour (2 x 16 B), Debye-Hueckel, "near" bit if any other term can reach
//   pass 2  near pairs: full particle record, the four excluded-volume site pairs, bits for "bases in range" / "stacking sites in range"
//   pass 3  base-base contacts: hydrogen bonding + cross stacking        pass 4  stack-stack contacts: coaxial stacking
// The p-side accumulators (force, lever sums, pure torque) are linear in the pairs: one PairAcc is carried through all passes and the two
// cross products of the torque are taken once.
template<int MB, int LPP>
__global__ void __launch_bounds__(128, MB) k_forces_dna3(const __grid_constant__ oxb_dna3_dev M, BoxF box, int N, const int4 *__restrict__ ipos,
		const int4 *__restrict__ iback, const float4 *__restrict__ axf, const double4 *__restrict__ posd, const double4 *__restrict__ quatd,
		const int2 *__restrict__ bonds, const int *__restrict__ nbr, const int *__restrict__ nnbr, int stride,
		float4 *__restrict__ F, float4 *__restrict__ T, int *__restrict__ flags, int hw) {
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(flags, flags[hw] ? OXB_PROF_WAIT : OXB_PROF_FORCE);
	if(flags[hw]) return;
	// LPP = 2: lanes 2m and 2m + 1 share particle m -- alternate neighbours in every pass, one bond each -- and add their sums with shuffles:
	// half the dependent chain per thread, twice the blocks (82k nucleotides are 1.4 waves of 128-thread blocks at 3 blocks per SM)
	const int gid = blockIdx.x * blockDim.x + threadIdx.x;
	const int i = LPP == 2 ? gid >> 1 : gid, sub = LPP == 2 ? gid & 1 : 0;
	if(i >= N) return;
	const unsigned am = __activemask(); // the lanes that take part in the final shuffles (both lanes of a particle leave together)
	const P3 P = load_p3(M, ipos, axf, i);
	const int4 ibp = __ldg(iback + i);
	const int2 b = __ldg(bonds + i);
	const bool p_end = (b.x < 0 || b.y < 0);
	float e = 0.f, ehb = 0.f;
	bool broken = false;
	ExclRefine R = refine3(M, box, posd, quatd);
	const bool refine = posd != nullptr; // backend_precision = mixed
	PairAcc acc; // everything in which this particle is "p"
	acc.clear();
	acc.refine = refine ? &R : nullptr;
	v3 fq = mk3(0.f, 0.f, 0.f), tq = mk3(0.f, 0.f, 0.f); // the bond in which it is "q"
	if(b.x >= 0 && sub == 0) { // I am the 5' side of the bond (p), q = my n3
		const P3 Q = load_p3(M, ipos, axf, b.x);
		R.sp = i; R.sq = b.x;
		e += bond3(M, box, P, Q, iback, i, b.x, refine, acc, broken, nullptr);
	}
	if(b.y >= 0 && sub == LPP - 1) { // my n5 neighbour is p, I am q
		const P3 Q = load_p3(M, ipos, axf, b.y);
		PairAcc a2;
		a2.clear();
		R.sp = b.y; R.sq = i; a2.refine = refine ? &R : nullptr;
		e += bond3(M, box, Q, P, iback, b.y, i, refine, a2, broken, nullptr);
		fq = a2.F;
		tq = a2.torque_q(P.ax, P.back);
	}
	R.sp = i;
	const int nn = __ldg(nnbr + i);
	for(int base = 0; base < nn; base += 64) {
		const int cnt = min(64, nn - base);
		const int *__restrict__ row = nbr + (size_t) base * stride + i;
		unsigned long long near = 0ull, mb = 0ull, ms = 0ull;
#pragma unroll 4
		for(int k = sub; k < cnt; k += LPP) {
			const int j = __ldg(row + (size_t) k * stride);
			const int4 ipq = __ldg(ipos + j);
			const int4 ibq = __ldg(iback + j);
			const v3 r = min_image_fixed(box, P.ip, ipq);
			const float r2 = dot(r, r);
			const v3 rbb = min_image_fixed(box, ibp, ibq);
			float fs;
			float en = dna2_dh_fast(M, dot(rbb, rbb), p_end, ibq.w & 1, fs);
			if(r2 >= M.rcut2) { en = 0.f; fs = 0.f; } // no interaction beyond the centre-centre cutoff (DNA2Interaction.cpp:46-48)
			e += en;
			acc.site_kk(rbb * fs);
			if(r2 < M.r2_near_max) near |= 1ull << k;
		}
		while(near) {
			const int k = __ffsll((long long) near) - 1;
			near &= near - 1ull;
			const int j = __ldg(row + (size_t) k * stride);
			const P3 Q = load_p3(M, ipos, axf, j);
			const v3 r = min_image_fixed(box, P.ip, Q.ip);
			const v3 a1d_b = Q.ax.a1 * M.pos_base[Q.n.si] - P.ax.a1 * M.pos_base[P.n.si];
			const v3 rb = r + a1d_b;
			const v3 rs = r + Q.ax.a1 * M.pos_stack[Q.n.si] - P.ax.a1 * M.pos_stack[P.n.si];
			if(dot(r, r) < M.r2_excl_max) {
				R.sq = j;
				e += dna3_excl4(M, r, r + Q.back - P.back, rb, P.ax, Q.ax, P.n, Q.n, P.back, Q.back, acc);
			}
			if(dot(rb, rb) < M.r2_base_max) mb |= 1ull << k;
			if(dot(rs, rs) < M.r2_stack_max) ms |= 1ull << k;
		}
		while(mb) {
			const int k = __ffsll((long long) mb) - 1;
			mb &= mb - 1ull;
			const int j = __ldg(row + (size_t) k * stride);
			const P3 Q = load_p3(M, ipos, axf, j);
			const v3 rb = min_image_fixed(box, P.ip, Q.ip) + Q.ax.a1 * M.pos_base[Q.n.si] - P.ax.a1 * M.pos_base[P.n.si];
			float eh;
			e += dna3_hbcr(M, rb, dot(rb, rb), P.ax, Q.ax, P.btype, Q.btype, P.n, Q.n, acc, eh);
			ehb += eh;
		}
		while(ms) {
			const int k = __ffsll((long long) ms) - 1;
			ms &= ms - 1ull;
			const int j = __ldg(row + (size_t) k * stride);
			const P3 Q = load_p3(M, ipos, axf, j);
			const v3 rs = min_image_fixed(box, P.ip, Q.ip) + Q.ax.a1 * M.pos_stack[Q.n.si] - P.ax.a1 * M.pos_stack[P.n.si];
			e += dna3_cxst(M, rs, dot(rs, rs), P.ax, Q.ax, P.n, Q.n, acc);
		}
	}
	// torque stays in the lab frame; the integrator rotates it into the body frame
	v3 f = fq - acc.F, t = tq + acc.torque_p(P.ax, P.back);
	if(LPP == 2) {
		f.x += __shfl_xor_sync(am, f.x, 1); f.y += __shfl_xor_sync(am, f.y, 1); f.z += __shfl_xor_sync(am, f.z, 1);
		t.x += __shfl_xor_sync(am, t.x, 1); t.y += __shfl_xor_sync(am, t.y, 1); t.z += __shfl_xor_sync(am, t.z, 1);
		e += __shfl_xor_sync(am, e, 1); ehb += __shfl_xor_sync(am, ehb, 1);
	}
	if(sub == 0) {
		F[i] = make_float4(f.x, f.y, f.z, e);
		T[i] = make_float4(t.x, t.y, t.z, ehb);
	}
	if(broken) atomicOr(flags + OXB_FLAG_ERROR, OXB_ERR_FENE_BROKEN);
}

// per-term energies (the CPU get_system_energy_split of the reference, BaseInteraction.cpp:61-90): every unique pair once, from its lower slot
__global__ void __launch_bounds__(128) k_energy_split_dna3(const __grid_constant__ oxb_dna3_dev M, BoxF box, int N, const int4 *__restrict__ ipos,
		const float4 *__restrict__ axf, const int2 *__restrict__ bonds, const int *__restrict__ nbr, const int *__restrict__ nnbr, int stride,
		double *__restrict__ out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	float e[OXB_NTERMS];
#pragma unroll
	for(int t = 0; t < OXB_NTERMS; t++) e[t] = 0.f;
	if(i < N) {
		const P3 P = load_p3(M, ipos, axf, i);
		const int2 b = __ldg(bonds + i);
		PairAcc acc;
		acc.clear();
		bool broken = false;
		if(b.x >= 0) {
			const P3 Q = load_p3(M, ipos, axf, b.x);
			bond3(M, box, P, Q, nullptr, i, b.x, false, acc, broken, e);
		}
		const int nn = __ldg(nnbr + i);
		for(int k = 0; k < nn; k++) {
			const int j = __ldg(nbr + (size_t) k * stride + i) & OXB_SLOT_MASK;
			if(j < i) continue;
			const P3 Q = load_p3(M, ipos, axf, j);
			dna3_nonbonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, P.n, Q.n, P.back, Q.back, acc, e);
		}
	}
	__shared__ double sh[OXB_NTERMS][4];
#pragma unroll
	for(int t = 0; t < OXB_NTERMS; t++) {
		double x = (double) e[t];
		for(int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
		if((threadIdx.x & 31) == 0) sh[t][threadIdx.x >> 5] = x;
	}
	__syncthreads();
	if(threadIdx.x < OXB_NTERMS) {
		const double x = sh[threadIdx.x][0] + sh[threadIdx.x][1] + sh[threadIdx.x][2] + sh[threadIdx.x][3];
		if(x != 0.) atomicAdd(out + threadIdx.x, x);
	}
}

} // namespace

namespace oxb {

void launch_forces_dna3(cudaStream_t s, const oxb_dna3_dev &M, BoxF box, int N, const int4 *ipos, const int4 *iback, const float4 *axf,
		const double4 *posd, const double4 *quatd, const int2 *bonds, const int *nbr, const int *nnbr, int stride, float4 *F, float4 *T, int *flags, int hw) {
	const int tpb = 128;
	// minimum resident blocks per SM asked of the compiler = register cap (4: 128 registers, 790 B of spills; 3: 168, 230 B; 2: 250, none).
	// Measured on B200 (profiles/sweeps_r02.txt, ag): OXB_DNA3_MB overrides
	static const int mb = [] { const char *v = getenv("OXB_DNA3_MB"); return (v != nullptr && v[0] != 0) ? atoi(v) : 3; }();
	static const int lpp = [] { const char *v = getenv("OXB_DNA3_LPP"); return (v != nullptr && v[0] != 0) ? atoi(v) : 1; }();
	const int blocks = (int) (((long long) N * (lpp == 2 ? 2 : 1) + tpb - 1) / tpb);
	if(lpp == 2) {
		if(mb >= 4) k_forces_dna3<4, 2><<<blocks, tpb, 0, s>>>(M, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, flags, hw);
		else k_forces_dna3<3, 2><<<blocks, tpb, 0, s>>>(M, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, flags, hw);
	}
	else if(mb >= 4) k_forces_dna3<4, 1><<<blocks, tpb, 0, s>>>(M, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, flags, hw);
	else if(mb == 3) k_forces_dna3<3, 1><<<blocks, tpb, 0, s>>>(M, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, flags, hw);
	else k_forces_dna3<2, 1><<<blocks, tpb, 0, s>>>(M, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, flags, hw);
}

void launch_energy_split_dna3(cudaStream_t s, const oxb_dna3_dev &M, BoxF box, int N, const int4 *ipos, const float4 *axf, const int2 *bonds,
		const int *nbr, const int *nnbr, int stride, double *out) {
	cudaMemsetAsync(out, 0, sizeof(double) * OXB_NTERMS, s);
	const int tpb = 128;
	k_energy_split_dna3<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(M, box, N, ipos, axf, bonds, nbr, nnbr, stride, out);
}

} // namespace oxb
