// Force kernels of the oxDNA2 step (sm_100a).  Replaces dna_forces, dna_forces_edge_nonbonded, dna_forces_edge_bonded,
// sum_edge_forces_torques and set_external_forces of the reference (src/CUDA/Interactions/CUDA_DNA.cuh:727-904,
// src/CUDA/Interactions/CUDABaseInteraction.cu:21-48, src/CUDA/Backends/CUDA_MD.cuh:97-554).
//
// Data: ipos int4 (fixed-point position, .w = btype<<22 | original index), quat float4, bonds int2 (n3, n5 slots),
// neighbour matrix column-major nbr[k * stride + i], forces/torques float4 (.w = energy / HB energy as in the reference).
#include "dna_model.cuh"
#include "kernels.h"

namespace {

struct Particle {
	int4 ip;
	Axes ax;
	v3 back;
	int btype;
};

__device__ __forceinline__ Particle load_particle(const oxb_dna2_params &M, const int4 *__restrict__ ipos, const float4 *__restrict__ quat, int i) {
	Particle P;
	P.ip = __ldg(ipos + i);
	P.ax = axes_from_quat(__ldg(quat + i));
	P.back = P.ax.a1 * M.back_a1 + P.ax.a2 * M.back_a2;
	P.btype = word_btype(P.ip.w);
	return P;
}

__device__ __forceinline__ v3 to_body(const Axes &A, v3 t) { return mk3(dot(A.a1, t), dot(A.a2, t), dot(A.a3, t)); }

// ------------------------------------------------------------------------------------------------------------
// Particle-centric: one thread per particle, every listed pair evaluated from both ends, no atomics, deterministic.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_forces_particle(const __grid_constant__ oxb_dna2_params M, BoxF box, int N, const int4 *__restrict__ ipos,
		const float4 *__restrict__ quat, const int2 *__restrict__ bonds, const int *__restrict__ nbr, const int *__restrict__ nnbr, int stride,
		float4 *__restrict__ F, float4 *__restrict__ T, int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;

	Particle P = load_particle(M, ipos, quat, i);
	int2 b = __ldg(bonds + i);
	bool p_end = (b.x < 0 || b.y < 0);

	v3 f = mk3(0.f, 0.f, 0.f), t = mk3(0.f, 0.f, 0.f);
	float e = 0.f, ehb = 0.f;
	bool broken = false;

	if(b.x >= 0) { // I am the 5' side of the bond (p), q = my n3
		Particle Q = load_particle(M, ipos, quat, b.x);
		PairAcc acc;
		acc.clear();
		e += dna2_bonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, P.back, Q.back, acc, broken);
		f -= acc.F;
		t += acc.torque_p(P.ax, P.back);
	}
	if(b.y >= 0) { // my n5 neighbour is p, I am q
		Particle Q = load_particle(M, ipos, quat, b.y);
		PairAcc acc;
		acc.clear();
		e += dna2_bonded(M, min_image_fixed(box, Q.ip, P.ip), Q.ax, P.ax, Q.btype, P.btype, Q.back, P.back, acc, broken);
		f += acc.F;
		t += acc.torque_q(P.ax, P.back);
	}

	int nn = __ldg(nnbr + i);
	for(int k = 0; k < nn; k++) {
		int j = __ldg(nbr + (size_t) k * stride + i);
		Particle Q = load_particle(M, ipos, quat, j);
		int2 bq = __ldg(bonds + j);
		PairAcc acc;
		acc.clear();
		PairEnergy pe = dna2_nonbonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, p_end, (bq.x < 0 || bq.y < 0), P.back,
				Q.back, acc);
		e += pe.total;
		ehb += pe.hb;
		f -= acc.F;
		t += acc.torque_p(P.ax, P.back);
	}

	v3 tb = to_body(P.ax, t);
	F[i] = make_float4(f.x, f.y, f.z, e);
	T[i] = make_float4(tb.x, tb.y, tb.z, ehb);
	if(broken) atomicOr(flags + OXB_FLAG_ERROR, OXB_ERR_FENE_BROKEN);
}

// ------------------------------------------------------------------------------------------------------------
// Edge-centric: one thread per unique pair.  Edges are grouped by `from` (list order), so lanes of a warp that share the
// same `from` particle first reduce among themselves with shuffles (segmented suffix sum); only segment heads issue
// the vector atomic for the `from` side.  The `to` side is scattered with one float4 atomic per array.
// F/T hold lab-frame sums here; k_forces_edge_bonded finishes the job.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_add4(float4 *dst, float x, float y, float z, float w) {
	// sm_90+ 128-bit vector reduction (red.global.add.v4.f32)
	atomicAdd(dst, make_float4(x, y, z, w));
}

__global__ void __launch_bounds__(128) k_forces_edge_nonbonded(const __grid_constant__ oxb_dna2_params M, BoxF box, const int *__restrict__ n_edges,
		const int4 *__restrict__ ipos, const float4 *__restrict__ quat, const int2 *__restrict__ bonds, const int2 *__restrict__ edges,
		float4 *__restrict__ F, float4 *__restrict__ T, const int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	const int ne = *n_edges;
	const unsigned lane = threadIdx.x & 31;
	for(int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < ne; base += gridDim.x * blockDim.x) {
		int eidx = base + lane;
		bool valid = eidx < ne;
		int2 ed = valid ? __ldg(edges + eidx) : make_int2(-1 - (int) lane, -1);
		float fx = 0.f, fy = 0.f, fz = 0.f, fe = 0.f, tx = 0.f, ty = 0.f, tz = 0.f, th = 0.f;
		if(valid) {
			Particle P = load_particle(M, ipos, quat, ed.x);
			Particle Q = load_particle(M, ipos, quat, ed.y);
			int2 bp = __ldg(bonds + ed.x), bq = __ldg(bonds + ed.y);
			PairAcc acc;
			acc.clear();
			PairEnergy pe = dna2_nonbonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, (bp.x < 0 || bp.y < 0),
					(bq.x < 0 || bq.y < 0), P.back, Q.back, acc);
			if(pe.total != 0.f || acc.F.x != 0.f || acc.F.y != 0.f || acc.F.z != 0.f) {
				v3 tq = acc.torque_q(Q.ax, Q.back);
				atomic_add4(F + ed.y, acc.F.x, acc.F.y, acc.F.z, pe.total);
				atomic_add4(T + ed.y, tq.x, tq.y, tq.z, pe.hb);
				v3 tp = acc.torque_p(P.ax, P.back);
				fx = -acc.F.x; fy = -acc.F.y; fz = -acc.F.z; fe = pe.total;
				tx = tp.x; ty = tp.y; tz = tp.z; th = pe.hb;
			}
		}
		// segmented suffix reduction over lanes with equal `from`
		int key = ed.x;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1) {
			int okey = __shfl_down_sync(0xffffffffu, key, d);
			float ofx = __shfl_down_sync(0xffffffffu, fx, d), ofy = __shfl_down_sync(0xffffffffu, fy, d);
			float ofz = __shfl_down_sync(0xffffffffu, fz, d), ofe = __shfl_down_sync(0xffffffffu, fe, d);
			float otx = __shfl_down_sync(0xffffffffu, tx, d), oty = __shfl_down_sync(0xffffffffu, ty, d);
			float otz = __shfl_down_sync(0xffffffffu, tz, d), oth = __shfl_down_sync(0xffffffffu, th, d);
			if(lane + d < 32 && okey == key) {
				fx += ofx; fy += ofy; fz += ofz; fe += ofe;
				tx += otx; ty += oty; tz += otz; th += oth;
			}
		}
		int pkey = __shfl_up_sync(0xffffffffu, key, 1);
		bool head = (lane == 0) || (pkey != key);
		if(valid && head && (fx != 0.f || fy != 0.f || fz != 0.f || fe != 0.f || tx != 0.f || ty != 0.f || tz != 0.f)) {
			atomic_add4(F + key, fx, fy, fz, fe);
			atomic_add4(T + key, tx, ty, tz, th);
		}
	}
}

// bonded terms per particle + lab->body rotation of the accumulated torque (must run after the edge kernel)
__global__ void __launch_bounds__(128) k_forces_edge_bonded(const __grid_constant__ oxb_dna2_params M, BoxF box, int N, const int4 *__restrict__ ipos,
		const float4 *__restrict__ quat, const int2 *__restrict__ bonds, float4 *__restrict__ F, float4 *__restrict__ T, int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	Particle P = load_particle(M, ipos, quat, i);
	int2 b = __ldg(bonds + i);
	float4 f4 = F[i], t4 = T[i];
	v3 f = mk3(f4.x, f4.y, f4.z), t = mk3(t4.x, t4.y, t4.z);
	float e = f4.w;
	bool broken = false;
	if(b.x >= 0) {
		Particle Q = load_particle(M, ipos, quat, b.x);
		PairAcc acc;
		acc.clear();
		e += dna2_bonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, P.back, Q.back, acc, broken);
		f -= acc.F;
		t += acc.torque_p(P.ax, P.back);
	}
	if(b.y >= 0) {
		Particle Q = load_particle(M, ipos, quat, b.y);
		PairAcc acc;
		acc.clear();
		e += dna2_bonded(M, min_image_fixed(box, Q.ip, P.ip), Q.ax, P.ax, Q.btype, P.btype, Q.back, P.back, acc, broken);
		f += acc.F;
		t += acc.torque_q(P.ax, P.back);
	}
	v3 tb = to_body(P.ax, t);
	F[i] = make_float4(f.x, f.y, f.z, e);
	T[i] = make_float4(tb.x, tb.y, tb.z, t4.w);
	if(broken) atomicOr(flags + OXB_FLAG_ERROR, OXB_ERR_FENE_BROKEN);
}

// ------------------------------------------------------------------------------------------------------------
// External forces: one thread per force entry (compact table, not 15 slots x N as in the reference).  Runs after the
// interaction kernels and adds into F.  Particle ids in the table are ORIGINAL ids, mapped through slot_of, so the
// Hilbert re-sort needs no table rewrite (the reference forbids sort + external forces, MD_CUDABackend.cu:110-112).
// ------------------------------------------------------------------------------------------------------------
__global__ void k_ext_forces(int n, const DevExtForce *__restrict__ ef, const int *__restrict__ slot_of, const int4 *__restrict__ ipos,
		const double4 *__restrict__ posd, BoxF box, long long step, float4 *__restrict__ F, const int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if(k >= n) return;
	DevExtForce e = ef[k];
	int i = slot_of[e.particle];
	float st = (float) step;
	v3 f;
	if(e.type == OXB_EXT_STRING) {
		// ConstantRateForce.cpp:52-61
		float s = e.F0 + e.rate * st;
		f = mk3(e.dir[0] * s, e.dir[1] * s, e.dir[2] * s);
	}
	else if(e.type == OXB_EXT_TRAP) {
		// MovingTrap.cpp:50-64: uses the absolute (unwrapped) position, kept in double
		double4 p = posd[i];
		double rs = (double) e.rate * (double) step;
		f = mk3((float) (-(double) e.stiff * (p.x - (e.pos0[0] + rs * e.dir[0]))), (float) (-(double) e.stiff * (p.y - (e.pos0[1] + rs * e.dir[1]))),
				(float) (-(double) e.stiff * (p.z - (e.pos0[2] + rs * e.dir[2]))));
	}
	else {
		// MutualTrap.cpp:54-66
		int j = slot_of[e.ref];
		v3 dr;
		if(e.pbc) dr = min_image_fixed(box, ipos[i], ipos[j]);
		else {
			double4 p = posd[i], q = posd[j];
			dr = mk3((float) (q.x - p.x), (float) (q.y - p.y), (float) (q.z - p.z));
		}
		float m = sqrtf(dot(dr, dr));
		float s = (m - (e.r0 + e.rate * st)) * (e.stiff + e.stiff_rate * st) / m;
		f = dr * s;
	}
	atomicAdd(&F[i].x, f.x);
	atomicAdd(&F[i].y, f.y);
	atomicAdd(&F[i].z, f.z);
}

} // namespace

namespace oxb {

void launch_forces_particle(cudaStream_t s, const oxb_dna2_params &M, BoxF box, int N, const int4 *ipos, const float4 *quat, const int2 *bonds,
		const int *nbr, const int *nnbr, int stride, float4 *F, float4 *T, int *flags, int hw) {
	int tpb = 128;
	k_forces_particle<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(M, box, N, ipos, quat, bonds, nbr, nnbr, stride, F, T, flags, hw);
}

void launch_forces_edge(cudaStream_t s, const oxb_dna2_params &M, BoxF box, int N, const int *n_edges, int edge_capacity_hint, const int4 *ipos,
		const float4 *quat, const int2 *bonds, const int2 *edges, float4 *F, float4 *T, int *flags, int hw, int n_sm) {
	cudaMemsetAsync(F, 0, sizeof(float4) * (size_t) N, s);
	cudaMemsetAsync(T, 0, sizeof(float4) * (size_t) N, s);
	int tpb = 128;
	// grid-stride over the device-side edge count: enough CTAs to cover the expected edge count, a multiple of the SM count
	long long want = ((long long) edge_capacity_hint + tpb - 1) / tpb;
	int per_sm = (int) ((want + n_sm - 1) / n_sm);
	if(per_sm < 1) per_sm = 1;
	int blocks = per_sm * n_sm;
	k_forces_edge_nonbonded<<<blocks, tpb, 0, s>>>(M, box, n_edges, ipos, quat, bonds, edges, F, T, flags, hw);
	k_forces_edge_bonded<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(M, box, N, ipos, quat, bonds, F, T, flags, hw);
}

void launch_ext_forces(cudaStream_t s, int n, const DevExtForce *ef, const int *slot_of, const int4 *ipos, const double4 *posd, BoxF box,
		long long step, float4 *F, const int *flags, int hw) {
	if(n <= 0) return;
	int tpb = 128;
	k_ext_forces<<<(n + tpb - 1) / tpb, tpb, 0, s>>>(n, ef, slot_of, ipos, posd, box, step, F, flags, hw);
}

} // namespace oxb
