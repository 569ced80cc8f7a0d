// Force kernels of the oxDNA2 step (sm_100a).  Replaces dna_forces, dna_forces_edge_nonbonded, dna_forces_edge_bonded,
// sum_edge_forces_torques and set_external_forces of the reference (src/CUDA/Interactions/CUDA_DNA.cuh:727-904,
// src/CUDA/Interactions/CUDABaseInteraction.cu:21-48, src/CUDA/Backends/CUDA_MD.cuh:97-554).
//
// Data: ipos int4 (fixed-point position, .w = btype<<22 | original index), axf = FP32 orientation record (a1, a3; common.cuh), bonds int2 (n3, n5 slots),
// neighbour matrix column-major nbr[k * stride + i], forces/torques float4 (.w = energy / HB energy as in the reference).
#include "models.cuh"
#include "dna3_model.cuh"
#include "kernels.h"

#include <cstdlib>

#include <algorithm>

// minimum resident blocks per SM asked of the compiler for the latency-bound gather kernels (register cap = 65536 / (threads x blocks)).
// Tuned on B200 with profiles/micro/occupancy_sweep.sh (profiles/occupancy_sweep_r01.txt): at 92-96 registers these kernels ran at
// 28 % occupancy with 46-62 % issue utilisation (ncu r01h); capping them at 64 (near, heavy) / 85 (bonded) registers costs a few
// spilled words and gives +10 % on the 1M-nt step, +5 % at 82k nt.  48 registers (10/20 blocks) is slower again.
#ifndef OXB_MB_NEAR
#define OXB_MB_NEAR 8
#endif
#ifndef OXB_MB_HEAVY
#define OXB_MB_HEAVY 16
#endif
#ifndef OXB_MB_PARTICLE
#define OXB_MB_PARTICLE 5
#endif
#ifndef OXB_MB_BONDED
#define OXB_MB_BONDED 6
#endif

namespace {

template<class MD>
__device__ __forceinline__ ExclRefine make_refine(const typename MD::Params &M, const BoxF &box, const double4 *posd, const double4 *quatd) {
	ExclRefine R;
	R.posd = posd; R.quatd = quatd; R.sp = R.sq = 0;
	R.L[0] = box.dsx * 4294967296.; R.L[1] = box.dsy * 4294967296.; R.L[2] = box.dsz * 4294967296.;
	R.b1 = (double) M.back_a1; R.b2 = (double) M.back_a2; R.b3 = (double) MD::back_a3(M);
	return R;
}

__device__ __forceinline__ void atomic_add4(float4 *dst, float x, float y, float z, float w) {
	atomicAdd(dst, make_float4(x, y, z, w));
}

// One pair whose excluded-volume site pairs `mask` are in range, evaluated entirely in double from the FP64 state and added to F / T
// (lab frame; .w of F = pair energy, as everywhere).  Runs block-locally AFTER the main loop of the producing kernel, on the few pairs
// that were parked in shared memory: the hot loops stay free of FP64 and of its registers.
template<class MD>
__device__ __noinline__ void excl_double_item(const typename MD::Params &M, const BoxF &box, const double4 *__restrict__ posd, const double4 *__restrict__ qd,
		int p, int q, int mask, float4 *__restrict__ F, float4 *__restrict__ T) {
	const double L[3] = { box.dsx * 4294967296., box.dsy * 4294967296., box.dsz * 4294967296. };
	const double b1 = (double) M.back_a1, b2 = (double) M.back_a2, b3 = (double) MD::back_a3(M), cb = (double) M.base_a1, eps = (double) M.excl_eps;
	const double4 pp = posd[p], pq = posd[q], qp = qd[p], qq = qd[q];
	double rd[3] = { pq.x - pp.x, pq.y - pp.y, pq.z - pp.z };
	for(int k = 0; k < 3; k++) rd[k] -= L[k] * rint(rd[k] / L[k]);
	double a1p[3], a2p[3], a3p[3], a1q[3], a2q[3], a3q[3];
	quatd Qp = { qp.x, qp.y, qp.z, qp.w }, Qq = { qq.x, qq.y, qq.z, qq.w };
	axes_from_quatd(Qp, a1p, a2p, a3p);
	axes_from_quatd(Qq, a1q, a2q, a3q);
	double kp[3], kq[3], ap[3], aq[3]; // backbone and base sites relative to the centres
	for(int k = 0; k < 3; k++) {
		kp[k] = b1 * a1p[k] + b2 * a2p[k] + b3 * a3p[k]; kq[k] = b1 * a1q[k] + b2 * a2q[k] + b3 * a3q[k];
		ap[k] = cb * a1p[k]; aq[k] = cb * a1q[k];
	}
	double Fq[3] = { 0., 0., 0. }, Tp[3] = { 0., 0., 0. }, Tq[3] = { 0., 0., 0. }, E = 0.;
	// (kind, parameter block): backbone-backbone 0, base-base 1, base(p)-back(q) 2, back(p)-base(q) 3 -- as in dna2_excl / bonded_fene_excl
	const int kinds[4] = { OXB_SITE_KK, OXB_SITE_AA, OXB_SITE_AK, OXB_SITE_KA };
	for(int t = 0; t < 4; t++) {
		if(!(mask & (1 << kinds[t]))) continue;
		const oxb_excl &e = M.excl[t];
		const double *sp = (kinds[t] == OXB_SITE_KK || kinds[t] == OXB_SITE_KA) ? kp : ap;
		const double *sq = (kinds[t] == OXB_SITE_KK || kinds[t] == OXB_SITE_AK) ? kq : aq;
		const double d[3] = { rd[0] + sq[0] - sp[0], rd[1] + sq[1] - sp[1], rd[2] + sq[2] - sp[2] };
		const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
		if(!(r2 < (double) e.rc2)) continue;
		double sd, en;
		if(r2 > (double) e.rstar2) {
			const double m = sqrt(r2), x = m - (double) e.rc;
			en = eps * (double) e.b * x * x;
			sd = -2. * eps * (double) e.b * x / m;
		}
		else {
			const double tt = (double) e.sigma2 / r2, lj = tt * tt * tt;
			en = 4. * eps * (lj * lj - lj);
			sd = -24. * eps * (lj - 2. * lj * lj) / r2;
		}
		const double f[3] = { d[0] * sd, d[1] * sd, d[2] * sd }; // on q, at site sq; -f on p at site sp
		E += en;
		for(int k = 0; k < 3; k++) Fq[k] += f[k];
		Tq[0] += sq[1] * f[2] - sq[2] * f[1]; Tq[1] += sq[2] * f[0] - sq[0] * f[2]; Tq[2] += sq[0] * f[1] - sq[1] * f[0];
		Tp[0] -= sp[1] * f[2] - sp[2] * f[1]; Tp[1] -= sp[2] * f[0] - sp[0] * f[2]; Tp[2] -= sp[0] * f[1] - sp[1] * f[0];
	}
	if(E == 0.) return;
	atomic_add4(F + q, (float) Fq[0], (float) Fq[1], (float) Fq[2], (float) E);
	atomic_add4(T + q, (float) Tq[0], (float) Tq[1], (float) Tq[2], 0.f);
	atomic_add4(F + p, (float) -Fq[0], (float) -Fq[1], (float) -Fq[2], (float) E);
	atomic_add4(T + p, (float) Tp[0], (float) Tp[1], (float) Tp[2], 0.f);
}

struct Particle {
	int4 ip;
	Axes ax;
	v3 back;
	int btype;
};

template<class MD>
__device__ __forceinline__ Particle load_particle(const typename MD::Params &M, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, int i) {
	Particle P;
	P.ip = __ldg(ipos + i);
	P.ax = load_axes(axf, i);
	P.back = MD::back(M, P.ax);
	P.btype = word_btype(P.ip.w);
	return P;
}

// one coaxial-stacking pair of a block's own work-list segment, evaluated in the TAIL of the producing kernel (fold = 1): the list is tiny
// (a few hundredths of an entry per particle), so a separate launch costs more than it computes.  Not inlined: the hot loop of the
// producer keeps its register footprint.
template<class MD>
__device__ __noinline__ void cxst_item(const typename MD::Params &M, const BoxF &box, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, int2 ed,
		float4 *__restrict__ F, float4 *__restrict__ T) {
	Particle P = load_particle<MD>(M, ipos, axf, ed.x);
	Particle Q = load_particle<MD>(M, ipos, axf, ed.y);
	const v3 r = min_image_fixed(box, P.ip, Q.ip);
	PairAcc acc;
	acc.clear();
	const v3 rs = r + (Q.ax.a1 - P.ax.a1) * M.stack_a1;
	const float en = MD::cxst(M, rs, dot(rs, rs), r + Q.back - P.back, P.ax, Q.ax, acc);
	if(en != 0.f) {
		const v3 tp = acc.torque_p(P.ax, P.back), tq = acc.torque_q(Q.ax, Q.back);
		atomic_add4(F + ed.x, -acc.F.x, -acc.F.y, -acc.F.z, en);
		atomic_add4(T + ed.x, tp.x, tp.y, tp.z, 0.f);
		atomic_add4(F + ed.y, acc.F.x, acc.F.y, acc.F.z, en);
		atomic_add4(T + ed.y, tq.x, tq.y, tq.z, 0.f);
	}
}

// one pair of a block's own hydrogen-bonding / cross-stacking segment, evaluated in the tail of the producing kernel (fold & 2: small
// systems, where the consumer launch behind the producer costs more in latency than the separate grid gains in balance)
template<class MD>
__device__ __noinline__ void hbcr_item(const typename MD::Params &M, const BoxF &box, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, int2 ed,
		float4 *__restrict__ F, float4 *__restrict__ T) {
	Particle P = load_particle<MD>(M, ipos, axf, ed.x);
	Particle Q = load_particle<MD>(M, ipos, axf, ed.y);
	const v3 r = min_image_fixed(box, P.ip, Q.ip);
	PairAcc acc;
	acc.clear();
	float ehb = 0.f;
	const v3 rb = r + (Q.ax.a1 - P.ax.a1) * M.base_a1;
	const float rbm2 = dot(rb, rb);
	const float en = MD::template hbcr<true>(M, rb, rbm2, P.ax, Q.ax, P.btype, Q.btype, MD::hb_in_range(M, rbm2, P.btype, Q.btype), MD::crst_in_range(M, rbm2), acc, ehb);
	if(en != 0.f) {
		const v3 tp = acc.torque_p(P.ax, P.back), tq = acc.torque_q(Q.ax, Q.back);
		atomic_add4(F + ed.x, -acc.F.x, -acc.F.y, -acc.F.z, en);
		atomic_add4(T + ed.x, tp.x, tp.y, tp.z, ehb);
		atomic_add4(F + ed.y, acc.F.x, acc.F.y, acc.F.z, en);
		atomic_add4(T + ed.y, tq.x, tq.y, tq.z, ehb);
	}
}

// ------------------------------------------------------------------------------------------------------------
// Particle-centric: one thread per particle, every listed pair evaluated from both ends, no atomics, deterministic.
// ------------------------------------------------------------------------------------------------------------
template<class MD>
__global__ void __launch_bounds__(128, OXB_MB_PARTICLE) k_forces_particle(const __grid_constant__ typename MD::Params M, BoxF box, int N, const int4 *__restrict__ ipos,
		const int4 *__restrict__ iback, const float4 *__restrict__ axf, const double4 *__restrict__ posd, const double4 *__restrict__ quatd,
		const int2 *__restrict__ bonds, const int *__restrict__ nbr, const int *__restrict__ nnbr, int stride,
		float4 *__restrict__ F, float4 *__restrict__ T, const oxb_replica_consts *__restrict__ rep, int n_per, int *__restrict__ flags, int hw) {
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(flags, flags[hw] ? OXB_PROF_WAIT : OXB_PROF_FORCE);
	if(flags[hw]) return;
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;

	Particle P = load_particle<MD>(M, ipos, axf, i);
	int2 b = __ldg(bonds + i);
	bool p_end = (b.x < 0 || b.y < 0);

	v3 f = mk3(0.f, 0.f, 0.f), t = mk3(0.f, 0.f, 0.f);
	float e = 0.f, ehb = 0.f;
	bool broken = false;
	ExclRefine R = make_refine<MD>(M, box, posd, quatd);
	const bool refine = posd != nullptr; // backend_precision = mixed
	if(rep != nullptr) rep += i / n_per;   // replica batching: slots are replica-contiguous

	if(b.x >= 0) { // I am the 5' side of the bond (p), q = my n3
		Particle Q = load_particle<MD>(M, ipos, axf, b.x);
		PairAcc acc;
		acc.clear();
		R.sp = i; R.sq = b.x; acc.refine = refine ? &R : nullptr;
		acc.rep = rep;
		FeneSite fs;
		if(refine) fs = fene_from_sites(M, box, __ldg(iback + i), __ldg(iback + b.x), broken);
		e += MD::bonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, P.back, Q.back, acc, broken, nullptr, refine ? &fs : nullptr);
		f -= acc.F;
		t += acc.torque_p(P.ax, P.back);
	}
	if(b.y >= 0) { // my n5 neighbour is p, I am q
		Particle Q = load_particle<MD>(M, ipos, axf, b.y);
		PairAcc acc;
		acc.clear();
		R.sp = b.y; R.sq = i; acc.refine = refine ? &R : nullptr;
		acc.rep = rep;
		FeneSite fs;
		if(refine) fs = fene_from_sites(M, box, __ldg(iback + b.y), __ldg(iback + i), broken);
		e += MD::bonded(M, min_image_fixed(box, Q.ip, P.ip), Q.ax, P.ax, Q.btype, P.btype, Q.back, P.back, acc, broken, nullptr, refine ? &fs : nullptr);
		f += acc.F;
		t += acc.torque_q(P.ax, P.back);
	}

	int nn = __ldg(nnbr + i);
	for(int k = 0; k < nn; k++) {
		int j = __ldg(nbr + (size_t) k * stride + i);
		Particle Q = load_particle<MD>(M, ipos, axf, j);
		int2 bq = __ldg(bonds + j);
		PairAcc acc;
		acc.clear();
		R.sp = i; R.sq = j; acc.refine = refine ? &R : nullptr;
		acc.rep = rep;
		PairEnergy pe = MD::nonbonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, p_end, (bq.x < 0 || bq.y < 0), P.back,
				Q.back, acc);
		e += pe.total;
		ehb += pe.hb;
		f -= acc.F;
		t += acc.torque_p(P.ax, P.back);
	}

	// torque stays in the lab frame; the integrator rotates it into the body frame
	F[i] = make_float4(f.x, f.y, f.z, e);
	T[i] = make_float4(t.x, t.y, t.z, ehb);
	if(broken) atomicOr(flags + OXB_FLAG_ERROR, OXB_ERR_FENE_BROKEN);
}

// The same pass with the neighbour loop split by COST into uniform sub-passes (as k_forces_dna3, forces_dna3.cu): a single loop runs with a
// third of the lanes active, because one lane inside the six-angle hydrogen-bonding code holds the 31 others.  Per chunk of 64 neighbours:
//   pass 1  every listed pair: fixed-point centre + backbone site of the neighbour (2 x 16 B), Debye-Hueckel, "near" bit
//   pass 2  near pairs: full particle record, the four excluded-volume site pairs, radial + cosine-window screening -> bits
//   pass 3  hydrogen bonding + cross stacking        pass 4  coaxial stacking
// The p-side sums are linear in the pairs: one PairAcc is carried through all passes, the torque's cross products are taken once.
// Still no atomics: deterministic.  OXB_PARTICLE_SPLIT=0 selects the single loop.
template<class MD, int MB>
__global__ void __launch_bounds__(128, MB) k_forces_particle_split(const __grid_constant__ typename MD::Params M, BoxF box, int N,
		const int4 *__restrict__ ipos, const int4 *__restrict__ iback, const float4 *__restrict__ axf, const double4 *__restrict__ posd,
		const double4 *__restrict__ quatd, const int2 *__restrict__ bonds, const int *__restrict__ nbr, const int *__restrict__ nnbr, int stride,
		float4 *__restrict__ F, float4 *__restrict__ T, const oxb_replica_consts *__restrict__ rep, int n_per, int *__restrict__ flags, int hw) {
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(flags, flags[hw] ? OXB_PROF_WAIT : OXB_PROF_FORCE);
	if(flags[hw]) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	const Particle P = load_particle<MD>(M, ipos, axf, i);
	const int4 ibp = __ldg(iback + i);
	const int2 b = __ldg(bonds + i);
	const bool p_end = (b.x < 0 || b.y < 0);
	float e = 0.f, ehb = 0.f;
	bool broken = false;
	ExclRefine R = make_refine<MD>(M, box, posd, quatd);
	const bool refine = posd != nullptr; // backend_precision = mixed
	if(rep != nullptr) rep += i / n_per;   // replica batching: slots are replica-contiguous
	const DhView D = dh_view(M, rep);
	const float rcut2 = rep ? rep->rcut2 : M.rcut * M.rcut;
	PairAcc acc; // everything in which this particle is "p"
	acc.clear();
	acc.refine = refine ? &R : nullptr;
	acc.rep = rep;
	v3 fq = mk3(0.f, 0.f, 0.f), tq = mk3(0.f, 0.f, 0.f); // the bond in which it is "q"
	if(b.x >= 0) { // I am the 5' side of the bond (p), q = my n3
		const Particle Q = load_particle<MD>(M, ipos, axf, b.x);
		R.sp = i; R.sq = b.x;
		FeneSite fs;
		if(refine) fs = fene_from_sites(M, box, ibp, __ldg(iback + b.x), broken);
		e += MD::bonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, P.back, Q.back, acc, broken, nullptr, refine ? &fs : nullptr);
	}
	if(b.y >= 0) { // my n5 neighbour is p, I am q
		const Particle Q = load_particle<MD>(M, ipos, axf, b.y);
		PairAcc a2;
		a2.clear();
		R.sp = b.y; R.sq = i; a2.refine = refine ? &R : nullptr;
		a2.rep = rep;
		FeneSite fs;
		if(refine) fs = fene_from_sites(M, box, __ldg(iback + b.y), ibp, broken);
		e += MD::bonded(M, min_image_fixed(box, Q.ip, P.ip), Q.ax, P.ax, Q.btype, P.btype, Q.back, P.back, a2, broken, nullptr, refine ? &fs : nullptr);
		fq = a2.F;
		tq = a2.torque_q(P.ax, P.back);
	}
	R.sp = i;
	const float rnear2 = M.rcut_near * M.rcut_near;
	const int nn = __ldg(nnbr + i);
	for(int base = 0; base < nn; base += 64) {
		const int cnt = min(64, nn - base);
		const int *__restrict__ row = nbr + (size_t) base * stride + i;
		unsigned long long near = 0ull, mb = 0ull, ms = 0ull;
#pragma unroll 4
		for(int k = 0; k < cnt; k++) {
			const int j = __ldg(row + (size_t) k * stride);
			const int4 ipq = __ldg(ipos + j);
			const int4 ibq = __ldg(iback + j);
			const v3 r = min_image_fixed(box, P.ip, ipq);
			const float r2 = dot(r, r);
			const v3 rbb = min_image_fixed(box, ibp, ibq);
			float fs;
			float en = dna2_dh_fast(D, dot(rbb, rbb), p_end, ibq.w & 1, fs);
			if(r2 >= rcut2) { en = 0.f; fs = 0.f; } // no interaction beyond the centre-centre cutoff (DNA2Interaction.cpp:46-48)
			e += en;
			acc.site_kk(rbb * fs);
			if(r2 < rnear2 && r2 < rcut2) near |= 1ull << k;
		}
		while(near) {
			const int k = __ffsll((long long) near) - 1;
			near &= near - 1ull;
			const int j = __ldg(row + (size_t) k * stride);
			const Particle Q = load_particle<MD>(M, ipos, axf, j);
			const v3 r = min_image_fixed(box, P.ip, Q.ip);
			const v3 da = Q.ax.a1 - P.ax.a1;
			const v3 rb = r + da * M.base_a1, rs = r + da * M.stack_a1;
			R.sq = j;
			e += dna2_excl(M, r, r + Q.back - P.back, rb, P.ax, Q.ax, P.back, Q.back, acc);
			const float rbm2 = dot(rb, rb), rs2 = dot(rs, rs);
			const bool hb_on = MD::hb_in_range(M, rbm2, P.btype, Q.btype), cr_on = MD::crst_in_range(M, rbm2);
			if((hb_on || cr_on) && MD::hbcr_may_act(M, rb * rsqrtf(rbm2), P.ax, Q.ax, hb_on, cr_on)) mb |= 1ull << k;
			if(MD::cxst_in_range(M, rs2) && MD::cxst_may_act(M, rs * rsqrtf(rs2), P.ax, Q.ax)) ms |= 1ull << k;
		}
		while(mb) {
			const int k = __ffsll((long long) mb) - 1;
			mb &= mb - 1ull;
			const int j = __ldg(row + (size_t) k * stride);
			const Particle Q = load_particle<MD>(M, ipos, axf, j);
			const v3 rb = min_image_fixed(box, P.ip, Q.ip) + (Q.ax.a1 - P.ax.a1) * M.base_a1;
			const float rbm2 = dot(rb, rb);
			float eh;
			e += MD::template hbcr<true>(M, rb, rbm2, P.ax, Q.ax, P.btype, Q.btype, MD::hb_in_range(M, rbm2, P.btype, Q.btype), MD::crst_in_range(M, rbm2), acc, eh);
			ehb += eh;
		}
		while(ms) {
			const int k = __ffsll((long long) ms) - 1;
			ms &= ms - 1ull;
			const int j = __ldg(row + (size_t) k * stride);
			const Particle Q = load_particle<MD>(M, ipos, axf, j);
			const v3 r = min_image_fixed(box, P.ip, Q.ip);
			const v3 rs = r + (Q.ax.a1 - P.ax.a1) * M.stack_a1;
			e += MD::cxst(M, rs, dot(rs, rs), r + Q.back - P.back, P.ax, Q.ax, acc);
		}
	}
	// torque stays in the lab frame; the integrator rotates it into the body frame
	const v3 f = fq - acc.F, t = tq + acc.torque_p(P.ax, P.back);
	F[i] = make_float4(f.x, f.y, f.z, e);
	T[i] = make_float4(t.x, t.y, t.z, ehb);
	if(broken) atomicOr(flags + OXB_FLAG_ERROR, OXB_ERR_FENE_BROKEN);
}

// ------------------------------------------------------------------------------------------------------------
// Edge-centric, staged.  One thread per unique pair, but the pair population is split by COST so that every kernel is
// (nearly) uniform across a warp -- the single-kernel version ran with 8 of 32 lanes active (ncu r01):
//   k_dh_particle Debye-Hueckel, one thread per particle over a backbone-site neighbour matrix (2 x 16 B per pair, no atomics)
//   k_edge_near   edges that can come within rcut_near before the next rebuild: 4 excluded-volume site pairs +
//                 detection of pairs inside the hydrogen-bonding / cross-stacking / coaxial-stacking radial ranges AND
//                 inside the cosine window of every angular factor; those are appended (warp-aggregated) to two
//                 compact work lists
//                 (excluded-volume terms that are in range are parked for k_excl_fix when the refinement is on)
//   k_edge_heavy<0>  dense list of base-base contacts: hydrogen bonding + cross stacking
//   k_edge_heavy<1>  dense list of stack-stack contacts: coaxial stacking
//   k_bonded      per particle: FENE + bonded excluded volume + stacking with its n3 neighbour (each bond once)
//   k_excl_fix    the parked excluded-volume pairs of k_edge_near and k_bonded in double (backend_precision = mixed)
// (the integrator folds the backbone-site force sum Fb of k_dh_particle into force and torque)
// Edges are grouped by `from`, so lanes sharing `from` first reduce with shuffles (segmented suffix sum) and only the
// segment head issues the 128-bit vector atomic (red.global.add.v4.f32).  F/T/Fb hold lab-frame sums; the integrator
// rotates the torque into the body frame.
// ------------------------------------------------------------------------------------------------------------

template<int NV>
__device__ __forceinline__ bool segmented_reduce(int key, unsigned lane, float (&v)[NV]) {
#pragma unroll
	for(int d = 1; d < 32; d <<= 1) {
		int okey = __shfl_down_sync(0xffffffffu, key, d);
		bool take = (lane + d < 32) && (okey == key);
#pragma unroll
		for(int k = 0; k < NV; k++) {
			float o = __shfl_down_sync(0xffffffffu, v[k], d);
			if(take) v[k] += o;
		}
	}
	int pkey = __shfl_up_sync(0xffffffffu, key, 1);
	return (lane == 0) || (pkey != key);
}

// Debye-Hueckel, particle-centric over its own neighbour matrix (selected on the backbone-site distance): per neighbour one
// coalesced index load + one 16-byte gather of a fixed-point backbone site; no atomics, deterministic.  Writes the
// backbone-site force sum Fb (.w = energy); the integrator folds it into force and torque.
// LPP lanes share one particle (lane s takes neighbours s, s + LPP, ...; partial sums folded with shuffles in a fixed order, so the
// result stays deterministic): systems that cannot fill the GPU with one thread per particle (C2: 17 warps per SM) get LPP times the
// loads in flight; index loads stay sector-efficient (LPP rows x 32 / LPP consecutive ints per warp instruction).
// REP (replica batching): the constants come from the row of the particle's replica, as register values; the single-system
// instantiation keeps them as constant-bank operands.
// HALF: the matrix lists every pair in ONE of its two rows; the thread evaluates it once, keeps its own share in registers and sends
// the partner's with one 128-bit red.global.add (half the evaluations and half the index traffic of the full matrix; Fb is then an
// accumulator that the integrator zeroes, and the sum is no longer order-deterministic -- like the rest of the edge pipeline).
template<class MD, int LPP, bool REP, bool HALF, bool SORTED = false>
__global__ void __launch_bounds__(128) k_dh_particle(const __grid_constant__ typename MD::Params M, BoxF box, int N, const int4 *__restrict__ iback,
		const int *__restrict__ dh_nbr, const int *__restrict__ dh_nnbr, float4 *__restrict__ Fb, const oxb_replica_consts *__restrict__ rep, int n_per,
		int *__restrict__ flags, int hw) {
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(flags, flags[hw] ? OXB_PROF_WAIT : OXB_PROF_FORCE);
	if(flags[hw]) return;
	const int gid = blockIdx.x * blockDim.x + threadIdx.x;
	const int sub = gid % LPP;
	int i = gid / LPP;
	const bool active = i < N;
	if(!active) i = N - 1; // whole groups stay in the shuffles
	DhView D;
	if(REP) D = dh_view(M, rep + i / n_per);
	const int4 bp = __ldg(iback + i);
	const int nn = active ? __ldg(dh_nnbr + i) : 0;
	const bool p_end = bp.w & 1;
	v3 f = mk3(0.f, 0.f, 0.f);
	float e = 0.f;
	if(LPP == 1) {
		// Row lengths differ a lot inside a warp -- with the half-shell lists a particle keeps the partners that come AFTER it on the Hilbert
		// curve, 0 to twice the mean -- and a warp lasts as long as its longest row (ncu r02m: 14.7 of 32 lanes active).  Lanes l and 31 - l
		// therefore share their two rows evenly: the one with the shorter row takes the tail of the other's and hands the partial force
		// back with one shuffle at the end.  The schedule is fixed by the row lengths, so the full-matrix variant stays deterministic.
		const unsigned lane = threadIdx.x & 31;
		int pl = 31 - (int) lane;
		if(SORTED) {
			// ... and which two lanes share is decided by the row lengths: the lane of rank r (ties by lane) pairs with the lane of rank 31 - r,
			// the longest row with the shortest.  With the fixed pairing l <-> 31 - l two long rows meet as often as not (ncu r02c: 17.9 lanes).
			int rank = 0;
#pragma unroll
			for(int j = 0; j < 32; j++) {
				const int o = __shfl_sync(0xffffffffu, nn, j);
				rank += (o < nn || (o == nn && j < (int) lane)) ? 1 : 0;
			}
			__shared__ unsigned char s_lane_of_rank[4][32]; // (the kernel is launched with 128 threads)
			s_lane_of_rank[threadIdx.x >> 5][rank] = (unsigned char) lane;
			__syncwarp();
			pl = s_lane_of_rank[threadIdx.x >> 5][31 - rank];
		}
		const int nn_o = __shfl_sync(0xffffffffu, nn, pl), i_o = __shfl_sync(0xffffffffu, i, pl);
		int4 bo;
		bo.x = __shfl_sync(0xffffffffu, bp.x, pl); bo.y = __shfl_sync(0xffffffffu, bp.y, pl); bo.z = __shfl_sync(0xffffffffu, bp.z, pl); bo.w = __shfl_sync(0xffffffffu, bp.w, pl);
		DhView Do;
		if(REP) Do = dh_view(M, rep + i_o / n_per);
		const int half = (nn + nn_o + 1) >> 1;
		const int n_own = (nn > nn_o) ? nn - (half - nn_o) : nn;  // the longer row gives its tail away ...
		const int n_help = (nn < nn_o) ? half - nn : 0;           // ... to the lane with the shorter one
		const int k_help0 = nn_o - n_help;
		v3 g = mk3(0.f, 0.f, 0.f);
		float eg = 0.f;
		const int n_all = n_own + n_help;
#pragma unroll 4
		for(int t = 0; t < n_all; t++) {
			const bool mine = t < n_own;
			const int j = __ldg(dh_nbr + (size_t) (mine ? t : k_help0 + (t - n_own)) * N + (mine ? i : i_o));
			const int4 b0 = mine ? bp : bo;
			const int4 bq = __ldg(iback + j);
			const v3 rbb = min_image_fixed(box, b0, bq);
			float fs;
			const float en = REP ? dna2_dh_fast(mine ? D : Do, dot(rbb, rbb), b0.w & 1, bq.w & 1, fs) : dna2_dh_fast(M, dot(rbb, rbb), b0.w & 1, bq.w & 1, fs);
			if(mine) { e += en; axpy(f, -fs, rbb); }
			else { eg += en; axpy(g, -fs, rbb); }
			if(HALF && en != 0.f) atomic_add4(Fb + j, fs * rbb.x, fs * rbb.y, fs * rbb.z, en);
		}
		f.x += __shfl_sync(0xffffffffu, g.x, pl); f.y += __shfl_sync(0xffffffffu, g.y, pl); f.z += __shfl_sync(0xffffffffu, g.z, pl);
		e += __shfl_sync(0xffffffffu, eg, pl);
	}
	else {
#pragma unroll 4
		for(int k = sub; k < nn; k += LPP) {
			int j = __ldg(dh_nbr + (size_t) k * N + i);
			int4 bq = __ldg(iback + j);
			v3 rbb = min_image_fixed(box, bp, bq);
			float fs;
			float en = REP ? dna2_dh_fast(D, dot(rbb, rbb), p_end, bq.w & 1, fs) : dna2_dh_fast(M, dot(rbb, rbb), p_end, bq.w & 1, fs);
			e += en;
			axpy(f, -fs, rbb);
			if(HALF && en != 0.f) atomic_add4(Fb + j, fs * rbb.x, fs * rbb.y, fs * rbb.z, en);
		}
#pragma unroll
		for(int o = LPP >> 1; o > 0; o >>= 1) {
			f.x += __shfl_xor_sync(0xffffffffu, f.x, o); f.y += __shfl_xor_sync(0xffffffffu, f.y, o);
			f.z += __shfl_xor_sync(0xffffffffu, f.z, o); e += __shfl_xor_sync(0xffffffffu, e, o);
		}
	}
	if(HALF) { if(active && sub == 0 && e != 0.f) atomic_add4(Fb + i, f.x, f.y, f.z, e); }
	else if(active && sub == 0) Fb[i] = make_float4(f.x, f.y, f.z, e);
}

// Work lists are SEGMENTED by producer block: block b of k_edge_near owns list[b * seg .. (b + 1) * seg) and publishes its
// length in seg_counts; block b of the consuming kernel walks exactly that segment.  A warp-aggregated append therefore
// only touches a shared-memory counter.  (A single global counter serialises ~1e5 same-address atomics per launch at 1M
// particles: measured ~50 us of the 81 us this kernel took, ncu r01e/r01g.)
__device__ __forceinline__ void block_append(bool flag, int2 item, int2 *__restrict__ seg_base, int *s_counter, int seg, int *flags) {
	unsigned mask = __ballot_sync(0xffffffffu, flag);
	if(mask == 0u) return;
	unsigned lane = threadIdx.x & 31;
	int leader = __ffs(mask) - 1;
	int base = 0;
	if((int) lane == leader) base = atomicAdd(s_counter, __popc(mask));
	base = __shfl_sync(0xffffffffu, base, leader);
	if(flag) {
		int pos = base + __popc(mask & ((1u << lane) - 1u));
		if(pos < seg) seg_base[pos] = item;
		else atomicOr(flags + OXB_FLAG_ERROR, OXB_ERR_SEG_OVERFLOW);
	}
}

// TILE (OXB_NEAR_TILE=1, an experiment kept switchable): block b owns the 128 consecutive `from` slots [128 b, 128 b + 128) and their edges
// (edge_offsets), stages those particles -- position, the three axes, the backbone site -- in shared memory once and takes both ends of an
// edge from there whenever they fall into the tile: the "shared-memory staging of neighbour data" variant of this kernel.
template<class MD, bool TILE>
__global__ void __launch_bounds__(128, OXB_MB_NEAR) k_edge_near(const __grid_constant__ typename MD::Params M, BoxF box, const int *__restrict__ n_edges,
		const int4 *__restrict__ edge_cnt, int N,
		const int2 *__restrict__ edges, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, float4 *__restrict__ F, float4 *__restrict__ T,
		int2 *__restrict__ hb_list, int2 *__restrict__ cx_list, int2 *__restrict__ cr_list, int *__restrict__ seg_counts, int hb_seg, int cx_seg,
		int cr_seg, int4 *__restrict__ ex_list, int *__restrict__ ex_counts, int ex_seg, int refine, int fold, const double4 *__restrict__ posd,
		const double4 *__restrict__ quatd, int *__restrict__ flags, int hw) {
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(flags, flags[hw] ? OXB_PROF_WAIT : OXB_PROF_FORCE);
	if(flags[hw]) return;
	__shared__ int s_cnt[3];
	if(threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
	__syncthreads();
	hb_list += (size_t) blockIdx.x * hb_seg;
	cx_list += (size_t) blockIdx.x * cx_seg;
	cr_list += (size_t) blockIdx.x * cr_seg;
	const int ne = *n_edges;
	const unsigned lane = threadIdx.x & 31;
	// pairs with an excluded-volume site pair in range are parked in this block's segment of ex_list; k_excl_fix evaluates them in double
	__shared__ int s_nex;
	if(threadIdx.x == 0) s_nex = 0;
	__syncthreads();
	ex_list += (size_t) blockIdx.x * ex_seg;
	// TILE: the staged particles, 64 bytes each: (a1, back.x), (a2, back.y), (a3, back.z) and the fixed-point position
	constexpr int TP = 128;
	__shared__ float4 s_ax[TILE ? 3 * TP : 1];
	__shared__ int4 s_ip[TILE ? TP : 1];
	const int p0 = blockIdx.x * TP;
	int e_begin = (blockIdx.x * blockDim.x + threadIdx.x) - lane, e_end = ne, e_step = gridDim.x * blockDim.x;
	if(TILE) {
		for(int t = threadIdx.x; t < TP && p0 + t < N; t += blockDim.x) {
			const Particle S = load_particle<MD>(M, ipos, axf, p0 + t);
			s_ip[t] = S.ip;
			s_ax[3 * t] = make_float4(S.ax.a1.x, S.ax.a1.y, S.ax.a1.z, S.back.x);
			s_ax[3 * t + 1] = make_float4(S.ax.a2.x, S.ax.a2.y, S.ax.a2.z, S.back.y);
			s_ax[3 * t + 2] = make_float4(S.ax.a3.x, S.ax.a3.y, S.ax.a3.z, S.back.z);
		}
		__syncthreads();
		const int o0 = __ldg(edge_cnt + min(p0, N)).x, o1 = __ldg(edge_cnt + min(p0 + TP, N)).x;
		e_begin = o0 + (int) threadIdx.x - (int) lane; e_end = min(o1, ne); e_step = blockDim.x;
	}
	auto staged = [&](int slot) {
		Particle S;
		const int t = slot - p0;
		const float4 u = s_ax[3 * t], v = s_ax[3 * t + 1], w = s_ax[3 * t + 2];
		S.ip = s_ip[t];
		S.ax.a1 = mk3(u.x, u.y, u.z); S.ax.a2 = mk3(v.x, v.y, v.z); S.ax.a3 = mk3(w.x, w.y, w.z);
		S.back = mk3(u.w, v.w, w.w);
		S.btype = word_btype(S.ip.w);
		return S;
	};
	for(int base = e_begin; base < e_end; base += e_step) {
		int eidx = base + lane;
		bool valid = eidx < e_end;
		int2 ed = valid ? __ldg(edges + eidx) : make_int2(-1 - (int) lane, -1);
		// the list builder's verdict on which families of site pairs can come into range before the next rebuild (common.cuh, OXB_CLS_*)
		const int cls = valid ? (int) ((unsigned) ed.y >> OXB_CLS_SHIFT) : 0;
		if(valid) ed.y &= OXB_SLOT_MASK;
		float v[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
		float ve = 0.f;
		bool want_hb = false, want_cx = false, hb_capable = false;
		if(valid) {
			Particle P, Q;
			if(TILE) {
				P = staged(ed.x);
				if((unsigned) (ed.y - p0) < (unsigned) TP) Q = staged(ed.y);
				else Q = load_particle<MD>(M, ipos, axf, ed.y);
			}
			else {
				P = load_particle<MD>(M, ipos, axf, ed.x);
				Q = load_particle<MD>(M, ipos, axf, ed.y);
			}
			v3 r = min_image_fixed(box, P.ip, Q.ip);
			if(dot(r, r) < M.rcut_near * M.rcut_near) {
				v3 rbb = r + Q.back - P.back;
				v3 rb = r + (Q.ax.a1 - P.ax.a1) * M.base_a1;
				PairAcc acc;
				acc.clear();
				// common path unchanged: the FP32 evaluation tells whether any site pair is in range at all.  If one is (a few per
				// cent of the edges) and the refinement is on, the pair is parked for k_excl_fix and the FP32 result is dropped
				float en = (cls & (OXB_CLS_BB | OXB_CLS_EB | OXB_CLS_BK)) ? dna2_excl_cls(M, cls, r, rbb, rb, P.ax, Q.ax, P.back, Q.back, acc) : 0.f;
				if(en != 0.f && refine) {
					const int slot = atomicAdd(&s_nex, 1);
					if(slot < ex_seg) {
						ex_list[slot] = make_int4(ed.x, ed.y, dna2_excl_mask(M, r, rbb, rb, P.ax, Q.ax, P.back, Q.back), 0);
						en = 0.f;
					}
				}
				if(en != 0.f) {
					v3 tq = acc.torque_q(Q.ax, Q.back), tp = acc.torque_p(P.ax, P.back);
					atomic_add4(F + ed.y, acc.F.x, acc.F.y, acc.F.z, en);
					atomic_add4(T + ed.y, tq.x, tq.y, tq.z, 0.f);
					v[0] = -acc.F.x; v[1] = -acc.F.y; v[2] = -acc.F.z;
					v[3] = tp.x; v[4] = tp.y; v[5] = tp.z;
					ve = en;
				}
				// radial range first, then the cosine windows of every angular factor: only pairs whose product can be
				// non-zero reach the heavy kernels
				float rbm2 = dot(rb, rb);
				const bool base_on = (cls & OXB_CLS_HBCR) != 0;
				bool hb_on = base_on && MD::hb_in_range(M, rbm2, P.btype, Q.btype), cr_on = base_on && MD::crst_in_range(M, rbm2);
				if(hb_on || cr_on) {
					// pairs that can hydrogen-bond go to the full kernel, the (more numerous) cross-stacking-only pairs to a
					// specialised one: both lists are warp-uniform
					// (splitting this list into a hydrogen-bonding and a cross-stacking-only list with specialised kernels was
					// measured SLOWER, 53 + 39 us against 83 us at 1M particles, and is kept only as MODE 2 below)
					want_hb = MD::hbcr_may_act(M, rb * rsqrtf(rbm2), P.ax, Q.ax, hb_on, cr_on);
					hb_capable = hb_on;
				}
				if(cls & OXB_CLS_ST) {
					v3 rs = r + (Q.ax.a1 - P.ax.a1) * M.stack_a1;
					float rs2 = dot(rs, rs);
					if(MD::cxst_in_range(M, rs2)) want_cx = MD::cxst_may_act(M, rs * rsqrtf(rs2), P.ax, Q.ax);
				}
			}
		}
		// the dense list is kept sorted by kind at no cost: pairs that can hydrogen-bond fill the first third of the block's segment,
		// cross-stacking-only pairs (about twice as many) the rest, so that the warps of k_edge_heavy<0> are uniform in which of the
		// two potentials they evaluate (one mixed warp per block at most; ncu r01k: 23.6 of 32 lanes active with the unsorted list)
		block_append(want_hb && hb_capable, ed, hb_list, &s_cnt[0], hb_seg / 3, flags);
		block_append(want_hb && !hb_capable, ed, hb_list + hb_seg / 3, &s_cnt[2], hb_seg - hb_seg / 3, flags);
		block_append(want_cx, ed, cx_list, &s_cnt[1], cx_seg, flags);
		// excluded volume between non-bonded nucleotides is rare: skip the shuffle reduction when the warp has none
		if(__any_sync(0xffffffffu, ve != 0.f)) {
			float w[7] = { v[0], v[1], v[2], v[3], v[4], v[5], ve };
			bool head = segmented_reduce<7>(ed.x, lane, w);
			if(valid && head && w[6] != 0.f) {
				atomic_add4(F + ed.x, w[0], w[1], w[2], w[6]);
				atomic_add4(T + ed.x, w[3], w[4], w[5], 0.f);
			}
		}
	}
	__syncthreads();
	if(fold && threadIdx.x < 3) {
		const int seg = (threadIdx.x == 0) ? hb_seg / 3 : (threadIdx.x == 1 ? cx_seg : hb_seg - hb_seg / 3);
		seg_counts[threadIdx.x * gridDim.x + blockIdx.x] = min(s_cnt[threadIdx.x], seg);
	}
	if(fold) {
		// tail: this block's own coaxial-stacking pairs and parked excluded-volume pairs (written above, visible after the barrier)
		const int ncx = min(s_cnt[1], cx_seg);
		for(int k = threadIdx.x; k < ncx; k += blockDim.x) cxst_item<MD>(M, box, ipos, axf, cx_list[k], F, T);
		const int nex = min(s_nex, ex_seg);
		for(int k = threadIdx.x; k < nex; k += blockDim.x) {
			const int4 it = ex_list[k];
			excl_double_item<MD>(M, box, posd, quatd, it.x, it.y, it.z, F, T);
		}
		if(fold & 2) {
			// ... and its hydrogen-bonding / cross-stacking pairs: front third of the segment, then the cross-stacking-only pairs behind it
			const int n_front = min(s_cnt[0], hb_seg / 3), n_hb = n_front + min(s_cnt[2], hb_seg - hb_seg / 3);
			for(int k = threadIdx.x; k < n_hb; k += blockDim.x) hbcr_item<MD>(M, box, ipos, axf, hb_list[(k >= n_front) ? hb_seg / 3 + (k - n_front) : k], F, T);
		}
		return;
	}
	if(threadIdx.x == 0) ex_counts[blockIdx.x] = min(s_nex, ex_seg);
	if(threadIdx.x < 3) {
		// list 0: hydrogen-bonding-capable pairs (front of the hb segment), 1: coaxial stacking, 2: cross-stacking-only pairs (rest of the hb segment)
		const int seg = (threadIdx.x == 0) ? hb_seg / 3 : (threadIdx.x == 1 ? cx_seg : hb_seg - hb_seg / 3);
		seg_counts[threadIdx.x * gridDim.x + blockIdx.x] = min(s_cnt[threadIdx.x], seg);
	}
}

// MODE 0: hydrogen bonding (+ cross stacking where also in range) | 1: coaxial stacking | 2: cross stacking only
template<class MD, int MODE>
__global__ void __launch_bounds__(64, OXB_MB_HEAVY) k_edge_heavy(const __grid_constant__ typename MD::Params M, BoxF box, const int *__restrict__ seg_counts,
		const int2 *__restrict__ list, int seg, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, float4 *__restrict__ F,
		float4 *__restrict__ T, const int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	// list index in seg_counts: 0 hydrogen bonding, 1 coaxial stacking, 2 cross stacking only (same order as MODE)
	// gridDim.y consumer blocks share one segment (long segments at large N would otherwise serialise on 64 threads)
	// MODE 0 walks both halves of its segment: [0, n_front) and [seg / 3, seg / 3 + n_back)
	const int n_front = seg_counts[MODE * gridDim.x + blockIdx.x];
	const int n = n_front + (MODE == 0 ? seg_counts[2 * gridDim.x + blockIdx.x] : 0);
	list += (size_t) blockIdx.x * seg;
	for(int k = blockIdx.y * blockDim.x + threadIdx.x; k < n; k += blockDim.x * gridDim.y) {
		int2 ed = __ldg(list + ((MODE == 0 && k >= n_front) ? seg / 3 + (k - n_front) : k));
		Particle P = load_particle<MD>(M, ipos, axf, ed.x);
		Particle Q = load_particle<MD>(M, ipos, axf, ed.y);
		v3 r = min_image_fixed(box, P.ip, Q.ip);
		PairAcc acc;
		acc.clear();
		float en, ehb = 0.f;
		if(MODE == 1) {
			v3 rs = r + (Q.ax.a1 - P.ax.a1) * M.stack_a1;
			en = MD::cxst(M, rs, dot(rs, rs), r + Q.back - P.back, P.ax, Q.ax, acc);
		}
		else {
			v3 rb = r + (Q.ax.a1 - P.ax.a1) * M.base_a1;
			float rbm2 = dot(rb, rb);
			if(MODE == 0) en = MD::template hbcr<true>(M, rb, rbm2, P.ax, Q.ax, P.btype, Q.btype, MD::hb_in_range(M, rbm2, P.btype, Q.btype), MD::crst_in_range(M, rbm2), acc, ehb);
			else en = MD::template hbcr<false>(M, rb, rbm2, P.ax, Q.ax, P.btype, Q.btype, false, true, acc, ehb);
		}
		if(en != 0.f) {
			v3 tp = acc.torque_p(P.ax, P.back), tq = acc.torque_q(Q.ax, Q.back);
			atomic_add4(F + ed.x, -acc.F.x, -acc.F.y, -acc.F.z, en);
			atomic_add4(T + ed.x, tp.x, tp.y, tp.z, ehb);
			atomic_add4(F + ed.y, acc.F.x, acc.F.y, acc.F.z, en);
			atomic_add4(T + ed.y, tq.x, tq.y, tq.z, ehb);
		}
	}
}

// per particle: bonded interaction with its n3 neighbour (each bond evaluated once).  Independent of the other kernels of
// the force pass (it only adds into F/T), so it runs concurrently with them on its own stream.
template<class MD>
__global__ void __launch_bounds__(128, OXB_MB_BONDED) k_bonded(const __grid_constant__ typename MD::Params M, BoxF box, int N, const int4 *__restrict__ ipos,
		const int4 *__restrict__ iback, const float4 *__restrict__ axf, const int2 *__restrict__ bonds, float4 *__restrict__ F, float4 *__restrict__ T,
		int *__restrict__ ex_bonded, int refine, int fold, const double4 *__restrict__ posd, const double4 *__restrict__ quatd,
		const oxb_replica_consts *__restrict__ rep, int n_per, int *__restrict__ flags, int hw) {
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(flags, flags[hw] ? OXB_PROF_WAIT : OXB_PROF_FORCE);
	if(flags[hw]) return;
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	int2 b = __ldg(bonds + i);
	if(b.x < 0) { if(refine && !fold) ex_bonded[i] = 0; return; }
	Particle P = load_particle<MD>(M, ipos, axf, i);
	Particle Q = load_particle<MD>(M, ipos, axf, b.x);
	PairAcc acc;
	acc.clear();
	if(rep != nullptr) acc.rep = rep + i / n_per; // replica batching: this replica's stacking strength
	bool broken = false;
	const v3 r = min_image_fixed(box, P.ip, Q.ip);
	float en;
	int ex_mask = 0;
	if(refine) {
		// d2V/dr2 of the FENE spring is eps / Delta^2 = 32 at rest and grows as (1 + u) / (1 - u)^2, u = x^2 / Delta^2: the FP32
		// distance (~1e-7) is good for 1e-5 of the force until u ~ 0.5; bonds stretched further take the distance in double
		FeneSite fs;
		fs.has_fene = 0;
		{
			const v3 dbb = r + Q.back - P.back;
			const float x = sqrtf(dot(dbb, dbb)) - M.fene_r0;
			if(x * x > 0.4f * M.fene_delta2) fs = fene_from_sites(M, box, __ldg(iback + i), __ldg(iback + b.x), broken);
		}
		// bonded excluded-volume site pairs in range are left to k_excl_fix (double); one mask per particle
		fs.excl_deferred = bonded_excl_mask(M, r, P.ax, Q.ax, P.back, Q.back);
		if(!fold) ex_bonded[i] = fs.excl_deferred;
		ex_mask = fs.excl_deferred;
		en = MD::bonded(M, r, P.ax, Q.ax, P.btype, Q.btype, P.back, Q.back, acc, broken, nullptr, &fs);
	}
	else en = MD::bonded(M, r, P.ax, Q.ax, P.btype, Q.btype, P.back, Q.back, acc, broken);
	v3 tp = acc.torque_p(P.ax, P.back), tq = acc.torque_q(Q.ax, Q.back);
	atomic_add4(F + i, -acc.F.x, -acc.F.y, -acc.F.z, en);
	atomic_add4(T + i, tp.x, tp.y, tp.z, 0.f);
	atomic_add4(F + b.x, acc.F.x, acc.F.y, acc.F.z, en);
	atomic_add4(T + b.x, tq.x, tq.y, tq.z, 0.f);
	if(broken) atomicOr(flags + OXB_FLAG_ERROR, OXB_ERR_FENE_BROKEN);
	// fold = 1: the few bonds with a bonded excluded-volume site pair in range take it in double right here instead of in k_excl_fix
	if(fold && ex_mask != 0) excl_double_item<MD>(M, box, posd, quatd, i, b.x, ex_mask, F, T);
}

// Excluded volume in double for the parked pairs: blocks [0, n_seg) walk the near-edge segments, the following blocks scan the
// per-particle masks of the bonds (256 particles each).  Runs after k_edge_near and k_bonded, next to the coaxial-stacking kernel.
template<class MD>
__global__ void __launch_bounds__(128) k_excl_fix(const __grid_constant__ typename MD::Params M, BoxF box, int N, int n_seg, const int4 *__restrict__ ex_list,
		const int *__restrict__ ex_counts, int ex_seg, const int *__restrict__ ex_bonded, const int2 *__restrict__ bonds, const double4 *__restrict__ posd,
		const double4 *__restrict__ quatd, float4 *__restrict__ F, float4 *__restrict__ T, const int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	if((int) blockIdx.x < n_seg) {
		const int n = ex_counts[blockIdx.x];
		const int4 *seg = ex_list + (size_t) blockIdx.x * ex_seg;
		for(int k = threadIdx.x; k < n; k += blockDim.x) {
			const int4 it = __ldg(seg + k);
			excl_double_item<MD>(M, box, posd, quatd, it.x, it.y, it.z, F, T);
		}
	}
	else {
		for(int i = ((int) blockIdx.x - n_seg) * 256 + threadIdx.x, e = min(N, ((int) blockIdx.x - n_seg + 1) * 256); i < e; i += blockDim.x) {
			const int mask = __ldg(ex_bonded + i);
			if(mask != 0) excl_double_item<MD>(M, box, posd, quatd, i, __ldg(bonds + i).x, mask, F, T);
		}
	}
}

// ------------------------------------------------------------------------------------------------------------
// oxDNA3 through the staged edge pipeline (use_edge = 1): the same lists (half Debye-Hueckel matrix, near edges, block-segmented work
// lists) and the same cost split as above, with the per-tetramer parameter records and per-type site offsets of dna3_model.cuh.
//   k_dh_particle<Dna3Dh>  Debye-Hueckel (needs the Debye-Hueckel scalars only)
//   k3_edge_near     near edges: the four excluded-volume site pairs (active ones re-evaluated in double through PairAcc::refine);
//                    pairs with the bases / the stacking sites inside the longest range of any tetramer go to the two work lists
//   k3_edge_heavy<0> hydrogen bonding + both cross-stacking diagonals      k3_edge_heavy<1> coaxial stacking
//   k3_bonded        FENE + bonded excluded volume + stacking, each bond once
// The near-edge classes of the list builder are not used (every family is screened here): they are thresholds of one parameter set.
// ------------------------------------------------------------------------------------------------------------
struct Dna3Dh { typedef oxb_dna3_dev Params; };

struct P3e {
	int4 ip;
	Axes ax;
	v3 back;
	int btype;
	Nuc3 n;
};
__device__ __forceinline__ P3e load_p3e(const oxb_dna3_dev &M, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, int i) {
	P3e P;
	P.ip = __ldg(ipos + i);
	P.ax = load_axes(axf, i);
	P.back = P.ax.a1 * M.back_a1 + P.ax.a2 * M.back_a2;
	P.btype = word_btype(P.ip.w);
	P.n = nuc3_from_code(__ldg(M.tcode + word_index(P.ip.w)));
	return P;
}
__device__ __forceinline__ ExclRefine refine3e(const oxb_dna3_dev &M, const BoxF &box, const double4 *posd, const double4 *quatd) {
	ExclRefine R;
	R.posd = posd; R.quatd = quatd; R.sp = R.sq = 0;
	R.L[0] = box.dsx * 4294967296.; R.L[1] = box.dsy * 4294967296.; R.L[2] = box.dsz * 4294967296.;
	R.b1 = (double) M.back_a1; R.b2 = (double) M.back_a2; R.b3 = 0.;
	return R;
}

template<int MB>
__global__ void __launch_bounds__(128, MB) k3_edge_near(const __grid_constant__ oxb_dna3_dev M, BoxF box, const int *__restrict__ n_edges,
		const int2 *__restrict__ edges, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, float4 *__restrict__ F, float4 *__restrict__ T,
		int2 *__restrict__ hb_list, int2 *__restrict__ cx_list, int *__restrict__ seg_counts, int hb_seg, int cx_seg, const double4 *__restrict__ posd,
		const double4 *__restrict__ quatd, int *__restrict__ flags, int hw) {
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(flags, flags[hw] ? OXB_PROF_WAIT : OXB_PROF_FORCE);
	if(flags[hw]) return;
	__shared__ int s_cnt[2];
	if(threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
	__syncthreads();
	hb_list += (size_t) blockIdx.x * hb_seg;
	cx_list += (size_t) blockIdx.x * cx_seg;
	const int ne = *n_edges;
	const unsigned lane = threadIdx.x & 31;
	ExclRefine R = refine3e(M, box, posd, quatd);
	for(int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < ne; base += gridDim.x * blockDim.x) {
		const int eidx = base + lane;
		const bool valid = eidx < ne;
		int2 ed = valid ? __ldg(edges + eidx) : make_int2(-1 - (int) lane, -1);
		if(valid) ed.y &= OXB_SLOT_MASK;
		float v[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
		float ve = 0.f;
		bool want_hb = false, want_cx = false;
		if(valid) {
			const P3e P = load_p3e(M, ipos, axf, ed.x), Q = load_p3e(M, ipos, axf, ed.y);
			const v3 r = min_image_fixed(box, P.ip, Q.ip);
			const float r2 = dot(r, r);
			if(r2 < M.r2_near_max && r2 < M.rcut2) {
				const v3 rb = r + Q.ax.a1 * M.pos_base[Q.n.si] - P.ax.a1 * M.pos_base[P.n.si];
				const v3 rs = r + Q.ax.a1 * M.pos_stack[Q.n.si] - P.ax.a1 * M.pos_stack[P.n.si];
				if(r2 < M.r2_excl_max) {
					PairAcc acc;
					acc.clear();
					R.sp = ed.x; R.sq = ed.y; acc.refine = posd != nullptr ? &R : nullptr;
					const float en = dna3_excl4(M, r, r + Q.back - P.back, rb, P.ax, Q.ax, P.n, Q.n, P.back, Q.back, acc);
					if(en != 0.f) {
						const v3 tq = acc.torque_q(Q.ax, Q.back), tp = acc.torque_p(P.ax, P.back);
						atomic_add4(F + ed.y, acc.F.x, acc.F.y, acc.F.z, en);
						atomic_add4(T + ed.y, tq.x, tq.y, tq.z, 0.f);
						v[0] = -acc.F.x; v[1] = -acc.F.y; v[2] = -acc.F.z;
						v[3] = tp.x; v[4] = tp.y; v[5] = tp.z;
						ve = en;
					}
				}
				want_hb = dot(rb, rb) < M.r2_base_max;
				want_cx = dot(rs, rs) < M.r2_stack_max;
			}
		}
		block_append(want_hb, ed, hb_list, &s_cnt[0], hb_seg, flags);
		block_append(want_cx, ed, cx_list, &s_cnt[1], cx_seg, flags);
		if(__any_sync(0xffffffffu, ve != 0.f)) {
			float w[7] = { v[0], v[1], v[2], v[3], v[4], v[5], ve };
			const bool head = segmented_reduce<7>(ed.x, lane, w);
			if(valid && head && w[6] != 0.f) {
				atomic_add4(F + ed.x, w[0], w[1], w[2], w[6]);
				atomic_add4(T + ed.x, w[3], w[4], w[5], 0.f);
			}
		}
	}
	__syncthreads();
	if(threadIdx.x < 3) seg_counts[threadIdx.x * gridDim.x + blockIdx.x] = threadIdx.x == 0 ? min(s_cnt[0], hb_seg) : (threadIdx.x == 1 ? min(s_cnt[1], cx_seg) : 0);
}

// MODE 0: hydrogen bonding + cross stacking | 1: coaxial stacking, on this producer block's segment of the list (gridDim.y blocks share it)
template<int MODE, int MB>
__global__ void __launch_bounds__(64, MB) k3_edge_heavy(const __grid_constant__ oxb_dna3_dev M, BoxF box, const int *__restrict__ seg_counts, const int2 *__restrict__ list,
		int seg, const int4 *__restrict__ ipos, const float4 *__restrict__ axf, float4 *__restrict__ F, float4 *__restrict__ T, int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	const int n = seg_counts[MODE * gridDim.x + blockIdx.x];
	list += (size_t) blockIdx.x * seg;
	for(int k = blockIdx.y * blockDim.x + threadIdx.x; k < n; k += blockDim.x * gridDim.y) {
		const int2 ed = __ldg(list + k);
		const P3e P = load_p3e(M, ipos, axf, ed.x), Q = load_p3e(M, ipos, axf, ed.y);
		const v3 r = min_image_fixed(box, P.ip, Q.ip);
		PairAcc acc;
		acc.clear();
		float en, ehb = 0.f;
		if(MODE == 1) {
			const v3 rs = r + Q.ax.a1 * M.pos_stack[Q.n.si] - P.ax.a1 * M.pos_stack[P.n.si];
			en = dna3_cxst(M, rs, dot(rs, rs), P.ax, Q.ax, P.n, Q.n, acc);
		}
		else {
			const v3 rb = r + Q.ax.a1 * M.pos_base[Q.n.si] - P.ax.a1 * M.pos_base[P.n.si];
			en = dna3_hbcr(M, rb, dot(rb, rb), P.ax, Q.ax, P.btype, Q.btype, P.n, Q.n, acc, ehb);
		}
		if(en != 0.f) {
			const v3 tp = acc.torque_p(P.ax, P.back), tq = acc.torque_q(Q.ax, Q.back);
			atomic_add4(F + ed.x, -acc.F.x, -acc.F.y, -acc.F.z, en);
			atomic_add4(T + ed.x, tp.x, tp.y, tp.z, ehb);
			atomic_add4(F + ed.y, acc.F.x, acc.F.y, acc.F.z, en);
			atomic_add4(T + ed.y, tq.x, tq.y, tq.z, ehb);
		}
	}
}

template<int MB>
__global__ void __launch_bounds__(128, MB) k3_bonded(const __grid_constant__ oxb_dna3_dev M, BoxF box, int N, const int4 *__restrict__ ipos,
		const int4 *__restrict__ iback, const float4 *__restrict__ axf, const int2 *__restrict__ bonds, float4 *__restrict__ F, float4 *__restrict__ T,
		const double4 *__restrict__ posd, const double4 *__restrict__ quatd, int *__restrict__ flags, int hw) {
	if(blockIdx.x == 0 && threadIdx.x == 0) prof_mark(flags, flags[hw] ? OXB_PROF_WAIT : OXB_PROF_FORCE);
	if(flags[hw]) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	const int2 b = __ldg(bonds + i);
	if(b.x < 0) return;
	const P3e P = load_p3e(M, ipos, axf, i), Q = load_p3e(M, ipos, axf, b.x);
	const float4 *rec = M.bonded + ix4(Q.n.n3t, Q.n.type, P.n.type, P.n.n5t) * (OXB3_REC_BONDED / 4);
	PairAcc acc;
	acc.clear();
	ExclRefine R = refine3e(M, box, posd, quatd);
	const bool refine = posd != nullptr;
	R.sp = i; R.sq = b.x; acc.refine = refine ? &R : nullptr;
	bool broken = false;
	FeneSite fs;
	if(refine) fs = fene_from_sites(fene3_of(M, rec), box, __ldg(iback + i), __ldg(iback + b.x), broken);
	const float en = dna3_bonded(M, rec, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.n, Q.n, P.back, Q.back, acc, broken, nullptr, refine ? &fs : nullptr);
	const v3 tp = acc.torque_p(P.ax, P.back), tq = acc.torque_q(Q.ax, Q.back);
	atomic_add4(F + i, -acc.F.x, -acc.F.y, -acc.F.z, en);
	atomic_add4(T + i, tp.x, tp.y, tp.z, 0.f);
	atomic_add4(F + b.x, acc.F.x, acc.F.y, acc.F.z, en);
	atomic_add4(T + b.x, tq.x, tq.y, tq.z, 0.f);
	if(broken) atomicOr(flags + OXB_FLAG_ERROR, OXB_ERR_FENE_BROKEN);
}

// Observable: potential energy split into the reference's eight terms (FENE, bonded excluded volume, stacking, non-bonded
// excluded volume, hydrogen bonding, cross stacking, coaxial stacking, Debye-Hueckel), summed on the device in double.
// Replaces the CPU get_system_energy_split() the reference runs after a D2H copy and a CPU list rebuild
// (src/Interactions/BaseInteraction.cpp:61-90).  One thread per particle over the full Verlet matrix,
// every unique pair once (from its lower slot); forces are not needed and are eliminated by the compiler.
// ------------------------------------------------------------------------------------------------------------
template<class MD>
__global__ void __launch_bounds__(128) k_energy_split(const __grid_constant__ typename MD::Params M, BoxF box, int N, const int4 *__restrict__ ipos,
		const float4 *__restrict__ axf, const int2 *__restrict__ bonds, const int *__restrict__ nbr, const int *__restrict__ nnbr, int stride,
		double *__restrict__ out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	float e[OXB_NTERMS];
#pragma unroll
	for(int t = 0; t < OXB_NTERMS; t++) e[t] = 0.f;
	if(i < N) {
		Particle P = load_particle<MD>(M, ipos, axf, i);
		int2 b = __ldg(bonds + i);
		bool p_end = (b.x < 0 || b.y < 0);
		PairAcc acc;
		acc.clear();
		if(b.x >= 0) {
			Particle Q = load_particle<MD>(M, ipos, axf, b.x);
			bool broken = false;
			MD::bonded(M, min_image_fixed(box, P.ip, Q.ip), P.ax, Q.ax, P.btype, Q.btype, P.back, Q.back, acc, broken, e);
		}
		int nn = __ldg(nnbr + i);
		for(int k = 0; k < nn; k++) {
			int j = __ldg(nbr + (size_t) k * stride + i) & OXB_SLOT_MASK; // (half-shell entries carry the near-edge class above the slot)
			if(j < i) continue;
			Particle Q = load_particle<MD>(M, ipos, axf, j);
			v3 r = min_image_fixed(box, P.ip, Q.ip);
			float r2 = dot(r, r);
			if(r2 >= M.rcut * M.rcut) continue;
			int2 bq = __ldg(bonds + j);
			v3 rbb = r + Q.back - P.back;
			float fs;
			e[OXB_TERM_DH] += dna2_dh(M, dot(rbb, rbb), p_end, (bq.x < 0 || bq.y < 0), fs);
			if(r2 >= M.rcut_near * M.rcut_near) continue;
			v3 rb = r + (Q.ax.a1 - P.ax.a1) * M.base_a1;
			e[OXB_TERM_NEXC] += dna2_excl(M, r, rbb, rb, P.ax, Q.ax, P.back, Q.back, acc);
			float rbm2 = dot(rb, rb);
			bool hb_on = MD::hb_in_range(M, rbm2, P.btype, Q.btype), cr_on = MD::crst_in_range(M, rbm2);
			if(hb_on || cr_on) {
				float ehb;
				float tot = MD::template hbcr<true>(M, rb, rbm2, P.ax, Q.ax, P.btype, Q.btype, hb_on, cr_on, acc, ehb);
				e[OXB_TERM_HB] += ehb;
				e[OXB_TERM_CRST] += tot - ehb;
			}
			v3 rs = r + (Q.ax.a1 - P.ax.a1) * M.stack_a1;
			float rs2 = dot(rs, rs);
			if(MD::cxst_in_range(M, rs2)) e[OXB_TERM_CXST] += MD::cxst(M, rs, rs2, rbb, P.ax, Q.ax, acc);
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
		double x = sh[threadIdx.x][0] + sh[threadIdx.x][1] + sh[threadIdx.x][2] + sh[threadIdx.x][3];
		if(x != 0.) atomicAdd(out + threadIdx.x, x);
	}
}

// ------------------------------------------------------------------------------------------------------------
// External forces: one thread per force entry (compact table, not 15 slots x N as in the reference).  Runs after the
// interaction kernels and adds into F.  Particle ids in the table are ORIGINAL ids, mapped through slot_of, so the
// Hilbert re-sort needs no table rewrite (the reference forbids sort + external forces, MD_CUDABackend.cu:110-112).
// ------------------------------------------------------------------------------------------------------------
// force of a single-particle entry on a particle at absolute position p (double, unwrapped: what the reference's CPU classes see)
__device__ __forceinline__ v3 ext_single(const DevExtForce &e, double4 p, int4 ip, BoxF box, long long step, const int *__restrict__ slot_of,
		const double4 *__restrict__ posd) {
	const float st = (float) step;
	switch(e.type) {
	case OXB_EXT_STRING: {
		// ConstantRateForce.cpp:52-61; dir_as_centre (flag in pbc): towards the point pos0 (CUDA_MD.cuh:114-130)
		float s = e.F0 + e.rate * st;
		if(e.pbc) {
			const double dx = e.pos0[0] - p.x, dy = e.pos0[1] - p.y, dz = e.pos0[2] - p.z;
			const double inv = (double) s / sqrt(dx * dx + dy * dy + dz * dz);
			return mk3((float) (dx * inv), (float) (dy * inv), (float) (dz * inv));
		}
		return mk3(e.dir[0] * s, e.dir[1] * s, e.dir[2] * s);
	}
	case OXB_EXT_TRAP:
	case OXB_EXT_LOWDIM_TRAP: {
		// MovingTrap.cpp:50-64, LowdimMovingTrap.cpp:68-82: absolute (unwrapped) position, kept in double
		double rs = (double) e.rate * (double) step;
		v3 f = mk3((float) (-(double) e.stiff * (p.x - (e.pos0[0] + rs * e.dir[0]))), (float) (-(double) e.stiff * (p.y - (e.pos0[1] + rs * e.dir[1]))),
				(float) (-(double) e.stiff * (p.z - (e.pos0[2] + rs * e.dir[2]))));
		if(e.type == OXB_EXT_LOWDIM_TRAP) {
			if(!(e.iaux & 1)) f.x = 0.f;
			if(!(e.iaux & 2)) f.y = 0.f;
			if(!(e.iaux & 4)) f.z = 0.f;
		}
		return f;
	}
	case OXB_EXT_REPULSION_PLANE: {
		// RepulsionPlane.cpp:44-58
		double pos = (double) e.aux[0] + (double) e.aux[1] * (double) step, end = e.aux[2], start = e.aux[0];
		if(end > start && pos > end) pos = end;
		if(end < start && pos < end) pos = end;
		double d = e.dir[0] * p.x + e.dir[1] * p.y + e.dir[2] * p.z + pos;
		if(d >= 0.) return mk3(0.f, 0.f, 0.f);
		float s = (float) (-d * (double) e.stiff);
		return mk3(e.dir[0] * s, e.dir[1] * s, e.dir[2] * s);
	}
	case OXB_EXT_ATTRACTION_PLANE: {
		// AttractionPlane.cpp:45-55
		double d = e.dir[0] * p.x + e.dir[1] * p.y + e.dir[2] * p.z + (double) e.aux[0];
		float s = (d >= 0.) ? -e.stiff : (float) (-d * (double) e.stiff);
		return mk3(e.dir[0] * s, e.dir[1] * s, e.dir[2] * s);
	}
	case OXB_EXT_SPHERE: {
		// RepulsiveSphere.cpp:46-53: minimum image of (pos - center)
		int4 ic;
		ic.x = (int) to_fixed(e.pos0[0], 1. / (double) box.lx); ic.y = (int) to_fixed(e.pos0[1], 1. / (double) box.ly); ic.z = (int) to_fixed(e.pos0[2], 1. / (double) box.lz);
		v3 d = min_image_fixed(box, ic, ip);
		float m = sqrtf(dot(d, d));
		float radius = e.r0 + e.rate * st;
		if(m <= radius || m >= e.aux[0]) return mk3(0.f, 0.f, 0.f);
		return d * (-e.stiff * (1.f - radius / m));
	}
	case OXB_EXT_LJ_WALL: {
		// LJWall.cpp:62-68
		double d = e.dir[0] * p.x + e.dir[1] * p.y + e.dir[2] * p.z + (double) e.aux[0];
		float rel = (float) d / e.aux[1];
		if(rel > e.aux[2]) return mk3(0.f, 0.f, 0.f);
		float lj = powf(rel, -(float) e.iaux);
		float s = 4.f * (float) e.iaux * e.stiff * (2.f * lj * lj - lj) / (float) d;
		return mk3(e.dir[0] * s, e.dir[1] * s, e.dir[2] * s);
	}
	case OXB_EXT_TWIST: {
		// ConstantRateTorque.cpp:76-103: trap at pos0 rotated by (base + rate t) about the axis through `center`, masked
		double t = (double) e.F0 + (double) e.rate * (double) step;
		double sn, cs;
		sincos(t, &sn, &cs);
		double oc = 1. - cs, ax = e.dir[0], ay = e.dir[1], az = e.dir[2];
		double vx = e.pos0[0] - (double) e.aux[0], vy = e.pos0[1] - (double) e.aux[1], vz = e.pos0[2] - (double) e.aux[2];
		double tx = (ax * ax * oc + cs) * vx + (ax * ay * oc - az * sn) * vy + (ax * az * oc + ay * sn) * vz + (double) e.aux[0];
		double ty = (ax * ay * oc + az * sn) * vx + (ay * ay * oc + cs) * vy + (ay * az * oc - ax * sn) * vz + (double) e.aux[1];
		double tz = (ax * az * oc - ay * sn) * vx + (ay * az * oc + ax * sn) * vy + (az * az * oc + cs) * vz + (double) e.aux[2];
		return mk3((float) (-(double) e.stiff * (p.x - tx) * (double) e.aux[3]), (float) (-(double) e.stiff * (p.y - ty) * (double) e.aux[4]),
				(float) (-(double) e.stiff * (p.z - tz) * (double) e.aux[5]));
	}
	case OXB_EXT_SPHERE_SMOOTH:
	case OXB_EXT_ELLIPSOID: {
		int4 ic;
		ic.x = (int) to_fixed(e.pos0[0], 1. / (double) box.lx); ic.y = (int) to_fixed(e.pos0[1], 1. / (double) box.ly); ic.z = (int) to_fixed(e.pos0[2], 1. / (double) box.lz);
		v3 d = min_image_fixed(box, ic, ip);
		float m = sqrtf(dot(d, d));
		if(e.type == OXB_EXT_SPHERE_SMOOTH) {
			// RepulsiveSphereSmooth.cpp:48-63
			if(m < e.r0 || m > e.aux[0]) return mk3(0.f, 0.f, 0.f);
			float s = (m >= e.aux[2]) ? -(e.stiff * 0.5f * expf((m - e.aux[2]) / e.aux[1])) / m
					: -(e.stiff * m - e.stiff * 0.5f * expf(-(m - e.aux[2]) / e.aux[1])) / m;
			return d * s;
		}
		// RepulsiveEllipsoid.cpp:55-67
		float in = d.x * d.x / (e.aux[0] * e.aux[0]) + d.y * d.y / (e.aux[1] * e.aux[1]) + d.z * d.z / (e.aux[2] * e.aux[2]);
		float out = d.x * d.x / (e.aux[3] * e.aux[3]) + d.y * d.y / (e.aux[4] * e.aux[4]) + d.z * d.z / (e.aux[5] * e.aux[5]);
		if(in < 1.f && out > 1.f) return mk3(0.f, 0.f, 0.f);
		return d * (-e.stiff / m);
	}
	case OXB_EXT_REPULSION_PLANE_MOVING: {
		// RepulsionPlaneMoving.cpp:58-66: one plane through every particle of the contiguous ref range (absolute positions)
		v3 f = mk3(0.f, 0.f, 0.f);
		for(int idx = e.ref; idx <= e.iaux; idx++) {
			double4 q = posd[slot_of[idx]];
			float d = (float) ((p.x - q.x) * (double) e.dir[0] + (p.y - q.y) * (double) e.dir[1] + (p.z - q.z) * (double) e.dir[2]);
			if(d < 0.f) {
				float s = -e.stiff * d;
				f += mk3(e.dir[0] * s, e.dir[1] * s, e.dir[2] * s);
			}
		}
		return f;
	}
	case OXB_EXT_GENERIC_CENTRAL: {
		// GenericCentralForce.cpp:174-193, force_type = gravity
		v3 d = mk3((float) (e.pos0[0] - p.x), (float) (e.pos0[1] - p.y), (float) (e.pos0[2] - p.z));
		float d2 = dot(d, d);
		if(d2 < e.aux[0] || (e.aux[1] > 0.f && d2 > e.aux[1])) return mk3(0.f, 0.f, 0.f);
		return d * (e.F0 * rsqrtf(d2));
	}
	case OXB_EXT_LJ_CONE: {
		// LJCone.cpp:69-92; aux[3..5] = sin, cos, tan of the half-opening angle
		v3 dir = mk3(e.dir[0], e.dir[1], e.dir[2]);
		v3 va = mk3((float) (p.x - e.pos0[0]), (float) (p.y - e.pos0[1]), (float) (p.z - e.pos0[2]));
		float d_along = dot(va, dir);
		v3 v_from_axis = dir * d_along - va;
		float d_from_axis = sqrtf(dot(v_from_axis, v_from_axis));
		float d_from_cone = d_along * e.aux[3] - d_from_axis * e.aux[4];
		float rel = d_from_cone / e.aux[0];
		if(rel > e.aux[1]) return mk3(0.f, 0.f, 0.f);
		v3 normal = dir * (d_along + d_from_axis * e.aux[5]) - va;
		normal = normal * rsqrtf(dot(normal, normal));
		float lj = powf(rel, -(float) e.iaux);
		return normal * (4.f * (float) e.iaux * e.stiff * (2.f * lj * lj - lj) / d_from_cone);
	}
	case OXB_EXT_YUKAWA_SPHERE:
	case OXB_EXT_SPHERE_MOVING: {
		double cx = e.pos0[0], cy = e.pos0[1], cz = e.pos0[2];
		if(e.type == OXB_EXT_SPHERE_MOVING && e.daux > 0.) {
			// RepulsiveSphereMoving.cpp:85-91: centre interpolated origin -> target over `steps` MD steps
			double t = (double) step / e.daux;
			t = t < 0. ? 0. : (t > 1. ? 1. : t);
			cx += ((double) e.aux[1] - cx) * t; cy += ((double) e.aux[2] - cy) * t; cz += ((double) e.aux[3] - cz) * t;
		}
		int4 ic;
		ic.x = (int) to_fixed(cx, 1. / (double) box.lx); ic.y = (int) to_fixed(cy, 1. / (double) box.ly); ic.z = (int) to_fixed(cz, 1. / (double) box.lz);
		// the steep WCA walls act on a DIFFERENCE of lengths (radius - |d|, |d| - radius): |d| in double, the rest in float
		const double dxd = (double) (int) ((unsigned) ip.x - (unsigned) ic.x) * ((double) box.lx * (1. / 4294967296.));
		const double dyd = (double) (int) ((unsigned) ip.y - (unsigned) ic.y) * ((double) box.ly * (1. / 4294967296.));
		const double dzd = (double) (int) ((unsigned) ip.z - (unsigned) ic.z) * ((double) box.lz * (1. / 4294967296.));
		const double md = sqrt(dxd * dxd + dyd * dyd + dzd * dzd);
		v3 d = mk3((float) dxd, (float) dyd, (float) dzd);
		float m = (float) md;
		if(e.type == OXB_EXT_YUKAWA_SPHERE) {
			// YukawaSphere.cpp:53-74 (the WCA exponent is 6 whatever WCA_n says, as there)
			float ds = (float) (e.r0d - md);
			if(!(ds < e.aux[4])) return mk3(0.f, 0.f, 0.f);
			float s = e.aux[3] * expf(-ds / e.aux[2]) * (1.f / (ds * e.aux[2]) + 1.f / (ds * ds));
			if(ds < e.aux[1]) {
				float w = e.aux[0] / ds;
				w = w * w * w; w = w * w;
				s += 4.f * e.stiff * (float) e.iaux * (2.f * w * w - w) / ds;
			}
			return d * (-s / m);
		}
		// RepulsiveSphereMoving.cpp:94-131: WCA (x = 2, sigma = 1, epsilon = stiff) in the surface gap r = |d| - radius
		float r = (float) (md - (e.r0d + (double) e.rate * (double) step));
		if(r >= e.aux[0] || m <= 0.f || r >= 1.41421356237f) return mk3(0.f, 0.f, 0.f);
		float rs = fmaxf(r, 1e-9f);
		float A = 1.f / (rs * rs);
		float fmag = 4.f * e.stiff * (2.f * A - 1.f) * (2.f / rs) * A;
		return d * (fmag / m);
	}
	default: return mk3(0.f, 0.f, 0.f);
	}
}

// COMForce.cpp:46-71: one block per force; centres of mass of com_list and ref_list from the absolute positions, then the
// spring force shared equally among the com_list particles (the reference recomputes both sums in every thread,
// CUDA_MD.cuh:441-468).  The ref_list particles feel nothing, as there.
// ---- metadynamics coordination bias (LTCoordination, src/Forces/Metadynamics/LTCoordination.cpp:94-221 over meta_utils.cpp:119-366;
// the reference's device version: src/CUDA/Forces/metad_forces.cuh:138-470).  Double precision from the FP64 state; constants are the
// float literals of src/model.h.  One block per force: the coordination (sum over the candidate pairs) by block reduction, then every
// pair adds bias force and lab-frame torque to both of its particles (the reference recomputes the whole sum in every thread).
__device__ __forceinline__ double coord_f4(double t, double t0, double a) {
	t = fabs(t - t0);
	return (t < 1.0 / sqrt(a)) ? 1.0 - a * t * t : 0.;
}
__device__ __forceinline__ double coord_f4Dsin(double t, double t0, double a) {
	double m = 1.0, tt0 = t - t0;
	if(tt0 < 0.0) { tt0 = -tt0; m = -1.0; }
	if(!(tt0 < 1.0 / sqrt(a))) return 0.;
	const double sint = sin(t);
	return (sint > 1e-10) ? m * 2.0 * a * tt0 / sint : m * 2.0 * a;
}
// geometry of a pair in double: centre separation r (minimum image), axes a1/a3 of p and b1/b3 of q
struct CoordPair { double r[3], a1[3], a3[3], b1[3], b3[3]; int pair_types_sum; };
__device__ inline CoordPair coord_load(const double4 *__restrict__ posd, const double4 *__restrict__ qd, const int4 *__restrict__ ipos, const double *L, int sp, int sq) {
	CoordPair P;
	const double4 pp = posd[sp], pq = posd[sq], qp = qd[sp], qq = qd[sq];
	P.r[0] = pq.x - pp.x; P.r[1] = pq.y - pp.y; P.r[2] = pq.z - pp.z;
	for(int k = 0; k < 3; k++) P.r[k] -= L[k] * rint(P.r[k] / L[k]);
	double a2[3];
	quatd Qp = { qp.x, qp.y, qp.z, qp.w }, Qq = { qq.x, qq.y, qq.z, qq.w };
	axes_from_quatd(Qp, P.a1, a2, P.a3);
	axes_from_quatd(Qq, P.b1, a2, P.b3);
	P.pair_types_sum = word_btype(ipos[sp].w) + word_btype(ipos[sq].w);
	return P;
}
// unsmoothed oxDNA2 hydrogen-bond energy of the pair; force on p and lab-frame torque on p if `want` (meta_utils.cpp:269-366)
__device__ inline double coord_hb(const CoordPair &P, double *force, double *torque, bool want) {
	const float PIf = 3.141592653589793238462643f;
	const double T0[6] = { 0.f, 0.f, 0.f, PIf, PIf * 0.5f, PIf * 0.5f }, A[6] = { 1.5f, 1.5f, 1.5f, 0.46f, 4.f, 4.f };
	for(int k = 0; k < 3; k++) force[k] = torque[k] = 0.;
	if(P.pair_types_sum != 3) return 0.;
	double rh[3];
	for(int k = 0; k < 3; k++) rh[k] = P.r[k] + 0.4f * P.b1[k] - 0.4f * P.a1[k];
	const double m = sqrt(rh[0] * rh[0] + rh[1] * rh[1] + rh[2] * rh[2]);
	if(!(0.276908f < m && m < 0.783775f)) return 0.;
	const double h[3] = { rh[0] / m, rh[1] / m, rh[2] / m };
	auto dot = [](const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
	const double c[6] = { -dot(P.a1, P.b1), -dot(P.b1, h), dot(P.a1, h), dot(P.a3, P.b3), -dot(P.b3, h), dot(P.a3, h) };
	// f1 without the smoothing branches (meta_utils.cpp:215-232); the shift is evaluated in float there
	const double shift = 1.0678f * (1.0 - (double) expf(-(0.75f - 0.4f) * 8.f)) * (1.0 - (double) expf(-(0.75f - 0.4f) * 8.f));
	const double ex = exp(-(m - 0.4f) * 8.f);
	const double f1 = 1.0678f * (1.0 - ex) * (1.0 - ex) - shift, f1D = 2.0 * 1.0678f * (1 - ex) * ex * 8.f;
	double t[6], f4[6], e = f1;
	for(int k = 0; k < 6; k++) { t[k] = acos(fmax(-1.0, fmin(1.0, c[k]))); f4[k] = coord_f4(t[k], T0[k], A[k]); e *= f4[k]; }
	if(!want || e == 0.) return e;
	const double sgn[6] = { 1., 1., -1., -1., 1., -1. };
	double pw[6];
	for(int k = 0; k < 6; k++) {
		pw[k] = f1 * sgn[k] * coord_f4Dsin(t[k], T0[k], A[k]);
		for(int j = 0; j < 6; j++) if(j != k) pw[k] *= f4[j];
	}
	const double all = f4[0] * f4[1] * f4[2] * f4[3] * f4[4] * f4[5];
	auto cross_add = [](double s, const double *a, const double *b, double *o) {
		o[0] += s * (a[1] * b[2] - a[2] * b[1]); o[1] += s * (a[2] * b[0] - a[0] * b[2]); o[2] += s * (a[0] * b[1] - a[1] * b[0]);
	};
	for(int k = 0; k < 3; k++) {
		force[k] = -h[k] * (f1D * all) + (P.b1[k] + h[k] * c[1]) * (pw[1] / m) + (P.a1[k] - h[k] * c[2]) * (pw[2] / m) + (P.b3[k] + h[k] * c[4]) * (pw[4] / m) +
				(P.a3[k] - h[k] * c[5]) * (pw[5] / m);
	}
	cross_add(-pw[3], P.a3, P.b3, torque);
	cross_add(-pw[0], P.a1, P.b1, torque);
	cross_add(pw[2], h, P.a1, torque);
	cross_add(pw[5], h, P.a3, torque);
	const double base[3] = { 0.4f * P.a1[0], 0.4f * P.a1[1], 0.4f * P.a1[2] };
	cross_add(1., base, force, torque);
	return e;
}
struct CoordCfg { int mode, n; double w, cut, width, d0, r0; };
__device__ inline double coord_switch_vec(const CoordPair &P, double *r) { // base(q) - base(p), returns its length
	for(int k = 0; k < 3; k++) r[k] = P.r[k] + 0.4f * P.b1[k] - 0.4f * P.a1[k];
	return sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
}
__device__ inline double coord_contribution(const CoordCfg &C, const CoordPair &P) {
	double f[3], t[3], hbc = 0., sw = 0.;
	if(C.mode != 1) {
		const double x = (C.cut - coord_hb(P, f, t, false)) / C.width;
		hbc = x > 10.0 ? 1.0 : (x < -10.0 ? 0.0 : 0.5 * (1.0 + tanh(x)));
	}
	if(C.mode != 0) {
		double r[3];
		sw = 1.0 / (1.0 + pow((coord_switch_vec(P, r) - C.d0) / C.r0, (double) C.n));
	}
	return C.mode == 0 ? hbc : (C.mode == 1 ? sw : C.w * hbc + (1.0 - C.w) * sw);
}
// d(contribution)/d(position of p) and d/d(rotation of p) (lab frame), meta_utils.cpp:159-205 with current = p
__device__ inline void coord_gradient(const CoordCfg &C, const CoordPair &P, double *force, double *torque) {
	double f[3] = { 0., 0., 0. }, t[3] = { 0., 0., 0. }, fs[3] = { 0., 0., 0. }, ts[3] = { 0., 0., 0. };
	if(C.mode != 1) {
		const double e = coord_hb(P, f, t, true), x = (C.cut - e) / C.width;
		double d = 0.;
		if(!(x > 10.0 || x < -10.0)) { const double th = tanh(x); d = -0.5 * (1.0 - th * th) / C.width; }
		for(int k = 0; k < 3; k++) { f[k] *= d; t[k] *= d; }
	}
	if(C.mode != 0) {
		double r[3];
		const double rm = coord_switch_vec(P, r), xx = (rm - C.d0) / C.r0, xn = pow(xx, (double) C.n);
		const double dcdr = ((double) C.n / C.r0) * pow(xx, (double) (C.n - 1)) / ((1.0 + xn) * (1.0 + xn));
		for(int k = 0; k < 3; k++) fs[k] = r[k] / rm * dcdr;
		const double b[3] = { 0.4f * P.a1[0], 0.4f * P.a1[1], 0.4f * P.a1[2] };
		ts[0] = b[1] * fs[2] - b[2] * fs[1]; ts[1] = b[2] * fs[0] - b[0] * fs[2]; ts[2] = b[0] * fs[1] - b[1] * fs[0];
	}
	const double w = C.mode == 2 ? C.w : (C.mode == 0 ? 1. : 0.);
	for(int k = 0; k < 3; k++) { force[k] = w * f[k] + (1. - w) * fs[k]; torque[k] = w * t[k] + (1. - w) * ts[k]; }
}

__global__ void k_ext_com(int n, const DevExtForce *__restrict__ ef, const int *__restrict__ pool, const float *__restrict__ grid, const int *__restrict__ slot_of,
		const double4 *__restrict__ posd, const double4 *__restrict__ qd, const int4 *__restrict__ ipos, double lx, double ly, double lz, long long step,
		const long long *__restrict__ cur_step, float4 *__restrict__ F, float4 *__restrict__ T, const int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	if(step < 0) step = cur_step[hw & 1];
	__shared__ double sh[6][4];
	const DevExtForce e = ef[blockIdx.x];
	if(e.type == OXB_EXT_META_COORDINATION) {
		CoordCfg C;
		C.mode = (int) e.aux[3]; C.n = e.pbc; C.w = (double) e.aux[5]; C.cut = (double) e.aux[6]; C.width = (double) e.aux[7]; C.d0 = (double) e.r0; C.r0 = (double) e.stiff;
		const double L[3] = { lx, ly, lz };
		const int n_pairs = e.iaux;
		double part = 0.;
		for(int k = threadIdx.x; k < n_pairs; k += blockDim.x)
			part += coord_contribution(C, coord_load(posd, qd, ipos, L, slot_of[pool[e.ref + 2 * k]], slot_of[pool[e.ref + 2 * k + 1]]));
		for(int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
		if((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = part;
		__syncthreads();
		double coord = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
		// LTCoordination::_coordination clamps to [coord_min, coord_max]; -dV/dcoord by finite difference on the grid, zero off it
		const double cmin = (double) e.aux[0], dc = (double) e.aux[1], cmax = (double) e.F0;
		coord = fmin(fmax(coord, cmin), cmax);
		const int il = (int) floor((coord - cmin) / dc);
		double df = 0.;
		if(il >= 0 && il + 1 <= (int) e.aux[2] - 1) {
			const float *g = grid + (int) e.aux[4];
			df = -((double) g[il + 1] - (double) g[il]) / dc;
		}
		for(int k = threadIdx.x; k < 2 * n_pairs; k += blockDim.x) {
			const int sp = slot_of[pool[e.ref + k]], sq = slot_of[pool[e.ref + (k ^ 1)]];
			double f[3], t[3];
			coord_gradient(C, coord_load(posd, qd, ipos, L, sp, sq), f, t);
			atomicAdd(&F[sp].x, (float) (df * f[0])); atomicAdd(&F[sp].y, (float) (df * f[1])); atomicAdd(&F[sp].z, (float) (df * f[2]));
			atomicAdd(&T[sp].x, (float) (df * t[0])); atomicAdd(&T[sp].y, (float) (df * t[1])); atomicAdd(&T[sp].z, (float) (df * t[2]));
		}
		return;
	}
	const int n_com = e.iaux, n_ref = e.pbc;
	double acc[6] = { 0., 0., 0., 0., 0., 0. };
	for(int k = threadIdx.x; k < n_com + n_ref; k += blockDim.x) {
		double4 q = posd[slot_of[pool[e.ref + k]]];
		int o = k < n_com ? 0 : 3;
		acc[o] += q.x; acc[o + 1] += q.y; acc[o + 2] += q.z;
	}
	for(int c = 0; c < 6; c++) {
		double x = acc[c];
		for(int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
		if((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = x;
	}
	__syncthreads();
	double d[3]; // second list's centre of mass - first list's
	for(int c = 0; c < 3; c++) d[c] = (sh[3 + c][0] + sh[3 + c][1] + sh[3 + c][2] + sh[3 + c][3]) / n_ref - (sh[c][0] + sh[c][1] + sh[c][2] + sh[c][3]) / n_com;
	int first = 0, count = n_com;
	double s;
	if(e.type == OXB_EXT_META_COM_TRAP) {
		// LTCOMTrap.cpp:52-78: dra = com(p1a) - com(p2a) (minimum image if PBC), bias force -dV/dx from the tabulated potential by finite
		// difference of the two grid points around x = |dra| (meta_utils.h:30-34), zero off the grid; mode 1 acts on p1a, mode 2 on p2a
		if(e.aux[5] != 0.f) { d[0] -= lx * rint(d[0] / lx); d[1] -= ly * rint(d[1] / ly); d[2] -= lz * rint(d[2] / lz); }
		const double m = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
		const int il = (int) floor((m - (double) e.aux[0]) / (double) e.aux[1]);
		double fx = 0.;
		if(il >= 0 && il + 1 <= (int) e.aux[2] - 1) {
			const float *g = grid + (int) e.aux[4];
			fx = -((double) g[il + 1] - (double) g[il]) / (double) e.aux[1];
		}
		// force on p1a: dra fx / |dra| / n1 with dra = -d; on p2a: the opposite / n2
		if((int) e.aux[3] == 1) s = -fx / m / n_com;
		else { s = fx / m / n_ref; first = n_com; count = n_ref; }
	}
	else {
		const double m = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
		s = (m - ((double) e.r0 + (double) e.rate * (double) step)) * (double) e.stiff / n_com / m;
	}
	for(int k = threadIdx.x; k < count; k += blockDim.x) {
		int i = slot_of[pool[e.ref + first + k]];
		atomicAdd(&F[i].x, (float) (d[0] * s));
		atomicAdd(&F[i].y, (float) (d[1] * s));
		atomicAdd(&F[i].z, (float) (d[2] * s));
	}
}

__global__ void k_ext_forces(int n, const DevExtForce *__restrict__ ef, const int *__restrict__ slot_of, const int4 *__restrict__ ipos,
		const double4 *__restrict__ posd, BoxF box, long long step, const long long *__restrict__ cur_step, float4 *__restrict__ F,
		const int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	// graph-launched batches: the step index lives on the device (the word the integrator of the previous launch wrote)
	if(step < 0) step = cur_step[hw & 1];
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if(k >= n) return;
	DevExtForce e = ef[k];
	int i = slot_of[e.particle];
	v3 f;
	if(e.type == OXB_EXT_MUTUAL_TRAP) {
		// MutualTrap.cpp:54-66
		float st = (float) step;
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
	else f = ext_single(e, posd[i], ipos[i], box, step, slot_of, posd);
	atomicAdd(&F[i].x, f.x);
	atomicAdd(&F[i].y, f.y);
	atomicAdd(&F[i].z, f.z);
}

__global__ void k_ext_forces_all(int N, int n_all, const DevExtForce *__restrict__ ef, const int *__restrict__ slot_of, const int4 *__restrict__ ipos,
		const double4 *__restrict__ posd, BoxF box, long long step, const long long *__restrict__ cur_step, float4 *__restrict__ F, const int *__restrict__ flags, int hw) {
	if(flags[hw]) return;
	if(step < 0) step = cur_step[hw & 1];
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	double4 p = posd[i];
	int4 ip = ipos[i];
	v3 f = mk3(0.f, 0.f, 0.f);
	for(int k = 0; k < n_all; k++) f += ext_single(ef[k], p, ip, box, step, slot_of, posd);
	// the interaction kernels add into F concurrently (other streams): atomics
	atomicAdd(&F[i].x, f.x);
	atomicAdd(&F[i].y, f.y);
	atomicAdd(&F[i].z, f.z);
}

} // namespace

namespace oxb {

static int env_int(const char *name, int dflt);
static bool particle_split() {
	static const bool on = [] { const char *v = getenv("OXB_PARTICLE_SPLIT"); return v == nullptr || v[0] != '0'; }();
	return on;
}

void launch_forces_particle(cudaStream_t s, const ModelRef &MR, BoxF box, int N, const int4 *ipos, const int4 *iback, const float4 *axf,
		const double4 *posd, const double4 *quatd, const int2 *bonds,
		const int *nbr, const int *nnbr, int stride, float4 *F, float4 *T, const oxb_replica_consts *rep, int n_per, int *flags, int hw) {
	int tpb = 128;
	if(MR.dna3) launch_forces_dna3(s, *MR.dna3, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, flags, hw);
	else if(particle_split()) {
		// register cap as resident blocks per SM: 3 = 168 registers (the kernel wants ~250; at 96 it spills 700 B).  OXB_PARTICLE_MB = 4: 128
		static const int mb = env_int("OXB_PARTICLE_MB", 3);
		const int blocks = (N + tpb - 1) / tpb;
		if(MR.rna) {
			if(mb >= 4) k_forces_particle_split<RnaModel, 4><<<blocks, tpb, 0, s>>>(*MR.rna, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, rep, n_per, flags, hw);
			else k_forces_particle_split<RnaModel, 3><<<blocks, tpb, 0, s>>>(*MR.rna, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, rep, n_per, flags, hw);
		}
		else if(mb >= 4) k_forces_particle_split<DnaModel, 4><<<blocks, tpb, 0, s>>>(*MR.dna, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, rep, n_per, flags, hw);
		else k_forces_particle_split<DnaModel, 3><<<blocks, tpb, 0, s>>>(*MR.dna, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, rep, n_per, flags, hw);
	}
	else if(MR.rna) k_forces_particle<RnaModel><<<(N + tpb - 1) / tpb, tpb, 0, s>>>(*MR.rna, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, rep, n_per, flags, hw);
	else k_forces_particle<DnaModel><<<(N + tpb - 1) / tpb, tpb, 0, s>>>(*MR.dna, box, N, ipos, iback, axf, posd, quatd, bonds, nbr, nnbr, stride, F, T, rep, n_per, flags, hw);
}

// the kernels of the edge pipeline, launched one by one so that the context can place them on concurrent streams:
//   which = 0 Debye-Hueckel (writes Fb) | 1 near edges (F, T, work lists) | 2 HB (+ cross stacking) | 3 coaxial stacking | 4 bonds
//           5 cross stacking only
static int env_int(const char *name, int dflt) {
	const char *v = getenv(name);
	return (v != nullptr && v[0] != 0) ? atoi(v) : dflt;
}

template<class MD>
static void launch_edge_stage_t(cudaStream_t s, int which, const typename MD::Params &M, BoxF box, const EdgeArgs &a, int *flags, int hw) {
	auto blocks_for = [&](long long items) { return (int) std::max<long long>(1, (items + 127) / 128); };
	switch(which) {
	case 0: {
		// lanes per particle: 1 by default.  2 and 4 were measured neutral to slower at 81,920 and 1M nucleotides
		// (profiles/smalln_sweep_r01.txt): the kernel is bound by L2 gather bandwidth, not by loads in flight.  OXB_DH_LPP overrides.
		static const int lpp_env = env_int("OXB_DH_LPP", 0);
		const int lpp = lpp_env > 0 ? lpp_env : 1;
		// OXB_DH_SORTED=0: fixed lane pairs l <-> 31 - l in the half-matrix kernel instead of pairs chosen by row length
		static const int dh_sorted = env_int("OXB_DH_SORTED", 1);
		if(a.dh_half && dh_sorted) {
			if(a.rep != nullptr) k_dh_particle<MD, 1, true, true, true><<<blocks_for(a.N), 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, a.rep, a.n_per, flags, hw);
			else k_dh_particle<MD, 1, false, true, true><<<blocks_for(a.N), 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, nullptr, 1, flags, hw);
		}
		else if(a.rep != nullptr) {
			if(a.dh_half) k_dh_particle<MD, 1, true, true><<<blocks_for(a.N), 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, a.rep, a.n_per, flags, hw);
			else k_dh_particle<MD, 1, true, false><<<blocks_for(a.N), 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, a.rep, a.n_per, flags, hw);
		}
		else if(a.dh_half) k_dh_particle<MD, 1, false, true><<<blocks_for(a.N), 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, nullptr, 1, flags, hw);
		else if(lpp >= 4) k_dh_particle<MD, 4, false, false><<<blocks_for(4ll * a.N), 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, nullptr, 1, flags, hw);
		else if(lpp == 2) k_dh_particle<MD, 2, false, false><<<blocks_for(2ll * a.N), 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, nullptr, 1, flags, hw);
		else k_dh_particle<MD, 1, false, false><<<blocks_for(a.N), 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, nullptr, 1, flags, hw);
		break;
	}
	// the producer and the three consumers of the segmented work lists share one fixed grid (a.n_seg blocks, grid-stride
	// inside): nothing here depends on device-side counts, so a captured graph stays valid across list rebuilds
	case 1:
		if(a.near_tile) k_edge_near<MD, true><<<a.n_seg, 128, 0, s>>>(M, box, a.n_edges, a.edge_cnt, a.N, a.edges, a.ipos, a.axf, a.F, a.T, a.hb_list, a.cx_list, a.cr_list,
				a.seg_counts, a.hb_seg, a.cx_seg, a.cr_seg, a.ex_list, a.ex_counts, a.ex_seg, a.refine, a.fold, a.posd, a.quatd, flags, hw);
		else k_edge_near<MD, false><<<a.n_seg, 128, 0, s>>>(M, box, a.n_edges, a.edge_cnt, a.N, a.edges, a.ipos, a.axf, a.F, a.T, a.hb_list, a.cx_list, a.cr_list,
				a.seg_counts, a.hb_seg, a.cx_seg, a.cr_seg, a.ex_list, a.ex_counts, a.ex_seg, a.refine, a.fold, a.posd, a.quatd, flags, hw);
		break;
	case 2: k_edge_heavy<MD, 0><<<dim3(a.n_seg, a.hb_split), 64, 0, s>>>(M, box, a.seg_counts, a.hb_list, a.hb_seg, a.ipos, a.axf, a.F, a.T, flags, hw); break;
	case 3: k_edge_heavy<MD, 1><<<a.n_seg, 64, 0, s>>>(M, box, a.seg_counts, a.cx_list, a.cx_seg, a.ipos, a.axf, a.F, a.T, flags, hw); break;
	case 6:
		k_excl_fix<MD><<<a.n_seg + (a.N + 255) / 256, 128, 0, s>>>(M, box, a.N, a.n_seg, a.ex_list, a.ex_counts, a.ex_seg, a.ex_bonded, a.bonds, a.posd, a.quatd, a.F, a.T,
				flags, hw);
		break;
	case 5: break; // the separate cross-stacking-only list is no longer produced (slot 2 of seg_counts now counts the back half of the hb segment)
	default: {
		static const int tpb_env = env_int("OXB_TPB_BONDED", 0);
		const int tpb = tpb_env > 0 ? tpb_env : 128;
		k_bonded<MD><<<(a.N + tpb - 1) / tpb, tpb, 0, s>>>(M, box, a.N, a.ipos, a.iback, a.axf, a.bonds, a.F, a.T, a.ex_bonded, a.refine, a.fold, a.posd, a.quatd, a.rep, a.n_per, flags, hw);
		break;
	}
	}
}

// oxDNA3: stage numbers as above; stage 6 (parked excluded volume) does not exist -- active terms are refined in place
static void launch_edge_stage_dna3(cudaStream_t s, int which, const oxb_dna3_dev &M, BoxF box, const EdgeArgs &a, int *flags, int hw) {
	const int nb = (a.N + 127) / 128;
	const double4 *posd = a.refine ? a.posd : nullptr;
	static const int cfg = std::min(2, std::max(0, env_int("OXB_DNA3_EDGE_MB", 0)));
	static const int dh_sorted = env_int("OXB_DH_SORTED", 1);
	switch(which == 0 ? 0 : 10 * cfg + which) {
	case 0:
		if(a.dh_half && dh_sorted) k_dh_particle<Dna3Dh, 1, false, true, true><<<nb, 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, nullptr, 1, flags, hw);
		else if(a.dh_half) k_dh_particle<Dna3Dh, 1, false, true><<<nb, 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, nullptr, 1, flags, hw);
		else k_dh_particle<Dna3Dh, 1, false, false><<<nb, 128, 0, s>>>(M, box, a.N, a.iback, a.dh_nbr, a.dh_nnbr, a.Fb, nullptr, 1, flags, hw);
		break;
	// register caps (resident blocks per SM asked of the compiler), OXB_DNA3_EDGE_MB = 0: 128 / 120 / 128 registers for near / heavy / bonded,
	// 1: 80 / 80 / 96, 2: 64 / 64 / 80 (measured: profiles/sweeps_r02.txt, aw)
#define OXB_K3(CFG, NEAR, HEAVY, BONDED)                                                                                                                           \
	case 10 * CFG + 1: k3_edge_near<NEAR><<<a.n_seg, 128, 0, s>>>(M, box, a.n_edges, a.edges, a.ipos, a.axf, a.F, a.T, a.hb_list, a.cx_list, a.seg_counts, a.hb_seg, a.cx_seg, posd, a.quatd, flags, hw); break; \
	case 10 * CFG + 2: k3_edge_heavy<0, HEAVY><<<dim3(a.n_seg, a.hb_split), 64, 0, s>>>(M, box, a.seg_counts, a.hb_list, a.hb_seg, a.ipos, a.axf, a.F, a.T, flags, hw); break;  \
	case 10 * CFG + 3: k3_edge_heavy<1, HEAVY><<<a.n_seg, 64, 0, s>>>(M, box, a.seg_counts, a.cx_list, a.cx_seg, a.ipos, a.axf, a.F, a.T, flags, hw); break;                    \
	case 10 * CFG + 4: k3_bonded<BONDED><<<nb, 128, 0, s>>>(M, box, a.N, a.ipos, a.iback, a.axf, a.bonds, a.F, a.T, posd, a.quatd, flags, hw); break;
	OXB_K3(0, 4, 6, 4)
	OXB_K3(1, 6, 12, 5)
	OXB_K3(2, 8, 16, 6)
#undef OXB_K3
	default: break;
	}
}

void launch_edge_stage(cudaStream_t s, int which, const ModelRef &MR, BoxF box, const EdgeArgs &a, int *flags, int hw) {
	if(MR.dna3) { launch_edge_stage_dna3(s, which, *MR.dna3, box, a, flags, hw); return; }
	if(MR.rna) launch_edge_stage_t<RnaModel>(s, which, *MR.rna, box, a, flags, hw);
	else launch_edge_stage_t<DnaModel>(s, which, *MR.dna, box, a, flags, hw);
}

void launch_energy_split(cudaStream_t s, const ModelRef &MR, BoxF box, int N, const int4 *ipos, const float4 *axf, const int2 *bonds, const int *nbr,
		const int *nnbr, int stride, double *out) {
	if(MR.dna3) { launch_energy_split_dna3(s, *MR.dna3, box, N, ipos, axf, bonds, nbr, nnbr, stride, out); return; }
	cudaMemsetAsync(out, 0, sizeof(double) * OXB_NTERMS, s);
	int tpb = 128;
	if(MR.rna) k_energy_split<RnaModel><<<(N + tpb - 1) / tpb, tpb, 0, s>>>(*MR.rna, box, N, ipos, axf, bonds, nbr, nnbr, stride, out);
	else k_energy_split<DnaModel><<<(N + tpb - 1) / tpb, tpb, 0, s>>>(*MR.dna, box, N, ipos, axf, bonds, nbr, nnbr, stride, out);
}

void launch_ext_forces(cudaStream_t s, int n, const DevExtForce *ef, const int *slot_of, const int4 *ipos, const double4 *posd, BoxF box,
		long long step, const long long *cur_step, float4 *F, const int *flags, int hw) {
	if(n <= 0) return;
	int tpb = 128;
	k_ext_forces<<<(n + tpb - 1) / tpb, tpb, 0, s>>>(n, ef, slot_of, ipos, posd, box, step, cur_step, F, flags, hw);
}

void launch_ext_forces_all(cudaStream_t s, int N, int n_all, const DevExtForce *ef_all, const int *slot_of, const int4 *ipos, const double4 *posd, BoxF box,
		long long step, const long long *cur_step, float4 *F, const int *flags, int hw) {
	if(n_all <= 0) return;
	int tpb = 128;
	k_ext_forces_all<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, n_all, ef_all, slot_of, ipos, posd, box, step, cur_step, F, flags, hw);
}

void launch_ext_com(cudaStream_t s, int n, const DevExtForce *ef_com, const int *pool, const float *grid, const int *slot_of, const double4 *posd,
		const double4 *quatd, const int4 *ipos, const double *box, long long step, const long long *cur_step, float4 *F, float4 *T, const int *flags, int hw) {
	if(n <= 0) return;
	k_ext_com<<<n, 128, 0, s>>>(n, ef_com, pool, grid, slot_of, posd, quatd, ipos, box[0], box[1], box[2], step, cur_step, F, T, flags, hw);
}

} // namespace oxb
