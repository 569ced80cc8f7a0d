// Velocity-Verlet quaternion integrator + thermostats, one fused streaming kernel (sm_100a, HBM-bound).
//
// Reference: first_step_mixed / second_step_mixed (src/CUDA/Backends/CUDA_mixed.cuh:8-93), brownian_thermostat
// (src/CUDA/Thermostats/CUDABrownianThermostat.cu:15-51), langevin_thermostat (CUDALangevinThermostat.cu:14-40),
// bussi_thermostat + host-side K update (CUDABussiThermostat.cu:55-122, src/Backends/Thermostats/BussiThermostat.cpp:45-150).
//
// Differences by design:
//  * the second half-kick of step n, the thermostat of step n and the first half-kick + drift + rotation of step n+1
//    act on the same per-particle data with the same forces, so they are ONE pass over the FP64 state instead of
//    three (the reference streams velocities/momenta 3x and converts precision around the thermostat);
//  * RNG is counter-based Philox keyed by (seed, original particle id, step): no 48-byte curandState per particle;
//  * the Bussi kinetic-energy dynamics runs on the device (one warp), no host round trip;
//  * list staleness raises a device flag that turns the rest of a speculative batch of launches into no-ops.
#include "kernels.h"

#include <cstdlib>

namespace {

__device__ __forceinline__ void philox_gauss4(unsigned long long seed, unsigned id, unsigned long long step, unsigned draw, float g[4]) {
	uint4 ctr = make_uint4(id, (unsigned) step, (unsigned) (step >> 32), draw);
	uint2 key = make_uint2((unsigned) seed, (unsigned) (seed >> 32));
	uint4 r = Philox::gen(ctr, key);
	float u0 = u01(r.x), u1 = u01(r.y), u2 = u01(r.z), u3 = u01(r.w);
	float m0 = sqrtf(-2.f * logf(u0)), m1 = sqrtf(-2.f * logf(u2));
	float s0, c0, s1, c1;
	sincospif(2.f * u1, &s0, &c0);
	sincospif(2.f * u3, &s1, &c1);
	g[0] = m0 * c0; g[1] = m0 * s0; g[2] = m1 * c1; g[3] = m1 * s1;
}

__device__ __forceinline__ uint4 philox_u4(unsigned long long seed, unsigned id, unsigned long long step, unsigned draw) {
	return Philox::gen(make_uint4(id, (unsigned) step, (unsigned) (step >> 32), draw), make_uint2((unsigned) seed, (unsigned) (seed >> 32)));
}

#ifndef OXB_MB_INTEGRATE
#define OXB_MB_INTEGRATE 1
#endif

template<int PH>
__global__ void __launch_bounds__(256, OXB_MB_INTEGRATE) k_integrate(oxb::IntegrateArgs a, int epoch) {
	int *flags = a.flags;
	const int rd = OXB_FLAG_COUNT + (epoch & 1), wr = OXB_FLAG_COUNT + ((epoch + 1) & 1);
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	// halted earlier in this batch: stay halted (sticky), do nothing.  An incomplete force pass (a work-list segment overflowed) halts the
	// batch the same way: the forces must not be integrated; the host enlarges the segments and repeats the pass
	if(flags[rd] || (flags[OXB_FLAG_ERROR] & OXB_ERR_SEG_OVERFLOW)) {
		if(i == 0) { flags[wr] = 1; prof_mark(flags, OXB_PROF_WAIT); }
		return;
	}
	if(i == 0) prof_mark(flags, OXB_PROF_INTEG);
	// step index: explicit (stream-launched paths) or the device-side counter (graph-launched batches, where kernel
	// arguments are frozen): word (epoch & 1) is read, word ((epoch + 1) & 1) is written -- never the same word
	const long long step = (a.step >= 0) ? a.step : a.cur_step[epoch & 1];
	if(i == 0) {
		if(PH & OXB_PH_COUNT_STEP) atomicAdd(flags + OXB_FLAG_STEPS_DONE, 1);
		a.cur_step[(epoch + 1) & 1] = step + ((PH & OXB_PH_COUNT_STEP) ? 1 : 0);
	}

	double sv[5] = { 0., 0., 0., 0., 0. };
	if(i < a.N) {
		const double hdt = 0.5 * a.dt;
		float4 F = a.F[i], T = a.T[i];
		double4 qd0 = make_double4(0., 0., 0., 1.);
		// every load of the launch is issued here, before the first store: the arrays of IntegrateArgs may alias as far as the compiler
		// knows, so a load written behind a store is issued behind it and its latency is exposed (position, fixed-point position and
		// backbone site were fetched one after the other in the middle of the arithmetic)
		double4 r0 = make_double4(0., 0., 0., 0.);
		int4 ip0 = make_int4(0, 0, 0, 0), ib0 = make_int4(0, 0, 0, 0);
		if(PH & OXB_PH_FIRST) { r0 = a.posd[i]; ip0 = a.ipos[i]; ib0 = a.iback[i]; }
		else if(PH & OXB_PH_THERMO) ip0 = a.ipos[i];
		const double4 v0 = a.veld[i], L0 = a.Ld[i];
		float4 fb0 = make_float4(0.f, 0.f, 0.f, 0.f);
		if((PH & (OXB_PH_SECOND | OXB_PH_FIRST)) && a.Fb != nullptr) fb0 = a.Fb[i];
		if(PH & (OXB_PH_SECOND | OXB_PH_FIRST)) {
			// the force kernels accumulate the torque in the lab frame: rotate it into the body frame (L is a body-frame
			// angular momentum with unit inertia, src/CUDA/Interactions/CUDA_DNA.cuh:896)
			Axes A;
			if(PH & OXB_PH_FIRST) {
				// the first-half phase streams the FP64 quaternion anyway: no second orientation read
				qd0 = a.quatd[i];
				A = axes_from_quat(make_float4((float) qd0.x, (float) qd0.y, (float) qd0.z, (float) qd0.w));
			}
			else {
				// plain loads (not __ldg): other variants of this kernel rewrite the record
				const float4 u = a.axf[2 * (size_t) i], w = a.axf[2 * (size_t) i + 1];
				A.a1 = mk3(u.x, u.y, u.z); A.a3 = mk3(u.w, w.x, w.y); A.a2 = cross(A.a3, A.a1);
			}
			v3 tl = mk3(T.x, T.y, T.z);
			if(a.Fb != nullptr) {
				// edge pipeline: the Debye-Hueckel kernel leaves its force sum, acting at the backbone site, in Fb
				const float4 fb = fb0;
				v3 g = mk3(fb.x, fb.y, fb.z);
				F.x += g.x; F.y += g.y; F.z += g.z;
				tl += cross(A.a1 * a.back_a1 + A.a2 * a.back_a2 + A.a3 * a.back_a3, g);
			}
			T.x = dot(A.a1, tl); T.y = dot(A.a2, tl); T.z = dot(A.a3, tl);
		}
		double4 v = v0, L = L0;
		if(PH & OXB_PH_SECOND) {
			v.x += F.x * hdt; v.y += F.y * hdt; v.z += F.z * hdt;
			L.x += T.x * hdt; L.y += T.y * hdt; L.z += T.z * hdt;
		}
		if(PH & OXB_PH_BUSSI_SUMS) {
			sv[0] = v.x; sv[1] = v.y; sv[2] = v.z;
			sv[3] = v.x * v.x + v.y * v.y + v.z * v.z;
			sv[4] = L.x * L.x + L.y * L.y + L.z * L.z;
		}
		if(PH & OXB_PH_BUSSI_APPLY) {
			const KinSums S = *a.sums;
			double cx = S.vx / a.N, cy = S.vy / a.N, cz = S.vz / a.N;
			v.x = (v.x - cx) * S.factor_t + cx; v.y = (v.y - cy) * S.factor_t + cy; v.z = (v.z - cz) * S.factor_t + cz;
			L.x *= S.factor_r; L.y *= S.factor_r; L.z *= S.factor_r;
		}
		if((PH & OXB_PH_THERMO) && (a.th.type == OXB_THERMOSTAT_LANGEVIN || (step % a.th.every) == 0)) {
			unsigned id = (unsigned) word_index(ip0.w);
			if(a.rep != nullptr) {
				// replica batching: the thermostat constants of this particle's replica (slots are replica-contiguous)
				const oxb_replica_consts *rc = a.rep + i / a.n_per;
				a.th.a = rc->th_a; a.th.b = rc->th_b; a.th.c = rc->th_c; a.th.d = rc->th_d;
			}
			if(a.th.type == OXB_THERMOSTAT_BROWNIAN) {
				uint4 u = philox_u4(a.th.seed, id, (unsigned long long) step, 0u);
				bool rt = u01(u.x) < a.th.a, rr = u01(u.y) < a.th.b;
				if(rt || rr) {
					float g0[4], g1[4];
					philox_gauss4(a.th.seed, id, (unsigned long long) step, 1u, g0);
					philox_gauss4(a.th.seed, id, (unsigned long long) step, 2u, g1);
					if(rt) { v.x = g0[0] * a.th.c; v.y = g0[1] * a.th.c; v.z = g0[2] * a.th.c; }
					if(rr) { L.x = g1[0] * a.th.c; L.y = g1[1] * a.th.c; L.z = g1[2] * a.th.c; }
				}
			}
			else if(a.th.type == OXB_THERMOSTAT_LANGEVIN) {
				float g0[4], g1[4];
				philox_gauss4(a.th.seed, id, (unsigned long long) step, 1u, g0);
				philox_gauss4(a.th.seed, id, (unsigned long long) step, 2u, g1);
				double dt = a.dt;
				v.x += dt * (-a.th.a * v.x + g0[0] * a.th.c); v.y += dt * (-a.th.a * v.y + g0[1] * a.th.c); v.z += dt * (-a.th.a * v.z + g0[2] * a.th.c);
				L.x += dt * (-a.th.b * L.x + g1[0] * a.th.d); L.y += dt * (-a.th.b * L.y + g1[1] * a.th.d); L.z += dt * (-a.th.b * L.z + g1[2] * a.th.d);
			}
		}
		if(PH & OXB_PH_FIRST) {
			v.x += F.x * hdt; v.y += F.y * hdt; v.z += F.z * hdt;
			L.x += T.x * hdt; L.y += T.y * hdt; L.z += T.z * hdt;
			double4 r = r0;
			r.x += v.x * a.dt; r.y += v.y * a.dt; r.z += v.z * a.dt;
			a.posd[i] = r;
			int4 ip = ip0;
			ip.x = (int) to_fixed(r.x, a.box_inv[0]); ip.y = (int) to_fixed(r.y, a.box_inv[1]); ip.z = (int) to_fixed(r.z, a.box_inv[2]);
			a.ipos[i] = ip;
			// body-frame rotation by |L| dt about L: q <- q (x) (Lhat sin(th/2), cos(th/2))
			double n2 = L.x * L.x + L.y * L.y + L.z * L.z;
			double4 qn = qd0;
			if(n2 > 0.) {
				// half angle th = dt |L| / 2 is ~1e-3: sin(th)/|L| and cos(th) from their Taylor series (remainder < 1e-20 for
				// th < 0.03) -- no square root, no division, no slow-path double sincos; the general path stays for huge |L|
				const double h = 0.5 * a.dt, t2 = h * h * n2;
				double k, ch;
				if(t2 < 9e-4) {
					k = h * (1. + t2 * (-1. / 6. + t2 * (1. / 120. + t2 * (-1. / 5040. + t2 * (1. / 362880.)))));
					ch = 1. + t2 * (-0.5 + t2 * (1. / 24. + t2 * (-1. / 720. + t2 * (1. / 40320. - t2 * (1. / 3628800.)))));
				}
				else {
					double n = sqrt(n2), sh;
					sincos(h * n, &sh, &ch);
					k = sh / n;
				}
				double bx = L.x * k, by = L.y * k, bz = L.z * k, bw = ch;
				double4 q = qn, o;
				o.w = q.w * bw - q.x * bx - q.y * by - q.z * bz;
				o.x = q.w * bx + q.x * bw + q.y * bz - q.z * by;
				o.y = q.w * by - q.x * bz + q.y * bw + q.z * bx;
				o.z = q.w * bz + q.x * by - q.y * bx + q.z * bw;
				a.quatd[i] = o;
				qn = o;
			}
			{
				// fixed-point backbone-site position for the Debye-Hueckel kernel: r + back_a1 a1 + back_a2 a2
				double sqx = qn.x * qn.x, sqy = qn.y * qn.y, sqz = qn.z * qn.z, sqw = qn.w * qn.w;
				double xy = qn.x * qn.y, xz = qn.x * qn.z, xw = qn.x * qn.w, yz = qn.y * qn.z, yw = qn.y * qn.w, zw = qn.z * qn.w;
				// FP32 orientation record for the pair kernels (a1, a3), from the same double products
				if(n2 > 0.) store_axes(a.axf, i, sqx - sqy - sqz + sqw, 2. * (xy + zw), 2. * (xz - yw), 2. * (xz + yw), 2. * (yz - xw), -sqx - sqy + sqz + sqw);
				double b1 = a.back_a1, b2 = a.back_a2, b3 = a.back_a3;
				double bx = r.x + b1 * (sqx - sqy - sqz + sqw) + b2 * (2. * (xy - zw)) + b3 * (2. * (xz + yw));
				double by = r.y + b1 * (2. * (xy + zw)) + b2 * (-sqx + sqy - sqz + sqw) + b3 * (2. * (yz - xw));
				double bz = r.z + b1 * (2. * (xz - yw)) + b2 * (2. * (yz + xw)) + b3 * (-sqx - sqy + sqz + sqw);
				int4 ib = ib0;
				ib.x = (int) to_fixed(bx, a.box_inv[0]); ib.y = (int) to_fixed(by, a.box_inv[1]); ib.z = (int) to_fixed(bz, a.box_inv[2]);
				a.iback[i] = ib;
				// rotational staleness: neither the backbone site nor the base site may have moved further than the skin
				v3 db = min_image_fixed(a.box, unpack_ref(v.w), ib);
				if(dot(db, db) > a.skin2) flags[wr] = 1;
				double c1 = a.base_a1;
				int4 is = ib;
				is.x = (int) to_fixed(r.x + c1 * (sqx - sqy - sqz + sqw), a.box_inv[0]);
				is.y = (int) to_fixed(r.y + c1 * (2. * (xy + zw)), a.box_inv[1]);
				is.z = (int) to_fixed(r.z + c1 * (2. * (xz - yw)), a.box_inv[2]);
				v3 ds = min_image_fixed(a.box, unpack_ref(L.w), is);
				if(dot(ds, ds) > a.skin2) flags[wr] = 1;
			}
			// forces are consumed: leave zeroed accumulators for the next force pass
			a.F[i] = make_float4(0.f, 0.f, 0.f, 0.f);
			a.T[i] = make_float4(0.f, 0.f, 0.f, 0.f);
			if(a.zero_Fb) a.Fb[i] = make_float4(0.f, 0.f, 0.f, 0.f); // half-matrix Debye-Hueckel kernel: Fb is an accumulator
			v3 d = min_image_fixed(a.box, unpack_ref(r.w), ip);
			if(dot(d, d) > a.skin2) flags[wr] = 1;
		}
		a.veld[i] = v;
		a.Ld[i] = L;
	}
	if(PH & OXB_PH_BUSSI_SUMS) {
		__shared__ double sh[5][8];
#pragma unroll
		for(int c = 0; c < 5; c++) {
			double x = sv[c];
			for(int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
			if((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = x;
		}
		__syncthreads();
		if(threadIdx.x < 5) {
			double x = 0.;
			for(int w = 0; w < (int) (blockDim.x >> 5); w++) x += sh[threadIdx.x][w];
			double *dst = &a.sums->vx + threadIdx.x;
			atomicAdd(dst, x);
		}
	}
}

// ---- Bussi stochastic velocity rescaling, kinetic-energy dynamics (BussiThermostat.cpp:45-52,105-150) on one thread.
struct SerialRng {
	unsigned long long seed, step;
	unsigned draw;
	__device__ float gauss() {
		float g[4];
		philox_gauss4(seed, 0xFFFFFFFFu, step, draw++, g);
		return g[0];
	}
	__device__ float unif() {
		uint4 u = philox_u4(seed, 0xFFFFFFFEu, step, draw++);
		return u01(u.x);
	}
	// Gamma(k, 1), Marsaglia-Tsang (k >= 1) -- same distribution as the reference's gamdev(ia) for integer ia
	__device__ double gamma(double k) {
		double d = k - 1. / 3., c = 1. / sqrt(9. * d);
		for(int it = 0; it < 64; it++) {
			double x = gauss(), v = 1. + c * x;
			if(v <= 0.) continue;
			v = v * v * v;
			double u = unif();
			if(log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
		}
		return d;
	}
	// sum of nn squared standard normals
	__device__ double sum_noises(int nn) {
		if(nn == 0) return 0.;
		if(nn == 1) { double r = gauss(); return r * r; }
		if(nn % 2 == 0) return 2. * gamma(nn / 2);
		double r = gauss();
		return 2. * gamma((nn - 1) / 2) + r * r;
	}
};

__device__ void bussi_update_K(SerialRng &R, double &K, int dof, double T, double ex) {
	double Kt = 0.5 * dof * T;
	double rr = R.gauss();
	double dK = (1.0 - ex) * (Kt * (R.sum_noises(dof - 1) + rr * rr) / dof - K) + 2.0 * rr * sqrt(K * Kt / dof * (1.0 - ex) * ex);
	K += dK;
}

__global__ void k_bussi_update(KinSums *S, int N, ThermostatCfg th, long long step, const int *flags, int epoch) {
	if(flags[OXB_FLAG_COUNT + (epoch & 1)]) return;
	if(threadIdx.x != 0 || blockIdx.x != 0) return;
	SerialRng R;
	R.seed = th.seed; R.step = (unsigned long long) step; R.draw = 16;
	double T = th.a, ex = th.b;
	int dof_t = 3 * (N - 1), dof_r = 3 * N;
	double cx = S->vx / N, cy = S->vy / N, cz = S->vz / N;
	double Know_t = 0.5 * (S->v2 - N * (cx * cx + cy * cy + cz * cz));
	double Know_r = 0.5 * S->L2;
	double Kt = S->K_t, Kr = S->K_r;
	bussi_update_K(R, Kt, dof_t, T, ex);
	bussi_update_K(R, Kr, dof_r, T, ex);
	S->K_t = Kt; S->K_r = Kr;
	S->factor_t = sqrt(Kt / Know_t);
	S->factor_r = sqrt(Kr / Know_r);
}

__global__ void k_clear_sums(KinSums *S, const int *flags, int epoch) {
	if(epoch >= 0 && flags[OXB_FLAG_COUNT + (epoch & 1)]) return;
	S->vx = S->vy = S->vz = S->v2 = S->L2 = 0.;
}

__global__ void __launch_bounds__(256) k_kinetic_sums(int N, const double4 *__restrict__ veld, const double4 *__restrict__ Ld, KinSums *S) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	double sv[5] = { 0., 0., 0., 0., 0. };
	if(i < N) {
		double4 v = veld[i], L = Ld[i];
		sv[0] = v.x; sv[1] = v.y; sv[2] = v.z;
		sv[3] = v.x * v.x + v.y * v.y + v.z * v.z;
		sv[4] = L.x * L.x + L.y * L.y + L.z * L.z;
	}
	__shared__ double sh[5][8];
#pragma unroll
	for(int c = 0; c < 5; c++) {
		double x = sv[c];
		for(int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
		if((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = x;
	}
	__syncthreads();
	if(threadIdx.x < 5) {
		double x = 0.;
		for(int w = 0; w < (int) (blockDim.x >> 5); w++) x += sh[threadIdx.x][w];
		atomicAdd(&S->vx + threadIdx.x, x);
	}
}

__global__ void __launch_bounds__(256) k_energy_sum(int N, const float4 *__restrict__ F, const float4 *__restrict__ Fb, double *out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	double x = (i < N) ? (double) F[i].w : 0.;
	if(Fb != nullptr && i < N) x += (double) Fb[i].w;
	for(int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
	__shared__ double sh[8];
	if((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
	__syncthreads();
	if(threadIdx.x == 0) {
		double s = 0.;
		for(int w = 0; w < (int) (blockDim.x >> 5); w++) s += sh[w];
		atomicAdd(out, s);
	}
}

// per-replica potential energy: slots [r * n_per, (r + 1) * n_per) belong to replica r; a block that lies inside one replica reduces
// in shared memory and issues one atomic, a block that straddles a boundary falls back to one atomic per warp / lane
__global__ void __launch_bounds__(256) k_energy_sum_replicas(int N, int n_per, const float4 *__restrict__ F, const float4 *__restrict__ Fb, double *out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	double x = (i < N) ? (double) F[i].w : 0.;
	if(Fb != nullptr && i < N) x += (double) Fb[i].w;
	const int first = blockIdx.x * blockDim.x, last = min(N, first + (int) blockDim.x) - 1;
	const int r = min(i, N - 1) / n_per;
	if(first / n_per == last / n_per) {
		for(int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
		__shared__ double sh[8];
		if((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
		__syncthreads();
		if(threadIdx.x == 0) {
			double s = 0.;
			for(int w = 0; w < (int) (blockDim.x >> 5); w++) s += sh[w];
			atomicAdd(out + r, s);
		}
	}
	else if(i < N) atomicAdd(out + r, x);
}

template<int PH>
void launch_ph(cudaStream_t s, const oxb::IntegrateArgs &a, int epoch) {
	// small systems: 128-thread blocks halve the tail of the last wave (C2: 320 blocks of 256 on 148 SMs = 2.2 blocks per SM)
	static const int tpb_env = [] { const char *v = getenv("OXB_TPB_INTEGRATE"); return (v != nullptr && v[0] != 0) ? atoi(v) : 0; }();
	int tpb = tpb_env > 0 ? tpb_env : 256;
	k_integrate<PH><<<(a.N + tpb - 1) / tpb, tpb, 0, s>>>(a, epoch);
}

// ---- MC barostat support (MD_CUDABackend::_rescale_positions / _rescale_molecular_positions, src/CUDA/Backends/MD_CUDABackend.cu:412-449;
// kernels compute_molecular_coms / rescale_molecular_positions / rescale_positions, src/CUDA/Backends/CUDA_MD.cuh:62-95), on the FP64 state
__global__ void k_mol_coms(int N, const int4 *__restrict__ ipos, const int *__restrict__ mol_of, const double *__restrict__ inv_size,
		const double4 *__restrict__ posd, double *__restrict__ coms) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	const int m = mol_of[word_index(ipos[i].w)];
	const double4 r = posd[i];
	const double w = inv_size[m];
	atomicAdd(coms + 3 * m, r.x * w); atomicAdd(coms + 3 * m + 1, r.y * w); atomicAdd(coms + 3 * m + 2, r.z * w);
}

// SimBackend::fix_diffusion (src/Backends/SimBackend.cpp:786-882) + CubicBox::shift_particle (src/Boxes/CubicBox.cpp:70-77): every strand is
// translated by whole box sides so that its centre of mass lies inside the box; the unit quaternion is renormalised (the reference
// orthonormalises the rotation matrix).  A shift by whole box sides leaves every minimum-image separation -- hence energy, forces and
// Verlet lists -- untouched, so the reference's before/after energy check has nothing to catch here.  shifts (may be null): floor(com / L)
// per original particle id, for the caller's BaseParticle::_pos_shift.
__global__ void k_fix_diffusion(int N, const int4 *__restrict__ ipos, const int *__restrict__ mol_of, const double *__restrict__ coms, double lx,
		double ly, double lz, double4 *__restrict__ posd, double4 *__restrict__ quatd, float4 *__restrict__ axf, int *__restrict__ shifts) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	const int id = word_index(ipos[i].w);
	const double *cm = coms + 3 * mol_of[id];
	const double sx = floor(cm[0] / lx), sy = floor(cm[1] / ly), sz = floor(cm[2] / lz);
	double4 r = posd[i];
	r.x -= lx * sx; r.y -= ly * sy; r.z -= lz * sz;
	posd[i] = r;
	double4 q = quatd[i];
	const double n = rsqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
	q.x *= n; q.y *= n; q.z *= n; q.w *= n;
	quatd[i] = q;
	store_axes_from_quatd(axf, i, q.x, q.y, q.z, q.w);
	if(shifts) { shifts[3 * id] = (int) sx; shifts[3 * id + 1] = (int) sy; shifts[3 * id + 2] = (int) sz; }
}

// molecular: r += com(molecule) * shift; atomic: r *= ratio.  Then the fixed-point centre and backbone site are re-encoded for the NEW box.
__global__ void k_rescale_positions(oxb::RescaleArgs a) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= a.N) return;
	double4 r = a.posd[i];
	int4 ip = a.ipos[i];
	if(a.backup) a.backup[word_index(ip.w)] = r; // snapshot by ORIGINAL id: a re-sort may reorder the slots before a rejection
	if(a.restore) r = a.restore[word_index(ip.w)];
	else if(a.molecular) {
		const double *cm = a.coms + 3 * a.mol_of[word_index(ip.w)];
		r.x += cm[0] * a.f[0]; r.y += cm[1] * a.f[1]; r.z += cm[2] * a.f[2];
	}
	else { r.x *= a.f[0]; r.y *= a.f[1]; r.z *= a.f[2]; }
	a.posd[i] = r;
	ip.x = (int) to_fixed(r.x, a.box_inv[0]); ip.y = (int) to_fixed(r.y, a.box_inv[1]); ip.z = (int) to_fixed(r.z, a.box_inv[2]);
	a.ipos[i] = ip;
	const double4 qn = a.quatd[i];
	double sqx = qn.x * qn.x, sqy = qn.y * qn.y, sqz = qn.z * qn.z, sqw = qn.w * qn.w;
	double xy = qn.x * qn.y, xz = qn.x * qn.z, xw = qn.x * qn.w, yz = qn.y * qn.z, yw = qn.y * qn.w, zw = qn.z * qn.w;
	double b1 = a.back_a1, b2 = a.back_a2, b3 = a.back_a3;
	int4 ib = a.iback[i];
	ib.x = (int) to_fixed(r.x + b1 * (sqx - sqy - sqz + sqw) + b2 * (2. * (xy - zw)) + b3 * (2. * (xz + yw)), a.box_inv[0]);
	ib.y = (int) to_fixed(r.y + b1 * (2. * (xy + zw)) + b2 * (-sqx + sqy - sqz + sqw) + b3 * (2. * (yz - xw)), a.box_inv[1]);
	ib.z = (int) to_fixed(r.z + b1 * (2. * (xz - yw)) + b2 * (2. * (yz + xw)) + b3 * (-sqx - sqy + sqz + sqw), a.box_inv[2]);
	a.iback[i] = ib;
}

} // namespace

namespace oxb {

// `phases` selects one of the instantiated variants; `epoch` is the running launch index inside the batch
void launch_integrate_epoch(cudaStream_t s, const IntegrateArgs &a, int phases, int epoch) {
	switch(phases) {
	case OXB_PH_FIRST: launch_ph<OXB_PH_FIRST>(s, a, epoch); break;
	case OXB_PH_SECOND | OXB_PH_COUNT_STEP: launch_ph<OXB_PH_SECOND | OXB_PH_COUNT_STEP>(s, a, epoch); break;
	case OXB_PH_SECOND | OXB_PH_THERMO | OXB_PH_COUNT_STEP: launch_ph<OXB_PH_SECOND | OXB_PH_THERMO | OXB_PH_COUNT_STEP>(s, a, epoch); break;
	case OXB_PH_SECOND | OXB_PH_FIRST | OXB_PH_COUNT_STEP: launch_ph<OXB_PH_SECOND | OXB_PH_FIRST | OXB_PH_COUNT_STEP>(s, a, epoch); break;
	case OXB_PH_SECOND | OXB_PH_THERMO | OXB_PH_FIRST | OXB_PH_COUNT_STEP:
		launch_ph<OXB_PH_SECOND | OXB_PH_THERMO | OXB_PH_FIRST | OXB_PH_COUNT_STEP>(s, a, epoch);
		break;
	case OXB_PH_SECOND | OXB_PH_BUSSI_SUMS | OXB_PH_COUNT_STEP: launch_ph<OXB_PH_SECOND | OXB_PH_BUSSI_SUMS | OXB_PH_COUNT_STEP>(s, a, epoch); break;
	case OXB_PH_BUSSI_APPLY: launch_ph<OXB_PH_BUSSI_APPLY>(s, a, epoch); break;
	case OXB_PH_BUSSI_APPLY | OXB_PH_FIRST: launch_ph<OXB_PH_BUSSI_APPLY | OXB_PH_FIRST>(s, a, epoch); break;
	case OXB_PH_THERMO: launch_ph<OXB_PH_THERMO>(s, a, epoch); break;
	default: break;
	}
}

void launch_bussi_update_epoch(cudaStream_t s, KinSums *sums, int N, ThermostatCfg th, long long step, const int *flags, int epoch) {
	k_bussi_update<<<1, 32, 0, s>>>(sums, N, th, step, flags, epoch);
}

void launch_clear_sums(cudaStream_t s, KinSums *sums, const int *flags, int epoch) {
	k_clear_sums<<<1, 1, 0, s>>>(sums, flags, epoch);
}

void launch_kinetic_sums(cudaStream_t s, int N, const double4 *veld, const double4 *Ld, KinSums *sums) {
	k_clear_sums<<<1, 1, 0, s>>>(sums, nullptr, -1);
	int tpb = 256;
	k_kinetic_sums<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, veld, Ld, sums);
}

void launch_mol_coms(cudaStream_t s, int N, int n_mol, const int4 *ipos, const int *mol_of, const double *inv_size, const double4 *posd, double *coms) {
	cudaMemsetAsync(coms, 0, sizeof(double) * 3 * (size_t) n_mol, s);
	int tpb = 256;
	k_mol_coms<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, ipos, mol_of, inv_size, posd, coms);
}

void launch_fix_diffusion(cudaStream_t s, int N, const int4 *ipos, const int *mol_of, const double *coms, const double *box, double4 *posd,
		double4 *quatd, float4 *axf, int *shifts) {
	int tpb = 256;
	k_fix_diffusion<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, ipos, mol_of, coms, box[0], box[1], box[2], posd, quatd, axf, shifts);
}

void launch_rescale_positions(cudaStream_t s, const RescaleArgs &a) {
	int tpb = 256;
	k_rescale_positions<<<(a.N + tpb - 1) / tpb, tpb, 0, s>>>(a);
}

void launch_energy_sum(cudaStream_t s, int N, const float4 *F, const float4 *Fb, double *out) {
	cudaMemsetAsync(out, 0, sizeof(double), s);
	int tpb = 256;
	k_energy_sum<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, F, Fb, out);
}

void launch_energy_sum_replicas(cudaStream_t s, int N, int n_rep, int n_per, const float4 *F, const float4 *Fb, double *out) {
	cudaMemsetAsync(out, 0, sizeof(double) * (size_t) n_rep, s);
	int tpb = 256;
	k_energy_sum_replicas<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, n_per, F, Fb, out);
}

} // namespace oxb
