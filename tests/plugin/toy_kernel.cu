// TEST INFRASTRUCTURE: the device side of the toy third-party interaction of tests/plugin/CUDAToy.cpp.  It knows nothing about
// oxdna_b200 but the oxb_force_views block (include/oxdna_b200.h): the reference-layout device arrays a plugin's compute_forces() gets.
// Potential: soft repulsion k (rc - r)^2 / 2 between every pair of the Verlet matrix closer than rc.
#include "../../include/oxdna_b200.h"

#include <cuda_runtime.h>

namespace {

__global__ void k_toy(int N, int stride, const float4 *__restrict__ poss, const int *__restrict__ matrix_neighs, const int *__restrict__ number_neighs,
		float4 *__restrict__ forces, float Lx, float Ly, float Lz, float k, float rc) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	const float4 p = poss[i];
	float fx = 0.f, fy = 0.f, fz = 0.f, e = 0.f;
	const int nn = number_neighs[i];
	for(int n = 0; n < nn; n++) {
		const float4 q = poss[matrix_neighs[n * stride + i]];
		float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
		dx -= Lx * rintf(dx / Lx); dy -= Ly * rintf(dy / Ly); dz -= Lz * rintf(dz / Lz);
		const float r = sqrtf(dx * dx + dy * dy + dz * dz);
		if(r < rc) {
			const float mag = k * (rc - r) / r;
			fx -= mag * dx; fy -= mag * dy; fz -= mag * dz;
			e += 0.5f * k * (rc - r) * (rc - r);
		}
	}
	float4 f = forces[i]; // accumulators: zero on entry, add
	forces[i] = make_float4(f.x + fx, f.y + fy, f.z + fz, f.w + e);
}

} // namespace

extern "C" int toy_force_pass(const oxb_force_views *v, float k, float rc) {
	k_toy<<<(v->N + 127) / 128, 128, 0, (cudaStream_t) v->stream>>>(v->N, v->stride, (const float4 *) v->poss, v->matrix_neighs, v->number_neighs,
			(float4 *) v->forces, (float) v->box[0], (float) v->box[1], (float) v->box[2], k, rc);
	return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
