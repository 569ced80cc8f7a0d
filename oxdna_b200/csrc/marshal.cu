// Host <-> device marshalling of the particle state on the device: replaces the serial host loops of
// MD_CUDABackend::apply_changes_to_simulation_data / apply_simulation_data_changes (src/CUDA/Backends/MD_CUDABackend.cu:231-394,
// mixed: MD_CUDAMixedBackend.cu:115-138).  The host side only moves flat N x 3 double arrays; orthonormalisation, quaternion
// conversion, fixed-point encoding, index packing and the inverse (slot order -> original order, quaternion -> a1/a3) run here.
#include "kernels.h"

namespace {

using namespace oxb;

__global__ void k_state_in(MarshalArgs a) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= a.N) return;
	const double *p = a.pos + 3 * (size_t) i;
	a.posd[i] = make_double4(p[0], p[1], p[2], 0.);
	a.veld[i] = a.vel ? make_double4(a.vel[3 * (size_t) i], a.vel[3 * (size_t) i + 1], a.vel[3 * (size_t) i + 2], 0.) : make_double4(0., 0., 0., 0.);
	a.Ld[i] = a.L ? make_double4(a.L[3 * (size_t) i], a.L[3 * (size_t) i + 1], a.L[3 * (size_t) i + 2], 0.) : make_double4(0., 0., 0., 0.);
	// orthonormalise exactly like the reference's configuration reader (src/Backends/SimBackend.cpp:623-629)
	double v1[3] = { a.a1[3 * (size_t) i], a.a1[3 * (size_t) i + 1], a.a1[3 * (size_t) i + 2] };
	double v3_[3] = { a.a3[3 * (size_t) i], a.a3[3 * (size_t) i + 1], a.a3[3 * (size_t) i + 2] }, v2[3];
	double n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]), n3 = sqrt(v3_[0] * v3_[0] + v3_[1] * v3_[1] + v3_[2] * v3_[2]);
	if(n1 == 0. || n3 == 0.) {
		atomicMin(a.err, i); // a null vector is an error in the reader as well
		n1 = n3 = 1.;
		v1[0] = 1.; v3_[2] = 1.;
	}
	for(int k = 0; k < 3; k++) { v1[k] /= n1; v3_[k] /= n3; }
	double d = v1[0] * v3_[0] + v1[1] * v3_[1] + v1[2] * v3_[2];
	for(int k = 0; k < 3; k++) v1[k] -= v3_[k] * d;
	n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
	for(int k = 0; k < 3; k++) v1[k] /= n1;
	v2[0] = v3_[1] * v1[2] - v3_[2] * v1[1]; v2[1] = v3_[2] * v1[0] - v3_[0] * v1[2]; v2[2] = v3_[0] * v1[1] - v3_[1] * v1[0];
	double n2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
	for(int k = 0; k < 3; k++) v2[k] /= n2;
	quatd q = quat_from_axes(v1, v2, v3_);
	a.quatd[i] = make_double4(q.x, q.y, q.z, q.w);
	store_axes_from_quatd(a.axf, i, q.x, q.y, q.z, q.w); // from the quaternion, exactly as the integrator derives it
	const int4 t = a.topo[i]; // btype, n3, n5, strand
	int4 ip;
	ip.x = (int) to_fixed(p[0], a.box_inv[0]); ip.y = (int) to_fixed(p[1], a.box_inv[1]); ip.z = (int) to_fixed(p[2], a.box_inv[2]);
	ip.w = pack_word(t.x, i);
	a.ipos[i] = ip;
	// backbone site (grooved): r + back_a1 a1 + back_a2 a2 (+ back_a3 a3 for RNA), from the same float-rounded constants the kernels use
	double b1 = a.back_a1, b2 = a.back_a2, b3 = a.back_a3;
	int4 ib;
	ib.x = (int) to_fixed(p[0] + b1 * v1[0] + b2 * v2[0] + b3 * v3_[0], a.box_inv[0]);
	ib.y = (int) to_fixed(p[1] + b1 * v1[1] + b2 * v2[1] + b3 * v3_[1], a.box_inv[1]);
	ib.z = (int) to_fixed(p[2] + b1 * v1[2] + b2 * v2[2] + b3 * v3_[2], a.box_inv[2]);
	ib.w = (t.y < 0 || t.z < 0) ? 1 : 0;
	a.iback[i] = ib;
	a.bonds[i] = make_int2(t.y, t.z);
	a.slot_of[i] = i;
}

// slot order -> original order; any output may be null
__global__ void k_state_out(int N, const int4 *__restrict__ ipos, const double4 *__restrict__ posd, const double4 *__restrict__ veld,
		const double4 *__restrict__ Ld, const double4 *__restrict__ qd, double *pos, double *a1, double *a3, double *vel, double *L) {
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= N) return;
	const size_t i = (size_t) word_index(ipos[s].w);
	if(pos) { double4 r = posd[s]; pos[3 * i] = r.x; pos[3 * i + 1] = r.y; pos[3 * i + 2] = r.z; }
	if(vel) { double4 v = veld[s]; vel[3 * i] = v.x; vel[3 * i + 1] = v.y; vel[3 * i + 2] = v.z; }
	if(L) { double4 l = Ld[s]; L[3 * i] = l.x; L[3 * i + 1] = l.y; L[3 * i + 2] = l.z; }
	if(a1 || a3) {
		double4 qq = qd[s];
		quatd q = { qq.x, qq.y, qq.z, qq.w };
		double x1[3], x2[3], x3[3];
		axes_from_quatd(q, x1, x2, x3);
		for(int d = 0; d < 3; d++) {
			if(a1) a1[3 * i + d] = x1[d];
			if(a3) a3[3 * i + d] = x3[d];
		}
	}
}

} // namespace

namespace oxb {

void launch_state_in(cudaStream_t s, const MarshalArgs &a) {
	int tpb = 256;
	k_state_in<<<(a.N + tpb - 1) / tpb, tpb, 0, s>>>(a);
}

void launch_state_out(cudaStream_t s, int N, const int4 *ipos, const double4 *posd, const double4 *veld, const double4 *Ld, const double4 *quatd,
		double *pos, double *a1, double *a3, double *vel, double *L) {
	int tpb = 256;
	k_state_out<<<(N + tpb - 1) / tpb, tpb, 0, s>>>(N, ipos, posd, veld, Ld, quatd, pos, a1, a3, vel, L);
}

} // namespace oxb
