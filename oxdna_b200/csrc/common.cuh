// Shared device helpers for the oxdna_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/oxdna_b200.h"

#define OXB_NO_PARTICLE (-1)
#define OXB_HD __host__ __device__ __forceinline__

// ---- error flags raised by kernels (read back lazily by the host)
enum : int {
	OXB_ERR_NEIGH_OVERFLOW = 1, // a neighbour row exceeded max_neigh
	OXB_ERR_FENE_BROKEN = 2,    // a backbone bond left the FENE range
	OXB_ERR_EDGE_OVERFLOW = 4,
	OXB_ERR_NAN = 8,
	OXB_ERR_SEG_OVERFLOW = 16,  // a work-list segment of the edge pipeline overflowed during a force pass: the pass is incomplete, the
	                            // integrator launch behind it turns into a halt and the host grows the segments and repeats the pass
};

struct v3 {
	float x, y, z;
};

__host__ __device__ __forceinline__ v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ v3 operator-(v3 a) { return mk3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ v3 operator*(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ v3 operator*(float s, v3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ void operator+=(v3 &a, v3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__host__ __device__ __forceinline__ void operator-=(v3 &a, v3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
__host__ __device__ __forceinline__ float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ v3 cross(v3 a, v3 b) {
	return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// y += s * x
__host__ __device__ __forceinline__ void axpy(v3 &y, float s, v3 x) { y.x += s * x.x; y.y += s * x.y; y.z += s * x.z; }

struct Axes {
	v3 a1, a2, a3;
};

// unit quaternion (x, y, z, w) -> body axes; same expansion as the reference (src/CUDA/cuda_utils/CUDA_lr_common.cuh:40-61)
__host__ __device__ __forceinline__ Axes axes_from_quat(float4 q) {
	float sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
	float xy = q.x * q.y, xz = q.x * q.z, xw = q.x * q.w, yz = q.y * q.z, yw = q.y * q.w, zw = q.z * q.w;
	Axes A;
	A.a1 = mk3(sqx - sqy - sqz + sqw, 2.f * (xy + zw), 2.f * (xz - yw));
	A.a2 = mk3(2.f * (xy - zw), -sqx + sqy - sqz + sqw, 2.f * (yz + xw));
	A.a3 = mk3(2.f * (xz + yw), 2.f * (yz - xw), -sqx - sqy + sqz + sqw);
	return A;
}

// ---- FP32 orientation record of a particle: its a1 and a3 axes in 32 bytes = ONE sector,
//   axf[2 i] = (a1.x, a1.y, a1.z, a3.x), axf[2 i + 1] = (a3.y, a3.z, 0, 0);  a2 = a3 x a1.
// The integrator writes it from the FP64 quaternion (it forms these products anyway for the backbone site); every pair kernel reads it
// with two 128-bit loads of the same sector instead of gathering a quaternion (half a sector) and expanding it (~35 instructions per
// particle and pair: the pair kernels are instruction-issue bound).
OXB_HD void store_axes(float4 *axf, int i, double a1x, double a1y, double a1z, double a3x, double a3y, double a3z) {
	axf[2 * (size_t) i] = make_float4((float) a1x, (float) a1y, (float) a1z, (float) a3x);
	axf[2 * (size_t) i + 1] = make_float4((float) a3y, (float) a3z, 0.f, 0.f);
}
OXB_HD void store_axes_from_quatd(float4 *axf, int i, double x, double y, double z, double w) {
	store_axes(axf, i, x * x - y * y - z * z + w * w, 2. * (x * y + z * w), 2. * (x * z - y * w), 2. * (x * z + y * w), 2. * (y * z - x * w), -x * x - y * y + z * z + w * w);
}
#ifdef __CUDACC__
__device__ __forceinline__ Axes load_axes(const float4 *__restrict__ axf, int i) {
	const float4 u = __ldg(axf + 2 * (size_t) i), w = __ldg(axf + 2 * (size_t) i + 1);
	Axes A;
	A.a1 = mk3(u.x, u.y, u.z);
	A.a3 = mk3(u.w, w.x, w.y);
	A.a2 = cross(A.a3, A.a1);
	return A;
}
__device__ __forceinline__ v3 load_a1(const float4 *__restrict__ axf, int i) {
	const float4 u = __ldg(axf + 2 * (size_t) i);
	return mk3(u.x, u.y, u.z);
}
#endif

// ---- classes of a near edge: which families of site pairs can come into range before the next list rebuild (decided by the list builder
// with the same 2 x skin padding that selects the near edges; the near-edge kernel evaluates only the flagged families).  In the edge list
// and -- half-shell builds only, where the matrix is private to the edge pipeline -- in the Verlet matrix entry the class sits above the
// 22-bit slot index.
enum : int { OXB_CLS_BB = 1, OXB_CLS_EB = 2, OXB_CLS_HBCR = 4, OXB_CLS_ST = 8, OXB_CLS_BK = 16, OXB_CLS_ALL = 31, OXB_CLS_SHIFT = 24, OXB_SLOT_MASK = 0x003FFFFF };

// The edge list is laid out in three segments by class GROUP, each grouped by `from` in slot order: 0 = hydrogen bonding / cross stacking
// only, 1 = + coaxial stacking, 2 = any excluded-volume family.  Warps of the near-edge kernel are then uniform in which families they test
// (with one list every warp holds some edge of every class and executes all of the screening code with a part of its lanes).
OXB_HD int cls_group(int cls) { return (cls & (OXB_CLS_BB | OXB_CLS_EB | OXB_CLS_BK)) ? 2 : ((cls & OXB_CLS_ST) ? 1 : 0); }

// ---- packed particle word: (btype << 22) | original index, as in the reference (MD_CUDABackend.cu:243-254)
__host__ __device__ __forceinline__ int pack_word(int btype, int index) { return (btype << 22) | (index & 0x003FFFFF); }
__host__ __device__ __forceinline__ int word_btype(int w) { return w >> 22; }
__host__ __device__ __forceinline__ int word_index(int w) { return w & 0x003FFFFF; }
// base type 0..3 (A, G, C, T) from a possibly "special" btype (src/Interactions/DNAInteraction.cpp:1543)
// (the dummy base 'D' of a topology has btype = type = 4: src/Utilities/TopologyParser.cpp:96-99, Utils.cpp:33-34; numeric bases have two digits or a sign)
__host__ __device__ __forceinline__ int btype_to_type(int b) { return (b == 4) ? 4 : ((b < 0) ? 3 - ((3 - b) % 4) : b % 4); }

// ---- periodic box in fixed point: a coordinate x is stored as u = frac(x / L) * 2^32, so that the
// minimum-image separation is the wrapped 32-bit difference (exact, branch-free) times L / 2^32.
struct BoxF {
	float sx, sy, sz;      // L / 2^32
	float lx, ly, lz;      // L
	double dsx, dsy, dsz;  // L / 2^32 in double (the FENE distance is taken in double from the fixed-point backbone sites)
};

OXB_HD v3 min_image_fixed(const BoxF &b, int4 p, int4 q) {
	return mk3((float) (int) ((unsigned) q.x - (unsigned) p.x) * b.sx, (float) (int) ((unsigned) q.y - (unsigned) p.y) * b.sy,
			(float) (int) ((unsigned) q.z - (unsigned) p.z) * b.sz);
}

__host__ __device__ __forceinline__ unsigned to_fixed(double x, double invL) {
	double f = x * invL;
	f -= floor(f);
	// round to nearest grid point (half the error of truncation; 2^32 wraps to 0, which is the same point)
	unsigned long long u = (unsigned long long) (f * 4294967296.0 + 0.5);
	return (unsigned) (u & 0xFFFFFFFFull);
}

// ---- staleness references of the Verlet lists.  The list builder records where the centre, the backbone site and the base site of
// every particle were; the integrator compares against them every step.  They live in the otherwise unused .w lanes of the FP64
// position / velocity / angular-momentum elements (which the integrator streams anyway: the three int4 arrays they replace cost it
// 48 B of reads per particle-step) as 3 x 21-bit fixed point, rounded to nearest: resolution L / 2^21, error <= L / 2^22 per axis,
// which the integrator's threshold is tightened by.
OXB_HD double pack_ref(int4 p) {
	const unsigned long long x = (((unsigned) p.x + 0x400u) >> 11) & 0x1FFFFFu, y = (((unsigned) p.y + 0x400u) >> 11) & 0x1FFFFFu, z = (((unsigned) p.z + 0x400u) >> 11) & 0x1FFFFFu;
	const unsigned long long w = x | (y << 21) | (z << 42);
#ifdef __CUDA_ARCH__
	return __longlong_as_double((long long) w);
#else
	double d; memcpy(&d, &w, 8); return d;
#endif
}
#ifdef __CUDACC__
__device__ __forceinline__ int4 unpack_ref(double d) {
	const unsigned long long w = (unsigned long long) __double_as_longlong(d);
	return make_int4((int) ((unsigned) (w & 0x1FFFFFull) << 11), (int) ((unsigned) ((w >> 21) & 0x1FFFFFull) << 11), (int) ((unsigned) ((w >> 42) & 0x1FFFFFull) << 11), 0);
}
#endif

// ---- Philox4x32-10 counter-based RNG (Salmon et al., SC'11).  Stateless: stream = (seed, original particle id),
// counter = (step, draw index), so the Hilbert re-sort and temperature changes never touch RNG state.
struct Philox {
	__host__ __device__ static inline uint4 round4(uint4 c, uint2 k) {
		const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
		unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
		unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
#else
		unsigned long long p0 = (unsigned long long) M0 * c.x, p1 = (unsigned long long) M1 * c.z;
		unsigned hi0 = (unsigned) (p0 >> 32), lo0 = (unsigned) p0, hi1 = (unsigned) (p1 >> 32), lo1 = (unsigned) p1;
#endif
		uint4 r;
		r.x = hi1 ^ c.y ^ k.x;
		r.y = lo1;
		r.z = hi0 ^ c.w ^ k.y;
		r.w = lo0;
		return r;
	}
	__host__ __device__ static inline uint4 gen(uint4 ctr, uint2 key) {
		const unsigned W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
		for(int i = 0; i < 10; i++) {
			ctr = round4(ctr, key);
			key.x += W0;
			key.y += W1;
		}
		return ctr;
	}
};

// uniform in (0, 1]
__host__ __device__ __forceinline__ float u01(unsigned x) { return ((float) (x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// rotation matrix with columns (a1, a2, a3) -> unit quaternion (x, y, z, w), double precision; branch on the largest
// diagonal element as the reference's host marshalling does (src/CUDA/Backends/MD_CUDABackend.cu:275-307)
struct quatd { double x, y, z, w; };
__host__ __device__ inline quatd quat_from_axes(const double *a1, const double *a2, const double *a3) {
	// m[r][c] = component r of axis c
	double m00 = a1[0], m10 = a1[1], m20 = a1[2];
	double m01 = a2[0], m11 = a2[1], m21 = a2[2];
	double m02 = a3[0], m12 = a3[1], m22 = a3[2];
	quatd q;
	double tr = m00 + m11 + m22;
	if(tr > 0) {
		double s = 0.5 / sqrt(tr + 1.0);
		q.w = 0.25 / s; q.x = (m21 - m12) * s; q.y = (m02 - m20) * s; q.z = (m10 - m01) * s;
	}
	else if(m00 > m11 && m00 > m22) {
		double s = 0.5 / sqrt(1.0 + m00 - m11 - m22);
		q.w = (m21 - m12) * s; q.x = 0.25 / s; q.y = (m01 + m10) * s; q.z = (m02 + m20) * s;
	}
	else if(m11 > m22) {
		double s = 0.5 / sqrt(1.0 + m11 - m00 - m22);
		q.w = (m02 - m20) * s; q.x = (m01 + m10) * s; q.y = 0.25 / s; q.z = (m12 + m21) * s;
	}
	else {
		double s = 0.5 / sqrt(1.0 + m22 - m00 - m11);
		q.w = (m10 - m01) * s; q.x = (m02 + m20) * s; q.y = (m12 + m21) * s; q.z = 0.25 / s;
	}
	double n = 1.0 / sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
	q.x *= n; q.y *= n; q.z *= n; q.w *= n;
	return q;
}

// double-precision axes from a double quaternion
__host__ __device__ inline void axes_from_quatd(quatd q, double *a1, double *a2, double *a3) {
	double sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
	double xy = q.x * q.y, xz = q.x * q.z, xw = q.x * q.w, yz = q.y * q.z, yw = q.y * q.w, zw = q.z * q.w;
	a1[0] = sqx - sqy - sqz + sqw; a1[1] = 2 * (xy + zw); a1[2] = 2 * (xz - yw);
	a2[0] = 2 * (xy - zw); a2[1] = -sqx + sqy - sqz + sqw; a2[2] = 2 * (yz + xw);
	a3[0] = 2 * (xz + yw); a3[1] = 2 * (yz - xw); a3[2] = -sqx - sqy + sqz + sqw;
}
