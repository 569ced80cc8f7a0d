// TEST SUPPORT: runs the product's FP32 pair-potential functions (oxdna_b200/csrc/dna_model.cuh, compiled for the
// host by nvcc) over a pair list, so their formulation can be checked against the oracle on a machine without a GPU.
// This is a unit test of device code's arithmetic, not a product code path.
#include "../../oxdna_b200/csrc/models.cuh"

#include <cmath>
#include <vector>

template<class MD>
static void host_forces(const typename MD::Params &M, int N, const double *pos, const double *axes, const int *btype, const int *n3,
		const int *n5, const double *box, const int *pairs, long long npairs, double *F, double *Tlab, double *epart) {
	BoxF b;
	b.lx = (float) box[0]; b.ly = (float) box[1]; b.lz = (float) box[2];
	b.sx = (float) (box[0] / 4294967296.0); b.sy = (float) (box[1] / 4294967296.0); b.sz = (float) (box[2] / 4294967296.0);
	std::vector<int4> ip(N);
	std::vector<Axes> ax(N);
	std::vector<v3> back(N);
	for(int i = 0; i < N; i++) {
		ip[i].x = (int) to_fixed(pos[3 * i], 1. / box[0]);
		ip[i].y = (int) to_fixed(pos[3 * i + 1], 1. / box[1]);
		ip[i].z = (int) to_fixed(pos[3 * i + 2], 1. / box[2]);
		ip[i].w = pack_word(btype[i], i);
		quatd q = quat_from_axes(axes + 9 * i, axes + 9 * i + 3, axes + 9 * i + 6);
		float4 qf = make_float4((float) q.x, (float) q.y, (float) q.z, (float) q.w);
		ax[i] = axes_from_quat(qf);
		back[i] = MD::back(M, ax[i]);
	}
	for(int i = 0; i < 3 * N; i++) F[i] = Tlab[i] = 0.;
	for(int i = 0; i < N; i++) epart[i] = 0.;
	auto scatter = [&](int p, int q, const PairAcc &acc, float e) {
		v3 tp = acc.torque_p(ax[p], back[p]), tq = acc.torque_q(ax[q], back[q]);
		F[3 * p] -= acc.F.x; F[3 * p + 1] -= acc.F.y; F[3 * p + 2] -= acc.F.z;
		F[3 * q] += acc.F.x; F[3 * q + 1] += acc.F.y; F[3 * q + 2] += acc.F.z;
		Tlab[3 * p] += tp.x; Tlab[3 * p + 1] += tp.y; Tlab[3 * p + 2] += tp.z;
		Tlab[3 * q] += tq.x; Tlab[3 * q + 1] += tq.y; Tlab[3 * q + 2] += tq.z;
		epart[p] += 0.5 * e; epart[q] += 0.5 * e;
	};
	for(int p = 0; p < N; p++) {
		int q = n3[p];
		if(q < 0) continue;
		v3 r = min_image_fixed(b, ip[p], ip[q]);
		PairAcc acc; acc.clear();
		bool broken = false;
		float e = MD::bonded(M, r, ax[p], ax[q], btype[p], btype[q], back[p], back[q], acc, broken);
		scatter(p, q, acc, e);
	}
	for(long long k = 0; k < npairs; k++) {
		int p = pairs[2 * k + 1], q = pairs[2 * k];
		v3 r = min_image_fixed(b, ip[p], ip[q]);
		PairAcc acc; acc.clear();
		PairEnergy e = MD::nonbonded(M, r, ax[p], ax[q], btype[p], btype[q], n3[p] < 0 || n5[p] < 0, n3[q] < 0 || n5[q] < 0, back[p], back[q], acc);
		scatter(p, q, acc, e.total);
	}
}

extern "C" void host_dna2_forces(const oxb_dna2_params *M, int N, const double *pos, const double *axes, const int *btype, const int *n3,
		const int *n5, const double *box, const int *pairs, long long npairs, double *F, double *Tlab, double *epart) {
	host_forces<DnaModel>(*M, N, pos, axes, btype, n3, n5, box, pairs, npairs, F, Tlab, epart);
}

extern "C" void host_rna2_forces(const oxb_rna2_params *M, int N, const double *pos, const double *axes, const int *btype, const int *n3,
		const int *n5, const double *box, const int *pairs, long long npairs, double *F, double *Tlab, double *epart) {
	host_forces<RnaModel>(*M, N, pos, axes, btype, n3, n5, box, pairs, npairs, F, Tlab, epart);
}

// ---- oxDNA3: the packed records (dna3_pack.h) + the FP32 device functions of dna3_model.cuh, on the host
#include "../../oxdna_b200/csrc/dna3_model.cuh"
#include "../../oxdna_b200/csrc/dna3_pack.h"

extern "C" void host_dna3_forces(const double *tables, const oxb_dna3_scalars *S, int N, const double *pos, const double *axes, const int *btype,
		const int *n3, const int *n5, const double *box, const int *pairs, long long npairs, double *F, double *Tlab, double *epart, double *esplit8) {
	std::vector<float> h;
	oxb_dna3_dev M;
	size_t off[5];
	dna3_pack(tables, S, h, M, off);
	const float4 *base = reinterpret_cast<const float4 *>(h.data());
	M.bonded = base + off[0]; M.crst = base + off[1]; M.cxst = base + off[2]; M.hb = base + off[3]; M.nexcl = base + off[4];
	BoxF b;
	b.lx = (float) box[0]; b.ly = (float) box[1]; b.lz = (float) box[2];
	b.sx = (float) (box[0] / 4294967296.0); b.sy = (float) (box[1] / 4294967296.0); b.sz = (float) (box[2] / 4294967296.0);
	std::vector<int4> ip(N);
	std::vector<Axes> ax(N);
	std::vector<v3> back(N);
	std::vector<Nuc3> nuc(N);
	for(int i = 0; i < N; i++) {
		ip[i].x = (int) to_fixed(pos[3 * i], 1. / box[0]);
		ip[i].y = (int) to_fixed(pos[3 * i + 1], 1. / box[1]);
		ip[i].z = (int) to_fixed(pos[3 * i + 2], 1. / box[2]);
		ip[i].w = pack_word(btype[i], i);
		quatd q = quat_from_axes(axes + 9 * i, axes + 9 * i + 3, axes + 9 * i + 6);
		ax[i] = axes_from_quat(make_float4((float) q.x, (float) q.y, (float) q.z, (float) q.w));
		back[i] = ax[i].a1 * M.back_a1 + ax[i].a2 * M.back_a2;
		auto ty = [](int b) { return b == 4 ? 4 : btype_to_type(b); };
		const int t3 = n3[i] >= 0 ? ty(btype[n3[i]]) : 5, t5 = n5[i] >= 0 ? ty(btype[n5[i]]) : 5;
		nuc[i] = nuc3_from_code(ty(btype[i]) | (t3 << 3) | (t5 << 6) | ((btype[i] == 4) ? (1 << 9) : 0));
	}
	for(int i = 0; i < 3 * N; i++) F[i] = Tlab[i] = 0.;
	for(int i = 0; i < N; i++) epart[i] = 0.;
	float es[8] = { 0.f };
	auto scatter = [&](int p, int q, const PairAcc &acc, float e) {
		v3 tp = acc.torque_p(ax[p], back[p]), tq = acc.torque_q(ax[q], back[q]);
		F[3 * p] -= acc.F.x; F[3 * p + 1] -= acc.F.y; F[3 * p + 2] -= acc.F.z;
		F[3 * q] += acc.F.x; F[3 * q + 1] += acc.F.y; F[3 * q + 2] += acc.F.z;
		Tlab[3 * p] += tp.x; Tlab[3 * p + 1] += tp.y; Tlab[3 * p + 2] += tp.z;
		Tlab[3 * q] += tq.x; Tlab[3 * q + 1] += tq.y; Tlab[3 * q + 2] += tq.z;
		epart[p] += 0.5 * e; epart[q] += 0.5 * e;
	};
	for(int p = 0; p < N; p++) {
		int q = n3[p];
		if(q < 0) continue;
		const float4 *rec = M.bonded + ix4(nuc[q].n3t, nuc[q].type, nuc[p].type, nuc[p].n5t) * (OXB3_REC_BONDED / 4);
		PairAcc acc; acc.clear();
		bool broken = false;
		float e = dna3_bonded(M, rec, min_image_fixed(b, ip[p], ip[q]), ax[p], ax[q], nuc[p], nuc[q], back[p], back[q], acc, broken, es);
		scatter(p, q, acc, e);
	}
	for(long long k = 0; k < npairs; k++) {
		int p = pairs[2 * k + 1], q = pairs[2 * k];
		PairAcc acc; acc.clear();
		PairEnergy e = dna3_nonbonded(M, min_image_fixed(b, ip[p], ip[q]), ax[p], ax[q], btype[p], btype[q], nuc[p], nuc[q], back[p], back[q], acc, es);
		scatter(p, q, acc, e.total);
	}
	for(int t = 0; t < 8; t++) esplit8[t] = es[t];
}
