// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product path.
//
// A thin C-ABI driver around the UNMODIFIED reference CPU implementation (compiled from
// /root/reference/src by oracle/Makefile.ref into oracle/_ref/liboxdna_ref.a).  It exposes what the
// reference's own `oxpy` bindings do not: full-precision forces/torques, the CPU Verlet pair set
// (src/Lists/VerletList.cpp:35-66, src/Lists/Cells.cpp:120-181), per-term energies
// (src/Interactions/BaseInteraction.cpp:61-90) and single MD steps (src/Backends/MD_CPUBackend.cpp:189-218).
// All code in this file is ours; it only *calls* the reference's public classes.
//
// Used by: tests/ (checker), oracle/make_golden.py (fixture generator), bench.py --impl reference.

#include <Managers/SimManager.h>
#include <Backends/SimBackend.h>
#include <Utilities/ConfigInfo.h>
#include <Utilities/Logger.h>
#include <Utilities/Timings.h>
#include <Utilities/oxDNAException.h>
#include <Interactions/BaseInteraction.h>
#include <Interactions/DNA3Interaction.h>
#include <Lists/BaseList.h>
#include <Boxes/BaseBox.h>
#include <Particles/BaseParticle.h>

#include <cstring>
#include <memory>
#include <sstream>
#include <string>

namespace {

struct Harness: public SimManager {
	explicit Harness(input_file inp) : SimManager(inp) {
	}
	void setup() {
		SimManager::load_options();
		SimManager::init();
	}
	SimBackend *backend() { return _backend.get(); }
	~Harness() override {
		SimManager::stop = false;
		SimManager::started = false;
	}
};

// raw pointer, deliberately leaked at process exit: the reference's singletons (Logger, ConfigInfo)
// are destroyed before a static unique_ptr would be, and ~SimBackend uses them.
Harness *g_h = nullptr;
std::string g_err;

LR_matrix matrix_from_axes(const double *a1, const double *a3) {
	LR_vector v1(a1[0], a1[1], a1[2]);
	LR_vector v3(a3[0], a3[1], a3[2]);
	v1.normalize();
	v3.normalize();
	LR_vector v2 = v3.cross(v1);
	// columns of `orientation` are the body axes (src/Backends/SimBackend.cpp, conf reader)
	return LR_matrix(v1.x, v2.x, v3.x, v1.y, v2.y, v3.y, v1.z, v2.z, v3.z);
}

} // namespace

extern "C" {

const char *oxref_last_error() {
	return g_err.c_str();
}

// `overrides` = newline-separated key = value lines appended to the input file.
int oxref_open(const char *input_path, const char *overrides) {
	try {
		if(g_h) {
			delete g_h;
			g_h = nullptr;
			TimingManager::clear();
			ConfigInfo::clear();
		}
		try {
			Logger::init();
		}
		catch(oxDNAException &e) {
			// the Logger may be initialised only once per process
		}
		TimingManager::init();
		input_file inp;
		inp.init_from_filename(input_path);
		if(inp.state == ERROR) {
			g_err = "cannot open input file";
			return -1;
		}
		inp.show_overwrite_warnings = false;
		if(overrides != nullptr && overrides[0] != '\0') {
			inp.add_input_source(std::string(overrides));
		}
		g_h = new Harness(inp);
		g_h->setup();
		return 0;
	}
	catch(oxDNAException &e) {
		g_err = e.what();
		return -2;
	}
	catch(std::exception &e) {
		g_err = e.what();
		return -3;
	}
}

void oxref_close() {
	if(g_h) {
		delete g_h;
		g_h = nullptr;
		TimingManager::clear();
		ConfigInfo::clear();
	}
}

int oxref_N() {
	return CONFIG_INFO->N();
}

void oxref_box(double *sides) {
	LR_vector s = CONFIG_INFO->box->box_sides();
	sides[0] = s.x;
	sides[1] = s.y;
	sides[2] = s.z;
}

double oxref_rcut() {
	return CONFIG_INFO->interaction->get_rcut();
}

double oxref_temperature() {
	return CONFIG_INFO->temperature();
}

long long oxref_current_step() {
	return CONFIG_INFO->curr_step;
}

void oxref_get_topology(int *btype, int *type, int *n3, int *n5, int *strand) {
	for(auto p : CONFIG_INFO->particles()) {
		int i = p->index;
		btype[i] = p->btype;
		type[i] = p->type;
		n3[i] = (p->n3 == P_VIRTUAL) ? -1 : p->n3->index;
		n5[i] = (p->n5 == P_VIRTUAL) ? -1 : p->n5->index;
		strand[i] = p->strand_id;
	}
}

void oxref_get_state(double *pos, double *a1, double *a3, double *vel, double *L) {
	for(auto p : CONFIG_INFO->particles()) {
		int i = p->index;
		for(int k = 0; k < 3; k++) {
			pos[3 * i + k] = p->pos[k];
			a1[3 * i + k] = p->orientationT.v1[k];
			a3[3 * i + k] = p->orientationT.v3[k];
			vel[3 * i + k] = p->vel[k];
			L[3 * i + k] = p->L[k];
		}
	}
}

// full 3x3 orientation (row-major) for tests that must not lose the a2 rounding
void oxref_get_orientation(double *m) {
	for(auto p : CONFIG_INFO->particles()) {
		int i = p->index;
		const LR_matrix &o = p->orientation;
		double v[9] = { o.v1.x, o.v1.y, o.v1.z, o.v2.x, o.v2.y, o.v2.z, o.v3.x, o.v3.y, o.v3.z };
		std::memcpy(m + 9 * i, v, sizeof(v));
	}
}

void oxref_set_state(const double *pos, const double *a1, const double *a3, const double *vel, const double *L) {
	for(auto p : CONFIG_INFO->particles()) {
		int i = p->index;
		p->pos = LR_vector(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		p->orientation = matrix_from_axes(a1 + 3 * i, a3 + 3 * i);
		p->orientationT = p->orientation.get_transpose();
		p->set_positions();
		if(vel != nullptr) p->vel = LR_vector(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
		if(L != nullptr) p->L = LR_vector(L[3 * i], L[3 * i + 1], L[3 * i + 2]);
	}
	CONFIG_INFO->lists->global_update(true);
}

// Same loop as MD_CPUBackend::_compute_forces (src/Backends/MD_CPUBackend.cpp:146-167), driven from outside
// because that method is protected.  Returns the potential energy.
double oxref_compute_forces() {
	BaseInteraction *inter = CONFIG_INFO->interaction;
	BaseList *lists = CONFIG_INFO->lists;
	inter->begin_energy_and_force_computation();
	for(auto p : CONFIG_INFO->particles()) {
		p->set_initial_forces(CONFIG_INFO->curr_step, CONFIG_INFO->box);
	}
	double U = 0.;
	for(auto p : CONFIG_INFO->particles()) {
		for(auto &pair : p->affected) {
			if(pair.first == p) {
				U += inter->pair_interaction_bonded(pair.first, pair.second, true, true);
			}
		}
		for(auto q : lists->get_neigh_list(p)) {
			U += inter->pair_interaction_nonbonded(p, q, true, true);
		}
	}
	return U;
}

// force: lab frame; torque_body: body frame (as stored by the reference); torque_lab = orientation * torque_body
void oxref_get_forces(double *force, double *torque_body, double *torque_lab) {
	for(auto p : CONFIG_INFO->particles()) {
		int i = p->index;
		LR_vector tl = p->orientation * p->torque;
		for(int k = 0; k < 3; k++) {
			force[3 * i + k] = p->force[k];
			if(torque_body != nullptr) torque_body[3 * i + k] = p->torque[k];
			if(torque_lab != nullptr) torque_lab[3 * i + k] = tl[k];
		}
	}
}

// Per-term total energies, in the order of the interaction map (for DNA2: FENE, BEXC, STCK, NEXC, HB, CRSTCK, CXSTCK, DH).
int oxref_energy_split(double *out, int max_terms) {
	auto m = CONFIG_INFO->interaction->get_system_energy_split(CONFIG_INFO->particles(), CONFIG_INFO->lists);
	int n = 0;
	for(auto &kv : m) {
		if(n < max_terms) out[n] = kv.second;
		n++;
	}
	return n;
}

double oxref_system_energy() {
	return CONFIG_INFO->interaction->get_system_energy(CONFIG_INFO->particles(), CONFIG_INFO->lists);
}

// Verlet pairs as stored by the CPU list: each pair once, on the higher-index particle (p > q).
long long oxref_get_pairs(int *pairs, long long max_pairs) {
	long long n = 0;
	for(auto p : CONFIG_INFO->particles()) {
		for(auto q : CONFIG_INFO->lists->get_neigh_list(p)) {
			if(n < max_pairs) {
				pairs[2 * n] = q->index;
				pairs[2 * n + 1] = p->index;
			}
			n++;
		}
	}
	return n;
}

void oxref_rebuild_lists() {
	CONFIG_INFO->lists->global_update(true);
}

int oxref_step(long long n) {
	try {
		SimBackend *b = g_h->backend();
		for(long long i = 0; i < n; i++) {
			b->sim_step();
			b->increment_current_step();
		}
		return 0;
	}
	catch(oxDNAException &e) {
		g_err = e.what();
		return -2;
	}
}

int oxref_N_updates() {
	return g_h->backend()->get_N_updates();
}

void oxref_update_temperature(double T) {
	CONFIG_INFO->update_temperature(T);
}

// ---- oxDNA3: the tetramer-indexed parameter tables of the live DNA3Interaction, in the table order of include/oxdna_b200.h
// (OXB_DNA3_*), i.e. what CUDADNA3Interaction::cuda_init uploads (src/CUDA/Interactions/CUDADNA3Interaction.cu:46-150).
namespace {
typedef MultiDimArray<TETRAMER_DIM_A, TETRAMER_DIM_B, TETRAMER_DIM_B, TETRAMER_DIM_A> Tab;
struct Dna3Peek: public DNA3Interaction {
	// pointers to protected members may be formed inside a derived class
	static Tab &r0(DNA3Interaction &d) { return d.*(&Dna3Peek::_fene_r0_SD); }
	static Tab &delta(DNA3Interaction &d) { return d.*(&Dna3Peek::_fene_delta_SD); }
	static Tab &delta2(DNA3Interaction &d) { return d.*(&Dna3Peek::_fene_delta2_SD); }
	static Tab &xmax(DNA3Interaction &d) { return d.*(&Dna3Peek::_mbf_xmax_SD); }
	static double fene_eps(DNA3Interaction &d) { return d.*(&Dna3Peek::_fene_eps); }
	static double hb_multi(DNA3Interaction &d) { return d.*(&Dna3Peek::_hb_multiplier); }
	static bool use_mbf(DNA3Interaction &d) { return d.*(&Dna3Peek::_use_mbf); }
	static double mbf_fmax(DNA3Interaction &d) { return d.*(&Dna3Peek::_mbf_fmax); }
	static double mbf_finf(DNA3Interaction &d) { return d.*(&Dna3Peek::_mbf_finf); }
	static double dh_rc(DNA3Interaction &d) { return d.*(&Dna3Peek::_debye_huckel_RC); }
	static double dh_rhigh(DNA3Interaction &d) { return d.*(&Dna3Peek::_debye_huckel_RHIGH); }
	static double dh_pref(DNA3Interaction &d) { return d.*(&Dna3Peek::_debye_huckel_prefactor); }
	static double dh_b(DNA3Interaction &d) { return d.*(&Dna3Peek::_debye_huckel_B); }
	static double dh_mk(DNA3Interaction &d) { return d.*(&Dna3Peek::_minus_kappa); }
	static bool dh_half(DNA3Interaction &d) { return d.*(&Dna3Peek::_debye_huckel_half_charged_ends); }
};
void put(double *&o, const Tab *t, int n) {
	for(int i = 0; i < n; i++) { std::memcpy(o, t[i].data, sizeof(double) * Tab::total_size); o += Tab::total_size; }
}
}

// tables: 215 x 900 doubles; scalars: 40 doubles (see oracle/oracle.py: DNA3_SCALARS).  Returns 0, or -1 if the interaction is not oxDNA3.
int oxref_dna3_tables(double *tables, double *scalars) {
	DNA3Interaction *d = dynamic_cast<DNA3Interaction *>(CONFIG_INFO->interaction);
	if(d == nullptr) return -1;
	double *o = tables;
	put(o, &Dna3Peek::r0(*d), 1); put(o, &Dna3Peek::delta(*d), 1); put(o, &Dna3Peek::delta2(*d), 1); put(o, &Dna3Peek::xmax(*d), 1);
	put(o, d->_excl_s, 7); put(o, d->_excl_r, 7); put(o, d->_excl_b, 7); put(o, d->_excl_rc, 7);
	put(o, d->F1_SD_EPS, 2); put(o, d->F1_SD_A, 2); put(o, d->F1_SD_RC, 2); put(o, d->F1_SD_R0, 2); put(o, d->F1_SD_BLOW, 2); put(o, d->F1_SD_BHIGH, 2);
	put(o, d->F1_SD_RLOW, 2); put(o, d->F1_SD_RHIGH, 2); put(o, d->F1_SD_RCLOW, 2); put(o, d->F1_SD_RCHIGH, 2); put(o, d->F1_SD_SHIFT, 2);
	put(o, d->F2_SD_K, 4); put(o, d->F2_SD_K_SYMM, 4); put(o, d->F2_SD_RC, 4); put(o, d->F2_SD_R0, 4); put(o, d->F2_SD_BLOW, 4); put(o, d->F2_SD_RLOW, 4);
	put(o, d->F2_SD_RCLOW, 4); put(o, d->F2_SD_BHIGH, 4); put(o, d->F2_SD_RCHIGH, 4); put(o, d->F2_SD_RHIGH, 4);
	put(o, d->F4_SD_THETA_A, 21); put(o, d->F4_SD_THETA_B, 21); put(o, d->F4_SD_THETA_T0, 21); put(o, d->F4_SD_THETA_TS, 21); put(o, d->F4_SD_THETA_TC, 21);
	put(o, d->F5_SD_PHI_A, 4); put(o, d->F5_SD_PHI_B, 4); put(o, d->F5_SD_PHI_XC, 4); put(o, d->F5_SD_PHI_XS, 4);
	double *s = scalars;
	int k = 0;
	s[k++] = Dna3Peek::fene_eps(*d); s[k++] = Dna3Peek::use_mbf(*d) ? 1. : 0.; s[k++] = Dna3Peek::mbf_fmax(*d); s[k++] = Dna3Peek::mbf_finf(*d);
	s[k++] = Dna3Peek::hb_multi(*d);
	s[k++] = Dna3Peek::dh_rc(*d); s[k++] = Dna3Peek::dh_rhigh(*d); s[k++] = Dna3Peek::dh_pref(*d); s[k++] = Dna3Peek::dh_b(*d); s[k++] = Dna3Peek::dh_mk(*d);
	s[k++] = Dna3Peek::dh_half(*d) ? 1. : 0.;
	s[k++] = d->get_rcut();
	// the scalar f4 set of the coaxial stacking (DNA2Interaction members): theta1 (+ pure harmonic), theta4, theta5 = theta6
	const int ids[3] = { CXST_F4_THETA1, CXST_F4_THETA4, CXST_F4_THETA5 };
	for(int i = 0; i < 3; i++) {
		s[k++] = d->F4_THETA_A[ids[i]]; s[k++] = d->F4_THETA_B[ids[i]]; s[k++] = d->F4_THETA_T0[ids[i]]; s[k++] = d->F4_THETA_TS[ids[i]]; s[k++] = d->F4_THETA_TC[ids[i]];
	}
	s[k++] = d->F4_THETA_SA[CXST_F4_THETA1]; s[k++] = d->F4_THETA_SB[CXST_F4_THETA1];
	return k;
}

} // extern "C"
