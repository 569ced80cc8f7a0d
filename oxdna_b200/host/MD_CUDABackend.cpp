// See MD_CUDABackend.h.  Call order and input keys follow src/CUDA/Backends/CUDABaseBackend.cu:110-242 and
// src/CUDA/Backends/MD_CUDABackend.cu:231-394,567-747; the device work goes through the C ABI (include/oxdna_b200.h).
#include "MD_CUDABackend.h"

#include "Forces/AttractionPlane.h"
#include "Forces/COMForce.h"
#include "Forces/GenericCentralForce.h"
#include "Forces/LJCone.h"
#include "Forces/Metadynamics/LTCOMTrap.h"
#include "Forces/Metadynamics/LTCoordination.h"
#include "Forces/RepulsionPlaneMoving.h"
#include "Forces/RepulsiveSphereMoving.h"
#include "Forces/YukawaSphere.h"
#include "Forces/ConstantRateTorque.h"
#include "Forces/RepulsiveEllipsoid.h"
#include "Forces/RepulsiveSphereSmooth.h"
#include "Forces/LJWall.h"
#include "Forces/LowdimMovingTrap.h"
#include "Forces/RepulsionPlane.h"
#include "Forces/RepulsiveSphere.h"

#include <cmath>
#include <map>
#include <set>

#include "Forces/ConstantRateForce.h"
#include "Forces/MovingTrap.h"
#include "Forces/MutualTrap.h"
#include "Lists/BaseList.h"
#include "Observables/ObservableOutput.h"
#include "Particles/BaseParticle.h"
#include "Utilities/ConfigInfo.h"
#include "Utilities/Logger.h"
#include "Utilities/Utils.h"

#include <typeinfo>
#include <vector>

// `type = total_energy` served from the device: same three columns as src/Observables/TotalEnergy.cpp:33-38
// (U, K, U + K per particle, BaseObservable's default "%10.6lf" formatter), no particle data needed on the CPU
class CUDATotalEnergy: public BaseObservable {
	const double *_UK; // potential, kinetic (totals), owned by the backend
	int _N;
public:
	CUDATotalEnergy(const double *UK, int N) : _UK(UK), _N(N) { }
	bool require_data_on_CPU() override { return false; }
	std::string get_output_string(llint curr_step) override {
		return Utils::sformat("%10.6lf %10.6lf %10.6lf", _UK[0] / _N, _UK[1] / _N, (_UK[0] + _UK[1]) / _N);
	}
};

MD_CUDABackend::MD_CUDABackend() :
				MDBackend() {
}

MD_CUDABackend::~MD_CUDABackend() {
	if(_ctx != nullptr) oxb_destroy(_ctx);
}

void MD_CUDABackend::get_settings(input_file &inp) {
	MDBackend::get_settings(inp);

	if(getInputInt(&inp, "CUDA_device", &_device_number, 0) == KEY_NOT_FOUND) {
		OX_LOG(Logger::LOG_INFO, "CUDA device not specified");
		_device_number = -1;
	}
	else {
		OX_LOG(Logger::LOG_INFO, "Using CUDA device %d", _device_number);
	}
	if(getInputInt(&inp, "CUDA_sort_every", &_sort_every, 0) == KEY_NOT_FOUND) {
		OX_LOG(Logger::LOG_INFO, "CUDA sort_every not specified, using 0");
	}
	getInputInt(&inp, "threads_per_block", &_threads_per_block, 0);

	_cuda_interaction = CUDAInteractionFactory::make_interaction(inp);
	_cuda_interaction->get_settings(inp);
	_cuda_interaction->get_cuda_settings(inp);

	_cuda_lists = CUDAListFactory::make_list(inp);
	_cuda_lists->get_settings(inp);

	std::string reload_from;
	if(getInputString(&inp, "reload_from", reload_from, 0) == KEY_FOUND) {
		throw oxDNAException("The CUDA backend does not support reloading checkpoints, owing to its intrinsically stochastic nature");
	}

	if(getInputBool(&inp, "use_edge", &_use_edge, 0) == KEY_FOUND) {
		if(_use_edge && _use_barostat) {
			throw oxDNAException("use_edge and use_barostat are not compatible");
		}
	}
	// the reference refuses use_edge + use_barostat (MD_CUDABackend.cu:629-631) only because its edge scratch buffers are sized
	// for the initial box; the check is kept for input compatibility
	getInputBool(&inp, "CUDA_barostat_always_refresh", &_barostat_always_refresh, 0);

	getInputBool(&inp, "CUDA_avoid_cpu_calculations", &_avoid_cpu_calculations, 0);
	getInputBool(&inp, "CUDA_print_energy", &_print_energy, 0);
	getInputLLInt(&inp, "CUDA_max_queued_steps", &_max_pending, 0);
	getInputBool(&inp, "CUDA_device_observables", &_device_observables, 0);

	_cuda_thermostat = CUDAThermostatFactory::make_thermostat(inp, _box.get());
	_cuda_thermostat->get_settings(inp);

	// same trimming of the default outputs as the reference (MD_CUDABackend.cu:652-671)
	if(_avoid_cpu_calculations) {
		_obs_output_file->clear();
		_obs_output_file->add_observable("type = step\nunits = MD");
		bool no_stdout_energy = false;
		getInputBool(&inp, "no_stdout_energy", &no_stdout_energy, 0);
		if(!no_stdout_energy) {
			_obs_output_stdout->clear();
			_obs_output_stdout->add_observable("type = step");
			_obs_output_stdout->add_observable("type = step\nunits = MD");
		}
	}
}

void MD_CUDABackend::init() {
	MDBackend::init();

	const int n = N();
	int dev = _device_number < 0 ? 0 : _device_number;
	int rc = oxb_create(&_ctx, dev, n, _precision);
	if(rc != 0) {
		std::string msg = _ctx ? oxb_last_error(_ctx) : "out of memory";
		throw oxDNAException("Cannot initialise the oxdna_b200 device context: %s", msg.c_str());
	}
	OX_LOG(Logger::LOG_INFO, "oxdna_b200 backend running on device %d", dev);

	LR_vector sides = _box->box_sides();
	double box[3] = { (double) sides.x, (double) sides.y, (double) sides.z };
	oxb_check(_ctx, oxb_set_box(_ctx, box), "set_box");

	std::vector<int> btype(n), n3(n), n5(n), strand(n);
	for(int i = 0; i < n; i++) {
		BaseParticle *p = _particles[i];
		btype[i] = p->btype;
		n3[i] = (p->n3 == P_VIRTUAL) ? -1 : p->n3->index;
		n5[i] = (p->n5 == P_VIRTUAL) ? -1 : p->n5->index;
		strand[i] = p->strand_id;
		if(p->btype > 511 || p->btype < -511) {
			throw oxDNAException("Could not treat the type (A, C, G, T or something specific) of particle %d; On CUDA, integer base types cannot be larger than 511 or smaller than -511", i);
		}
	}
	oxb_check(_ctx, oxb_set_topology(_ctx, btype.data(), n3.data(), n5.data(), strand.data()), "set_topology");

	_cuda_interaction->cuda_init(_ctx, n);
	_cuda_lists->init(_ctx, n, _cuda_interaction->get_cuda_rcut(), _sort_every);
	oxb_check(_ctx, oxb_set_dt(_ctx, (double) _dt), "set_dt");

	_cuda_thermostat->set_seed(lrand48());
	_cuda_thermostat->init();
	_cuda_thermostat->attach(_ctx);

	if(_device_observables) _install_device_observables();

	// copy all the particle related stuff to device memory, then lists and forces for the first step
	apply_changes_to_simulation_data();
	oxb_check(_ctx, oxb_set_step(_ctx, current_step()), "set_step");
	_cuda_lists->update();
	_cuda_interaction->compute_forces(_ctx);
}

void MD_CUDABackend::_host_to_gpu() {
	const int n = N();
	std::vector<double> pos(3 * n), a1(3 * n), a3(3 * n), vel(3 * n), L(3 * n);
	for(int i = 0; i < n; i++) {
		BaseParticle *p = _particles[i];
		const LR_vector &v1 = p->orientationT.v1, &v3 = p->orientationT.v3;
		pos[3 * i] = p->pos.x; pos[3 * i + 1] = p->pos.y; pos[3 * i + 2] = p->pos.z;
		a1[3 * i] = v1.x; a1[3 * i + 1] = v1.y; a1[3 * i + 2] = v1.z;
		a3[3 * i] = v3.x; a3[3 * i + 1] = v3.y; a3[3 * i + 2] = v3.z;
		vel[3 * i] = p->vel.x; vel[3 * i + 1] = p->vel.y; vel[3 * i + 2] = p->vel.z;
		L[3 * i] = p->L.x; L[3 * i + 1] = p->L.y; L[3 * i + 2] = p->L.z;
	}
	{
		// CUDABaseBackend::_host_to_gpu: the device box follows the CPU one (CUDABaseBackend.cu:196)
		double box[3];
		LR_vector sides = _box->box_sides();
		if(oxb_get_box(_ctx, box) == 0 && (box[0] != (double) sides.x || box[1] != (double) sides.y || box[2] != (double) sides.z)) {
			double nb[3] = { (double) sides.x, (double) sides.y, (double) sides.z };
			oxb_check(_ctx, oxb_set_box(_ctx, nb), "set_box");
		}
	}
	oxb_check(_ctx, oxb_set_state(_ctx, pos.data(), a1.data(), a3.data(), vel.data(), L.data()), "set_state");
}

void MD_CUDABackend::_gpu_to_host() {
	const int n = N();
	std::vector<double> pos(3 * n), a1(3 * n), a3(3 * n), vel(3 * n), L(3 * n);
	oxb_check(_ctx, oxb_get_state(_ctx, pos.data(), a1.data(), a3.data(), vel.data(), L.data()), "get_state");
	if(_use_barostat) {
		// CUDABaseBackend::_gpu_to_host: the box follows the device (CUDABaseBackend.cu:106-107)
		double box[3];
		oxb_check(_ctx, oxb_get_box(_ctx, box), "get_box");
		LR_vector sides = _box->box_sides();
		if(box[0] != (double) sides.x || box[1] != (double) sides.y || box[2] != (double) sides.z) _box->init(box[0], box[1], box[2]);
	}
	std::vector<double> F, T;
	if(!_avoid_cpu_calculations) {
		// superset of the reference, which never writes the GPU forces back
		F.resize(3 * n);
		T.resize(3 * n);
		oxb_check(_ctx, oxb_get_forces(_ctx, F.data(), T.data(), nullptr, nullptr, nullptr), "get_forces");
	}
	for(int i = 0; i < n; i++) {
		BaseParticle *p = _particles[i];
		p->pos = LR_vector(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		p->vel = LR_vector(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
		p->L = LR_vector(L[3 * i], L[3 * i + 1], L[3 * i + 2]);
		LR_vector v1(a1[3 * i], a1[3 * i + 1], a1[3 * i + 2]), v3(a3[3 * i], a3[3 * i + 1], a3[3 * i + 2]);
		LR_vector v2 = v3.cross(v1);
		// orientationT has the axes as rows, orientation as columns
		p->orientationT = LR_matrix(v1, v2, v3);
		p->orientation = p->orientationT.get_transpose();
		p->set_positions();
		if(!F.empty()) {
			p->force = LR_vector(F[3 * i], F[3 * i + 1], F[3 * i + 2]);
			p->torque = LR_vector(T[3 * i], T[3 * i + 1], T[3 * i + 2]);
		}
	}
}

void MD_CUDABackend::_apply_external_forces_changes() {
	if(!_external_forces) return;
	std::vector<oxb_ext_force> table;
	std::vector<int> pool;    // com_list / ref_list (p1a / p2a) indices of the COM forces
	std::vector<double> grid; // tabulated bias potentials of the metadynamics COM traps
	// a force given with `particle = all` is the same object attached to every particle: keep it as ONE table entry
	// (particle = -1) instead of N copies (the reference keeps 15 union slots per particle)
	std::map<BaseForce *, int> uses;
	for(int i = 0; i < N(); i++) for(auto f : _particles[i]->ext_forces) uses[f]++;
	std::set<BaseForce *> emitted;
	for(int i = 0; i < N(); i++) {
		BaseParticle *p = _particles[i];
		for(auto f : p->ext_forces) {
			oxb_ext_force e;
			memset(&e, 0, sizeof(e));
			e.particle = i;
			auto &ft = typeid(*f);
			bool single_particle_type = true;
			if(ft == typeid(ConstantRateForce)) {
				ConstantRateForce *cf = static_cast<ConstantRateForce *>(f);
				e.type = OXB_EXT_STRING;
				e.F0 = cf->_F0; e.rate = cf->_rate;
				e.dir[0] = cf->_direction.x; e.dir[1] = cf->_direction.y; e.dir[2] = cf->_direction.z;
				if(cf->dir_as_centre) {
					// the "direction" is a point: the force pulls towards it (src/CUDA/Backends/CUDA_MD.cuh:117-122)
					e.pbc = 1;
					e.pos0[0] = cf->_direction.x; e.pos0[1] = cf->_direction.y; e.pos0[2] = cf->_direction.z;
				}
			}
			else if(ft == typeid(MutualTrap)) {
				MutualTrap *mf = static_cast<MutualTrap *>(f);
				single_particle_type = false;
				e.type = OXB_EXT_MUTUAL_TRAP;
				e.ref = mf->_p_ptr->index;
				e.pbc = mf->PBC ? 1 : 0;
				e.stiff = mf->_stiff; e.r0 = mf->_r0; e.rate = mf->_rate; e.stiff_rate = mf->_stiff_rate;
			}
			else if(ft == typeid(MovingTrap)) {
				MovingTrap *tf = static_cast<MovingTrap *>(f);
				e.type = OXB_EXT_TRAP;
				e.stiff = tf->_stiff; e.rate = tf->_rate;
				e.dir[0] = tf->_direction.x; e.dir[1] = tf->_direction.y; e.dir[2] = tf->_direction.z;
				e.pos0[0] = tf->_pos0.x; e.pos0[1] = tf->_pos0.y; e.pos0[2] = tf->_pos0.z;
			}
			else if(ft == typeid(LowdimMovingTrap)) {
				LowdimMovingTrap *tf = static_cast<LowdimMovingTrap *>(f);
				e.type = OXB_EXT_LOWDIM_TRAP;
				e.stiff = tf->_stiff; e.rate = tf->_rate;
				e.dir[0] = tf->_direction.x; e.dir[1] = tf->_direction.y; e.dir[2] = tf->_direction.z;
				e.pos0[0] = tf->_pos0.x; e.pos0[1] = tf->_pos0.y; e.pos0[2] = tf->_pos0.z;
				e.iaux = (tf->_visX ? 1 : 0) | (tf->_visY ? 2 : 0) | (tf->_visZ ? 4 : 0);
			}
			else if(ft == typeid(RepulsionPlane)) {
				RepulsionPlane *pf = static_cast<RepulsionPlane *>(f);
				e.type = OXB_EXT_REPULSION_PLANE;
				e.stiff = pf->_stiff;
				e.dir[0] = pf->_direction.x; e.dir[1] = pf->_direction.y; e.dir[2] = pf->_direction.z;
				e.aux[0] = pf->_starting_position; e.aux[1] = pf->_v; e.aux[2] = pf->_end_position;
			}
			else if(ft == typeid(AttractionPlane)) {
				AttractionPlane *pf = static_cast<AttractionPlane *>(f);
				e.type = OXB_EXT_ATTRACTION_PLANE;
				e.stiff = pf->_stiff;
				e.dir[0] = pf->_direction.x; e.dir[1] = pf->_direction.y; e.dir[2] = pf->_direction.z;
				e.aux[0] = pf->_position;
			}
			else if(ft == typeid(RepulsiveSphere)) {
				RepulsiveSphere *sf = static_cast<RepulsiveSphere *>(f);
				e.type = OXB_EXT_SPHERE;
				e.stiff = sf->_stiff; e.r0 = sf->_r0; e.rate = sf->_rate;
				e.pos0[0] = sf->_center.x; e.pos0[1] = sf->_center.y; e.pos0[2] = sf->_center.z;
				e.aux[0] = sf->_r_ext;
			}
			else if(ft == typeid(LJWall)) {
				LJWall *wf = static_cast<LJWall *>(f);
				e.type = OXB_EXT_LJ_WALL;
				e.stiff = wf->_stiff;
				e.dir[0] = wf->_direction.x; e.dir[1] = wf->_direction.y; e.dir[2] = wf->_direction.z;
				e.aux[0] = wf->_position; e.aux[1] = wf->_sigma; e.aux[2] = wf->_cutoff;
				e.iaux = wf->_n;
			}
			else if(ft == typeid(ConstantRateTorque)) {
				ConstantRateTorque *tf = static_cast<ConstantRateTorque *>(f);
				e.type = OXB_EXT_TWIST;
				e.stiff = tf->_stiff; e.rate = tf->_rate; e.F0 = tf->_F0;
				e.dir[0] = tf->_axis.x; e.dir[1] = tf->_axis.y; e.dir[2] = tf->_axis.z;
				e.pos0[0] = tf->_pos0.x; e.pos0[1] = tf->_pos0.y; e.pos0[2] = tf->_pos0.z;
				e.aux[0] = tf->_center.x; e.aux[1] = tf->_center.y; e.aux[2] = tf->_center.z;
				e.aux[3] = tf->_mask.x; e.aux[4] = tf->_mask.y; e.aux[5] = tf->_mask.z;
			}
			else if(ft == typeid(RepulsiveSphereSmooth)) {
				RepulsiveSphereSmooth *sf = static_cast<RepulsiveSphereSmooth *>(f);
				e.type = OXB_EXT_SPHERE_SMOOTH;
				e.stiff = sf->_stiff; e.r0 = sf->_r0;
				e.pos0[0] = sf->_center.x; e.pos0[1] = sf->_center.y; e.pos0[2] = sf->_center.z;
				e.aux[0] = sf->_r_ext; e.aux[1] = sf->_smooth; e.aux[2] = sf->_alpha;
			}
			else if(ft == typeid(RepulsiveEllipsoid)) {
				RepulsiveEllipsoid *ef = static_cast<RepulsiveEllipsoid *>(f);
				e.type = OXB_EXT_ELLIPSOID;
				e.stiff = ef->_stiff;
				e.pos0[0] = ef->_centre.x; e.pos0[1] = ef->_centre.y; e.pos0[2] = ef->_centre.z;
				e.aux[0] = ef->_r_2.x; e.aux[1] = ef->_r_2.y; e.aux[2] = ef->_r_2.z;
				e.aux[3] = ef->_r_1.x; e.aux[4] = ef->_r_1.y; e.aux[5] = ef->_r_1.z;
			}
			else if(ft == typeid(RepulsionPlaneMoving)) {
				RepulsionPlaneMoving *pf = static_cast<RepulsionPlaneMoving *>(f);
				e.type = OXB_EXT_REPULSION_PLANE_MOVING;
				e.stiff = pf->_stiff;
				e.dir[0] = pf->_direction.x; e.dir[1] = pf->_direction.y; e.dir[2] = pf->_direction.z;
				e.ref = pf->low_idx; e.iaux = pf->high_idx;
			}
			else if(ft == typeid(GenericCentralForce)) {
				// as in the reference's CUDA backend (forces_defs.cuh:300-310) only the gravity flavour reaches the device:
				// an `interpolated` force has F0 = 0 there as well
				GenericCentralForce *gf = static_cast<GenericCentralForce *>(f);
				e.type = OXB_EXT_GENERIC_CENTRAL;
				e.F0 = gf->_F0;
				e.pos0[0] = gf->center.x; e.pos0[1] = gf->center.y; e.pos0[2] = gf->center.z;
				e.aux[0] = gf->inner_cut_off_sqr; e.aux[1] = gf->outer_cut_off_sqr;
			}
			else if(ft == typeid(LJCone)) {
				LJCone *cf = static_cast<LJCone *>(f);
				e.type = OXB_EXT_LJ_CONE;
				e.stiff = cf->_stiff;
				e.dir[0] = cf->_direction.x; e.dir[1] = cf->_direction.y; e.dir[2] = cf->_direction.z;
				e.pos0[0] = cf->_pos0.x; e.pos0[1] = cf->_pos0.y; e.pos0[2] = cf->_pos0.z;
				e.aux[0] = cf->_sigma; e.aux[1] = cf->_cutoff; e.aux[2] = cf->_alpha;
				e.iaux = cf->_n;
			}
			else if(ft == typeid(COMForce)) {
				// one table entry per force object (it is attached to every particle of its com_list)
				if(emitted.count(f)) continue;
				emitted.insert(f);
				COMForce *cf = static_cast<COMForce *>(f);
				single_particle_type = false;
				e.type = OXB_EXT_COM;
				e.particle = -1;
				e.stiff = cf->_stiff; e.r0 = cf->_r0; e.rate = cf->_rate;
				e.ref = (int) pool.size(); e.iaux = (int) cf->_com_list.size(); e.pbc = (int) cf->_ref_list.size();
				for(auto q : cf->_com_list) pool.push_back(q->index);
				for(auto q : cf->_ref_list) pool.push_back(q->index);
			}
			else if(ft == typeid(LTCOMTrap)) {
				// meta_com_trap: one table entry per force object (it is attached to every particle of p1a or of p2a, according to its mode)
				if(emitted.count(f)) continue;
				emitted.insert(f);
				LTCOMTrap *mf = static_cast<LTCOMTrap *>(f);
				single_particle_type = false;
				e.type = OXB_EXT_META_COM_TRAP;
				e.particle = -1;
				e.ref = (int) pool.size(); e.iaux = (int) mf->_p1a_ptr.size(); e.pbc = (int) mf->_p2a_ptr.size();
				for(auto q : mf->_p1a_ptr) pool.push_back(q->index);
				for(auto q : mf->_p2a_ptr) pool.push_back(q->index);
				e.aux[0] = mf->xmin; e.aux[1] = mf->dX; e.aux[2] = mf->N_grid; e.aux[3] = mf->_mode; e.aux[4] = (double) grid.size(); e.aux[5] = mf->PBC ? 1. : 0.;
				for(auto v : mf->potential_grid) grid.push_back(v);
			}
			else if(ft == typeid(LTCoordination)) {
				// meta_coordination: one table entry per force object (it is attached to every particle of its hydrogen-bond pairs)
				if(emitted.count(f)) continue;
				emitted.insert(f);
				LTCoordination *cf = static_cast<LTCoordination *>(f);
				single_particle_type = false;
				e.type = OXB_EXT_META_COORDINATION;
				e.particle = -1;
				e.ref = (int) pool.size(); e.iaux = (int) cf->all_pairs.size();
				for(auto &pr : cf->all_pairs) { pool.push_back(pr.first->index); pool.push_back(pr.second->index); }
				const int mode = static_cast<int>(cf->settings.coord_mode); // HB_ENERGY, SWITCHING_FUNCTION, MIXED = 0, 1, 2
				e.pbc = cf->settings.n; e.r0 = cf->settings.d0; e.stiff = cf->settings.r0; e.F0 = cf->coord_max;
				e.aux[0] = cf->coord_min; e.aux[1] = cf->d_coord; e.aux[2] = cf->N_grid; e.aux[3] = mode; e.aux[4] = (double) grid.size();
				e.aux[5] = (mode == 2) ? cf->settings.mixed_weight : 0.; // only set by the parser for the mixed type
				e.aux[6] = cf->settings.hb_energy_cutoff; e.aux[7] = cf->settings.hb_transition_width;
				for(auto v : cf->potential_grid) grid.push_back(v);
			}
			else if(ft == typeid(YukawaSphere)) {
				YukawaSphere *yf = static_cast<YukawaSphere *>(f);
				e.type = OXB_EXT_YUKAWA_SPHERE;
				e.pos0[0] = yf->_center.x; e.pos0[1] = yf->_center.y; e.pos0[2] = yf->_center.z;
				e.r0 = yf->_radius; e.stiff = yf->_epsilon;
				e.aux[0] = yf->_sigma; e.aux[1] = yf->_WCA_cutoff; e.aux[2] = yf->_debye_length; e.aux[3] = yf->_debye_A; e.aux[4] = yf->_cutoff;
				e.iaux = yf->_WCA_n;
			}
			else if(ft == typeid(RepulsiveSphereMoving)) {
				RepulsiveSphereMoving *sf = static_cast<RepulsiveSphereMoving *>(f);
				e.type = OXB_EXT_SPHERE_MOVING;
				e.stiff = sf->stiff(); e.r0 = sf->r0(); e.rate = sf->rate();
				LR_vector o = sf->origin(), t = sf->target();
				e.pos0[0] = o.x; e.pos0[1] = o.y; e.pos0[2] = o.z;
				e.aux[0] = sf->r_ext(); e.aux[1] = t.x; e.aux[2] = t.y; e.aux[3] = t.z; e.aux[4] = (double) sf->steps();
			}
			else {
				throw oxDNAException("Only string, trap, mutual_trap, lowdim_trap, twist, repulsion_plane, repulsion_plane_moving, attraction_plane, sphere, sphere_smooth, repulsive_sphere_moving, ellipsoid, LJ_wall, LJ_cone, generic_central_force, com, meta_com_trap, meta_coordination and yukawa_sphere forces are supported by the oxdna_b200 CUDA backend at the moment.\n");
			}
			if(single_particle_type && N() > 1 && uses[f] == N()) {
				if(emitted.count(f)) continue;
				emitted.insert(f);
				e.particle = -1;
			}
			table.push_back(e);
		}
	}
	// unlike the reference (MD_CUDABackend.cu:110-112) external forces may be combined with CUDA_sort_every > 0:
	// the table holds original particle ids, the device maps them through its slot table
	oxb_check(_ctx, oxb_set_ext_forces(_ctx, 0, nullptr), "set_ext_forces"); // COM entries refer to the pool: drop them before replacing it
	oxb_check(_ctx, oxb_set_ext_index_pool(_ctx, (int) pool.size(), pool.data()), "set_ext_index_pool");
	oxb_check(_ctx, oxb_set_ext_grid_pool(_ctx, (int) grid.size(), grid.data()), "set_ext_grid_pool");
	oxb_check(_ctx, oxb_set_ext_forces(_ctx, (int) table.size(), table.data()), "set_ext_forces");
}

void MD_CUDABackend::_flush() {
	if(_pending_steps == 0) return;
	oxb_check(_ctx, oxb_set_step(_ctx, _first_pending_step), "set_step");
	llint n = _pending_steps;
	_pending_steps = 0;
	int rc = oxb_run(_ctx, n);
	long long updates = 0;
	oxb_get_stats(_ctx, &updates, nullptr, nullptr, nullptr);
	_N_updates = (int) updates;
	if(rc != 0) throw oxDNAException("%s", oxb_last_error(_ctx));
	if(_print_energy) {
		double U = 0., K = 0.;
		oxb_check(_ctx, oxb_energy(_ctx, &U, &K), "energy");
		_backend_info = Utils::sformat("\tCUDA_energy: %lf", U / N());
	}
}

void MD_CUDABackend::_install_device_observables() {
	// the default energy streams of MDBackend::get_settings (src/Backends/MDBackend.cpp:60-93) with total_energy evaluated on the device
	auto fill = [&](ObservableOutputPtr out, bool with_plain_step) {
		if(out == nullptr) return;
		out->clear();
		if(with_plain_step) out->add_observable("type = step");
		out->add_observable("type = step\nunits = MD");
		out->add_observable(std::make_shared<CUDATotalEnergy>(_device_UK, N()));
		if(_use_barostat) out->add_observable("type = density");
		out->add_observable("type = backend_info");
	};
	fill(_obs_output_file, false);
	fill(_obs_output_stdout, true);
}

void MD_CUDABackend::print_observables() {
	if(!_device_observables) {
		MDBackend::print_observables();
		return;
	}
	// SimBackend::print_observables (src/Backends/SimBackend.cpp:728-768) brackets every print with a download, a CPU list rebuild and
	// an upload.  With device-side energies the default streams need none of that: only the queued steps have to run.
	bool any = false, only_default = true;
	for(auto const &out : _obs_outputs) {
		if(out->is_ready(current_step())) {
			any = true;
			if(out != _obs_output_file && out != _obs_output_stdout) only_default = false;
		}
	}
	if(!any) {
		_backend_info = std::string("");
		return;
	}
	_flush();
	oxb_check(_ctx, oxb_energy(_ctx, &_device_UK[0], &_device_UK[1]), "energy");
	if(!only_default) {
		MDBackend::print_observables(); // some stream wants the particles on the CPU: the reference path
		return;
	}
	_mytimer->resume();
	_obs_timer->resume();
	if(_use_barostat) _backend_info.insert(0, Utils::sformat(" %5.3lf", _barostat_acceptance));
	for(auto const &out : _obs_outputs) {
		if(out->is_ready(current_step())) out->print_output(current_step());
	}
	_backend_info = std::string("");
	_obs_timer->pause();
	_mytimer->pause();
}

void MD_CUDABackend::sim_step() {
	_mytimer->resume();
	if(_pending_steps == 0) _first_pending_step = current_step();
	_pending_steps++;
	// MD_CUDABackend.cu:595-599: one draw per step decides whether the barostat fires (MDBackend::_is_barostat_active); the move is
	// applied once the queued steps, this one included, have run
	const bool barostat_now = _is_barostat_active();
	if(_pending_steps >= _max_pending || barostat_now) _flush();
	if(barostat_now) {
		_timer_barostat->resume();
		_apply_barostat();
		_timer_barostat->pause();
	}
	_mytimer->pause();
}

void MD_CUDABackend::fix_diffusion() {
	// SimBackend::fix_diffusion (src/Backends/SimBackend.cpp:786-882) without the round trip through the CPU particles and its two CPU
	// energy evaluations (seconds at 1M nucleotides): strands are translated by whole box sides on the device
	if(!_enable_fix_diffusion) return;
	if(_reset_com_momentum) {
		MDBackend::fix_diffusion(); // the momentum reset works on the CPU copies: keep the reference path
		return;
	}
	_flush();
	std::vector<int> shifts(3 * (size_t) N());
	oxb_check(_ctx, oxb_fix_diffusion(_ctx, shifts.data()), "fix_diffusion");
	for(int i = 0; i < N(); i++) {
		int cur[3];
		_particles[i]->get_pos_shift(cur);
		_particles[i]->set_pos_shift(cur[0] + shifts[3 * i], cur[1] + shifts[3 * i + 1], cur[2] + shifts[3 * i + 2]);
	}
	OX_LOG(Logger::LOG_INFO, "diffusion fixed on the device");
}

void MD_CUDABackend::_apply_barostat() {
	// MD_CUDABackend::_apply_barostat (src/CUDA/Backends/MD_CUDABackend.cu:451-516): box draw and acceptance number on the host
	// (drand48, same order of draws), energies / rescaling / list rebuild on the device
	_barostat_attempts++;
	double old_box[3], new_box[3];
	oxb_check(_ctx, oxb_get_box(_ctx, old_box), "get_box");
	if(_barostat_isotropic) {
		double dL = _delta_L * (drand48() - 0.5);
		for(int k = 0; k < 3; k++) new_box[k] = old_box[k] + dL;
	}
	else {
		for(int k = 0; k < 3; k++) new_box[k] = old_box[k] + _delta_L * (drand48() - 0.5);
	}
	int accepted = 0;
	const double u = drand48();
	oxb_check(_ctx, oxb_barostat_move(_ctx, new_box, _barostat_molecular ? 1 : 0, (double) _P, (double) _T, u, &accepted, nullptr), "barostat_move");
	if(accepted) {
		_barostat_accepted++;
		_box->init(new_box[0], new_box[1], new_box[2]);
		CONFIG_INFO->notify(CONFIG_INFO->box->UPDATE_EVENT);
	}
	_barostat_acceptance = _barostat_accepted / (number) _barostat_attempts;
	if(_barostat_always_refresh) {
		// CUDA_barostat_always_refresh (MD_CUDABackend.cu:512-515,646-651): a Brownian thermostat with newtonian_steps = 1, pt = 1
		// (hence pr = 1/2, BrownianThermostat.cpp:43-54) redraws every velocity after each attempt
		oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_BROWNIAN, 1, 1.0, 0.5, std::sqrt((double) _T), 0., (unsigned long long) lrand48()), "set_thermostat");
		oxb_check(_ctx, oxb_thermostat(_ctx), "thermostat");
		_cuda_thermostat->attach(_ctx);
	}
}

void MD_CUDABackend::apply_simulation_data_changes() {
	_flush();
	_gpu_to_host();
	_cuda_interaction->sync_host();
	if(!_avoid_cpu_calculations) {
		_lists->global_update(true);
	}
}

void MD_CUDABackend::apply_changes_to_simulation_data() {
	_host_to_gpu();
	_apply_external_forces_changes();
	_cuda_interaction->sync_GPU();
}

void MD_CUDABackend::_on_T_update() {
	// steps queued at the old temperature must run before the operators re-derive their constants
	if(_ctx != nullptr) _flush();
	MDBackend::_on_T_update();
}
