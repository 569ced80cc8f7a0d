// Implementation of the host-side operator classes (see CUDAOperators.h).  Nothing here computes physics on the CPU:
// settings are parsed with the reference's own input helpers, constants come from the reference's CPU classes, and all
// work is forwarded to liboxdna_b200.so.
#include "CUDAOperators.h"

#include "Utilities/ConfigInfo.h"
#include "Utilities/Logger.h"
#include "Utilities/Utils.h"
#include "model.h"

#include <cmath>
#include <cstring>

void oxb_check(oxb_ctx *ctx, int rc, const char *what) {
	if(rc != 0) throw oxDNAException("%s: %s (oxdna_b200 status %d)", what, ctx ? oxb_last_error(ctx) : "no context", rc);
}

// ---------------------------------------------------------------------------------------------------- interactions
void CUDABaseInteraction::get_cuda_settings(input_file &inp) {
	// src/CUDA/Interactions/CUDABaseInteraction.cu:79-98
	int tmpi;
	if(getInputBoolAsInt(&inp, "use_edge", &tmpi, 0) == KEY_FOUND) {
		if(tmpi > 0) {
			_use_edge = true;
			getInputInt(&inp, "edge_n_forces", &_n_forces, 0);
			if(_n_forces < 1) throw oxDNAException("edge_n_forces must be > 0");
		}
	}
}

void CUDABaseInteraction::cuda_init(oxb_ctx *ctx, int N) {
	_ctx = ctx;
	_N = N;
}

void CUDABaseInteraction::compute_forces(oxb_ctx *ctx) {
	oxb_check(ctx, oxb_compute_forces(ctx), "compute_forces");
}

CUDADNAInteraction::CUDADNAInteraction() {}

CUDADNAInteraction::~CUDADNAInteraction() {}

void CUDADNAInteraction::get_settings(input_file &inp) {
	std::string inter_type("DNA");
	getInputString(&inp, "interaction_type", inter_type, 0);
	if(inter_type != "DNA2") {
		throw oxDNAException("interaction_type = %s is not available in the oxdna_b200 CUDA backend (DNA2 only in this build)", inter_type.c_str());
	}
	// every key of the CPU class: salt_concentration, dh_lambda, dh_strength, dh_half_charged_ends, use_average_seq,
	// seq_dep_file, hb_multiplier, max_backbone_force[_far], major-minor grooving
	DNA2Interaction::get_settings(inp);
	if(!this->_grooving) throw oxDNAException("major_minor_grooving = false is not available in the oxdna_b200 CUDA backend");
}

void CUDADNAInteraction::cuda_init(oxb_ctx *ctx, int N) {
	CUDABaseInteraction::cuda_init(ctx, N);
	Logger::instance()->disable_log("CUDADNAInteraction");
	DNA2Interaction::init();
	Logger::instance()->enable_log("CUDADNAInteraction");
	_upload();
}

void CUDADNAInteraction::_upload() {
	if(_ctx == nullptr) return;
	oxb_dna2_params P;
	double rcut = 0.;
	int rc = oxb_dna2_params_init(&P, (double) this->_T, (double) _salt_concentration, _debye_huckel_half_charged_ends ? 1 : 0, _use_mbf ? 1 : 0,
			(double) _mbf_fmax, (double) _mbf_finf, &rcut);
	if(rc != 0) throw oxDNAException("oxb_dna2_params_init failed (T = %lf, salt = %f)", (double) this->_T, _salt_concentration);
	// what depends on the input file comes from the CPU class that parsed it (CUDADNAInteraction.cu:67-153 copies the same members)
	for(int i = 0; i < 5; i++) {
		for(int j = 0; j < 5; j++) {
			P.hb_eps[5 * i + j] = (float) F1_EPS[HYDR_F1][i][j];
			P.hb_shift[5 * i + j] = (float) F1_SHIFT[HYDR_F1][i][j];
			P.stck_eps[5 * i + j] = (float) F1_EPS[STCK_F1][i][j];
			P.stck_shift[5 * i + j] = (float) F1_SHIFT[STCK_F1][i][j];
		}
	}
	P.hb_multiplier = (float) _hb_multiplier;
	P.dh_minus_kappa = (float) _minus_kappa;
	P.dh_prefactor = (float) _debye_huckel_prefactor;
	P.dh_rhigh = (float) _debye_huckel_RHIGH;
	P.dh_rc = (float) _debye_huckel_RC;
	P.dh_b = (float) _debye_huckel_B;
	if(_use_mbf) {
		P.mbf_xmax = (float) _mbf_xmax;
	}
	// the Verlet radius must be the CPU class's own double-precision cutoff (bit-exact pair sets)
	rcut = (double) this->_rcut;
	P.rcut = (float) rcut;
	oxb_check(_ctx, oxb_set_model_dna2(_ctx, &P, rcut), "set_model_dna2");
}

void CUDADNAInteraction::_on_T_update() {
	// the reference's GPU class re-runs cuda_init -> DNAInteraction::init() (CUDADNAInteraction.cu:156-158)
	this->_T = CONFIG_INFO->temperature();
	if(_ctx != nullptr) {
		Logger::instance()->disable_log("CUDADNAInteraction");
		DNA2Interaction::init();
		Logger::instance()->enable_log("CUDADNAInteraction");
		_upload();
	}
}

std::shared_ptr<CUDABaseInteraction> CUDAInteractionFactory::make_interaction(input_file &inp) {
	std::string inter_type("DNA");
	getInputString(&inp, "interaction_type", inter_type, 0);
	if(inter_type == "DNA2") return std::make_shared<CUDADNAInteraction>();
	throw oxDNAException("CUDA interaction '%s' not found in the oxdna_b200 backend (available: DNA2). Aborting", inter_type.c_str());
}

// ---------------------------------------------------------------------------------------------------------- lists
void CUDABaseList::init(oxb_ctx *ctx, int N, number rcut, int sort_every) {
	_ctx = ctx;
	_N = N;
}

void CUDASimpleVerletList::get_settings(input_file &inp) {
	// src/CUDA/Lists/CUDASimpleVerletList.cu:47-56
	getInputBool(&inp, "cells_auto_optimisation", &_auto_optimisation, 0);
	getInputBool(&inp, "print_problematic_ids", &_print_problematic_ids, 0);
	getInputNumber(&inp, "verlet_skin", &_verlet_skin, 1);
	getInputNumber(&inp, "max_density_multiplier", &_max_density_multiplier, 0);
	getInputBool(&inp, "use_edge", &_use_edge, 0);
	if(_use_edge) {
		OX_LOG(Logger::LOG_INFO, "Using edge-based approach");
	}
}

void CUDASimpleVerletList::init(oxb_ctx *ctx, int N, number rcut, int sort_every) {
	CUDABaseList::init(ctx, N, rcut, sort_every);
	oxb_check(ctx, oxb_set_lists(ctx, (double) _verlet_skin, _use_edge ? 1 : 0, sort_every, (double) _max_density_multiplier), "set_lists");
}

void CUDASimpleVerletList::update() {
	int rc = oxb_update_lists(_ctx);
	if(rc != 0) throw oxDNAException("A cell or neighbour row contains too many particles (%s). Please increase the value of max_density_multiplier (which defaults to 3) in the input file", oxb_last_error(_ctx));
}

std::shared_ptr<CUDABaseList> CUDAListFactory::make_list(input_file &inp) {
	// src/CUDA/Lists/CUDAListFactory.cu
	std::string list_type("verlet");
	getInputString(&inp, "CUDA_list", list_type, 0);
	if(list_type == "verlet") return std::make_shared<CUDASimpleVerletList>();
	if(list_type == "no") {
		bool use_edge = false;
		getInputBool(&inp, "use_edge", &use_edge, 0);
		if(use_edge) throw oxDNAException("'CUDA_list = no' and 'use_edge = true' are incompatible");
		throw oxDNAException("CUDA_list = no (all-pairs) is not available in the oxdna_b200 backend: use CUDA_list = verlet");
	}
	throw oxDNAException("CUDA_list '%s' is not supported", list_type.c_str());
}

// ----------------------------------------------------------------------------------------------------- thermostats
void CUDABaseThermostat::apply_cuda(llint curr_step) {
	oxb_check(_ctx, oxb_set_step(_ctx, curr_step), "set_step");
	oxb_check(_ctx, oxb_thermostat(_ctx), "thermostat");
}

void CUDANoThermostat::_upload() {
	if(_ctx) oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_NONE, 1, 0., 0., 0., 0., (unsigned long long) _seed), "set_thermostat");
}

void CUDABrownianThermostat::get_settings(input_file &inp) {
	BrownianThermostat::get_settings(inp);
}

void CUDABrownianThermostat::init() {
	BrownianThermostat::init();
	_upload();
}

void CUDABrownianThermostat::_upload() {
	if(_ctx) oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_BROWNIAN, _newtonian_steps, (double) _pt, (double) _pr, (double) _rescale_factor, 0., (unsigned long long) _seed), "set_thermostat");
}

bool CUDABrownianThermostat::would_activate(llint curr_step) {
	return (curr_step % _newtonian_steps == 0);
}

void CUDALangevinThermostat::get_settings(input_file &inp) {
	LangevinThermostat::get_settings(inp);
}

void CUDALangevinThermostat::init() {
	LangevinThermostat::init();
	_upload();
}

void CUDALangevinThermostat::_upload() {
	if(_ctx) oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_LANGEVIN, 1, (double) _gamma_trans, (double) _gamma_rot, (double) _rescale_factor_trans, (double) _rescale_factor_rot, (unsigned long long) _seed), "set_thermostat");
}

void CUDABussiThermostat::get_settings(input_file &inp) {
	BussiThermostat::get_settings(inp);
}

void CUDABussiThermostat::init() {
	BussiThermostat::init();
	_upload();
}

void CUDABussiThermostat::_upload() {
	if(_ctx) oxb_check(_ctx, oxb_set_thermostat(_ctx, OXB_THERMOSTAT_BUSSI, _newtonian_steps, (double) this->_T, (double) _exp_dt_tau, 0., 0., (unsigned long long) _seed), "set_thermostat");
}

bool CUDABussiThermostat::would_activate(llint curr_step) {
	return (curr_step % _newtonian_steps == 0);
}

std::shared_ptr<CUDABaseThermostat> CUDAThermostatFactory::make_thermostat(input_file &inp, BaseBox *box) {
	// src/CUDA/Thermostats/CUDAThermostatFactory.cu:18-45
	char thermostat_type[512] = "no";
	getInputString(&inp, "thermostat", thermostat_type, 0);
	if(!strncmp(thermostat_type, "john", 512) || !strncmp(thermostat_type, "brownian", 512)) return std::make_shared<CUDABrownianThermostat>();
	else if(!strncmp(thermostat_type, "bussi", 512) || !strncmp(thermostat_type, "Bussi", 512)) return std::make_shared<CUDABussiThermostat>();
	else if(!strncmp(thermostat_type, "langevin", 512)) return std::make_shared<CUDALangevinThermostat>();
	else if(!strncmp(thermostat_type, "no", 512)) return std::make_shared<CUDANoThermostat>();
	else throw oxDNAException("Invalid CUDA thermostat '%s'", thermostat_type);
}
