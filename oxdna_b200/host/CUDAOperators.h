// Host-side operator classes of the B200 MD path, written against the reference's own headers (compiled with
// -I<reference>/src, linked against its unmodified CPU objects).  They keep the names, input keys, error messages and
// call order of the reference's CUDA-side classes, but own no device memory and launch no kernels themselves: every
// method forwards to the C ABI of liboxdna_b200.so (include/oxdna_b200.h).
//
//   class here                      replaces (reference file)
//   CUDABaseInteraction             src/CUDA/Interactions/CUDABaseInteraction.h:20-64
//   CUDADNAInteraction              src/CUDA/Interactions/CUDADNAInteraction.{h,cu}
//   CUDARNAInteraction              src/CUDA/Interactions/CUDARNAInteraction.{h,cu}
//   CUDAInteractionFactory          src/CUDA/Interactions/CUDAInteractionFactory.cu:28-53
//   CUDABaseList / CUDASimpleVerletList / CUDAListFactory   src/CUDA/Lists/*.{h,cu}
//   CUDABaseThermostat, CUDA{No,Brownian,Langevin,Bussi}Thermostat, CUDAThermostatFactory   src/CUDA/Thermostats/*.{h,cu}
//
// Difference in the signatures: the reference passes raw device pointers (poss, orientations, forces ...) between its
// classes; here the opaque oxb_ctx owns the device arrays, so the operators take the context instead.  Third-party
// kernels that need the raw arrays get them from oxb_device_views() in the reference's layouts.
#pragma once

#include "../../include/oxdna_b200.h"

#include "Backends/Thermostats/BaseThermostat.h"
#include "Backends/Thermostats/BrownianThermostat.h"
#include "Backends/Thermostats/BussiThermostat.h"
#include "Backends/Thermostats/LangevinThermostat.h"
#include "Backends/Thermostats/NoThermostat.h"
#include "Boxes/BaseBox.h"
#include "Interactions/DNA2Interaction.h"
#include "Interactions/DNA3Interaction.h"
#include "Interactions/RNAInteraction2.h"
#include "Utilities/oxDNAException.h"

#include <memory>
#include <string>

/// throws oxDNAException with the context's message if a C-ABI call failed (the reference's CUDA_SAFE_CALL exits instead)
void oxb_check(oxb_ctx *ctx, int rc, const char *what);

// ---------------------------------------------------------------------------------------------------- interactions
class CUDABaseInteraction {
protected:
	bool _use_edge = false;
	/// force slots per particle (edge_n_forces): accepted for compatibility, the kernels need no slots
	int _n_forces = 1;
	int _N = -1;
	oxb_ctx *_ctx = nullptr;

public:
	CUDABaseInteraction() {}
	virtual ~CUDABaseInteraction() {}

	virtual void get_settings(input_file &inp) = 0;
	virtual void get_cuda_settings(input_file &inp);
	/// uploads the model constants to the device context (the reference copies them to __constant__ memory here)
	virtual void cuda_init(oxb_ctx *ctx, int N);
	virtual number get_cuda_rcut() = 0;

	virtual void sync_host() {}
	virtual void sync_GPU() {}

	bool use_edge() const { return _use_edge; }
	/// forces, torques and per-particle energies for the context's current positions and lists
	virtual void compute_forces(oxb_ctx *ctx);

	// ---- plugin seam (reference: CUDAInteractionFactory falls back to PluginManager::get_interaction("CUDA" + type),
	// src/CUDA/Interactions/CUDAInteractionFactory.cu:44-51, and the backend calls compute_forces(lists, d_poss, d_orientations,
	// d_forces, d_torques, d_bonds, d_box) on whatever came back, CUDABaseInteraction.h:60).  A third-party interaction derives from
	// this class and from a CPU BaseInteraction, overrides compute_forces_views() -- the same raw device arrays in the reference's
	// layouts, see oxb_force_views in include/oxdna_b200.h -- and calls attach_as_plugin() from its cuda_init().
	/// enqueue the force pass on views.stream; forces / torques are zeroed accumulators (lab frame, .w = energy share)
	virtual void compute_forces_views(const oxb_force_views &views);
	/// makes compute_forces_views() the context's force pass, with get_cuda_rcut() as the Verlet cutoff
	void attach_as_plugin(oxb_ctx *ctx);
};

/// interaction_type = DNA2: the CPU DNA2Interaction supplies every constant (sequence dependence, salt, dh_* keys,
/// hb_multiplier, max_backbone_force), exactly as the reference's CUDADNAInteraction inherits them from DNAInteraction
class CUDADNAInteraction: public CUDABaseInteraction, public DNA2Interaction {
protected:
	void _upload();
	void _on_T_update() override;

public:
	CUDADNAInteraction();
	virtual ~CUDADNAInteraction();

	void get_settings(input_file &inp) override;
	void cuda_init(oxb_ctx *ctx, int N) override;
	number get_cuda_rcut() override {
		return this->get_rcut();
	}
};

/// interaction_type = DNA / DNA_nomesh (first-generation oxDNA): constants from the CPU DNAInteraction
class CUDADNA1Interaction: public CUDABaseInteraction, public DNAInteraction {
protected:
	void _upload();
	void _on_T_update() override;

public:
	CUDADNA1Interaction() {}
	virtual ~CUDADNA1Interaction() {}

	void get_settings(input_file &inp) override { DNAInteraction::get_settings(inp); }
	void cuda_init(oxb_ctx *ctx, int N) override;
	number get_cuda_rcut() override {
		return this->get_rcut();
	}
};

/// interaction_type = DNA3 (oxDNA3): the CPU DNA3Interaction_nomesh parses the sequence-dependent parameter file and derives the 214
/// tetramer-indexed tables; they are handed to the device library as they stand (the reference's CUDADNA3Interaction::cuda_init uploads the
/// same member arrays, src/CUDA/Interactions/CUDADNA3Interaction.cu:46-150).  use_edge = 1: staged edge pipeline, use_edge = 0: particle-centric kernel.
class CUDADNA3Interaction: public CUDABaseInteraction, public DNA3Interaction_nomesh {
protected:
	void _upload();
	void _on_T_update() override;

public:
	CUDADNA3Interaction() {}
	virtual ~CUDADNA3Interaction() {}

	void get_settings(input_file &inp) override;
	void cuda_init(oxb_ctx *ctx, int N) override;
	number get_cuda_rcut() override {
		return this->get_rcut();
	}
};

/// interaction_type = RNA2: constants from the CPU RNA2Interaction and its Model block (rna_model.h; `external_model`,
/// `use_average_seq` / `seq_dep_file`, `mismatch_repulsion[_strength]`, salt and dh_* keys, max_backbone_force), as the
/// reference's CUDARNAInteraction copies them into its `CUDAModel` (CUDARNAInteraction.cu:43-230,278-385)
class CUDARNAInteraction: public CUDABaseInteraction, public RNA2Interaction {
protected:
	/// interaction_type = RNA (first-generation oxRNA, class RNAInteraction): the same model without the Debye-Hueckel and
	/// mismatch terms; only the RNAInteraction part of the CPU class is configured and initialised
	bool _v1 = false;
	void _upload();
	void _on_T_update() override;

public:
	explicit CUDARNAInteraction(bool first_generation = false) : _v1(first_generation) {}
	virtual ~CUDARNAInteraction() {}

	void get_settings(input_file &inp) override;
	void cuda_init(oxb_ctx *ctx, int N) override;
	number get_cuda_rcut() override {
		return this->get_rcut();
	}
};

class CUDAInteractionFactory {
public:
	static std::shared_ptr<CUDABaseInteraction> make_interaction(input_file &inp);
};

// ---------------------------------------------------------------------------------------------------------- lists
class CUDABaseList {
protected:
	bool _use_edge = false;
	int _N = -1;
	oxb_ctx *_ctx = nullptr;

public:
	CUDABaseList() {}
	virtual ~CUDABaseList() {}

	virtual void get_settings(input_file &inp) = 0;
	virtual void init(oxb_ctx *ctx, int N, number rcut, int sort_every);
	/// rebuilds cells + Verlet (and edge) lists for the current positions; throws on overflow like the reference
	virtual void update() = 0;
	virtual void clean() = 0;
	bool use_edge() { return _use_edge; }
};

class CUDASimpleVerletList: public CUDABaseList {
protected:
	number _verlet_skin = 0.;
	number _max_density_multiplier = 3.;
	bool _auto_optimisation = true;
	bool _print_problematic_ids = false;

public:
	void get_settings(input_file &inp) override;
	void init(oxb_ctx *ctx, int N, number rcut, int sort_every) override;
	void update() override;
	void clean() override {}
};

class CUDAListFactory {
public:
	static std::shared_ptr<CUDABaseList> make_list(input_file &inp);
};

// ----------------------------------------------------------------------------------------------------- thermostats
class CUDABaseThermostat: public virtual IBaseThermostat {
protected:
	llint _seed = 0;
	oxb_ctx *_ctx = nullptr;
	/// sends the derived parameters to the device context (no RNG state: Philox is keyed by seed, particle id and step)
	virtual void _upload() = 0;

public:
	CUDABaseThermostat() {}
	virtual ~CUDABaseThermostat() {}

	virtual void set_seed(llint seed) { _seed = seed; }
	virtual void get_cuda_settings(input_file &inp) {}
	virtual void attach(oxb_ctx *ctx) { _ctx = ctx; _upload(); }
	/// applies the thermostat at curr_step on its own (inside oxb_run it is fused into the integrator launch)
	virtual void apply_cuda(llint curr_step);
	virtual bool would_activate(llint curr_step) = 0;
};

class CUDANoThermostat: public CUDABaseThermostat, public NoThermostat {
protected:
	void _upload() override;
public:
	void get_settings(input_file &inp) override { NoThermostat::get_settings(inp); }
	void init() override { NoThermostat::init(); }
	bool would_activate(llint) override { return false; }
};

class CUDABrownianThermostat: public CUDABaseThermostat, public BrownianThermostat {
protected:
	void _upload() override;
public:
	void get_settings(input_file &inp) override;
	void init() override;
	bool would_activate(llint curr_step) override;
};

class CUDALangevinThermostat: public CUDABaseThermostat, public LangevinThermostat {
protected:
	void _upload() override;
public:
	void get_settings(input_file &inp) override;
	void init() override;
	bool would_activate(llint) override { return true; }
};

class CUDABussiThermostat: public CUDABaseThermostat, public BussiThermostat {
protected:
	void _upload() override;
public:
	void get_settings(input_file &inp) override;
	void init() override;
	bool would_activate(llint curr_step) override;
};

class CUDAThermostatFactory {
public:
	static std::shared_ptr<CUDABaseThermostat> make_thermostat(input_file &inp, BaseBox *box);
};
