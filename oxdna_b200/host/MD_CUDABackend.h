// MD_CUDABackend / CUDAMixedBackend for the B200 path: the SimBackend the reference's SimManager drives when
// `backend = CUDA`.  Replaces src/CUDA/Backends/{CUDABaseBackend,MD_CUDABackend,MD_CUDAMixedBackend}.{h,cu}.
//
// sim_step() is lazy: SimManager calls it once per MD step, but nothing on the CPU side can look at the particles
// without going through apply_simulation_data_changes() first (SimBackend::print_observables, update_observables_data,
// print_conf, fix_diffusion all do).  So steps are only counted, and a whole run of them is handed to oxb_run() -- one
// device-resident batch, no per-step synchronisation -- when the data is actually needed or the queue limit is reached.
#pragma once

#include "CUDAOperators.h"

#include "Backends/MDBackend.h"

class MD_CUDABackend: public MDBackend {
protected:
	oxb_ctx *_ctx = nullptr;
	int _precision = OXB_PRECISION_FLOAT; // plain MD_CUDABackend = backend_precision float, as in the reference's factory
	int _device_number = -1;
	int _sort_every = 0;
	int _threads_per_block = 0; // accepted, unused: launch shapes are fixed per kernel
	bool _use_edge = false;
	bool _avoid_cpu_calculations = false;
	bool _print_energy = false;
	bool _barostat_always_refresh = false;
	bool _device_observables = false; // CUDA_device_observables: total_energy of the default streams from the device, no CPU round trip
	double _device_UK[2] = { 0., 0. };
	llint _barostat_attempts = 0, _barostat_accepted = 0;
	llint _pending_steps = 0;
	llint _first_pending_step = 0;
	llint _max_pending = 100000;

	std::shared_ptr<CUDABaseInteraction> _cuda_interaction;
	std::shared_ptr<CUDABaseList> _cuda_lists;
	std::shared_ptr<CUDABaseThermostat> _cuda_thermostat;

	virtual void _host_to_gpu();
	virtual void _gpu_to_host();
	virtual void _apply_external_forces_changes();
	virtual void _flush();
	virtual void _apply_barostat();
	void _install_device_observables();
	void _on_T_update() override;

public:
	MD_CUDABackend();
	virtual ~MD_CUDABackend();

	void get_settings(input_file &inp) override;
	void init() override;
	void sim_step() override;
	void fix_diffusion() override;
	void print_observables() override;
	void apply_simulation_data_changes() override;
	void apply_changes_to_simulation_data() override;
};

/// backend_precision = mixed (the default): FP32 pair arithmetic, FP64 state and integration
class CUDAMixedBackend: public MD_CUDABackend {
public:
	CUDAMixedBackend() { _precision = OXB_PRECISION_MIXED; }
};
