// Our translation unit for BackendFactory::make_backend (declared in the reference's src/Backends/BackendFactory.h).
// Linked INSTEAD of the reference's src/Backends/BackendFactory.cpp -- the only file of the reference that names its CUDA
// classes -- so that `backend = CUDA` constructs the oxdna_b200 backend while every other backend stays the reference's.
#include "Backends/BackendFactory.h"

#include "MD_CUDABackend.h"

#include "Backends/FFS_MD_CPUBackend.h"
#include "Backends/FIREBackend.h"
#include "Backends/MC_CPUBackend.h"
#include "Backends/MC_CPUBackend2.h"
#include "Backends/MD_CPUBackend.h"
#include "Backends/MinBackend.h"
#include "Backends/VMMC_CPUBackend.h"

namespace {

template<typename B>
SimBackend *cpu_only(const std::string &backend_opt, const std::string &sim_type) {
	if(backend_opt != "CPU") throw oxDNAException("Backend '%s' not supported with sim_type = %s", backend_opt.c_str(), sim_type.c_str());
	return new B();
}

} // namespace

std::shared_ptr<SimBackend> BackendFactory::make_backend(input_file &inp) {
	std::string backend_opt, backend_prec, sim_type("MD");
	getInputString(&inp, "backend", backend_opt, 1);
	int precision_state = getInputString(&inp, "backend_precision", backend_prec, 0);
	if(precision_state == KEY_FOUND && backend_opt == "CPU") {
		OX_LOG(Logger::LOG_WARNING, "The 'backend_precision' option cannot be set by input file when running on CPU\n");
	}
	if(getInputString(&inp, "sim_type", sim_type, 0) == KEY_NOT_FOUND) {
		OX_LOG(Logger::LOG_INFO, "Simulation type not specified, using MD");
	}
	else {
		OX_LOG(Logger::LOG_INFO, "Simulation type: %s", sim_type.c_str());
	}

	SimBackend *new_backend = nullptr;
	if(sim_type == "MD") {
		if(backend_opt == "CPU") new_backend = new MD_CPUBackend();
		else if(backend_opt == "CUDA") {
			if(precision_state == KEY_NOT_FOUND) backend_prec = "mixed";
			if(backend_prec == "mixed") new_backend = new CUDAMixedBackend();
			else if(backend_prec == "float") new_backend = new MD_CUDABackend(); // BackendFactory.cpp:62-64; served by the mixed kernels
			else {
				throw oxDNAException("Backend precision '%s' is not allowed, as the oxdna_b200 backend has been compiled with 'float' and 'mixed' support only", backend_prec.c_str());
			}
			OX_LOG(Logger::LOG_INFO, "CUDA backend precision: %s", backend_prec.c_str());
		}
		else throw oxDNAException("Backend '%s' not supported", backend_opt.c_str());
	}
	else if(sim_type == "MC") new_backend = cpu_only<MC_CPUBackend>(backend_opt, sim_type);
	else if(sim_type == "MC2") new_backend = cpu_only<MC_CPUBackend2>(backend_opt, sim_type);
	else if(sim_type == "VMMC") new_backend = cpu_only<VMMC_CPUBackend>(backend_opt, sim_type);
	else if(sim_type == "min") new_backend = cpu_only<MinBackend>(backend_opt, sim_type);
	else if(sim_type == "FIRE") new_backend = cpu_only<FIREBackend>(backend_opt, sim_type);
	else if(sim_type == "FFS_MD") new_backend = cpu_only<FFS_MD_CPUBackend>(backend_opt, sim_type);
	else throw oxDNAException("Simulation type '%s' not supported", sim_type.c_str());

	return std::shared_ptr<SimBackend>(new_backend);
}
