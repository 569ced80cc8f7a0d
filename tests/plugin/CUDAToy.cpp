// TEST INFRASTRUCTURE: a toy third-party interaction for the plugin seam of the drop-in backend (tests/test_dropin.py).
// `interaction_type = Toy` makes the reference's InteractionFactory ask its PluginManager for Toy.so / make_Toy (the CPU object) and our
// CUDAInteractionFactory for CUDAToy.so / make_CUDAToy (src/CUDA/Interactions/CUDAInteractionFactory.cu:44-51,
// src/PluginManagement/PluginManager.cpp:89-180): both entry points live in this one shared object, installed under both names.
// The CPU side simply is the stock DNA2Interaction (topology, observables); the GPU side replaces the force field by a soft repulsion
// between all listed pairs, computed by its own kernel (toy_kernel.cu) from the raw device arrays of the seam.
#include "CUDAOperators.h"

#include "Interactions/DNA2Interaction.h"

extern "C" int toy_force_pass(const oxb_force_views *v, float k, float rc);

class CUDAToyInteraction: public CUDABaseInteraction, public DNA2Interaction {
	float _k = 3.f, _rc = 1.4f;

public:
	void get_settings(input_file &inp) override {
		DNA2Interaction::get_settings(inp);
		getInputFloat(&inp, "toy_k", &_k, 0);
		getInputFloat(&inp, "toy_rc", &_rc, 0);
	}
	void cuda_init(oxb_ctx *ctx, int N) override {
		CUDABaseInteraction::cuda_init(ctx, N);
		DNA2Interaction::init();
		attach_as_plugin(ctx);
	}
	number get_cuda_rcut() override { return (number) _rc; }
	void compute_forces_views(const oxb_force_views &views) override {
		if(toy_force_pass(&views, _k, _rc) != 0) throw oxDNAException("toy kernel launch failed");
	}
};

extern "C" BaseInteraction *make_Toy() { return new DNA2Interaction(); }
extern "C" BaseInteraction *make_CUDAToy() { return new CUDAToyInteraction(); }
