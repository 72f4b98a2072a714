/*
 * Test shim for include/relion_b200_adapter.hpp: drives the C++ host classes exactly as MlOptimiser drives the reference's
 * accelerator objects during one E-step (/root/reference/src/ml_optimiser.cpp:3577-3632 create, :4280-4282 fan-out,
 * :3805-3869 drain) and exposes that sequence to the Python tests through one extern "C" function.
 * Built by __graft_entry__.build() into tests/cpp/libadapter_shim.so (g++, links librelion_b200.so).
 */
#include "relion_b200_adapter.hpp"

#include <cstdio>
#include <cstring>

using namespace relion_b200;

extern "C" int adapter_estep(int device, const rb_model *model, const rb_sampling *sampling, int nr_classes,
                             const double *const *PPref, const int *ref_dims /* [K][6]: x y z inity initz r_max */,
                             const int *bp_dims /* [K][6] */, float padding_factor,
                             const rb_particles *pool, rb_pool_out *out, int skip_maximization, int nr_threads,
                             float *const *bp_real, float *const *bp_imag, float *const *bp_weight,
                             char *err, int errlen)
{
	try
	{
		EStepView view;
		view.model = *model; view.sampling = *sampling; view.do_skip_maximization = skip_maximization != 0;
		for (int k = 0; k < nr_classes; k++)
		{
			ClassGeometry g;
			g.PPref_data = PPref[k];
			g.xdim = ref_dims[6 * k]; g.ydim = ref_dims[6 * k + 1]; g.zdim = ref_dims[6 * k + 2];
			g.inity = ref_dims[6 * k + 3]; g.initz = ref_dims[6 * k + 4]; g.r_max = ref_dims[6 * k + 5];
			g.padding_factor = padding_factor;
			g.bp_xdim = bp_dims[6 * k]; g.bp_ydim = bp_dims[6 * k + 1]; g.bp_zdim = bp_dims[6 * k + 2];
			g.bp_inity = bp_dims[6 * k + 3]; g.bp_initz = bp_dims[6 * k + 4]; g.bp_r_max = bp_dims[6 * k + 5];
			view.classes.push_back(g);
		}

		// src/ml_optimiser.cpp:3577-3596
		MlDeviceBundle *b = new MlDeviceBundle(&view);
		b->setDevice(device);
		b->setupFixedSizedObjects();
		std::vector<MlOptimiserCuda *> gpuOptimisers;
		for (int t = 0; t < nr_threads; t++) gpuOptimisers.push_back(new MlOptimiserCuda(&view, b, "shim"));
		b->setupTunableSizedObjects(b->checkFixedSizedObjects(1));

		// :4280-4282 (the OpenMP fan-out, here sequential): every thread calls in, thread 0 carries the pool
		for (int t = 0; t < nr_threads; t++)
		{
			gpuOptimisers[t]->resetData();
			gpuOptimisers[t]->setPool(pool, out);
		}
		for (int t = nr_threads - 1; t >= 0; t--) gpuOptimisers[t]->doThreadExpectationSomeParticles(t);

		// :3805-3869
		b->syncAllBackprojects();
		for (int k = 0; k < nr_classes; k++)
		{
			if (bp_real && bp_real[k]) b->backprojectors[k].getMdlData(bp_real[k], bp_imag[k], bp_weight[k]);
			b->projectors[k].clear();
			b->backprojectors[k].clear();
		}
		for (size_t t = 0; t < gpuOptimisers.size(); t++) delete gpuOptimisers[t];
		delete b;
		return 0;
	}
	catch (const RelionError &e)
	{
		if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
		return -1;
	}
}
