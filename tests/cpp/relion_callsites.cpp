/*
 * Compile check (syntax only, CPU): include/relion_b200_adapter.hpp against the REAL reference headers, with the three call
 * sites of /root/reference/src/ml_optimiser.cpp that touch the accelerator objects (:3577-3632 create, :77-97 fan-out,
 * :3805-3869 drain) spelled with the replacement classes.  FFTW / libtiff / libpng are not installed here: tests/cpp/relion_stubs
 * holds declaration-only stand-ins for fftw3.h, tiffio.h and png.h, nothing is linked.
 *   g++ -std=c++17 -fsyntax-only -Itests/cpp/relion_stubs -I/root/reference -Iinclude tests/cpp/relion_callsites.cpp
 */
#include "src/ml_optimiser.h"
#include "relion_b200_adapter.hpp"

using namespace relion_b200;

// src/ml_optimiser.cpp:3577-3596
void callsite_create(MlOptimiser *mlo, int device)
{
	MlDeviceBundle *b = new MlDeviceBundle(mlo);
	b->setDevice(device);
	b->setupFixedSizedObjects();
	mlo->accDataBundles.push_back((void *) b);
	MlOptimiserCuda *o = new MlOptimiserCuda(mlo, b, "gpu_timing");
	o->resetData();
	mlo->gpuOptimisers.push_back((void *) o);
	size_t allocationSize = b->checkFixedSizedObjects(1);
	b->setupTunableSizedObjects(allocationSize);
}

// src/ml_optimiser.cpp:77-97 (globalThreadExpectationSomeParticles)
void callsite_fanout(MlOptimiser *mlo, int thread_id)
{
	((MlOptimiserCuda *) mlo->gpuOptimisers[thread_id])->doThreadExpectationSomeParticles(thread_id);
}

// src/ml_optimiser.cpp:3805-3869
void callsite_drain(MlOptimiser *mlo)
{
	for (size_t i = 0; i < mlo->accDataBundles.size(); i++)
	{
		MlDeviceBundle *b = (MlDeviceBundle *) mlo->accDataBundles[i];
		b->syncAllBackprojects();
		for (size_t j = 0; j < b->backprojectors.size(); j++)
		{
			unsigned long s = mlo->wsum_model.BPref[j].data.nzyxdim;
			XFLOAT *reals = new XFLOAT[s], *imags = new XFLOAT[s], *weights = new XFLOAT[s];
			b->backprojectors[j].getMdlData(reals, imags, weights);
			for (unsigned long n = 0; n < s; n++)
			{
				mlo->wsum_model.BPref[j].data.data[n].real += (RFLOAT) reals[n];
				mlo->wsum_model.BPref[j].data.data[n].imag += (RFLOAT) imags[n];
				mlo->wsum_model.BPref[j].weight.data[n] += (RFLOAT) weights[n];
			}
			delete[] reals; delete[] imags; delete[] weights;
			b->projectors[j].clear();
			b->backprojectors[j].clear();
		}
	}
	for (size_t i = 0; i < mlo->gpuOptimisers.size(); i++) delete (MlOptimiserCuda *) mlo->gpuOptimisers[i];
	mlo->gpuOptimisers.clear();
	for (size_t i = 0; i < mlo->accDataBundles.size(); i++) delete (MlDeviceBundle *) mlo->accDataBundles[i];
	mlo->accDataBundles.clear();
}

// src/ml_optimiser_mpi.cpp:2028-2185 over NCCL
void callsite_combine(MlOptimiser *mlo, rb_comm *comm)
{
	combineAllWeightedSums(mlo, (MlDeviceBundle *) mlo->accDataBundles[0], comm);
}
