/*
 * relion_b200_adapter.hpp — C++ host side above the C-ABI (include/relion_b200.h), header-only.
 *
 * Mirrors the accelerator boundary MlOptimiser talks to in the reference, with the same class and member names,
 * argument meaning and error behaviour, so that RELION's own call sites keep compiling when library target
 * relion_gpu_util is replaced for the E-step:
 *
 *   AccProjector       /root/reference/src/acc/acc_projector.h:17-104,  acc_projector_impl.h:5-312
 *   AccBackprojector   src/acc/acc_backprojector.h:24-99,               acc_backprojector_impl.h:13-186
 *   MlDeviceBundle     src/acc/cuda/cuda_ml_optimiser.h:18-76,          cuda_ml_optimiser.cu:67-226
 *   MlOptimiserCuda    src/acc/cuda/cuda_ml_optimiser.h:77-144,         cuda_ml_optimiser.cu:228-298
 *
 * What differs, and why:
 *  * The reference classes read their inputs through an `MlOptimiser *baseMLO`.  MlOptimiser cannot be compiled
 *    outside RELION's build (MPI / FFTW / TIFF headers), so the classes here read the same fields through
 *    `EStepView`, a plain struct holding exactly the data contract of SURVEY.md §8b (rb_model + rb_sampling +
 *    PPref / BPref geometry).  Inside RELION's tree `EStepView` is filled from `baseMLO`.
 *  * Device memory belongs to the library (one rb_ctx per device): AccProjector / AccBackprojector are handles
 *    (context, class index) with the reference's geometry members; initMdl uploads, getMdlData downloads.
 *  * The particles of a pool are processed by ONE batched call instead of one accDoExpectationOneParticle per
 *    host thread: MlOptimiserCuda::doThreadExpectationSomeParticles(thread_id) lets thread 0 run the pool that
 *    setPool() staged and returns immediately on every other thread.
 *  * Errors: the reference's HANDLE_ERROR / CRITICAL end in REPORT_ERROR, which throws RelionError
 *    (src/error.h, src/acc/cuda/cuda_settings.h:48-68).  Here RB_REPORT_ERROR throws relion_b200::RelionError with
 *    the library's message; compile with -DRB_REPORT_ERROR=REPORT_ERROR inside RELION to throw its own type.
 *    There is no CPU fallback: without an sm_100 device setDevice() throws.
 */
#ifndef RELION_B200_ADAPTER_HPP_
#define RELION_B200_ADAPTER_HPP_

#include "relion_b200.h"

#include <complex>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

namespace relion_b200 {

typedef float XFLOAT;    // src/acc/settings.h:6-18 (ACC_DOUBLE_PRECISION off)
typedef double RFLOAT;   // src/macros.h (RELION_SINGLE_PRECISION off)

class RelionError : public std::runtime_error   // src/error.h:60-80
{
public:
	std::string msg, file; long line;
	RelionError(const std::string &what, const std::string &fileArg, long lineArg)
	    : std::runtime_error(what + " (" + fileArg + ":" + std::to_string(lineArg) + ")"), msg(what), file(fileArg), line(lineArg) {}
};

#ifndef RB_REPORT_ERROR
#define RB_REPORT_ERROR(message) throw ::relion_b200::RelionError((message), __FILE__, __LINE__)
#endif
#define RB_TRY(call) do { if ((call) != RB_OK) RB_REPORT_ERROR(std::string("relion_b200: ") + rb_last_error()); } while (0)

/* What setupFixedSizedObjects / doThreadExpectationSomeParticles read from MlOptimiser ("data contract read
 * by the driver"): mymodel (PPref, sigma2_noise, pdf_class, pdf_direction, scale_correction, sigma2_offset), sampling,
 * flags, image_*_size — as the C-ABI structs — plus the geometry of every class' projector and back-projector. */
struct ClassGeometry {
	const RFLOAT *PPref_data;          // MlModel::PPref[k].data: MultidimArray<Complex>, (re, im) pairs, [Z][Y][X]
	int xdim, ydim, zdim;              // XSIZE / YSIZE / ZSIZE (zdim == 1: 2D reference)
	int inity, initz;                  // STARTINGY / STARTINGZ
	int r_max; XFLOAT padding_factor;  // Projector::r_max, padding_factor
	int bp_xdim, bp_ydim, bp_zdim, bp_inity, bp_initz, bp_r_max;   // wsum_model.BPref[k].data geometry
};

struct EStepView {
	rb_model model;
	rb_sampling sampling;
	std::vector<ClassGeometry> classes;    // mymodel.nr_classes entries
	bool do_skip_maximization;             // MlOptimiser::do_skip_maximization
	EStepView() : model(), sampling(), do_skip_maximization(false) {}
};

class MlDeviceBundle;

class AccProjector
{
	friend class MlDeviceBundle;
	rb_ctx *ctx; int iclass;
	int mdlX, mdlY, mdlZ, mdlMaxR, mdlInitY, mdlInitZ;
	XFLOAT padding_factor;
	size_t mdlXYZ;
	bool loaded;

public:
	AccProjector() : ctx(NULL), iclass(-1), mdlX(0), mdlY(0), mdlZ(0), mdlMaxR(0), mdlInitY(0), mdlInitZ(0), padding_factor(0), mdlXYZ(0), loaded(false) {}

	/* true when the geometry changed and the model has to be (re)initialised — acc_projector_impl.h:5-22 */
	bool setMdlDim(int xdim, int ydim, int zdim, int inity, int initz, int maxr, XFLOAT paddingFactor)
	{
		if (zdim == 1) zdim = 0;
		if (xdim == mdlX && ydim == mdlY && zdim == mdlZ && inity == mdlInitY && initz == mdlInitZ && maxr == mdlMaxR &&
		    paddingFactor == padding_factor)
			return false;
		clear();
		mdlX = xdim; mdlY = ydim; mdlZ = zdim;
		mdlXYZ = zdim == 0 ? (size_t) xdim * ydim : (size_t) xdim * ydim * zdim;
		mdlInitY = inity; mdlInitZ = initz; mdlMaxR = maxr; padding_factor = paddingFactor;
		return true;
	}

	/* MlModel::PPref[k].data.data (Complex with RFLOAT members) — acc_projector_impl.h:214-248 */
	void initMdl(const std::complex<RFLOAT> *data)
	{
		if (!ctx) RB_REPORT_ERROR("AccProjector::initMdl: projector is not attached to a device bundle");
		RB_TRY(rb_set_reference(ctx, iclass, (const double *) data, mdlX, mdlY, mdlZ == 0 ? 1 : mdlZ, mdlInitY, mdlInitZ, mdlMaxR, padding_factor));
		loaded = true;
	}

	/* separate real / imaginary arrays in XFLOAT — acc_projector_impl.h:112-212 */
	void initMdl(const XFLOAT *real, const XFLOAT *imag)
	{
		if (!ctx) RB_REPORT_ERROR("AccProjector::initMdl: projector is not attached to a device bundle");
		std::vector<XFLOAT> tmp(2 * mdlXYZ);
		for (size_t i = 0; i < mdlXYZ; i++) { tmp[2 * i] = real[i]; tmp[2 * i + 1] = imag[i]; }
		RB_TRY(rb_set_reference_f32(ctx, iclass, tmp.data(), mdlX, mdlY, mdlZ == 0 ? 1 : mdlZ, mdlInitY, mdlInitZ, mdlMaxR, padding_factor));
		loaded = true;
	}

	bool isLoaded() const { return loaded; }
	size_t voxels() const { return mdlXYZ; }

	/* device memory lives in the bundle's context; forgetting the geometry forces the next setMdlDim to report a change */
	void clear() { mdlX = mdlY = mdlZ = mdlMaxR = mdlInitY = mdlInitZ = 0; padding_factor = 0; mdlXYZ = 0; loaded = false; }
};

class AccBackprojector
{
	friend class MlDeviceBundle;
	rb_ctx *ctx; int iclass;

public:
	int mdlX, mdlY, mdlZ, mdlInitY, mdlInitZ, maxR, maxR2;
	XFLOAT padding_factor;
	size_t mdlXYZ;
	size_t voxelCount;

	AccBackprojector() : ctx(NULL), iclass(-1), mdlX(0), mdlY(0), mdlZ(0), mdlInitY(0), mdlInitZ(0), maxR(0), maxR2(0), padding_factor(0), mdlXYZ(0), voxelCount(0) {}

	/* returns the bytes the accumulators take on the device — acc_backprojector_impl.h:13-62 */
	size_t setMdlDim(int xdim, int ydim, int zdim, int inity, int initz, int max_r, XFLOAT paddingFactor)
	{
		if (!ctx) RB_REPORT_ERROR("AccBackprojector::setMdlDim: back-projector is not attached to a device bundle");
		if (zdim < 1) zdim = 1;
		if (xdim != mdlX || ydim != mdlY || zdim != mdlZ || inity != mdlInitY || initz != mdlInitZ || max_r != maxR || paddingFactor != padding_factor)
		{
			mdlX = xdim; mdlY = ydim; mdlZ = zdim;
			mdlXYZ = (size_t) xdim * ydim * zdim;
			mdlInitY = inity; mdlInitZ = initz; maxR = max_r; maxR2 = max_r * max_r; padding_factor = paddingFactor;
			RB_TRY(rb_bp_init(ctx, iclass, mdlX, mdlY, mdlZ, mdlInitY, mdlInitZ, maxR, padding_factor));   // allocates and zeroes
		}
		return mdlXYZ * 4 * sizeof(XFLOAT);   // interleaved (re, im, weight, pad) voxels
	}

	/* zero the accumulators — acc_backprojector_impl.h:64-107 */
	void initMdl()
	{
		RB_TRY(rb_bp_clear(ctx, iclass));
		voxelCount = mdlXYZ;
	}

	/* caller-allocated XFLOAT[mdlXYZ] arrays, as src/ml_optimiser.cpp:3812-3831 hands them in — acc_backprojector_impl.h:109-138 */
	void getMdlData(XFLOAT *real, XFLOAT *imag, XFLOAT *weights) { RB_TRY(rb_bp_get(ctx, iclass, real, imag, weights)); }

	/* the device accumulator itself, for an in-place NCCL all-reduce (interleaved floats) */
	void getMdlDevicePtr(void *&dptr, size_t &n_floats) { RB_TRY(rb_bp_device_buffer(ctx, iclass, &dptr, &n_floats)); }

	void clear() { mdlX = mdlY = mdlZ = mdlInitY = mdlInitZ = maxR = maxR2 = 0; padding_factor = 0; mdlXYZ = 0; voxelCount = 0; }
};

class MlDeviceBundle
{
public:
	std::vector<AccProjector> projectors;          // one per class (per body in multi-body refinement: not covered)
	std::vector<AccBackprojector> backprojectors;
	const EStepView *baseMLO;
	rb_ctx *ctx;
	int device_id;
	int rank_shared_count;

	MlDeviceBundle(const EStepView *baseMLOptimiser) : baseMLO(baseMLOptimiser), ctx(NULL), device_id(-1), rank_shared_count(1) {}

	/* the reference only records the id and calls cudaSetDevice later; the context is created here so that a missing
	 * device fails where the reference's first HANDLE_ERROR(cudaSetDevice) would */
	void setDevice(int did)
	{
		if (ctx) { rb_ctx_destroy(ctx); ctx = NULL; }
		device_id = did;
		RB_TRY(rb_ctx_create(did, &ctx));
	}

	/* bytes the fixed-size objects need on the device — cuda_ml_optimiser.cu:67-83 */
	size_t checkFixedSizedObjects(int shares)
	{
		size_t bytes = 0;
		for (size_t k = 0; k < baseMLO->classes.size(); k++)
		{
			const ClassGeometry &g = baseMLO->classes[k];
			bytes += (size_t) g.xdim * g.ydim * (g.zdim < 1 ? 1 : g.zdim) * (8 + 64 + 16);     // compact + expanded + x-pair reference
			bytes += (size_t) g.bp_xdim * g.bp_ydim * (g.bp_zdim < 1 ? 1 : g.bp_zdim) * 16;  // interleaved accumulator
		}
		return bytes * (size_t) (shares < 1 ? 1 : shares);
	}

	/* cuda_ml_optimiser.cu:85-152: model + sampling tables, then per class projector and back-projector */
	void setupFixedSizedObjects()
	{
		if (!ctx) RB_REPORT_ERROR("MlDeviceBundle::setupFixedSizedObjects: setDevice() has not been called");
		const int K = baseMLO->model.nr_classes;
		if ((int) baseMLO->classes.size() != K) RB_REPORT_ERROR("MlDeviceBundle::setupFixedSizedObjects: one ClassGeometry per class expected");
		RB_TRY(rb_set_model(ctx, &baseMLO->model));
		RB_TRY(rb_set_sampling(ctx, &baseMLO->sampling));
		if (baseMLO->model.pdf_direction) RB_TRY(rb_set_pdf_direction(ctx, baseMLO->model.pdf_direction));
		projectors.resize(K); backprojectors.resize(K);
		for (int k = 0; k < K; k++)
		{
			const ClassGeometry &g = baseMLO->classes[k];
			projectors[k].ctx = ctx; projectors[k].iclass = k;
			projectors[k].setMdlDim(g.xdim, g.ydim, g.zdim, g.inity, g.initz, g.r_max, g.padding_factor);
			projectors[k].initMdl((const std::complex<RFLOAT> *) g.PPref_data);
			backprojectors[k].ctx = ctx; backprojectors[k].iclass = k;
			backprojectors[k].setMdlDim(g.bp_xdim, g.bp_ydim, g.bp_zdim, g.bp_inity, g.bp_initz, g.bp_r_max, g.padding_factor);
			backprojectors[k].initMdl();
		}
	}

	/* allocator sizing and coarse projection plans of the reference (cuda_ml_optimiser.cu:154-226): the library grows its
	 * buffers on demand and builds the coarse matrices on the device, so there is nothing to size */
	void setupTunableSizedObjects(size_t /*allocationSize*/) {}

	void syncAllBackprojects() { RB_TRY(rb_sync(ctx)); }

	~MlDeviceBundle()
	{
		projectors.clear();
		backprojectors.clear();
		if (ctx) rb_ctx_destroy(ctx);
	}

private:
	MlDeviceBundle(const MlDeviceBundle &);
	MlDeviceBundle &operator=(const MlDeviceBundle &);
};

class MlOptimiserCuda
{
public:
	const EStepView *baseMLO;
	MlDeviceBundle *bundle;
	int device_id;
	std::string timing_name;

	MlOptimiserCuda(const EStepView *baseMLOptimiser, MlDeviceBundle *b, const char *timing_fnm)
	    : baseMLO(baseMLOptimiser), bundle(b), device_id(b->device_id), timing_name(timing_fnm ? timing_fnm : ""), pool(NULL), out(NULL) {}

	void resetData() { pool = NULL; out = NULL; }   // cuda_ml_optimiser.cu:228-248 (per-thread streams / buffers: none here)

	/* the pool the next doThreadExpectationSomeParticles works on: what getFourierTransformsAndCtfs left for the particles
	 * exp_my_first_part_id .. exp_my_last_part_id, and where the per-particle results go */
	void setPool(const rb_particles *p, rb_pool_out *o) { pool = p; out = o; }

	/* cuda_ml_optimiser.cu:250-298.  Called concurrently from nr_threads OpenMP threads (src/ml_optimiser.cpp:4280); the
	 * whole pool is one batched device call, so thread 0 runs it and the other threads have nothing to pull. */
	void doThreadExpectationSomeParticles(int thread_id)
	{
		if (thread_id != 0) return;
		if (!pool || !out) RB_REPORT_ERROR("MlOptimiserCuda::doThreadExpectationSomeParticles: setPool() has not been called");
		RB_TRY(rb_estep_pool(bundle->ctx, pool, out, baseMLO->do_skip_maximization ? 1u : 0u));
	}

private:
	const rb_particles *pool;
	rb_pool_out *out;
};

} // namespace relion_b200

#endif
