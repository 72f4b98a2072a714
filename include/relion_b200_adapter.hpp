/*
 * relion_b200_adapter.hpp — C++ host side above the C-ABI (include/relion_b200.h), header-only.
 *
 * Replacement for the accelerator objects MlOptimiser talks to in the reference, with the same class and member names,
 * constructor signatures, call sequence and error behaviour, so that RELION's own call sites
 *   /root/reference/src/ml_optimiser.cpp:3577-3632 (create), :77-97 + :4280 (fan-out), :3805-3869 (drain)
 * keep compiling when library target relion_gpu_util is replaced for the E-step:
 *
 *   AccProjector       src/acc/acc_projector.h:17-104,          acc_projector_impl.h:5-312
 *   AccBackprojector   src/acc/acc_backprojector.h:24-99,       acc_backprojector_impl.h:13-186
 *   MlDeviceBundle     src/acc/cuda/cuda_ml_optimiser.h:18-76,  cuda_ml_optimiser.cu:67-226       MlDeviceBundle(MlOptimiser *)
 *   MlOptimiserCuda    src/acc/cuda/cuda_ml_optimiser.h:77-144, cuda_ml_optimiser.cu:228-298      MlOptimiserCuda(MlOptimiser *, MlDeviceBundle *, const char *)
 *
 * Include RELION's "src/ml_optimiser.h" BEFORE this header (it provides MlOptimiser, MlModel, MlWsumModel, Experiment,
 * HealpixSampling, CTF, MultidimArray and the XSIZE / DIRECT_A2D_ELEM / METADATA_* macros used below).  The tests of this
 * repository include tests/cpp/mock_relion/src/ml_optimiser.h instead, a from-scratch stand-in with the same member names;
 * tests/test_adapter_cpp.py also compiles this header against the real reference headers when /root/reference is there.
 *
 * What the classes do, per E-step:
 *   MlDeviceBundle::setupFixedSizedObjects()           mymodel / sampling / flags -> rb_set_model, rb_set_sampling; PPref[k] ->
 *                                                      AccProjector::initMdl; wsum_model.BPref[k] geometry -> AccBackprojector
 *   MlOptimiserCuda::doThreadExpectationSomeParticles  the pool exp_my_first_part_id .. exp_my_last_part_id in ONE batched call:
 *        exp_imagedata + exp_metadata (+ the prior-selected orientation lists of local searches) -> rb_pool_prepare
 *        (getFourierTransformsAndCtfs on the device) -> rb_estep_slot -> the host bookkeeping of storeWeightedSums
 *        (acc_ml_optimiser_impl.h:2858-2931 metadata row, :3466-3656 weighted sums under omp critical) in fp64.
 *        Called concurrently from nr_threads OpenMP threads like the reference's; thread 0 carries the pool.
 *   pullBackprojectors()                               src/ml_optimiser.cpp:3805-3838: getMdlData += wsum_model.BPref[k]
 *   combineAllWeightedSums()                           src/ml_optimiser_mpi.cpp:2028-2185 over NCCL instead of MPI: accumulators
 *                                                      summed on the devices (rb_bp_allreduce), everything else as one fp64
 *                                                      vector in MlWsumModel::pack order (WsumPack, src/ml_model.cpp:1881-2049)
 * Scope limits (RB_REPORT_ERROR when violated): nr_bodies == 1, one image per particle, 2D images, no helical refinement
 * (--helix: translations in helical coordinates are per-particle tables), no tomo, no --only_sample_tilt.  Covered beyond the plain
 * case: optics groups with their own box / pixel size and anisotropic magnification (setGeometry, mat_left), --skip_align /
 * --skip_rotate (setSampling per pool, pre_shift), gradient refinement (do_grad, pseudo half-sets), both criteria.
 * Errors: the reference's HANDLE_ERROR / CRITICAL end in REPORT_ERROR, which throws RelionError (src/error.h,
 * src/acc/cuda/cuda_settings.h:48-68).  RB_REPORT_ERROR throws relion_b200::RelionError; compile with
 * -DRB_REPORT_ERROR=REPORT_ERROR inside RELION to throw its own type.  There is no CPU fallback.
 */
#ifndef RELION_B200_ADAPTER_HPP_
#define RELION_B200_ADAPTER_HPP_

#include "relion_b200.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace relion_b200 {

#ifndef XFLOAT
typedef float XFLOAT;    // src/acc/settings.h:6-18 (ACC_DOUBLE_PRECISION off); RELION's own header defines it as a macro
#endif

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

class MlDeviceBundle;

// ---------------------------------------------------------------------------------------------------------------------
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
	void initMdl(const Complex *data)
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

// ---------------------------------------------------------------------------------------------------------------------
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

	/* the device accumulator itself (interleaved floats) */
	void getMdlDevicePtr(void *&dptr, size_t &n_floats) { RB_TRY(rb_bp_device_buffer(ctx, iclass, &dptr, &n_floats)); }

	void clear() { mdlX = mdlY = mdlZ = mdlInitY = mdlInitZ = maxR = maxR2 = 0; padding_factor = 0; mdlXYZ = 0; voxelCount = 0; }
};

// ---------------------------------------------------------------------------------------------------------------------
// MlWsumModel::pack / unpack (src/ml_model.cpp:1881-2049) without the BPref volumes (those are summed on the devices):
// LL, ave_Pmax, sigma2_offset, avg_norm_correction, sigma2_rot, sigma2_tilt, sigma2_psi; per optics group sigma2_noise,
// sumw_ctf2, sumw_stMulti, sumw_group; per group wsum_signal_product, wsum_reference_power; per class pdf_direction; per class
// pdf_class (+ prior_offset_class x, y for 2D references).
struct WsumPack
{
	template <typename Op> static void walk(MlWsumModel &w, Op op)
	{
		op(w.LL); op(w.ave_Pmax); op(w.sigma2_offset); op(w.avg_norm_correction); op(w.sigma2_rot); op(w.sigma2_tilt); op(w.sigma2_psi);
		for (int g = 0; g < w.nr_optics_groups; g++)
		{
			for (long int n = 0; n < MULTIDIM_SIZE(w.sigma2_noise[g]); n++) op(DIRECT_MULTIDIM_ELEM(w.sigma2_noise[g], n));
			if ((int) w.sumw_ctf2.size() > g) for (long int n = 0; n < MULTIDIM_SIZE(w.sumw_ctf2[g]); n++) op(DIRECT_MULTIDIM_ELEM(w.sumw_ctf2[g], n));
			if ((int) w.sumw_stMulti.size() > g) for (long int n = 0; n < MULTIDIM_SIZE(w.sumw_stMulti[g]); n++) op(DIRECT_MULTIDIM_ELEM(w.sumw_stMulti[g], n));
			op(w.sumw_group[g]);
		}
		for (int g = 0; g < w.nr_groups; g++) { op(w.wsum_signal_product[g]); op(w.wsum_reference_power[g]); }
		for (int k = 0; k < w.nr_classes * w.nr_bodies; k++)
			for (long int n = 0; n < MULTIDIM_SIZE(w.pdf_direction[k]); n++) op(DIRECT_MULTIDIM_ELEM(w.pdf_direction[k], n));
		for (int k = 0; k < w.nr_classes; k++)
		{
			op(w.pdf_class[k]);
			if (w.ref_dim == 2) { op(XX(w.prior_offset_class[k])); op(YY(w.prior_offset_class[k])); }
		}
	}
	struct Push { std::vector<double> *v; void operator()(RFLOAT &x) { v->push_back((double) x); } };
	struct Pull { const double *p; void operator()(RFLOAT &x) { x = (RFLOAT) *p++; } };
	static void pack(MlWsumModel &w, std::vector<double> &packed) { packed.clear(); Push p = {&packed}; walk(w, p); }
	static void unpack(MlWsumModel &w, const std::vector<double> &packed) { Pull p = {packed.data()}; walk(w, p); }
};

// ---------------------------------------------------------------------------------------------------------------------
class MlDeviceBundle
{
public:
	std::vector<AccProjector> projectors;          // one per class (per body in multi-body refinement: not covered)
	std::vector<AccBackprojector> backprojectors;
	MlOptimiser *baseMLO;
	rb_ctx *ctx;
	int device_id;
	int rank_shared_count;

	int geometry_og;                               // the optics group whose image geometry the library currently holds (-1: none)

	MlDeviceBundle(MlOptimiser *baseMLOptimiser) : baseMLO(baseMLOptimiser), ctx(NULL), device_id(-1), rank_shared_count(1), geometry_og(-1) {}

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
		for (int k = 0; k < baseMLO->mymodel.nr_classes; k++)
		{
			const MultidimArray<Complex> &d = baseMLO->mymodel.PPref[k].data;
			bytes += (size_t) XSIZE(d) * YSIZE(d) * (ZSIZE(d) < 1 ? 1 : ZSIZE(d)) * (8 + 64 + 16);     // compact + expanded + x-pair reference
			const MultidimArray<Complex> &b = baseMLO->wsum_model.BPref[k].data;
			bytes += (size_t) XSIZE(b) * YSIZE(b) * (ZSIZE(b) < 1 ? 1 : ZSIZE(b)) * 16;                 // interleaved accumulator
		}
		return bytes * (size_t) (shares < 1 ? 1 : shares);
	}

	/* Optics groups a and b share one image geometry (box, pixel size, current and coarse window: src/ml_optimiser.cpp:5735-5777) */
	bool sameGeometry(int a, int b)
	{
		MlOptimiser &o = *baseMLO;
		return o.image_full_size[a] == o.image_full_size[b] && o.image_current_size[a] == o.image_current_size[b] &&
		       o.image_coarse_size[a] == o.image_coarse_size[b] &&
		       std::fabs(o.mydata.getOpticsPixelSize(a) - o.mydata.getOpticsPixelSize(b)) <= 1e-6 * o.mydata.getOpticsPixelSize(b);
	}

	/* (image box * image pixel) / (model box * model pixel) of optics group og, the factor of applyScaleDifference */
	RFLOAT remapSizes(int og)
	{
		MlOptimiser &o = *baseMLO;
		return (o.mydata.getOpticsPixelSize(og) * o.mydata.getOpticsImageSize(og)) / (o.mymodel.pixel_size * o.mymodel.ori_size);
	}

	/* Model and sampling tables in the image geometry of optics group og (include/relion_b200.h "Optics groups"): its box is the
	 * library's ori_size, its image_current_size / image_coarse_size the windows, its pixel size converts the translations
	 * (getTranslationsInPixel(..., my_pixel_size), acc_ml_optimiser_impl.h:1203, 1511) and every sigma2_noise spectrum is read at
	 * ROUND(remap * ires) (src/ml_optimiser.cpp:6840, 6875).  A no-op when the loaded geometry already is that of og. */
	void setGeometry(int og)
	{
		if (!ctx) RB_REPORT_ERROR("MlDeviceBundle::setGeometry: setDevice() has not been called");
		if (geometry_og >= 0 && sameGeometry(og, geometry_og)) return;
		MlOptimiser &o = *baseMLO;
		MlModel &m = o.mymodel;
		const int K = m.nr_classes, nog = m.nr_optics_groups;
		const int full = o.image_full_size[og], nshell = full / 2 + 1;
		const RFLOAT my_pixel_size = o.mydata.getOpticsPixelSize(og);

		// ---- model / flags (data contract of SURVEY.md 8b) ----
		h_sigma2.assign((size_t) nog * nshell, std::numeric_limits<double>::infinity());      // 1 / inf = 0: Minvsigma2 stays zero (:6877)
		for (int g = 0; g < nog; g++)
		{
			const RFLOAT remap = 1. / remapSizes(g);                                             // (ori * pixel) / (my_image_size * my_pixel_size), :6840
			for (int i = 0; i < nshell; i++)
			{
				const int ir = ROUND(remap * i);
				if (ir < (int) XSIZE(m.sigma2_noise[g])) h_sigma2[(size_t) g * nshell + i] = DIRECT_A1D_ELEM(m.sigma2_noise[g], ir);
			}
		}
		h_scale.assign(m.scale_correction.begin(), m.scale_correction.end());
		if (h_scale.empty()) h_scale.assign(m.nr_groups < 1 ? 1 : m.nr_groups, 1.);
		h_pdf_class.assign(m.pdf_class.begin(), m.pdf_class.end());
		h_dvp.assign((size_t) K * nshell, 0.);
		for (int k = 0; k < K && k < (int) m.data_vs_prior_class.size(); k++)
			for (int i = 0; i < nshell && i < (int) XSIZE(m.data_vs_prior_class[k]); i++) h_dvp[(size_t) k * nshell + i] = DIRECT_A1D_ELEM(m.data_vs_prior_class[k], i);
		const int n_dir = (int) o.sampling.NrDirections();
		const bool skip = o.do_skip_align || o.do_skip_rotate;
		const bool use_priors = m.orientational_prior_mode != NOPRIOR && !skip;
		h_pdf_dir.clear();
		if (!use_priors && !skip)
		{
			h_pdf_dir.assign((size_t) K * n_dir, 0.);
			for (int k = 0; k < K; k++)
				for (int d = 0; d < n_dir && d < (int) MULTIDIM_SIZE(m.pdf_direction[k]); d++) h_pdf_dir[(size_t) k * n_dir + d] = DIRECT_MULTIDIM_ELEM(m.pdf_direction[k], d);
		}
		rb_model rm;
		memset(&rm, 0, sizeof(rm));
		rm.nr_classes = K; rm.ori_size = full; rm.coarse_size = o.image_coarse_size[og]; rm.current_size = o.image_current_size[og];
		rm.pixel_size = my_pixel_size;
		// references that end inside this group's window (its box is bigger than the model's): the fine-pass row rule
		{
			int ref_r = m.PPref.empty() ? 0 : m.PPref[0].r_max;
			for (int k = 1; k < K && k < (int) m.PPref.size(); k++) ref_r = std::min(ref_r, (int) m.PPref[k].r_max);
			rm.ref_max_r = (ref_r > 0 && ref_r < rm.current_size / 2) ? ref_r : 0;
		}
		rm.nr_optics_groups = nog; rm.sigma2_noise = h_sigma2.data();
		rm.nr_groups = (int) h_scale.size(); rm.scale_correction = h_scale.data();
		rm.pdf_class = h_pdf_class.data(); rm.pdf_direction = h_pdf_dir.empty() ? NULL : h_pdf_dir.data(); rm.data_vs_prior_class = h_dvp.data();
		rm.sigma2_offset = m.sigma2_offset; rm.offset_range = o.offset_range_x;            // acc_ml_optimiser_impl.h:1916-1926
		rm.sigma2_fudge = o.sigma2_fudge; rm.adaptive_fraction = o.adaptive_fraction; rm.maximum_significants = o.maximum_significants;
		rm.do_ctf_correction = o.do_ctf_correction; rm.refs_are_ctf_corrected = o.refs_are_ctf_corrected;
		rm.do_scale_correction = o.do_scale_correction; rm.do_map = o.do_map;
		rm.ctf_premultiplied = o.mydata.obsModel.getCtfPremultiplied(0);
		rm.bp_circle_bound = 1;
		rm.do_cc = (o.iter == 1 && o.do_firstiter_cc) || o.do_always_cc;                   // :1164
		rm.do_grad = o.do_grad ? 1 : 0;                                                    // acc_ml_optimiser_impl.h:3418
		rm.do_skip_rotate = skip ? 1 : 0;                                                  // :1966-1967, :3752-3766
		h_prior_class.clear();
		if (m.ref_dim == 2 && m.nr_bodies == 1)                                             // :2100-2104, :2673-2677
		{
			for (int k = 0; k < K; k++) { h_prior_class.push_back(XX(m.prior_offset_class[k])); h_prior_class.push_back(YY(m.prior_offset_class[k])); }
			rm.prior_offset_class = h_prior_class.data();
		}
		RB_TRY(rb_set_model(ctx, &rm));
		geometry_og = og;
		setSampling();
	}

	/* Sampling tables of the loaded geometry (HealpixSampling stays RELION's: getOrientations :1832, getTranslationsInPixel :1724).
	 * With --skip_align / --skip_rotate RELION refills the sampling object with the pool's own orientations before every pool
	 * (src/ml_optimiser.cpp:4180-4225), so the pool path calls this again for every pool; particle row p then uses entry p of the
	 * direction and psi tables, and with --skip_align its own translation (entry p) travels as rb_particles.pre_shift while the
	 * library samples the single translation (0, 0). */
	void setSampling()
	{
		MlOptimiser &o = *baseMLO;
		const int og = geometry_og;
		const RFLOAT my_pixel_size = o.mydata.getOpticsPixelSize(og);
		const bool skip = o.do_skip_align || o.do_skip_rotate;
		const int n_dir = (int) o.sampling.NrDirections(), n_psi = (int) o.sampling.NrPsiSamplings();
		const int ov = o.adaptive_oversampling;
		if (skip && ov != 0) RB_REPORT_ERROR("relion_b200: --skip_align / --skip_rotate run without oversampling (src/ml_optimiser.cpp:2382-2389)");
		const int nor = o.sampling.oversamplingFactorOrientations(ov), not_ = o.sampling.oversamplingFactorTranslations(ov);
		const int n_trans = o.do_skip_align ? 1 : (int) o.sampling.NrTranslationalSamplings();      // --skip_align: (0, 0), the particle's own goes into pre_shift
		std::vector<RFLOAT> a, b, c;
		std::vector<int> no_ptr; std::vector<RFLOAT> no_prior;
		s_rot.assign(n_dir, 0.); s_tilt.assign(n_dir, 0.); s_psi.assign(n_psi, 0.);
		s_orot.clear(); s_otilt.clear(); s_opsi.clear();
		if (skip)
		{
			// only (d, d) is ever used: one pass over the directions and one over the psi angles
			for (int d = 0; d < n_dir; d++) { o.sampling.getOrientations(d, 0, 0, a, b, c, no_ptr, no_prior, no_ptr, no_prior); s_rot[d] = a[0]; s_tilt[d] = b[0]; }
			for (int p = 0; p < n_psi; p++) { o.sampling.getOrientations(0, p, 0, a, b, c, no_ptr, no_prior, no_ptr, no_prior); s_psi[p] = c[0]; }
		}
		else
		{
		s_orot.assign((size_t) n_dir * n_psi * nor, 0.); s_otilt = s_orot; s_opsi = s_orot;
		for (int d = 0; d < n_dir; d++)
			for (int p = 0; p < n_psi; p++)
			{
				o.sampling.getOrientations(d, p, 0, a, b, c, no_ptr, no_prior, no_ptr, no_prior);
				s_rot[d] = a[0]; s_tilt[d] = b[0]; s_psi[p] = c[0];
				o.sampling.getOrientations(d, p, ov, a, b, c, no_ptr, no_prior, no_ptr, no_prior);
				for (int i = 0; i < nor; i++) { const size_t g = ((size_t) d * n_psi + p) * nor + i; s_orot[g] = a[i]; s_otilt[g] = b[i]; s_opsi[g] = c[i]; }
			}
		}
		s_tx.assign(n_trans, 0.); s_ty = s_tx; s_otx.assign((size_t) n_trans * not_, 0.); s_oty = s_otx;
		for (int t = 0; t < n_trans && !o.do_skip_align; t++)
		{
			o.sampling.getTranslationsInPixel(t, 0, my_pixel_size, a, b, c, false);
			s_tx[t] = a[0]; s_ty[t] = b[0];
			o.sampling.getTranslationsInPixel(t, ov, my_pixel_size, a, b, c, false);
			for (int i = 0; i < not_; i++) { s_otx[(size_t) t * not_ + i] = a[i]; s_oty[(size_t) t * not_ + i] = b[i]; }
		}
		rb_sampling rs;
		memset(&rs, 0, sizeof(rs));
		rs.n_dir = n_dir; rs.n_psi = n_psi; rs.rot = s_rot.data(); rs.tilt = s_tilt.data(); rs.psi = s_psi.data();
		rs.n_over_rot = nor;
		if (!s_orot.empty()) { rs.over_rot = s_orot.data(); rs.over_tilt = s_otilt.data(); rs.over_psi = s_opsi.data(); }
		rs.n_trans = n_trans; rs.trans_x = s_tx.data(); rs.trans_y = s_ty.data();
		rs.n_over_trans = not_; rs.over_trans_x = s_otx.data(); rs.over_trans_y = s_oty.data();
		RB_TRY(rb_set_sampling(ctx, &rs));
		if (!h_pdf_dir.empty()) RB_TRY(rb_set_pdf_direction(ctx, h_pdf_dir.data()));
	}

	/* cuda_ml_optimiser.cu:85-152: model + sampling tables, then per class projector and back-projector */
	void setupFixedSizedObjects()
	{
		if (!ctx) RB_REPORT_ERROR("MlDeviceBundle::setupFixedSizedObjects: setDevice() has not been called");
		MlOptimiser &o = *baseMLO;
		MlModel &m = o.mymodel;
		if (m.nr_bodies != 1) RB_REPORT_ERROR("relion_b200: multi-body refinement is not covered");
		if (o.do_helical_refine) RB_REPORT_ERROR("relion_b200: helical refinement (translations in helical coordinates) is not covered");
		if (m.data_dim != 2) RB_REPORT_ERROR("relion_b200: 3D data (subtomograms) are not covered");
		const int K = m.nr_classes;
		geometry_og = -1;
		setGeometry(0);

		// ---- projectors / back-projectors (cuda_ml_optimiser.cu:98-152) ----
		// gradient refinement with pseudo half-sets: wsum_model.BPref holds 2 K accumulators, particle part_id goes into
		// iclass + (part_id % 2) * K (acc_ml_optimiser_impl.h:3395-3400, cuda_ml_optimiser.cu:121-141)
		const int nbp = o.grad_pseudo_halfsets ? 2 * K : K;
		if ((int) o.wsum_model.BPref.size() < nbp) RB_REPORT_ERROR("MlDeviceBundle::setupFixedSizedObjects: wsum_model.BPref holds fewer accumulators than the pseudo half-sets need");
		projectors.resize(K); backprojectors.resize(nbp);
		for (int k = K; k < nbp; k++)
		{
			BackProjector &bp = o.wsum_model.BPref[k];
			backprojectors[k].ctx = ctx; backprojectors[k].iclass = k;
			backprojectors[k].setMdlDim((int) XSIZE(bp.data), (int) YSIZE(bp.data), (int) ZSIZE(bp.data), (int) STARTINGY(bp.data), (int) STARTINGZ(bp.data),
			                            bp.r_max, (XFLOAT) bp.padding_factor);
			backprojectors[k].initMdl();
		}
		for (int k = 0; k < K; k++)
		{
			Projector &pp = m.PPref[k];
			projectors[k].ctx = ctx; projectors[k].iclass = k;
			projectors[k].setMdlDim((int) XSIZE(pp.data), (int) YSIZE(pp.data), (int) ZSIZE(pp.data), (int) STARTINGY(pp.data), (int) STARTINGZ(pp.data),
			                        pp.r_max, (XFLOAT) pp.padding_factor);
			projectors[k].initMdl(MULTIDIM_ARRAY(pp.data));
			BackProjector &bp = o.wsum_model.BPref[k];
			backprojectors[k].ctx = ctx; backprojectors[k].iclass = k;
			backprojectors[k].setMdlDim((int) XSIZE(bp.data), (int) YSIZE(bp.data), (int) ZSIZE(bp.data), (int) STARTINGY(bp.data), (int) STARTINGZ(bp.data),
			                            bp.r_max, (XFLOAT) bp.padding_factor);
			backprojectors[k].initMdl();
		}
	}

	/* allocator sizing and coarse projection plans of the reference (cuda_ml_optimiser.cu:154-226): the library grows its
	 * buffers on demand and builds the coarse matrices on the device, so there is nothing to size */
	void setupTunableSizedObjects(size_t /*allocationSize*/) {}

	void syncAllBackprojects() { RB_TRY(rb_sync(ctx)); }

	/* src/ml_optimiser.cpp:3805-3838: every class' device accumulator added to wsum_model.BPref[k] (double) on the host */
	void pullBackprojectors()
	{
		syncAllBackprojects();
		for (size_t k = 0; k < backprojectors.size(); k++)
		{
			BackProjector &bp = baseMLO->wsum_model.BPref[k];
			const size_t n = backprojectors[k].mdlXYZ;
			std::vector<XFLOAT> re(n), im(n), w(n);
			backprojectors[k].getMdlData(re.data(), im.data(), w.data());
			for (size_t i = 0; i < n; i++)
			{
				DIRECT_MULTIDIM_ELEM(bp.data, i).real += (RFLOAT) re[i];
				DIRECT_MULTIDIM_ELEM(bp.data, i).imag += (RFLOAT) im[i];
				DIRECT_MULTIDIM_ELEM(bp.weight, i) += (RFLOAT) w[i];
			}
		}
	}

	~MlDeviceBundle()
	{
		projectors.clear();
		backprojectors.clear();
		if (ctx) rb_ctx_destroy(ctx);
	}

private:
	// host copies the C-ABI structs point into (alive until the next setupFixedSizedObjects)
	std::vector<double> h_sigma2, h_scale, h_pdf_class, h_prior_class, h_pdf_dir, h_dvp, s_rot, s_tilt, s_psi, s_orot, s_otilt, s_opsi, s_tx, s_ty, s_otx, s_oty;
	MlDeviceBundle(const MlDeviceBundle &);
	MlDeviceBundle &operator=(const MlDeviceBundle &);
};

// ---------------------------------------------------------------------------------------------------------------------
class MlOptimiserCuda
{
public:
	MlOptimiser *baseMLO;
	MlDeviceBundle *bundle;
	int device_id;
	std::string timing_name;

	MlOptimiserCuda(MlOptimiser *baseMLOptimiser, MlDeviceBundle *b, const char *timing_fnm)
	    : baseMLO(baseMLOptimiser), bundle(b), device_id(b->device_id), timing_name(timing_fnm ? timing_fnm : "") {}

	void resetData() {}   // cuda_ml_optimiser.cu:228-248 (per-thread streams / buffers: none here, the library owns them)

	/* cuda_ml_optimiser.cu:250-298.  Called concurrently from nr_threads OpenMP threads (src/ml_optimiser.cpp:4280, :77-97); the
	 * pool is ONE batched device call, so thread 0 carries it and the other threads have nothing to pull. */
	void doThreadExpectationSomeParticles(int thread_id)
	{
		if (thread_id != 0) return;
		MlOptimiser &o = *baseMLO;
		const long int first = o.exp_my_first_part_id, last = o.exp_my_last_part_id;
		// one device call per run of particles that share an image geometry (src/ml_optimiser.cpp:6802-6821: every particle is
		// processed at its optics group's sizes; exp_imagedata itself holds one common box, :10312-10323)
		long int run_first = first;
		while (run_first <= last)
		{
			const int og0 = o.mydata.getOpticsGroup(run_first);
			long int run_last = run_first;
			while (run_last + 1 <= last && bundle->sameGeometry(o.mydata.getOpticsGroup(run_last + 1), og0)) run_last++;
			bundle->setGeometry(og0);
			if (o.do_skip_align || o.do_skip_rotate) bundle->setSampling();       // the sampling object holds THIS pool's orientations (:4180-4225)
			expectationRun(run_first, run_last);
			run_first = run_last + 1;
		}
	}

private:
	/* particles run_first .. run_last of the pool (exp_metadata / exp_imagedata rows run_first - exp_my_first_part_id ...), all of
	 * the image geometry the bundle currently holds */
	void expectationRun(long int first, long int last)
	{
		MlOptimiser &o = *baseMLO;
		MlModel &m = o.mymodel;
		const int row0 = (int) (first - o.exp_my_first_part_id);
		const int P = (int) (last - first + 1);
		if (P <= 0) return;
		const int og0 = o.mydata.getOpticsGroup(first);
		const int n = o.image_full_size[og0], nshell = n / 2 + 1, K = m.nr_classes;
		const int cur = o.image_current_size[og0];
		const RFLOAT my_pixel_size = o.mydata.getOpticsPixelSize(og0);
		if ((int) YSIZE(o.exp_imagedata) != n) RB_REPORT_ERROR("relion_b200: exp_imagedata does not have the box size of the particles' optics group");
		const bool skip = o.do_skip_align || o.do_skip_rotate;
		const bool use_priors = m.orientational_prior_mode != NOPRIOR && !skip;
		const bool do_cc = (o.iter == 1 && o.do_firstiter_cc) || o.do_always_cc;
		const int nog = m.nr_optics_groups;

		// ---- the pool: what getFourierTransformsAndCtfs starts from (acc_ml_optimiser_impl.h:11-430) ----
		std::vector<float> images((size_t) P * n * n);
		std::vector<double> norm_factor(P, 1.), old_offset(2 * (size_t) P), prior_offset(2 * (size_t) P), defU(P), defV(P), defA(P), bfac(P), kfac(P), phs(P);
		std::vector<double> og_kV(nog, 300.), og_Cs(nog, 2.7), og_Q0(nog, 0.1);
		std::vector<int> group_id(P), optics_group(P), dir_off, psi_off, dir_idx, psi_idx;
		std::vector<double> dir_prior, psi_prior;
		std::vector<std::vector<int> > ptr_dir(P), ptr_psi(P);
		std::vector<std::vector<RFLOAT> > pri_dir(P), pri_psi(P);
		if (use_priors || skip) { dir_off.push_back(0); psi_off.push_back(0); }
		std::vector<double> pre_shift;
		std::vector<RFLOAT> ta, tb, tc;
		for (int p = 0; p < P; p++)
		{
			const long int part_id = first + p;                       // exp_metadata row p: one image per particle
			if (o.mydata.numberOfImagesInParticle(part_id) != 1) RB_REPORT_ERROR("relion_b200: particles with several images (tomo) are not covered");
			for (size_t i = 0; i < (size_t) n * n; i++) images[(size_t) p * n * n + i] = (float) DIRECT_MULTIDIM_ELEM(o.exp_imagedata, (size_t) (row0 + p) * n * n + i);
			group_id[p] = (int) o.mydata.getGroupId(part_id); optics_group[p] = o.mydata.getOpticsGroup(part_id);
			const RFLOAT normcorr = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_NORM);
			norm_factor[p] = o.do_norm_correction ? m.avg_norm_correction / normcorr : 1.;                                    // :421
			old_offset[2 * p] = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_XOFF); old_offset[2 * p + 1] = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_YOFF);
			// op.prior (:130-160): the offset prior, 999 = none -> zero
			RFLOAT px = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_XOFF_PRIOR), py = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_YOFF_PRIOR);
			if (px > 998.99 && px < 999.01) px = 0.;
			if (py > 998.99 && py < 999.01) py = 0.;
			prior_offset[2 * p] = px; prior_offset[2 * p + 1] = py;
			defU[p] = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_CTF_DEFOCUS_U); defV[p] = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_CTF_DEFOCUS_V);
			defA[p] = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_CTF_DEFOCUS_ANGLE); bfac[p] = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_CTF_BFACTOR);
			kfac[p] = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_CTF_KFACTOR); phs[p] = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_CTF_PHASE_SHIFT);
			if (o.do_ctf_correction)
			{
				CTF ctf;                                                                                                         // :822-840
				ctf.setValuesByGroup(&o.mydata.obsModel, optics_group[p], defU[p], defV[p], defA[p], bfac[p], kfac[p], phs[p], -1.);
				og_kV[optics_group[p]] = ctf.kV; og_Cs[optics_group[p]] = ctf.Cs; og_Q0[optics_group[p]] = ctf.Q0;
			}
			if (skip)
			{
				// :3752-3766: idir = ipsi (= itrans with --skip_align) = the particle's row in the pool; prior = pdf_class (:1966)
				dir_idx.push_back(row0 + p); dir_prior.push_back(1.); psi_idx.push_back(row0 + p); psi_prior.push_back(1.);
				dir_off.push_back((int) dir_idx.size()); psi_off.push_back((int) psi_idx.size());
				if (o.do_skip_align)
				{
					o.sampling.getTranslationsInPixel(row0 + p, 0, my_pixel_size, ta, tb, tc, false);
					pre_shift.push_back(ta[0]); pre_shift.push_back(tb[0]);
				}
			}
			if (use_priors)
			{
				// :135-172: prior angles, 999 = none -> the current angles; local searches always centre on the current angles
				RFLOAT prior_rot = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_ROT_PRIOR), prior_tilt = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_TILT_PRIOR),
				       prior_psi = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_PSI_PRIOR);
				const bool local_auto = o.do_auto_refine && o.sampling.healpix_order >= o.autosampling_hporder_local_searches;
				const bool local_class = !o.do_auto_refine && m.orientational_prior_mode == PRIOR_ROTTILT_PSI && m.sigma2_rot > 0. && m.sigma2_tilt > 0. && m.sigma2_psi > 0.;
				const bool local = local_auto || local_class;
				if ((prior_rot > 998.99 && prior_rot < 999.01) || local) prior_rot = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_ROT);
				if ((prior_tilt > 998.99 && prior_tilt < 999.01) || local) prior_tilt = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_TILT);
				if ((prior_psi > 998.99 && prior_psi < 999.01) || local) prior_psi = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_PSI);
				o.sampling.selectOrientationsWithNonZeroPriorProbability(prior_rot, prior_tilt, prior_psi, sqrt(m.sigma2_rot), sqrt(m.sigma2_tilt), sqrt(m.sigma2_psi),
				                                                         ptr_dir[p], pri_dir[p], ptr_psi[p], pri_psi[p]);
				if (ptr_dir[p].empty() || ptr_psi[p].empty()) RB_REPORT_ERROR("relion_b200: zero orientations fall within the local angular search");   // :176-183
				dir_idx.insert(dir_idx.end(), ptr_dir[p].begin(), ptr_dir[p].end()); dir_prior.insert(dir_prior.end(), pri_dir[p].begin(), pri_dir[p].end());
				psi_idx.insert(psi_idx.end(), ptr_psi[p].begin(), ptr_psi[p].end()); psi_prior.insert(psi_prior.end(), pri_psi[p].begin(), pri_psi[p].end());
				dir_off.push_back((int) dir_idx.size()); psi_off.push_back((int) psi_idx.size());
			}
		}
		rb_raw_particles raw;
		memset(&raw, 0, sizeof(raw));
		raw.n_particles = P; raw.image_size = n; raw.images = images.data(); raw.norm_factor = norm_factor.data();
		raw.old_offset = old_offset.data(); raw.prior_offset = prior_offset.data(); raw.group_id = group_id.data(); raw.optics_group = optics_group.data();
		std::vector<int> bp_offset;
		if (o.grad_pseudo_halfsets)
		{
			bp_offset.resize(P);
			for (int p = 0; p < P; p++) bp_offset[p] = (int) ((first + p) % 2) * m.nr_classes;       // acc_ml_optimiser_impl.h:3397-3399
			raw.bp_offset = bp_offset.data();
		}
		// beam-tilt demodulation and MTF division of both transforms (acc_ml_optimiser_impl.h:535-536 ->
		// ObservationModel::demodulatePhase / divideByMtf(do_multiply_instead = false, do_correct_average_mtf = true),
		// src/jaz/single_particle/obs_model.cpp:528-626): one factor image per optics group, applied on the device after windowing
		std::vector<float> og_factor;
		{
			// the reference's own functions applied to an image of ones give the factor exactly as RELION would apply it
			ObservationModel &obs = o.mydata.obsModel;
			if (obs.hasOddZernike || obs.hasMultipleMtfs)
			{
				const int nog = obs.numberOfOpticsGroups(), xs = cur / 2 + 1;
				og_factor.assign((size_t) nog * cur * xs * 2, 0.f);
				for (int g = 0; g < nog; g++)
				{
					MultidimArray<Complex> one;
					one.resize(cur, xs);
					FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(one) DIRECT_MULTIDIM_ELEM(one, n) = Complex(1., 0.);
					obs.demodulatePhase(g, one);
					obs.divideByMtf(g, one);
					FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(one)
					{
						og_factor[((size_t) g * cur * xs + n) * 2] = (float) DIRECT_MULTIDIM_ELEM(one, n).real;
						og_factor[((size_t) g * cur * xs + n) * 2 + 1] = (float) DIRECT_MULTIDIM_ELEM(one, n).imag;
					}
				}
				raw.og_fourier_factor = og_factor.data();
			}
		}
		std::vector<int64_t> noise_seed;
		if (!o.do_zero_mask)                                                                          // noise-filled soft mask, seed as acc_ml_optimiser_impl.h:371
		{
			noise_seed.resize(P);
			for (int p = 0; p < P; p++) noise_seed[p] = (int64_t) o.random_seed + (int64_t) (first + p);
			raw.noise_seed = noise_seed.data();
		}
		raw.ctf_defU = defU.data(); raw.ctf_defV = defV.data(); raw.ctf_defAngle = defA.data(); raw.ctf_Bfac = bfac.data(); raw.ctf_scale = kfac.data();
		raw.ctf_phase_shift = phs.data(); raw.og_kV = og_kV.data(); raw.og_Cs = og_Cs.data(); raw.og_Q0 = og_Q0.data();
		raw.mask_radius = o.particle_diameter / (2. * my_pixel_size);                                                          // :550-552
		// MBL: anisotropic magnification and the scale difference between this optics group and the model (:1098-1103)
		double mat_left[9];
		{
			Matrix2D<RFLOAT> mag;
			mag.initIdentity(3);
			mag = o.mydata.obsModel.applyAnisoMag(mag, og0);
			mag = o.mydata.obsModel.applyScaleDifference(mag, og0, m.ori_size, m.pixel_size);
			if (!mag.isIdentity())
			{
				for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) mat_left[3 * i + j] = mag(i, j);
				raw.mat_left = mat_left;
			}
		}
		// the spectrum of the noise-filled mask: mymodel.sigma2_noise scattered onto this box (:374-384)
		std::vector<double> noise_sigma2;
		if (!o.do_zero_mask)
		{
			noise_sigma2.assign((size_t) nog * nshell, 0.);
			for (int g = 0; g < nog; g++)
			{
				const RFLOAT remap = bundle->remapSizes(g);
				for (int i = 0; i < (int) XSIZE(m.sigma2_noise[g]); i++)
				{
					const int ir = ROUND(remap * i);
					if (ir < nshell) noise_sigma2[(size_t) g * nshell + ir] = DIRECT_A1D_ELEM(m.sigma2_noise[g], i);
				}
			}
			raw.noise_sigma2 = noise_sigma2.data();
		}
		raw.width_mask_edge = (double) o.width_mask_edge;
		if (!pre_shift.empty()) raw.pre_shift = pre_shift.data();
		if (use_priors || skip)
		{
			raw.dir_off = dir_off.data(); raw.dir_idx = dir_idx.data(); raw.dir_prior = dir_prior.data();
			raw.psi_off = psi_off.data(); raw.psi_idx = psi_idx.data(); raw.psi_prior = psi_prior.data();
		}

		// ---- the E-step on the device ----
		std::vector<float> power_img((size_t) P * nshell);
		std::vector<rb_particle_out> parts(P);
		std::vector<float> wsum_sigma2((size_t) P * nshell);
		const int n_dir = (int) o.sampling.NrDirections();
		std::vector<double> wsum_pdf_dir((size_t) K * n_dir, 0.), wsum_pdf_class(K, 0.), wsum_prior_class(2 * (size_t) K, 0.);
		rb_pool_out out;
		memset(&out, 0, sizeof(out));
		out.wsum_prior_offset_class = wsum_prior_class.data();
		out.particles = parts.data(); out.wsum_sigma2_noise = wsum_sigma2.data(); out.wsum_pdf_direction = wsum_pdf_dir.data(); out.wsum_pdf_class = wsum_pdf_class.data();
		RB_TRY(rb_pool_prepare(bundle->ctx, 0, &raw, power_img.data()));
		RB_TRY(rb_estep_slot(bundle->ctx, 0, &out, o.do_skip_maximization ? 1u : 0u));

		// ---- host bookkeeping of storeWeightedSums, in fp64 (acc_ml_optimiser_impl.h:2858-2931, 3466-3656) ----
		std::vector<double> logsigma2(nog, 0.);
		for (int g = 0; g < nog; g++)                                                                                          // :3556-3566
		{
			if (!bundle->sameGeometry(g, og0)) continue;
			const RFLOAT remap_g = 1. / bundle->remapSizes(g);                                                                  // :3559
			FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(o.Mresol_fine[g])
			{
				const int ires = DIRECT_MULTIDIM_ELEM(o.Mresol_fine[g], n);
				const int ires_remapped = ROUND(remap_g * ires);
				if (ires > 0 && ires_remapped < (int) XSIZE(m.sigma2_noise[g])) logsigma2[g] += log(2. * PI * DIRECT_A1D_ELEM(m.sigma2_noise[g], ires_remapped));
			}
		}
		std::vector<RFLOAT> rot, tilt, psi, tx, ty, tz;
		RFLOAT thr_avg_norm_correction = 0., thr_sum_dLL = 0., thr_sum_Pmax = 0., thr_wsum_sigma2_offset = 0.;
		MlWsumModel &w = o.wsum_model;
		for (int p = 0; p < P; p++)
		{
			const rb_particle_out &r = parts[p];
			const int og = optics_group[p], ig = group_id[p];
			// metadata row (:2858-2931)
			// with --skip_align / --skip_rotate the reference's indices are the particle's row (:2874-2876)
			o.sampling.getOrientations(skip ? row0 + p : r.best_idir, skip ? row0 + p : r.best_ipsi, o.adaptive_oversampling, rot, tilt, psi, ptr_dir[p], pri_dir[p], ptr_psi[p], pri_psi[p]);
			o.sampling.getTranslationsInPixel(o.do_skip_align ? row0 + p : r.best_itrans, o.adaptive_oversampling, my_pixel_size, tx, ty, tz, false);   // :2696
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_ROT) = rot[r.best_iover_rot];
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_TILT) = tilt[r.best_iover_rot];
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_PSI) = psi[r.best_iover_rot];
			const RFLOAT ox = ROUND(old_offset[2 * p]), oy = ROUND(old_offset[2 * p + 1]);                                      // op.old_offset is the ROUNDED one (:216)
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_XOFF) = ox + tx[r.best_iover_trans];
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_YOFF) = oy + ty[r.best_iover_trans];
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_CLASS) = (RFLOAT) r.best_class + 1;
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_PMAX) = (RFLOAT) r.pmax;
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_NR_SIGN) = (RFLOAT) r.nr_significant_coarse;                            // :2319-2320
			if (o.do_skip_maximization) continue;
			// sigma2_noise and the norm correction extended beyond the current size with the image's own power spectrum (:3505-3515)
			RFLOAT exp_wsum_norm_correction = r.wsum_norm_correction;
			std::vector<RFLOAT> thr_sigma2(nshell, 0.);
			for (int i = 0; i < nshell; i++) thr_sigma2[i] = (RFLOAT) wsum_sigma2[(size_t) p * nshell + i];
			for (int i = cur / 2 + 1; i < nshell; i++) { thr_sigma2[i] += (RFLOAT) power_img[(size_t) p * nshell + i]; exp_wsum_norm_correction += (RFLOAT) power_img[(size_t) p * nshell + i]; }
			if (o.do_norm_correction)                                                                                           // :3519-3538
			{
				RFLOAT old_norm_correction = DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_NORM) / m.avg_norm_correction;
				const RFLOAT normcorr = old_norm_correction * sqrt(exp_wsum_norm_correction * 2.);
				thr_avg_norm_correction += normcorr;
				DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_NORM) = normcorr;
			}
			const RFLOAT dLL = do_cc ? r.dLL_nolog : r.dLL_nolog - logsigma2[og];                                              // :3568-3574
			DIRECT_A2D_ELEM(o.exp_metadata, row0 + p, METADATA_DLL) = dLL;
			thr_sum_dLL += dLL; thr_sum_Pmax += (RFLOAT) r.pmax;
			// the critical section (:3585-3656); this thread is the only writer
			const RFLOAT remap = 1. / bundle->remapSizes(og);                                                                   // :3617-3618
			for (int i = 0; i < nshell; i++)
			{
				const int i_resam = ROUND(i * remap);
				if (i_resam < (int) XSIZE(w.sigma2_noise[og])) DIRECT_A1D_ELEM(w.sigma2_noise[og], i_resam) += thr_sigma2[i];
			}
			w.sumw_group[og] += r.sumw;
			if (o.do_scale_correction)                                                                                          // :3546-3554, :3624-3627
			{
				const RFLOAT sc = m.scale_correction[ig];
				w.wsum_signal_product[ig] += r.wsum_XA / sc;
				w.wsum_reference_power[ig] += r.wsum_AA / (sc * sc);
			}
			thr_wsum_sigma2_offset += r.wsum_sigma2_offset;
		}
		if (!o.do_skip_maximization)
		{
			for (int k = 0; k < K; k++)
			{
				w.pdf_class[k] += wsum_pdf_class[k];
				if (o.mymodel.ref_dim == 2)                                                                                     // :3639-3642
				{
					XX(w.prior_offset_class[k]) += wsum_prior_class[2 * k];
					YY(w.prior_offset_class[k]) += wsum_prior_class[2 * k + 1];
				}
				if (!(o.do_skip_align || o.do_skip_rotate))
					for (int d = 0; d < n_dir && d < (int) MULTIDIM_SIZE(w.pdf_direction[k]); d++) DIRECT_MULTIDIM_ELEM(w.pdf_direction[k], d) += wsum_pdf_dir[(size_t) k * n_dir + d];
			}
			w.sigma2_offset += thr_wsum_sigma2_offset;
			if (o.do_norm_correction) w.avg_norm_correction += thr_avg_norm_correction;
			w.LL += thr_sum_dLL;
			w.ave_Pmax += thr_sum_Pmax;
			if (o.mydata.obsModel.getCtfPremultiplied(0))                                                                       // :3590-3603: sumw_ctf2 from the CTF images
			{
				const int xs = cur / 2 + 1;
				std::vector<float> Fctf((size_t) P * cur * xs);
				RB_TRY(rb_pool_download(bundle->ctx, 0, NULL, NULL, Fctf.data(), NULL));
				for (int p = 0; p < P; p++)
				{
					const int og = optics_group[p];
					const RFLOAT myscale = XMIPP_MAX(0.001, m.scale_correction[group_id[p]]);
					FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(o.Mresol_fine[og])
					{
						const int ires = DIRECT_MULTIDIM_ELEM(o.Mresol_fine[og], n);
						if (ires > -1 && ires < (int) XSIZE(w.sumw_ctf2[og])) DIRECT_A1D_ELEM(w.sumw_ctf2[og], ires) += myscale * (RFLOAT) Fctf[(size_t) p * cur * xs + n];
					}
				}
			}
		}
	}
};

// ---------------------------------------------------------------------------------------------------------------------
/* MlOptimiserMpi::combineAllWeightedSums (src/ml_optimiser_mpi.cpp:2028-2185) over NCCL: every rank's device accumulators are
 * summed in place (rb_bp_allreduce, on the bundle's stream), every other weighted sum travels as one fp64 vector in
 * MlWsumModel::pack order.  Afterwards every rank holds the totals; pullBackprojectors() then adds the accumulators to
 * wsum_model.BPref on ONE rank (or on every rank of a half-set that reconstructs). */
inline void combineAllWeightedSums(MlOptimiser *baseMLO, MlDeviceBundle *bundle, rb_comm *comm)
{
	RB_TRY(rb_bp_allreduce(bundle->ctx, comm));
	std::vector<double> packed;
	WsumPack::pack(baseMLO->wsum_model, packed);
	RB_TRY(rb_wsum_allreduce(bundle->ctx, comm, packed.data(), packed.size()));
	WsumPack::unpack(baseMLO->wsum_model, packed);
}

} // namespace relion_b200

#endif
