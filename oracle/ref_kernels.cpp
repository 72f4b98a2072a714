/*
 * TEST INFRASTRUCTURE ONLY — builds oracle/_ref/librefkernels.so (git-ignored).
 *
 * This translation unit contains NO algorithm of its own: it #includes the reference's own ALTCPU
 * kernel headers from where they lie under /root/reference (never copied into this repo) and wraps
 * them behind the C table declared in oracle/oracle_kernels.h, so that
 *   (a) oracle/port_kernels.cpp (our restatement) can be pinned against the real reference code, and
 *   (b) bench.py --impl reference can time the reference's own CPU implementation of the path.
 *
 * Build recipe: oracle/Makefile (target _ref), flags as verified in SURVEY.md §8(c)/Appendix D:
 *   g++ -O3 -march=native -std=c++17 -fopenmp -DALTCPU=1 -DACC_CUDA=2 -DACC_CPU=1 -DPROJECTOR_NO_TEXTURES
 *       -Ioracle/shim -I/root/reference
 * Header closure: acc_ptr.h, complex.h, error.h, macros.h, parallel.h, pipeline_control.h,
 * jaz/single_particle/t_complex.h, cpu_settings.h, settings.h — no FFTW/MPI/TIFF/TBB needed
 * (tbb::spin_mutex comes from oracle/shim/tbb/spin_mutex.h).
 *
 * The two helpers that live in src/acc/cpu/cpu_kernels/helper.cpp (exponentiate_weights_fine,
 * cpu_kernel_make_eulers_3D) come from that file itself, compiled as a second translation unit of this
 * library (oracle/Makefile): it includes src/acc/utilities.h and acc_helper_functions.h, which pull in
 * MlOptimiser's header, so it is built against declaration-level stand-ins for fftw3.h / tiffio.h / png.h
 * (tests/cpp/relion_stubs) and mpi.h / tbb (oracle/shim); the only symbol it needs from elsewhere is
 * rnd_gaus (noise fill), defined below as a call that aborts.  The same file also gives the image-preparation
 * helpers (cpu_translate2D, softMaskBackgroundValue, cosineFilter; powerClass is a template of helper.h)
 * that pin oracle/prepare.py (refk_prep_* below).
 */
#include "src/acc/cpu/device_stubs.h"
#include "src/acc/acc_ptr.h"
#include "src/acc/acc_projector.h"
#include "src/acc/acc_backprojector.h"
#include "src/acc/cpu/cpu_kernels/cpu_utils.h"
#include "src/acc/acc_projectorkernel_impl.h"
#include "src/acc/cpu/cpu_kernels/helper.h"
#include "src/acc/cpu/cpu_kernels/diff2.h"
#include "src/acc/cpu/cpu_kernels/wavg.h"
#include "src/acc/cpu/cpu_kernels/BP.h"

#include <vector>
#include <complex>
#include <cstring>

#include "oracle_kernels.h"

// needed by helper.cpp's noise-fill helpers only (never reached from here)
float rnd_gaus(float, float) { abort(); }

namespace {

AccProjectorKernel make_kernel(const ok_projector *p, int imgX, int imgY)
{
	// AccProjectorKernel::makeKernel (acc_projectorkernel_impl.h:301-319) with imgMaxR = imgX-1
	// as every caller passes (acc_ml_optimiser_impl.h:1322-1327)
	int imgMaxR = imgX - 1;
	int maxR = p->mdlMaxR >= imgMaxR ? imgMaxR : p->mdlMaxR;
	return AccProjectorKernel(p->mdlX, p->mdlY, p->mdlZ, imgX, imgY, 1,
	                          p->mdlInitY, p->mdlInitZ, p->padding_factor, maxR,
	                          (std::complex<XFLOAT> *) p->mdl);
}

void refk_project(const ok_projector *p, int imgX, int imgY, const float *e, float *out_re, float *out_im)
{
	AccProjectorKernel k = make_kernel(p, imgX, imgY);
	for (int iy = 0; iy < imgY; iy++)
	{
		// fine-pass row handling of diff2_fine_2D / wavg_ref3D (diff2.h:344-355, wavg.h:71-82)
		int xstart = 0, xend = imgX, y = iy;
		if (iy > k.maxR)
		{
			if (iy >= imgY - k.maxR) y = iy - imgY;
			else { xstart = k.maxR; xend = xstart + 1; }
		}
		for (int x = 0; x < imgX; x++) { out_re[iy * imgX + x] = 0.f; out_im[iy * imgX + x] = 0.f; }
		for (int x = xstart; x < xend; x++)
			k.project3Dmodel(x, y, e[0], e[1], e[3], e[4], e[6], e[7], out_re[iy * imgX + x], out_im[iy * imgX + x]);
	}
}

void refk_diff2_coarse(const ok_projector *p, int imgX, int imgY,
		const float *eulers, unsigned long O,
		const float *trans_x, const float *trans_y, unsigned long T,
		const float *img_re, const float *img_im, const float *corr, float *diff2s)
{
	AccProjectorKernel k = make_kernel(p, imgX, imgY);
	unsigned long image_size = (unsigned long) imgX * imgY;
	std::vector<float> tz(T, 0.f);
	// dispatch as runDiff2KernelCoarse does (acc_helper_functions_impl.h:1247-1300):
	unsigned long rest = O % D2C_BLOCK_SIZE_REF3D;
	unsigned long even = O - rest;
	if (even)
		CpuKernels::diff2_coarse<true, false, D2C_BLOCK_SIZE_REF3D, D2C_EULERS_PER_BLOCK_REF3D, PREFETCH_FRACTION_3D>(
			even / D2C_EULERS_PER_BLOCK_REF3D, (XFLOAT *) eulers, (XFLOAT *) trans_x, (XFLOAT *) trans_y, tz.data(),
			(XFLOAT *) img_re, (XFLOAT *) img_im, k, (XFLOAT *) corr, diff2s, T, image_size);
	if (rest)
		CpuKernels::diff2_coarse<true, false, D2C_BLOCK_SIZE_REF3D, 1, PREFETCH_FRACTION_3D>(
			rest, (XFLOAT *) &eulers[9 * even], (XFLOAT *) trans_x, (XFLOAT *) trans_y, tz.data(),
			(XFLOAT *) img_re, (XFLOAT *) img_im, k, (XFLOAT *) corr, &diff2s[T * even], T, image_size);
}

void refk_diff2_fine(const ok_projector *p, int imgX, int imgY, const float *eulers,
		const float *trans_x, const float *trans_y,
		const float *img_re, const float *img_im, const float *corr, float sum_init,
		unsigned long orientation_num, unsigned long translation_num, unsigned long num_jobs,
		const unsigned long *rot_idx, const unsigned long *trans_idx,
		const unsigned long *job_idx, const unsigned long *job_num, float *diff2s)
{
	AccProjectorKernel k = make_kernel(p, imgX, imgY);
	CpuKernels::diff2_fine_2D<true>(num_jobs, (XFLOAT *) eulers, (XFLOAT *) img_re, (XFLOAT *) img_im,
		(XFLOAT *) trans_x, (XFLOAT *) trans_y, (XFLOAT *) trans_x /*unused z*/, k, (XFLOAT *) corr, diff2s,
		(unsigned long) imgX * imgY, sum_init, orientation_num, translation_num, num_jobs,
		(unsigned long *) rot_idx, (unsigned long *) trans_idx, (unsigned long *) job_idx, (unsigned long *) job_num);
}

void refk_weights_exponent_coarse(const float *pdf_o, const unsigned char *pdf_oz,
		const float *pdf_t, const unsigned char *pdf_tz, float *w, float min_diff2,
		unsigned long no, unsigned long nt, size_t max_idx)
{
	static_assert(sizeof(bool) == 1, "bool must be 1 byte");
	CpuKernels::weights_exponent_coarse<XFLOAT>((XFLOAT *) pdf_o, (bool *) pdf_oz, (XFLOAT *) pdf_t, (bool *) pdf_tz,
		w, min_diff2, no, nt, max_idx);
}

void refk_exponentiate(float *a, float add, size_t n) { CpuKernels::exponentiate<XFLOAT>(a, add, n); }

void refk_collect2jobs(int grid_size, const float *ox, const float *oy, const float *o2,
		const float *w, float sig, float sum, unsigned long ct, unsigned long ot, unsigned long oo,
		unsigned long ov, float *o_w, float *px, float *py, float *s2,
		const unsigned long *rot_idx, const unsigned long *trans_idx,
		const unsigned long *job_idx, const unsigned long *job_num)
{
	CpuKernels::collect2jobs<false>(grid_size, SUMW_BLOCK_SIZE, (XFLOAT *) ox, (XFLOAT *) oy, (XFLOAT *) ox,
		(XFLOAT *) o2, (XFLOAT *) w, sig, sum, ct, ot, oo, ov, false, o_w, px, py, px /*z unused*/, s2,
		(unsigned long *) rot_idx, (unsigned long *) trans_idx, (unsigned long *) job_idx, (unsigned long *) job_num);
}

void refk_wavg(const ok_projector *p, int imgX, int imgY, const float *eulers, unsigned long orientation_num,
		const float *img_re, const float *img_im, const float *trans_x, const float *trans_y,
		const float *weights, const float *ctfs, float *parts, float *AA, float *XA,
		unsigned long trans_num, float weight_norm, float significant_weight, float part_scale)
{
	AccProjectorKernel k = make_kernel(p, imgX, imgY);
	// refs_are_ctf_corrected branch of runWavgKernel (acc_helper_functions_impl.h:343-370)
	CpuKernels::wavg_ref3D<true, true>((XFLOAT *) eulers, k, (unsigned long) imgX * imgY, orientation_num,
		(XFLOAT *) img_re, (XFLOAT *) img_im, (XFLOAT *) trans_x, (XFLOAT *) trans_y, (XFLOAT *) trans_x,
		(XFLOAT *) weights, (XFLOAT *) ctfs, parts, AA, XA, trans_num, weight_norm, significant_weight, part_scale);
}

void refk_backproject(const ok_backprojector *bp, int imgX, int imgY,
		const float *img_re, const float *img_im, const float *trans_x, const float *trans_y,
		const float *weights, const float *Minvsigma2s, const float *ctfs,
		unsigned long trans_num, float significant_weight, float weight_norm,
		const float *eulers, unsigned long image_count)
{
	std::vector<tbb::spin_mutex> local;
	tbb::spin_mutex *mutexes = (tbb::spin_mutex *) bp->sync;
	if (!mutexes) { local = std::vector<tbb::spin_mutex>((size_t) bp->mdlZ * bp->mdlY); mutexes = local.data(); }
	CpuKernels::backprojectRef3D<false>(image_count, (XFLOAT *) img_re, (XFLOAT *) img_im,
		(XFLOAT *) trans_x, (XFLOAT *) trans_y, (XFLOAT *) weights, (XFLOAT *) Minvsigma2s, (XFLOAT *) ctfs,
		trans_num, significant_weight, weight_norm, (XFLOAT *) eulers,
		bp->real, bp->imag, bp->weight, bp->maxR, bp->maxR * bp->maxR, bp->padding_factor,
		(unsigned) imgX, (unsigned) imgY, 1u, (size_t) imgX * imgY,
		(unsigned) bp->mdlX, (unsigned) bp->mdlY, bp->mdlInitY, bp->mdlInitZ, mutexes);
}

void refk_backproject2d(const ok_backprojector *bp, int imgX, int imgY,
		const float *img_re, const float *img_im, const float *trans_x, const float *trans_y,
		const float *weights, const float *Minvsigma2s, const float *ctfs,
		unsigned long trans_num, float significant_weight, float weight_norm,
		const float *eulers, unsigned long image_count)
{
	std::vector<tbb::spin_mutex> local;
	tbb::spin_mutex *mutexes = (tbb::spin_mutex *) bp->sync;
	if (!mutexes) { local = std::vector<tbb::spin_mutex>((size_t) bp->mdlY); mutexes = local.data(); }
	CpuKernels::backproject2D<false>(image_count, 128, (XFLOAT *) img_re, (XFLOAT *) img_im, (XFLOAT *) trans_x, (XFLOAT *) trans_y,
		(XFLOAT *) weights, (XFLOAT *) Minvsigma2s, (XFLOAT *) ctfs, trans_num, significant_weight, weight_norm, (XFLOAT *) eulers,
		bp->real, bp->imag, bp->weight, bp->maxR, bp->maxR * bp->maxR, bp->padding_factor,
		(unsigned) imgX, (unsigned) imgY, (unsigned) (imgX * imgY), (unsigned) bp->mdlX, bp->mdlInitY, mutexes);
}

// gradient refinement: runBackProjectKernel with do_grad (acc_helper_functions_impl.h:870-900, 3D reference, 2D data)
void refk_backproject_sgd(const ok_backprojector *bp, const ok_projector *p, int imgX, int imgY,
		const float *img_re, const float *img_im, const float *trans_x, const float *trans_y,
		const float *weights, const float *Minvsigma2s, const float *ctfs,
		unsigned long trans_num, float significant_weight, float weight_norm,
		const float *eulers, unsigned long image_count)
{
	std::vector<tbb::spin_mutex> local;
	tbb::spin_mutex *mutexes = (tbb::spin_mutex *) bp->sync;
	if (!mutexes) { local = std::vector<tbb::spin_mutex>((size_t) bp->mdlZ * bp->mdlY); mutexes = local.data(); }
	AccProjectorKernel k = make_kernel(p, imgX, imgY);
	CpuKernels::backproject3D_SGD<false, false>(image_count, BP_REF3D_BLOCK_SIZE, k, (XFLOAT *) img_re, (XFLOAT *) img_im,
		(XFLOAT *) trans_x, (XFLOAT *) trans_y, (XFLOAT *) trans_x, (XFLOAT *) weights, (XFLOAT *) Minvsigma2s, (XFLOAT *) ctfs,
		trans_num, significant_weight, weight_norm, (XFLOAT *) eulers,
		bp->real, bp->imag, bp->weight, bp->maxR, bp->maxR * bp->maxR, bp->padding_factor,
		(unsigned) imgX, (unsigned) imgY, 1u, (size_t) imgX * imgY,
		(unsigned) bp->mdlX, (unsigned) bp->mdlY, bp->mdlInitY, bp->mdlInitZ, mutexes);
}

void refk_backproject2d_sgd(const ok_backprojector *bp, const ok_projector *p, int imgX, int imgY,
		const float *img_re, const float *img_im, const float *trans_x, const float *trans_y,
		const float *weights, const float *Minvsigma2s, const float *ctfs,
		unsigned long trans_num, float significant_weight, float weight_norm,
		const float *eulers, unsigned long image_count)
{
	std::vector<tbb::spin_mutex> local;
	tbb::spin_mutex *mutexes = (tbb::spin_mutex *) bp->sync;
	if (!mutexes) { local = std::vector<tbb::spin_mutex>((size_t) bp->mdlY); mutexes = local.data(); }
	AccProjectorKernel k = make_kernel(p, imgX, imgY);
	CpuKernels::backproject2D_SGD<false>(image_count, 128, k, (XFLOAT *) img_re, (XFLOAT *) img_im, (XFLOAT *) trans_x, (XFLOAT *) trans_y,
		(XFLOAT *) weights, (XFLOAT *) Minvsigma2s, (XFLOAT *) ctfs, trans_num, significant_weight, weight_norm, (XFLOAT *) eulers,
		bp->real, bp->imag, bp->weight, bp->maxR, bp->maxR * bp->maxR, bp->padding_factor,
		(unsigned) imgX, (unsigned) imgY, (unsigned) (imgX * imgY), (unsigned) bp->mdlX, bp->mdlInitY, mutexes);
}

// first-iteration cross-correlation kernels, dispatched as runDiff2KernelCoarse / runDiff2KernelFine do for a 3D reference
// and 2D data (acc_helper_functions_impl.h:1761-1777, :1950-1973 -> AccUtilities::diff2_CC_coarse / diff2_CC_fine)
void refk_diff2_cc_coarse(const ok_projector *p, int imgX, int imgY,
		const float *eulers, unsigned long O,
		const float *trans_x, const float *trans_y, unsigned long T,
		const float *img_re, const float *img_im, const float *corr, float *diff2s)
{
	AccProjectorKernel k = make_kernel(p, imgX, imgY);
	CpuKernels::diff2_CC_coarse_2D<true>(O, (XFLOAT *) eulers, (XFLOAT *) img_re, (XFLOAT *) img_im,
		(XFLOAT *) trans_x, (XFLOAT *) trans_y, k, (XFLOAT *) corr, diff2s, T, (unsigned long) imgX * imgY, (XFLOAT) 0);
}

void refk_diff2_cc_fine(const ok_projector *p, int imgX, int imgY, const float *eulers,
		const float *trans_x, const float *trans_y,
		const float *img_re, const float *img_im, const float *corr,
		unsigned long orientation_num, unsigned long translation_num, unsigned long num_jobs,
		const unsigned long *rot_idx, const unsigned long *trans_idx,
		const unsigned long *job_idx, const unsigned long *job_num, float *diff2s)
{
	AccProjectorKernel k = make_kernel(p, imgX, imgY);
	CpuKernels::diff2_CC_fine_2D<true>(num_jobs, (XFLOAT *) eulers, (XFLOAT *) img_re, (XFLOAT *) img_im,
		(XFLOAT *) trans_x, (XFLOAT *) trans_y, k, (XFLOAT *) corr, diff2s, (unsigned long) imgX * imgY,
		(XFLOAT) 0 /*sum_init: unused*/, (XFLOAT) 0 /*sqrtXi2: unused*/, orientation_num, translation_num, num_jobs,
		(unsigned long *) rot_idx, (unsigned long *) trans_idx, (unsigned long *) job_idx, (unsigned long *) job_num);
}

// cpu_kernel_make_eulers_3D<invert = true, doL, doR> (helper.cpp:744-875), the instantiations the E-step uses
// (acc_projector_plan_impl.h:246-262 picks them by the shapes of MBL / MBR)
void refk_make_eulers_3d(const float *alphas, const float *betas, const float *gammas, float *eulers, unsigned long n, const float *L,
                         const float *R)
{
	const int bs = 128;                                                           // BLOCK_SIZE of the call site (acc_ml_optimiser_impl.h generateEulerMatrices)
	const int grid = (int) ((n + bs - 1) / bs);
	XFLOAT *a = (XFLOAT *) alphas, *b = (XFLOAT *) betas, *g = (XFLOAT *) gammas, *e = (XFLOAT *) eulers, *l = (XFLOAT *) L, *r = (XFLOAT *) R;
	if (L && R) CpuKernels::cpu_kernel_make_eulers_3D<true, true, true>(grid, bs, a, b, g, e, n, l, r);
	else if (L) CpuKernels::cpu_kernel_make_eulers_3D<true, true, false>(grid, bs, a, b, g, e, n, l, NULL);
	else if (R) CpuKernels::cpu_kernel_make_eulers_3D<true, false, true>(grid, bs, a, b, g, e, n, NULL, r);
	else CpuKernels::cpu_kernel_make_eulers_3D<true, false, false>(grid, bs, a, b, g, e, n, NULL, NULL);
}

// CpuKernels::exponentiate_weights_fine (helper.cpp:27-61)
void refk_exponentiate_weights_fine(const float *pdf_orientation, const unsigned char *pdf_orientation_zeros, const float *pdf_offset,
		const unsigned char *pdf_offset_zeros, float *weights, float min_diff2, unsigned long oversamples_orient,
		unsigned long oversamples_trans, const unsigned long *rot_id, const unsigned long *trans_idx, const unsigned long *job_idx,
		const unsigned long *job_num, long n_jobs)
{
	static_assert(sizeof(bool) == 1, "bool flags are passed as bytes");
	CpuKernels::exponentiate_weights_fine((XFLOAT *) pdf_orientation, (bool *) pdf_orientation_zeros, (XFLOAT *) pdf_offset,
		(bool *) pdf_offset_zeros, weights, min_diff2, oversamples_orient, oversamples_trans, (unsigned long *) rot_id,
		(unsigned long *) trans_idx, (unsigned long *) job_idx, (unsigned long *) job_num, n_jobs);
}

void *refk_bp_sync_alloc(int mdlY, int mdlZ) { return new tbb::spin_mutex[(size_t) mdlY * mdlZ]; }
void refk_bp_sync_free(void *s) { delete[] (tbb::spin_mutex *) s; }

const ok_kernel_table table = {
	"reference",
	refk_make_eulers_3d,
	refk_project,
	refk_diff2_coarse,
	refk_diff2_fine,
	refk_weights_exponent_coarse,
	refk_exponentiate,
	refk_exponentiate_weights_fine,
	refk_collect2jobs,
	refk_wavg,
	refk_backproject,
	refk_bp_sync_alloc,
	refk_bp_sync_free,
	refk_backproject2d,
	refk_diff2_cc_coarse,
	refk_diff2_cc_fine,
	refk_backproject_sgd,
	refk_backproject2d_sgd,
};

} // namespace

extern "C" const ok_kernel_table *refk_kernel_table(void) { return &table; }

// ---- the reference's own 2D-reference kernels (2D classification), used only to PIN the two-plane embedding that the
// oracle driver and the CUDA library use for 2D references (tests/test_oracle.py::test_2d_embedding_matches_reference_2d_kernels)
extern "C" {

// AccProjectorKernel::project2Dmodel (acc_projectorkernel_impl.h:233-300) over a half image, coarse-pass y wrap
void refk2d_project(const float *mdl_complex, int mdlX, int mdlY, int mdlInitY, int mdlMaxR, float padding_factor,
                    int imgX, int imgY, const float *e, float *out_re, float *out_im)
{
	int imgMaxR = imgX - 1;
	int maxR = mdlMaxR >= imgMaxR ? imgMaxR : mdlMaxR;
	AccProjectorKernel k(mdlX, mdlY, 0, imgX, imgY, 1, mdlInitY, 0, padding_factor, maxR, (std::complex<XFLOAT> *) mdl_complex);
	for (int iy = 0; iy < imgY; iy++)
	{
		int y = iy > k.maxR ? iy - imgY : iy;
		for (int x = 0; x < imgX; x++)
			k.project2Dmodel(x, y, e[0], e[1], e[3], e[4], out_re[iy * imgX + x], out_im[iy * imgX + x]);
	}
}

// CpuKernels::diff2_coarse<REF3D = false> as runDiff2KernelCoarse dispatches it for 2D references
void refk2d_diff2_coarse(const float *mdl_complex, int mdlX, int mdlY, int mdlInitY, int mdlMaxR, float padding_factor,
                         int imgX, int imgY, const float *eulers, unsigned long O,
                         const float *trans_x, const float *trans_y, unsigned long T,
                         const float *img_re, const float *img_im, const float *corr, float *diff2s)
{
	int imgMaxR = imgX - 1;
	int maxR = mdlMaxR >= imgMaxR ? imgMaxR : mdlMaxR;
	AccProjectorKernel k(mdlX, mdlY, 0, imgX, imgY, 1, mdlInitY, 0, padding_factor, maxR, (std::complex<XFLOAT> *) mdl_complex);
	unsigned long image_size = (unsigned long) imgX * imgY;
	std::vector<float> tz(T, 0.f);
	unsigned long rest = O % D2C_BLOCK_SIZE_2D;
	unsigned long even = O - rest;
	if (even)
		CpuKernels::diff2_coarse<false, false, D2C_BLOCK_SIZE_2D, D2C_EULERS_PER_BLOCK_2D, PREFETCH_FRACTION_2D>(
			even / D2C_EULERS_PER_BLOCK_2D, (XFLOAT *) eulers, (XFLOAT *) trans_x, (XFLOAT *) trans_y, tz.data(),
			(XFLOAT *) img_re, (XFLOAT *) img_im, k, (XFLOAT *) corr, diff2s, T, image_size);
	if (rest)
		CpuKernels::diff2_coarse<false, false, D2C_BLOCK_SIZE_2D, 1, PREFETCH_FRACTION_2D>(
			rest, (XFLOAT *) &eulers[9 * even], (XFLOAT *) trans_x, (XFLOAT *) trans_y, tz.data(),
			(XFLOAT *) img_re, (XFLOAT *) img_im, k, (XFLOAT *) corr, &diff2s[T * even], T, image_size);
}

} // extern "C"

// ---- the reference's own image-preparation helpers (getFourierTransformsAndCtfs, acc_ml_optimiser_impl.h:216-772), used only
// to PIN oracle/prepare.py (tests/test_reference_host.py) ----------------------------------------------------------------
extern "C" {

// cpu_translate2D (helper.cpp:255-282): out must be zero-filled by the caller, as the call site does (utilities_impl.h:374-436)
void refk_prep_translate2d(const float *in, float *out, int n, int dx, int dy)
{
	CpuKernels::cpu_translate2D<XFLOAT>((XFLOAT *) in, out, (size_t) n * n, n, n, dx, dy);
}

// softMaskBackgroundValue + cosineFilter with the zero mask (acc_ml_optimiser_impl.h:610-668; launch shape utilities_impl.h:494, 571)
void refk_prep_soft_mask(float *img, int n, float radius, float cosine_width, float *bg_out)
{
	if (radius < 0) radius = (float) n / 2.f;
	const float radius_p = radius + cosine_width;
	std::vector<XFLOAT> sum(SOFTMASK_BLOCK_SIZE, 0), sum_bg(SOFTMASK_BLOCK_SIZE, 0);
	CpuKernels::softMaskBackgroundValue(128, SOFTMASK_BLOCK_SIZE, img, (long) n * n, n, n, 1, n / 2, n / 2, 0, radius, radius_p, cosine_width,
	                                    sum.data(), sum_bg.data());
	double s = 0., sb = 0.;
	for (int i = 0; i < SOFTMASK_BLOCK_SIZE; i++) { s += sum[i]; sb += sum_bg[i]; }
	const XFLOAT bg = (XFLOAT) (sb / s);
	CpuKernels::cosineFilter(128, SOFTMASK_BLOCK_SIZE, img, (long) n * n, n, n, 1, n / 2, n / 2, 0, false, img, radius, radius_p, cosine_width, bg);
	if (bg_out) *bg_out = bg;
}

// powerClass<false> (helper.h:467-540) on a full-size transform [n][n/2+1] (interleaved complex)
void refk_prep_power_class(const float *F, int n, int current_size, float *spectrum, float *highres_Xi2)
{
	const int xdim = n / 2 + 1;
	const size_t sz = (size_t) n * xdim;
	for (int i = 0; i < xdim; i++) spectrum[i] = 0.f;
	*highres_Xi2 = 0.f;
	CpuKernels::powerClass<false>((int) ((sz + POWERCLASS_BLOCK_SIZE - 1) / POWERCLASS_BLOCK_SIZE), (ACCCOMPLEX *) F, spectrum, sz, (size_t) xdim,
	                              xdim, n, 1, current_size / 2 + 1, highres_Xi2);
}

}  // extern "C"
