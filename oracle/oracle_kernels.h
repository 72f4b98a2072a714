/*
 * TEST INFRASTRUCTURE ONLY — nothing under relion_b200/ may include, link or load this.
 *
 * Kernel-level C interface shared by the two CPU "kernel providers" of the oracle:
 *
 *   oracle/_ref/librefkernels.so   the reference's own ALTCPU kernels, compiled from the sources
 *                                  where they lie under /root/reference (oracle/ref_kernels.cpp,
 *                                  recipe in oracle/Makefile).  kind = "reference".
 *   oracle/liboracle.so            a plain scalar restatement of the same algorithms
 *                                  (oracle/port_kernels.cpp), no reference code.  kind = "port".
 *
 * Both export `const ok_kernel_table *<prefix>_kernel_table(void)`; the E-step driver
 * (oracle/estep_driver.cpp) runs on whichever table it is handed, so the restatement is validated
 * against the compiled reference through identical orchestration.
 *
 * All arrays are fp32 (XFLOAT=float, src/acc/settings.h:6-18), 3D reference / 2D data.
 */
#ifndef ORACLE_KERNELS_H_
#define ORACLE_KERNELS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Padded Fourier reference volume as AccProjector holds it (src/acc/acc_projector.h:17-60):
 * complex interleaved, [mdlZ][mdlY][mdlX], x>=0 half, yinit=zinit=-(mdl-1)/2. */
typedef struct {
	const float *mdl;       /* 2*mdlX*mdlY*mdlZ floats (re,im) */
	int mdlX, mdlY, mdlZ;
	int mdlInitY, mdlInitZ;
	int mdlMaxR;            /* r_max of the projector */
	float padding_factor;
} ok_projector;

/* Back-projection accumulators as AccBackprojector holds them (src/acc/acc_backprojector.h:24-60) */
typedef struct {
	float *real, *imag, *weight; /* mdlX*mdlY*mdlZ each */
	int mdlX, mdlY, mdlZ;
	int mdlInitY, mdlInitZ;
	int maxR;
	float padding_factor;
	void *sync;             /* provider-specific row locks from bp_sync_alloc (NULL: single-threaded use) */
} ok_backprojector;

typedef struct {
	const char *kind; /* "reference" | "port" */

	/* cpu_kernel_make_eulers_3D<invert=true,doL,doR>  (src/acc/cpu/cpu_kernels/helper.cpp, helper.h:725); L / R: [9] or NULL */
	void (*make_eulers_3d)(const float *alphas, const float *betas, const float *gammas,
	                       float *eulers, unsigned long n, const float *L, const float *R);

	/* AccProjectorKernel::project3Dmodel (2D-data overload) over a whole half-image with the fine-pass
	 * y-wrap (src/acc/acc_projectorkernel_impl.h:161-231, cpu_kernels/diff2.h:344-370) */
	void (*project)(const ok_projector *p, int imgX, int imgY, const float *euler9,
	                float *out_re, float *out_im);

	/* CpuKernels::diff2_coarse<true,false,256,16,4> (+<...,1,4> for the O%256 rest)
	 * (src/acc/cpu/cpu_kernels/diff2.h:32-282; dispatch acc_helper_functions_impl.h:1139-1400).
	 * diff2s[o*T+t] += sum_pix 0.5*corr*|ref_o - shift_t(img)|^2 */
	void (*diff2_coarse)(const ok_projector *p, int imgX, int imgY,
	                     const float *eulers, unsigned long O,
	                     const float *trans_x, const float *trans_y, unsigned long T,
	                     const float *img_re, const float *img_im, const float *corr,
	                     float *diff2s);

	/* CpuKernels::diff2_fine_2D<true> (src/acc/cpu/cpu_kernels/diff2.h:284-430) */
	void (*diff2_fine)(const ok_projector *p, int imgX, int imgY,
	                   const float *eulers,
	                   const float *trans_x, const float *trans_y,
	                   const float *img_re, const float *img_im, const float *corr,
	                   float sum_init,
	                   unsigned long orientation_num, unsigned long translation_num,
	                   unsigned long num_jobs,
	                   const unsigned long *rot_idx, const unsigned long *trans_idx,
	                   const unsigned long *job_idx, const unsigned long *job_num,
	                   float *diff2s);

	/* CpuKernels::weights_exponent_coarse<float> (src/acc/cpu/cpu_kernels/helper.h:16-39) */
	void (*weights_exponent_coarse)(const float *pdf_orientation, const unsigned char *pdf_orientation_zeros,
	                                const float *pdf_offset, const unsigned char *pdf_offset_zeros,
	                                float *weights, float min_diff2,
	                                unsigned long nr_coarse_orient, unsigned long nr_coarse_trans,
	                                size_t max_idx);

	/* CpuKernels::exponentiate<float> (helper.h:42-63) */
	void (*exponentiate)(float *array, float add, size_t size);

	/* CpuKernels::exponentiate_weights_fine (src/acc/cpu/cpu_kernels/helper.cpp:27-61) */
	void (*exponentiate_weights_fine)(const float *pdf_orientation, const unsigned char *pdf_orientation_zeros,
	                                  const float *pdf_offset, const unsigned char *pdf_offset_zeros,
	                                  float *weights, float min_diff2,
	                                  unsigned long oversamples_orient, unsigned long oversamples_trans,
	                                  const unsigned long *rot_id, const unsigned long *trans_idx,
	                                  const unsigned long *job_idx, const unsigned long *job_num,
	                                  long job_count);

	/* CpuKernels::collect2jobs<false> (helper.h:65-153) */
	void (*collect2jobs)(int grid_size,
	                     const float *oo_otrans_x, const float *oo_otrans_y,
	                     const float *myp_oo_otrans_x2y2z2,
	                     const float *i_weights, float significant_weight, float sum_weight,
	                     unsigned long coarse_trans, unsigned long oversamples_trans,
	                     unsigned long oversamples_orient, unsigned long oversamples,
	                     float *o_weights, float *wsum_prior_offsetx, float *wsum_prior_offsety,
	                     float *wsum_sigma2_offset,
	                     const unsigned long *rot_idx, const unsigned long *trans_idx,
	                     const unsigned long *job_idx, const unsigned long *job_num);

	/* CpuKernels::wavg_ref3D<REFCTF=true,REF3D=true> (src/acc/cpu/cpu_kernels/wavg.h:22-199) */
	void (*wavg)(const ok_projector *p, int imgX, int imgY,
	             const float *eulers, unsigned long orientation_num,
	             const float *img_re, const float *img_im,
	             const float *trans_x, const float *trans_y,
	             const float *weights, const float *ctfs,
	             float *wdiff2s_parts, float *wdiff2s_AA, float *wdiff2s_XA,
	             unsigned long trans_num, float weight_norm, float significant_weight, float part_scale);

	/* CpuKernels::backprojectRef3D<CTF_PREMULTIPLIED=false> (src/acc/cpu/cpu_kernels/BP.h:497-753) */
	void (*backproject)(const ok_backprojector *bp, int imgX, int imgY,
	                    const float *img_re, const float *img_im,
	                    const float *trans_x, const float *trans_y,
	                    const float *weights, const float *Minvsigma2s, const float *ctfs,
	                    unsigned long trans_num, float significant_weight, float weight_norm,
	                    const float *eulers, unsigned long image_count);

	/* row locks shared by concurrent backproject() calls on one accumulator
	 * (AccBackprojector::mutexes, src/acc/acc_backprojector_impl.h:61) */
	void *(*bp_sync_alloc)(int mdlY, int mdlZ);
	void (*bp_sync_free)(void *sync);

	/* CpuKernels::backproject2D<CTF_PREMULTIPLIED=false> (src/acc/cpu/cpu_kernels/BP.h:11-218): 2D classification,
	 * accumulators [mdlY][mdlX] (ok_backprojector with mdlZ == 1).  Projection-type kernels need no 2D twin: a 2D
	 * reference is handed to them as a two-plane volume whose second plane is zero (see oracle/bindings.py Projector). */
	void (*backproject2d)(const ok_backprojector *bp, int imgX, int imgY,
	                      const float *img_re, const float *img_im,
	                      const float *trans_x, const float *trans_y,
	                      const float *weights, const float *Minvsigma2s, const float *ctfs,
	                      unsigned long trans_num, float significant_weight, float weight_norm,
	                      const float *eulers, unsigned long image_count);

	/* CpuKernels::diff2_CC_coarse_2D<REF3D=true> (src/acc/cpu/cpu_kernels/diff2.h:611-742): first-iteration cross-correlation
	 * (--firstiter_cc / --always_cc).  diff2s[o*T+t] += -sum(corr Re(ref conj(shift_t img))) / sqrt(sum(corr |ref|^2)) */
	void (*diff2_cc_coarse)(const ok_projector *p, int imgX, int imgY,
	                        const float *eulers, unsigned long O,
	                        const float *trans_x, const float *trans_y, unsigned long T,
	                        const float *img_re, const float *img_im, const float *corr,
	                        float *diff2s);

	/* CpuKernels::diff2_CC_fine_2D<REF3D=true> (src/acc/cpu/cpu_kernels/diff2.h:904-1050) */
	void (*diff2_cc_fine)(const ok_projector *p, int imgX, int imgY,
	                      const float *eulers,
	                      const float *trans_x, const float *trans_y,
	                      const float *img_re, const float *img_im, const float *corr,
	                      unsigned long orientation_num, unsigned long translation_num,
	                      unsigned long num_jobs,
	                      const unsigned long *rot_idx, const unsigned long *trans_idx,
	                      const unsigned long *job_idx, const unsigned long *job_num,
	                      float *diff2s);

	/* CpuKernels::backproject3D_SGD<DATA3D=false, CTF_PREMULTIPLIED=false> (src/acc/cpu/cpu_kernels/BP.h:757-1047; CUDA twin
	 * BP.cuh:406-656): gradient refinement (baseMLO->do_grad).  Per pixel the weighted RESIDUAL
	 * sum_t w_t (shift_t(img) - ctf * ref) is back-projected instead of the weighted image; every pixel of the half image
	 * (no circle bound), phases by sincosf per (pixel, translation). */
	void (*backproject_sgd)(const ok_backprojector *bp, const ok_projector *p, int imgX, int imgY,
	                        const float *img_re, const float *img_im,
	                        const float *trans_x, const float *trans_y,
	                        const float *weights, const float *Minvsigma2s, const float *ctfs,
	                        unsigned long trans_num, float significant_weight, float weight_norm,
	                        const float *eulers, unsigned long image_count);

	/* CpuKernels::backproject2D_SGD<CTF_PREMULTIPLIED=false> (src/acc/cpu/cpu_kernels/BP.h:1047-1262; CUDA twin BP.cuh:659-826):
	 * gradient refinement of 2D references (RELION 4's default 2D classification), accumulators [mdlY][mdlX] */
	void (*backproject2d_sgd)(const ok_backprojector *bp, const ok_projector *p, int imgX, int imgY,
	                          const float *img_re, const float *img_im,
	                          const float *trans_x, const float *trans_y,
	                          const float *weights, const float *Minvsigma2s, const float *ctfs,
	                          unsigned long trans_num, float significant_weight, float weight_norm,
	                          const float *eulers, unsigned long image_count);
} ok_kernel_table;

#ifdef __cplusplus
}
#endif
#endif
