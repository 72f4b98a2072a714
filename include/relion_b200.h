/*
 * relion_b200 — C-ABI of the B200-native (sm_100a) expectation-step library.
 *
 * This is the drop-in boundary for RELION's accelerator path (library target `relion_gpu_util`,
 * /root/reference/src/apps/CMakeLists.txt:246-248): the functions below are what a replacement for
 * MlDeviceBundle / MlOptimiserCuda / AccProjector / AccBackprojector
 * (src/acc/cuda/cuda_ml_optimiser.h:18-144, src/acc/acc_projector.h:17-104,
 * src/acc/acc_backprojector.h:24-99) binds to.  INTEGRATION.md shows the C++ adapter.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  HOST pointers unless a name says `_dev`.
 *   - every function returns 0 (RB_OK) or a negative rb_status; rb_last_error() gives the text
 *     (thread-local).  The adapter turns a non-zero status into RelionError, matching
 *     HANDLE_ERROR -> CRITICAL -> REPORT_ERROR (src/acc/cuda/cuda_settings.h:48-68).
 *   - the library owns all device memory; the caller owns host buffers.
 *   - one rb_ctx per device.  A ctx is thread-compatible (calls on one ctx must be serialised);
 *     the batched rb_estep_pool() replaces the reference's "one particle per OpenMP thread" fan-out
 *     (src/ml_optimiser.cpp:4280), so no concurrent entry is needed.
 *   - there is NO CPU fallback: without a CUDA device rb_ctx_create() fails with RB_ERR_CUDA.
 *   - scope: 3D or 2D references / 2D images, no helical-segment or tomo branches of the E-step; both criteria (Gaussian
 *     squared difference and the first-iteration / --always_cc cross-correlation, rb_model.do_cc); weighted-image and
 *     gradient (SGD / VDAM residual, rb_model.do_grad) back-projection; pool-level left / right matrices (MBL / MBR:
 *     anisotropic magnification, optics groups with their own box or pixel size, body matrices: rb_particles.mat_left /
 *     mat_right, rb_model.ref_max_r, "Optics groups" below); reconstruction with point-group and helical symmetry.
 *
 * Index conventions follow the reference:
 *   coarse hidden index  ihidden      = ((iclass*n_dir + idir)*n_psi + ipsi)*n_trans + itrans
 *   fine hidden index    ihidden_over = (ihidden*n_over_rot + iover_rot)*n_over_trans + iover_trans
 *   where (idir, ipsi) index the particle's own (prior-selected) lists when local searches are on.
 *   Fourier half-image: pixel = iy*(n/2+1) + x, y = iy <= n/2 ? iy : iy-n  (FFTW layout, x fastest).
 *   Volume: voxel = (z-zinit)*Y*X + (y-yinit)*X + x, x in [0, X).
 */
#ifndef RELION_B200_H_
#define RELION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_VERSION 100

typedef enum {
	RB_OK = 0,
	RB_ERR_CUDA = -1,          /* CUDA runtime error (message has file:line)                      */
	RB_ERR_ARG = -2,           /* invalid argument / unsupported configuration                    */
	RB_ERR_STATE = -3,         /* call order (e.g. estep before set_reference)                    */
	RB_ERR_CAPACITY = -4,      /* fine-pass workspace too small for this pool: split the pool     */
	RB_ERR_TRANSLIM = -5,      /* more translations than supported (ERR_TRANSLIM analogue,
	                              acc_helper_functions_impl.h:1247-1253)                          */
	RB_ERR_NO_SIGNIFICANT = -6,/* ERRFILTEREDZERO / ERRNOSIGNIFS (acc_ml_optimiser_impl.h:2242-2299) */
	RB_ERR_SUMWEIGHT_ZERO = -7,/* ERRSUMWEIGHTZERO (acc_ml_optimiser_impl.h:2505-2520)            */
	RB_ERR_PMAX = -8           /* Pmax > 1 (acc_ml_optimiser_impl.h:2924-2929)                    */
} rb_status;

typedef struct rb_ctx rb_ctx;

/* ------------------------------------------------------------------------------------------------
 * Context  (MlDeviceBundle ctor/setDevice/dtor — cuda_ml_optimiser.h:18-76, cuda_ml_optimiser.cu:85)
 * ---------------------------------------------------------------------------------------------- */
int rb_ctx_create(int device, rb_ctx **out);
void rb_ctx_destroy(rb_ctx *ctx);
const char *rb_last_error(void);
int rb_version(void);
/* device synchronise (MlDeviceBundle::syncAllBackprojects, cuda_ml_optimiser.h:60-64) */
int rb_sync(rb_ctx *ctx);
/* number of kernel launches issued by this ctx since creation (bench.py "gpu_launches") */
long long rb_launch_count(rb_ctx *ctx);
/* device time (ms) spent in the named stage during the last rb_estep_pool call; stages:
 * "coarse", "weights_coarse", "fine_setup", "fine", "weights_fine", "store", "total".
 * Valid after rb_sync().  Returns <0 if unknown. */
double rb_stage_ms(rb_ctx *ctx, const char *stage);
/* CUDA-event stopwatch on the stream the kernels are launched on (bench.py's timed region):
 * rb_timer_start records an event; rb_timer_stop records another, waits for it and returns the
 * elapsed device time in milliseconds through *ms. */
int rb_timer_start(rb_ctx *ctx);
int rb_timer_stop(rb_ctx *ctx, double *ms);

/* ------------------------------------------------------------------------------------------------
 * Reference volumes (AccProjector::setMdlDim + initMdl, acc_projector_impl.h:5-312; fed from
 * MlModel::PPref[k].data, cuda_ml_optimiser.cu:116-127).
 * vol: complex (re,im) pairs, [mdlZ][mdlY][mdlX], x >= 0 half, mdlInitY = mdlInitZ = -(mdlY-1)/2.
 * mdlZ == 1: a 2D reference (2D classification, project2Dmodel acc_projectorkernel_impl.h:233-300); pass
 * rot = tilt = 0 in rb_sampling and n_over_rot = 2^oversampling.  rb_bp_init with mdlZ == 1 likewise gives a 2D
 * accumulator (backproject2D, BP.cuh:22-172) and rb_bp_get then returns [mdlY][mdlX] arrays.
 * ---------------------------------------------------------------------------------------------- */
int rb_set_reference(rb_ctx *ctx, int iclass, const double *vol_complex,
                     int mdlX, int mdlY, int mdlZ, int mdlInitY, int mdlInitZ,
                     int mdlMaxR, double padding_factor);
int rb_set_reference_f32(rb_ctx *ctx, int iclass, const float *vol_complex,
                         int mdlX, int mdlY, int mdlZ, int mdlInitY, int mdlInitZ,
                         int mdlMaxR, double padding_factor);

/* Reference of class iclass from a real-space map on the device (SURVEY.md 8f "next" row 3): Projector::
 * computeFourierTransformMap (src/projector.cpp:116-592) for a 3D reference used with 2D images: gridding correction,
 * zero-padding, forward FFT, centring, windowing to r_max = current_size / 2, normfft.  map: [ori_size]^3 floats, origin
 * at ori_size/2; power_spectrum: [ori_size/2+1] radial power of the reference (tau2 input of the M-step), may be NULL. */
int rb_set_reference_from_map(rb_ctx *ctx, int iclass, const float *map, int ori_size, int current_size, double padding_factor,
                              double *power_spectrum);

/* ------------------------------------------------------------------------------------------------
 * Back-projection accumulators (AccBackprojector::setMdlDim/initMdl/clear/getMdlData,
 * acc_backprojector_impl.h:13-186).  Device layout is one interleaved float4 (re, im, weight, 0)
 * per voxel so that each trilinear corner is one 16-byte vector reduction.
 * ---------------------------------------------------------------------------------------------- */
int rb_bp_init(rb_ctx *ctx, int iclass, int mdlX, int mdlY, int mdlZ,
               int mdlInitY, int mdlInitZ, int maxR, double padding_factor);
int rb_bp_clear(rb_ctx *ctx, int iclass);
/* getMdlData: three SoA arrays of mdlX*mdlY*mdlZ floats (src/ml_optimiser.cpp:3813-3838) */
int rb_bp_get(rb_ctx *ctx, int iclass, float *real, float *imag, float *weight);
/* device pointer + float count of the interleaved accumulator, for the per-iteration sum over GPUs
 * (ncclAllReduce(ncclFloat, ncclSum) replaces MlOptimiserMpi::combineAllWeightedSums,
 * src/ml_optimiser_mpi.cpp:2028-2185).  The pointer stays valid until rb_bp_init/rb_ctx_destroy. */
int rb_bp_device_buffer(rb_ctx *ctx, int iclass, void **dptr, size_t *n_floats);

/* ------------------------------------------------------------------------------------------------
 * Per-iteration reduction over the GPUs of one box (replaces MlOptimiserMpi::combineAllWeightedSums,
 * src/ml_optimiser_mpi.cpp:2028-2185, and MlWsumModel::pack / unpack, src/ml_model.cpp:1881-2049): NCCL over NVLink,
 * loaded at run time (libnccl.so.2, or RB_NCCL_LIB).  One rb_comm per context (rank).
 *   several processes, one GPU each : rank 0 calls rb_comm_unique_id, ships the 128 bytes to the others by any means,
 *                                     every rank calls rb_comm_create
 *   one process, several GPUs       : rb_comm_create_all (what RELION's threads-per-device layout needs); the collective
 *                                     calls of the ranks then come from different threads, or from one thread between
 *                                     rb_comm_group_start / rb_comm_group_end
 * rb_bp_allreduce sums the accumulators of all nr_classes classes in place, on the context's own stream (ordered after
 * the E-step, waits for completion): the zero pad lane of the float4 voxels does not travel.  rb_wsum_allreduce sums one
 * fp64 host vector holding every other weighted sum in MlWsumModel::pack order (relion_b200::WsumPack in the C++ adapter).
 * rb_bp_allreduce_nccl takes a caller-owned ncclComm_t (e.g. a half-set communicator made with ncclCommSplit) and may
 * return without waiting (wait == 0: later library calls on the context are stream-ordered behind it anyway).
 * ---------------------------------------------------------------------------------------------- */
typedef struct rb_comm rb_comm;
int rb_comm_unique_id(void *id128);
int rb_comm_create(rb_ctx *ctx, int nranks, int rank, const void *id128, rb_comm **out);
int rb_comm_create_all(rb_ctx *const *ctxs, int n, rb_comm **out);
void rb_comm_destroy(rb_comm *comm);
int rb_comm_size(const rb_comm *comm);
int rb_comm_rank(const rb_comm *comm);
void *rb_comm_handle(rb_comm *comm);            /* the ncclComm_t */
int rb_comm_group_start(void);
int rb_comm_group_end(void);
int rb_bp_allreduce(rb_ctx *ctx, rb_comm *comm);
int rb_bp_allreduce_nccl(rb_ctx *ctx, void *nccl_comm, int nr_classes, int wait);
int rb_wsum_allreduce(rb_ctx *ctx, rb_comm *comm, double *wsums, size_t n);

/* BackProjector::symmetrise (src/backprojector.cpp:2136-2480) on accumulator iclass: enforceHermitianSymmetry of the x = 0
 * plane, then applyPointGroupSymmetry with the nsym rotation matrices R ([nsym][9] row-major, the R of
 * SymList::get_matrices; nsym == 0 for C1).  RELION calls this before every reconstruct (src/ml_optimiser.cpp:4930-5044). */
int rb_bp_symmetrise(rb_ctx *ctx, int iclass, const double *R, int nsym);
/* The same with helical symmetry (BackProjector::symmetrise(nr_helical_asu, helical_twist, helical_rise),
 * applyHelicalSymmetry src/backprojector.cpp:2167-2322) applied between the Hermitian fold and the point group: the
 * accumulator also receives its copies rotated about Z by hh * -helical_twist degrees and phase-shifted along z by
 * hh * helical_rise, hh in [-nr_helical_asu/2, nr_helical_asu/2 + nr_helical_asu%2) without 0.  helical_rise is in pixels (the
 * caller divides by the pixel size, src/reconstructor.cpp:771), ori_size the unpadded box.  nr_helical_asu < 2: no helical part. */
int rb_bp_symmetrise_helical(rb_ctx *ctx, int iclass, const double *R, int nsym, int nr_helical_asu, double helical_twist,
                             double helical_rise, int ori_size);

/* Reconstruction of a map from accumulator iclass on the device (SURVEY.md 8f "next" row 2): BackProjector::reconstruct,
 * default skip_gridding branch (src/backprojector.cpp:1379-1575) + windowToOridimRealSpace (:2530-2665) + griddingCorrect
 * (src/projector.cpp:595-628).  tau2: [n_tau2] spectrum for the MAP term (NULL: plain weighted average, do_map == false);
 * vol_out: [ori_size]^3 floats, origin at ori_size/2.  The accumulator is left untouched (all-reduce it first on several GPUs). */
int rb_reconstruct(rb_ctx *ctx, int iclass, int ori_size, const double *tau2, int n_tau2, double tau2_fudge, int minres_map,
                   float *vol_out);
/* The same with the iterative gridding of Pipe & Menon instead of the plain division (--dont_skip_gridding,
 * src/backprojector.cpp:1577-1700 + convoluteBlobRealSpace :2483-2528): max_iter_preweight iterations (RELION: 10) of
 * Fnewweight /= |blob-convolution of Fnewweight * Fweight| in double precision, 3D transforms of the padded volume by cuFFT;
 * normalise as in BackProjector::reconstruct (1 unless the caller rescales the weights). */
int rb_reconstruct_gridding(rb_ctx *ctx, int iclass, int ori_size, const double *tau2, int n_tau2, double tau2_fudge, int minres_map,
                            int max_iter_preweight, double normalise, float *vol_out);

/* BackProjector::updateSSNRarrays (src/backprojector.cpp:1041-1204) on accumulator iclass (2D or 3D): sigma2 = shell
 * average of the inverse noise power of the reconstruction; tau2 recomputed from `fsc` (gold-standard FSC between the
 * half-maps) when update_tau2_with_fsc, else kept; data_vs_prior and fourier_coverage for model.star.  All spectra are
 * [ori_size/2 + 1] doubles; fsc / avgctf2 may be NULL.  Call it on the all-reduced accumulator, before rb_reconstruct. */
int rb_update_ssnr(rb_ctx *ctx, int iclass, int ori_size, double tau2_fudge, double *tau2_io, double *sigma2_out,
                   double *data_vs_prior_out, double *fourier_coverage_out, const double *fsc, const double *avgctf2,
                   int update_tau2_with_fsc, int is_whole_instead_of_half);

/* ------------------------------------------------------------------------------------------------
 * Sampling tables for this iteration (outputs of HealpixSampling, src/healpix_sampling.cpp:
 * getDirection/getPsiAngle :1662-1700, getOrientations :1832, getTranslationsInPixel :1724).
 * Host-side list generation stays RELION's; the library turns angles into matrices itself:
 * coarse in fp32 on the device (cuda_kernel_make_eulers_3D, helper.cuh:713-840), fine in fp64
 * then cast (generateEulerMatrices, acc_helper_functions_impl.h:198-262).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
	int n_dir, n_psi;            /* full coarse grid                                         */
	const double *rot, *tilt;    /* [n_dir] degrees                                          */
	const double *psi;           /* [n_psi] degrees                                          */
	int n_over_rot;              /* oversamplingFactorOrientations: 1 or 8                   */
	/* oversampled triplets [(idir*n_psi+ipsi)*n_over_rot + iover] degrees; may be NULL when
	 * n_over_rot == 1 (then the coarse angles are used)                                     */
	const double *over_rot, *over_tilt, *over_psi;
	int n_trans;                 /* coarse translations                                      */
	const double *trans_x, *trans_y;           /* [n_trans] pixels (oversampling 0)          */
	int n_over_trans;            /* oversamplingFactorTranslations: 1 or 4                   */
	const double *over_trans_x, *over_trans_y; /* [n_trans*n_over_trans] pixels              */
} rb_sampling;
int rb_set_sampling(rb_ctx *ctx, const rb_sampling *s);

/* ------------------------------------------------------------------------------------------------
 * Model / optimiser state read by the E-step (data contract of SURVEY.md §8b):
 * MlModel (src/ml_model.h) + MlOptimiser flags (src/ml_optimiser.h).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
	int nr_classes;
	int ori_size;                /* mymodel.ori_size == image_full_size (single optics-group size) */
	int coarse_size;             /* image_coarse_size  (src/ml_optimiser.cpp:5743-5777)      */
	int current_size;            /* image_current_size                                       */
	double pixel_size;           /* Angstrom / pixel                                         */
	int nr_optics_groups;
	const double *sigma2_noise;  /* [nr_optics_groups][ori_size/2+1]                         */
	int nr_groups;
	const double *scale_correction;   /* [nr_groups]                                         */
	const double *pdf_class;          /* [nr_classes]                                        */
	const double *pdf_direction;      /* [nr_classes][n_dir] (used when no orientational prior) */
	const double *data_vs_prior_class;/* [nr_classes][ori_size/2+1] (scale-correction shells) */
	double sigma2_offset;        /* Angstrom^2 (mymodel.sigma2_offset)                       */
	double offset_range;         /* >0: sigma2 = range^2/9 (acc_ml_optimiser_impl.h:1916-1921) */
	double sigma2_fudge;
	double adaptive_fraction;    /* --adaptive_fraction, default 0.999                       */
	int maximum_significants;    /* --maxsig, <=0: off                                       */
	int do_ctf_correction, refs_are_ctf_corrected, do_scale_correction, do_map;
	int ctf_premultiplied;
	int bp_circle_bound;         /* 1: drop x >= floor(sqrt((n/2)^2-y^2)) like ALTCPU BP.h:565 (default);
	                                0: CUDA BP.cuh behaviour (no extra bound)                */
	int do_cc;                   /* (iter == 1 && do_firstiter_cc) || do_always_cc (acc_ml_optimiser_impl.h:1164):
	                                normalised cross-correlation instead of the Gaussian squared difference in both
	                                passes (cuda_kernel_diff2_CC_coarse / _fine, diff2.cuh:336-640), weight one for the
	                                best pose and zero elsewhere (:2012-2071); rb_particle_out.dLL_nolog is then
	                                -min_diff2, which IS the particle's dLL (:3571-3572, no logsigma2 term)       */
	const double *prior_offset_class; /* [nr_classes][2] pixels: mymodel.prior_offset_class, the per-class centre of the
	                                translation prior of 2D references (ref_dim == 2: acc_ml_optimiser_impl.h:2100-2104
	                                for the fine-pass weights, :2673-2677 for wsum_sigma2_offset).  As in the reference
	                                the coarse-pass weights use the FIRST class' centre for every class (the coarse kernel
	                                indexes pdf_offset by translation only, :2187-2196).  Pass it for 2D references
	                                (zeros at iteration 1); NULL (3D references): rb_particles.prior_offset is the centre
	                                and rb_pool_out.wsum_prior_offset_class stays untouched                       */
	int do_grad;                 /* gradient (SGD / VDAM) refinement, baseMLO->do_grad (acc_ml_optimiser_impl.h:3418): the
	                                back-projection accumulates the weighted RESIDUAL sum_t w_t (X_t - CTF A) instead of the
	                                weighted image (cuda_kernel_backproject3D_SGD / _2D_SGD, BP.cuh:406-826; ALTCPU BP.h:757-1262: every
	                                pixel, no circle bound); 3D and 2D references.  With grad_pseudo_halfsets (= do_grad in
	                                src/ml_optimiser.cpp:1192) particles go into accumulator iclass + (part_id % 2) * nr_classes
	                                (:3395-3400): initialise 2 * nr_classes accumulators and pass rb_particles.bp_offset    */
	int do_skip_rotate;          /* do_skip_align || do_skip_rotate (--skip_align / --skip_rotate: only classify): every particle keeps
	                                its own orientation, passed as ONE-entry lists (dir_idx / psi_idx -> the rb_sampling tables, which
	                                then hold the pool's orientations, MlOptimiser::expectationSomeParticles src/ml_optimiser.cpp:4180-4190),
	                                and the orientation prior is pdf_class[iclass] (acc_ml_optimiser_impl.h:1966-1967).  With
	                                --skip_align the sampling holds the single translation (0, 0) and the particle's own fractional
	                                offset (:4196-4225) goes into rb_particles.pre_shift                                   */
	int ref_max_r;               /* 0, or the references' r_max (rb_set_reference maxR) when it is SMALLER than current_size / 2 —
	                                an optics group whose box is bigger than the model's (see "Optics groups" below).  The
	                                reference's fine-pass and wavg kernels then skip the image rows maxR < iy < imgY - maxR
	                                except their pixel x = maxR (cpu_kernels/diff2.h:347-355, wavg.h:74-82; the coarse kernel keeps
	                                every pixel, diff2.h:109-110), and so do the pixel sets built here.  rb_estep_* refuses pools
	                                whose references end inside the window when this is not set                               */
} rb_model;
/* Optics groups.  One rb_set_model holds ONE image geometry: ori_size is image_full_size[optics_group], current_size /
 * coarse_size its image_current_size / image_coarse_size (src/ml_optimiser.cpp:5735-5777), pixel_size its pixel size, sigma2_noise
 * indexed by the IMAGE's shells.  For an optics group whose box or pixel size differs from the model's
 * (src/ml_optimiser.cpp:6802-6821, 6840-6875) call rb_set_model + rb_set_sampling again before its pools, with that group's sizes,
 * sigma2_noise resampled as sigma2[ROUND(remap * ires)] (an infinite sigma2 where the remapped shell leaves the spectrum: Minvsigma2
 * stays zero there), and pass rb_particles.mat_left = applyScaleDifference(applyAnisoMag(I)) — the references keep their own box
 * (rb_set_reference) and are read through the scaled matrices.  include/relion_b200_adapter.hpp does exactly this. */
int rb_set_model(rb_ctx *ctx, const rb_model *m);
/* Call order: rb_set_model, rb_set_sampling, then (without orientational priors) rb_set_pdf_direction:
 * [nr_classes][n_dir] needs both the model (values) and the sampling (n_dir).  rb_set_model picks
 * m->pdf_direction up itself when the sampling is already known. */
int rb_set_pdf_direction(rb_ctx *ctx, const double *pdf_direction);

/* ------------------------------------------------------------------------------------------------
 * One pool of particles (what getFourierTransformsAndCtfs leaves in OptimisationParamters,
 * acc_ml_optimiser_impl.h:11-1010: Fimg, Fimg_nomask, Fctf at image_current_size, highres_Xi2,
 * old_offset, prior; plus the per-particle prior-selected orientation lists of local searches,
 * HealpixSampling::selectOrientationsWithNonZeroPriorProbability, healpix_sampling.cpp:695).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
	int n_particles;
	/* [P][current_size][current_size/2+1] complex (re,im) fp32: masked image FT, unmasked image FT */
	const float *Fimg, *Fimg_nomask;
	/* [P][current_size][current_size/2+1] fp32 CTF (ignored if !do_ctf_correction) */
	const float *Fctf;
	const int *group_id;         /* [P] */
	const int *optics_group;     /* [P] */
	const double *highres_Xi2;   /* [P]  power beyond current_size (op.highres_Xi2_img)          */
	const double *old_offset;    /* [P][2] pixels (already applied to the image)                  */
	const double *prior_offset;  /* [P][2] pixels (op.prior)                                      */
	/* local angular searches: per-particle lists; all NULL => every (dir,psi) with
	 * pdf_direction > 0 is searched (NOPRIOR, global search)                                     */
	const int *dir_off;          /* [P+1] offsets into dir_idx/dir_prior                          */
	const int *dir_idx;          /* pointer_dir_nonzeroprior: indices into the full n_dir grid    */
	const double *dir_prior;     /* directions_prior                                              */
	const int *psi_off;          /* [P+1]                                                         */
	const int *psi_idx;          /* pointer_psi_nonzeroprior                                      */
	const double *psi_prior;     /* psi_prior                                                     */
	const int *bp_offset;        /* [P] or NULL (0): added to the class index to select the accumulator the particle is
	                                back-projected into (pseudo half-sets of gradient refinement: (part_id % 2) * nr_classes) */
	const double *pre_shift;     /* [P][2] pixels or NULL: a per-particle translation applied on top of every sampled one (the images
	                                are multiplied by its phase ramp once, on the device; the translation prior and the sigma2_offset
	                                sums see old_offset + pre_shift + sampled translation).  --skip_align: the fractional part of the
	                                old offset, which the reference samples as the particle's only translation                 */
	const double *mat_left;      /* [9] row-major or NULL: MBL of the pool, every orientation matrix becomes
	                                inverse(mat_left * A(rot, tilt, psi) * mat_right) in both passes and the store stage
	                                (cuda_kernel_make_eulers_3D<invert, doL, doR>, helper.cuh:713-840; generateEulerMatrices(...,
	                                L, R), acc_helper_functions_impl.h:198-262).  The reference passes mag =
	                                applyScaleDifference(applyAnisoMag(I)) here (acc_ml_optimiser_impl.h:1098-1103): optics groups
	                                with anisotropic magnification or with a box / pixel size that differs from the model's, and
	                                Aori * orient_bodies^T * A_rot90 for multi-body refinement.  3D references only. */
	const double *mat_right;     /* [9] row-major or NULL: MBR (orient_bodies[ibody], :1085) */
} rb_particles;

/* Per-particle results (what storeWeightedSums writes to exp_metadata and folds into wsum_model,
 * acc_ml_optimiser_impl.h:2858-2931, 3466-3656).  The host adapter finishes the bookkeeping in
 * double exactly where the reference does (norm correction, dLL - logsigma2, scale XA/AA / scale). */
typedef struct {
	int64_t best_ihidden_over;   /* op.max_index.fineIdx                                          */
	int best_class, best_idir, best_ipsi, best_iover_rot, best_itrans, best_iover_trans;
	int nr_significant_coarse;   /* METADATA_NR_SIGN                                              */
	int n_fine_orient, n_fine_samples; /* sizes of the fine pass (diagnostics / roofline)         */
	float min_diff2_coarse;
	float sum_weight_coarse;
	float significant_weight_coarse;
	float min_diff2;             /* op.min_diff2 after the fine pass (+50 - max shift applied)    */
	float max_weight;            /* op.max_weight                                                 */
	float sum_weight;            /* op.sum_weight (fine pass)                                     */
	float significant_weight;    /* op.significant_weight (fine pass)                             */
	float pmax;                  /* METADATA_PMAX = max_weight / sum_weight                       */
	int n_bp_orient;             /* fine orientations holding >= 1 significant sample: the ones wavg / back-projection
	                                actually process (diagnostics / roofline; 0 with do_skip_maximization)          */
	double dLL_nolog;            /* log(sum_weight) - min_diff2  (host subtracts logsigma2)       */
	double wsum_norm_correction; /* sum over shells ires>-1 of wdiff2                              */
	double wsum_XA, wsum_AA;     /* exp_wsum_scale_correction_XA/AA before the /scale division     */
	double sumw;                 /* thr_sumw_group                                                */
	double wsum_sigma2_offset;   /* thr_wsum_sigma2_offset (Angstrom^2)                           */
} rb_particle_out;

typedef struct {
	rb_particle_out *particles;  /* [P]                                                           */
	float *wsum_sigma2_noise;    /* [P][ori_size/2+1] per-particle thr_wsum_sigma2_noise (may be NULL) */
	double *wsum_pdf_direction;  /* [nr_classes][n_dir]  += over the pool (may be NULL)           */
	double *wsum_pdf_class;      /* [nr_classes]         += over the pool (may be NULL)           */
	double *wsum_prior_offset_class; /* [nr_classes][2] += over the pool, Angstrom: thr_wsum_prior_offsetx/y_class
	                                (acc_ml_optimiser_impl.h:2847-2851), accumulated when rb_model.prior_offset_class is
	                                given (2D references); may be NULL                                              */
} rb_pool_out;

/* Batched E-step over a pool: coarse diff2 -> weights/significance -> fine diff2 -> weights ->
 * weighted sums + back-projection (accDoExpectationOneParticle for every particle of the pool,
 * acc_ml_optimiser_impl.h:3672-3962).  Back-projection lands in the ctx's accumulators.
 * flags: bit0 = skip back-projection/wavg (do_skip_maximization). */
int rb_estep_pool(rb_ctx *ctx, const rb_particles *pool, rb_pool_out *out, unsigned flags);

/* Same, split in two so a caller can overlap the H2D upload of pool i+1 with the compute of pool i:
 * rb_pool_upload stages the pool into one of two device slots asynchronously (pinned host memory
 * recommended); rb_estep_slot runs the E-step on a staged slot. */
int rb_pool_upload(rb_ctx *ctx, int slot, const rb_particles *pool);
int rb_estep_slot(rb_ctx *ctx, int slot, rb_pool_out *out, unsigned flags);
/* Asynchronous pair: rb_estep_slot_nocopy only enqueues the E-step of a staged slot (no host<->device copy, returns
 * immediately); rb_estep_fetch waits for THAT slot's completion event on a separate stream and reads its results, so a
 * pipelined caller runs   launch(i) ; fetch(i-1) ; upload(i+1)   and keeps the GPU busy across pools.  A slot must be
 * fetched before it is uploaded again.  (Also the device-resident timing entry of bench.py.) */
int rb_estep_slot_nocopy(rb_ctx *ctx, int slot, unsigned flags);
int rb_estep_fetch(rb_ctx *ctx, int slot, rb_pool_out *out);

/* Test hook: the coarse-pass posterior weights of one particle of a slot after its E-step (the reference's Mweight block
 * after convertAllSquaredDifferencesToWeights(0), acc_ml_optimiser_impl.h:2188-2345): n = nr_classes * nd * np * n_trans
 * floats in coarse hidden-index order.  Parity tests feed them to the reference's own significance rule
 * (acc_helper_functions.h:226-232).  Valid until the slot is uploaded again; *n_out receives n (out may be NULL to query). */
int rb_debug_coarse_weights(rb_ctx *ctx, int slot, int particle, float *out, long long capacity, long long *n_out);
/* Test hook: the coarse-pass Euler matrices the device built in rb_set_sampling (cuda_kernel_make_eulers_3D<invert = true>,
 * src/acc/cuda/cuda_kernels/helper.cuh:713-840; ALTCPU cpu_kernel_make_eulers_3D, cpu_kernels/helper.cpp): [n_dir][n_psi][9] fp32. */
int rb_debug_coarse_eulers(rb_ctx *ctx, float *out, long long capacity);
/* Test hook: the prepared coarse-window image of one particle of a slot after its E-step: [coarse_size][coarse_size/2+1] float4
 * (X'.re, X'.im, corr / 2, 0) = the pixel_correction / corr_img step (acc_ml_optimiser_impl.h:1251-1268, buildCorrImage
 * acc_helper_functions_impl.h:164-196, Minvsigma2 src/ml_optimiser.cpp:6868-6879) applied once per particle. */
int rb_debug_prepared_coarse_image(rb_ctx *ctx, int slot, int particle, float *out);
/* Test hook: the noise images [n_particles][image_size][image_size] the last rb_pool_prepare with noise_seed blended into
 * (centred like the particle images).  Tests rebuild the masked image from them exactly. */
int rb_debug_prep_noise(rb_ctx *ctx, int n_particles, int image_size, float *out);

/* ------------------------------------------------------------------------------------------------
 * Image preparation on the device ("next" row 1): getFourierTransformsAndCtfs
 * (acc_ml_optimiser_impl.h:11-1010) for a whole pool: integer translation by the rounded old offset and norm
 * correction, transform of the unmasked image (Fimg_nomask), soft circular zero-mask, transform of the masked
 * image (Fimg), power spectrum / highres_Xi2 beyond the current size, CTF image from the CTF parameters.
 * Covered branch: 2D images, one body, zero- or noise-filled soft mask, beam tilt / MTF as a per-optics-group factor image, no helix / tomo,
 * CTF without phase flipping.
 * The slot is then ready for rb_estep_slot exactly as after rb_pool_upload.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
	int n_particles;
	int image_size;              /* box size n == rb_model.ori_size                                */
	const float *images;         /* [P][n][n] real-space particle images (exp_imagedata as XFLOAT)  */
	const double *norm_factor;   /* [P] avg_norm_correction / normcorr; NULL = 1 (:421)             */
	const double *old_offset;    /* [P][2] METADATA_XOFF/YOFF, rounded inside like :216             */
	const double *prior_offset;  /* [P][2]                                                          */
	const int *group_id, *optics_group;   /* [P]                                                    */
	/* CTF (CTF::setValuesByGroup + getFftwImage, :822-840): per particle; NULL Bfac/scale/phase = 0/1/0 */
	const double *ctf_defU, *ctf_defV, *ctf_defAngle, *ctf_Bfac, *ctf_scale, *ctf_phase_shift;
	const double *og_kV, *og_Cs, *og_Q0;  /* [nr_optics_groups] kV, mm, fraction                    */
	double mask_radius;          /* particle_diameter / (2 pixel_size) in pixels; < 0: n/2 (:556)   */
	double width_mask_edge;      /* --mask_edge width in pixels                                     */
	/* local angular searches, as in rb_particles (all NULL: global search)                         */
	const int *dir_off, *dir_idx; const double *dir_prior;
	const int *psi_off, *psi_idx; const double *psi_prior;
	const int *bp_offset;        /* as in rb_particles                                              */
	const float *og_fourier_factor; /* [nr_optics_groups][current_size][current_size/2+1] complex (re, im) or NULL: per optics group,
	                                the factor both transforms are multiplied with after windowing: conj(phase correction) of the
	                                beam tilt / odd Zernike terms (ObservationModel::demodulatePhase,
	                                src/jaz/single_particle/obs_model.cpp:598-626) times avgMTF / MTF (divideByMtf, :528-584), as
	                                acc_ml_optimiser_impl.h:535-536 applies them; the host builds the images from the
	                                observation model once per E-step                                                  */
	const int64_t *noise_seed;   /* [P] random_seed + part_id, or NULL.  NULL: --zero_mask (soft mask towards the background
	                                value of the image's own edge).  Non-NULL: RELION's default, the soft mask blends into a NOISE
	                                image with the spectrum sqrt(sigma2_fudge * sigma2_noise[optics group]) (makeNoiseImage,
	                                src/acc/utilities_impl.h:231-371: independent complex normals per Fourier pixel, inverse FFT;
	                                cosineFilter with the noise as the fill value, acc_ml_optimiser_impl.h:355-400, 660-668).
	                                The generator is counter-based on (seed, pixel): reproducible, but - like the reference's
	                                curand and CPU generators among themselves - not the same random numbers as RELION's */
	const double *mat_left, *mat_right; /* as in rb_particles (NULL: none) */
	const double *pre_shift;     /* as in rb_particles */
	const double *noise_sigma2;  /* [nr_optics_groups][image_size/2+1] or NULL (rb_model.sigma2_noise): the spectrum the noise of
	                                the noise-filled mask is drawn from — remapped_sigma2_noise of acc_ml_optimiser_impl.h:374-384, which
	                                for an optics group with its own box / pixel size is NOT the gather rb_model.sigma2_noise holds */
} rb_raw_particles;
/* power_img: [P][n/2+1] spectrum of the masked full-size transform (op.power_img, used by the host for sigma2_noise
 * beyond the current size), may be NULL. */
int rb_pool_prepare(rb_ctx *ctx, int slot, const rb_raw_particles *raw, float *power_img);
/* read a staged pool back (parity tests of the preparation): [P][cs][cs/2+1] arrays, any pointer may be NULL */
int rb_pool_download(rb_ctx *ctx, int slot, float *Fimg, float *Fimg_nomask, float *Fctf, double *highres_Xi2);

/* ------------------------------------------------------------------------------------------------
 * Particle image feed: MRC stacks -> page-locked pool buffers, ahead of the GPU.
 * Replaces the image half of MlOptimiser::getMetaAndImageDataSubset (src/ml_optimiser.cpp:10285-10406)
 * and Image<T>::readMRC (src/rwMRC.h:140-283).  Host-only code, usable without a context.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rb_mrc rb_mrc;
/* open an .mrc / .mrcs file: modes 0 (signed 8-bit), 1, 2, 6, 12; byte-swapped files are detected like rwMRC.h:149 */
int rb_mrc_open(const char *path, rb_mrc **out);
/* pixel_size = cell / sampling along x (rwMRC.h:238), 0 when the header does not say; any pointer may be NULL */
int rb_mrc_info(const rb_mrc *m, int *nx, int *ny, int *nz, int *mode, float *pixel_size);
/* images `indices[i]` (0-BASED; rlnImageName counts from 1) -> dst[i][ny][nx] as float */
int rb_mrc_read_images(const rb_mrc *m, const long long *indices, int count, float *dst);
void rb_mrc_close(rb_mrc *m);
/* float32 (mode 2) stack / volume with the header fields of Image<T>::writeMRC (rwMRC.h:286-530) */
int rb_mrc_write(const char *path, const float *data, int nx, int ny, int nz, float pixel_size);

typedef struct rb_feed rb_feed;
/* `depth` staging buffers of max_particles images of image_size^2 floats (page-locked when a CUDA device is present),
 * `n_threads` reader threads */
int rb_feed_create(int image_size, int max_particles, int depth, int n_threads, rb_feed **out);
/* queue one pool: particle i is image index[i] (0-based) of stack paths[i]; fails with RB_ERR_STATE when every
 * staging buffer is in use */
int rb_feed_submit(rb_feed *f, const char *const *paths, const long long *index, int n_particles, int *ticket);
/* block until the pool is read; *images = [n_particles][n][n], valid until rb_feed_release(ticket) - pass it as
 * rb_raw_particles.images */
int rb_feed_wait(rb_feed *f, int ticket, const float **images);
int rb_feed_release(rb_feed *f, int ticket);
void rb_feed_destroy(rb_feed *f);

/* ------------------------------------------------------------------------------------------------
 * Stage-level entry points (kernel-granularity twins of AccUtilities::* / run*Kernel,
 * src/acc/utilities.h:946-1620, acc_helper_functions_impl.h).  Used by the parity tests and usable
 * by an adapter that keeps the reference's per-particle driver.
 * ---------------------------------------------------------------------------------------------- */
/* AccProjectorKernel::project3Dmodel over a half image, fine-pass y-wrap; eulers [n][9] fp32,
 * out [n][imgY][imgX] complex interleaved */
int rb_project(rb_ctx *ctx, int iclass, int img_size, const float *eulers, int n, float *out_complex);

/* runDiff2KernelCoarse (acc_helper_functions_impl.h:1139): diff2s[o*T+t] = sum 0.5*corr*|A_o - S_t X|^2 */
int rb_diff2_coarse(rb_ctx *ctx, int iclass, int img_size,
                    const float *eulers, int n_orient,
                    const float *trans_x, const float *trans_y, int n_trans,
                    const float *img_re, const float *img_im, const float *corr,
                    float *diff2s);

/* The same with the first-iteration cross-correlation criterion (runDiff2KernelCoarse with do_CC,
 * acc_helper_functions_impl.h:1737-1809 -> cuda_kernel_diff2_CC_coarse, diff2.cuh:336-460):
 * diff2s[o*T+t] += -sum(corr Re(A_o conj(S_t X))) / sqrt(sum(corr |A_o|^2)) */
int rb_diff2_cc_coarse(rb_ctx *ctx, int iclass, int img_size,
                       const float *eulers, int n_orient,
                       const float *trans_x, const float *trans_y, int n_trans,
                       const float *img_re, const float *img_im, const float *corr,
                       float *diff2s);

/* The coarse-pass contraction of GLOBAL searches, exposed for parity tests: C[M][N] = A[M][K] . B[N][K]^T on the
 * tcgen05 tensor cores with 3xTF32 operand splitting (FP32-equivalent accuracy; north_star "tensor cores are used
 * only for the coarse-pass cross term Re<X_t, CTF*A_r>").  Inside rb_estep_pool the same kernel computes
 * cuda_kernel_diff2_coarse (diff2.cuh:24-189) for every particle of a pool at once when the pool has no
 * per-particle orientation lists.  Environment RB_COARSE_GEMM=0 forces the SIMT kernel, 2 forces the tensor path. */
int rb_gemm_tf32x3(rb_ctx *ctx, const float *A, const float *B, int M, int N, int K, float *C);

/* runDiff2KernelFine (acc_helper_functions_impl.h:1813) with the reference's job lists */
int rb_diff2_fine(rb_ctx *ctx, int iclass, int img_size,
                  const float *eulers, int n_orient,
                  const float *trans_x, const float *trans_y, int n_trans,
                  const float *img_re, const float *img_im, const float *corr, float sum_init,
                  const uint64_t *rot_idx, const uint64_t *trans_idx,
                  const uint64_t *job_idx, const uint64_t *job_num, int n_jobs,
                  float *diff2s, int n_weights);

/* runDiff2KernelFine with do_CC (acc_helper_functions_impl.h:1922-1997 -> cuda_kernel_diff2_CC_fine, diff2.cuh:464-640) */
int rb_diff2_cc_fine(rb_ctx *ctx, int iclass, int img_size,
                     const float *eulers, int n_orient,
                     const float *trans_x, const float *trans_y, int n_trans,
                     const float *img_re, const float *img_im, const float *corr,
                     const uint64_t *rot_idx, const uint64_t *trans_idx,
                     const uint64_t *job_idx, const uint64_t *job_num, int n_jobs,
                     float *diff2s, int n_weights);

/* convertAllSquaredDifferencesToWeights, dense form (acc_ml_optimiser_impl.h:2188-2345):
 * in: diff2 [n_orient*n_trans] (Mweight), priors; out: weights in place, significance flags.
 * filter_zero=1 is the coarse pass (weights>0 only enter the sort), 0 the fine pass. */
typedef struct {
	float min_diff2;
	float max_weight; int64_t max_index;
	float sum_weight; float significant_weight;
	int nr_significant; int n_nonzero;
} rb_weights_out;
int rb_convert_weights(rb_ctx *ctx, float *weights, int64_t n_orient, int n_trans,
                       const float *pdf_orientation, const unsigned char *pdf_orientation_zeros,
                       const float *pdf_offset, const unsigned char *pdf_offset_zeros,
                       double adaptive_fraction, int maximum_significants, int filter_zero,
                       unsigned char *significant, rb_weights_out *out);

/* runWavgKernel (acc_helper_functions_impl.h:316): per-pixel sums over orientations/translations */
int rb_wavg(rb_ctx *ctx, int iclass, int img_size,
            const float *eulers, int n_orient,
            const float *trans_x, const float *trans_y, int n_trans,
            const float *img_re, const float *img_im, const float *weights, const float *ctfs,
            float weight_norm, float significant_weight,
            float *wdiff2s_parts, float *wdiff2s_AA, float *wdiff2s_XA);

/* runBackProjectKernel (acc_helper_functions_impl.h:505) into class iclass' accumulator */
int rb_backproject(rb_ctx *ctx, int iclass, int img_size,
                   const float *eulers, int n_orient,
                   const float *trans_x, const float *trans_y, int n_trans,
                   const float *img_re, const float *img_im,
                   const float *weights, const float *Minvsigma2s, const float *ctfs,
                   float weight_norm, float significant_weight);

/* relion_reconstruct-style posed back-projection (BASELINE config #2; Reconstructor::backprojectOneParticle,
 * src/reconstructor.cpp:328-744 -> BackProjector::backproject2Dto3D, src/backprojector.cpp:55-357, TRILINEAR, no Ewald
 * sphere): n images [n][img_size][img_size/2+1] complex already multiplied by their CTF, weights Fctf = ctf^2 (pixels
 * with weight <= 0 are skipped), one INVERTED 3x3 matrix per image ([n][9] fp32, row-major).  Host buffers (pinned
 * recommended); uploads are chunked and overlapped with the scatter. */
int rb_backproject_posed(rb_ctx *ctx, int iclass, int img_size, int n,
                         const float *F2D_complex, const float *Fctf, const float *eulers);
/* The same from RAW images: what Reconstructor::backprojectOneParticle (src/reconstructor.cpp:428-745) does per particle before
 * backproject2Dto3D - FourierTransform, CenterFFTbySign, shiftImageInFourierTransform by the particle's origin offset, CTF image
 * (CTF::getFftwImage with damping, no flips), F2D *= CTF (unless premultiplied), Fctf = CTF^2, F2D(0, 0) = 0 - runs on the
 * device for a chunk of images at a time (cuFFT + one kernel writing the band-ordered staging buffer of the scatter), so a
 * particle crosses PCIe as 4 bytes per pixel instead of 12 per Fourier pixel.  Branch covered: 2D images, no Ewald sphere, no
 * FOM / per-pixel weights, no reference subtraction. */
typedef struct {
	int n_images, image_size;
	const float *images;         /* [n][image_size][image_size] real space                                          */
	const float *eulers;         /* [n][9] INVERTED 3x3 matrices, as rb_backproject_posed                           */
	const double *shift;         /* [n][2] origin offset in pixels (rlnOriginX/YAngst / pixel size) or NULL         */
	/* CTF per image (NULL ctf_defU: no CTF, weight 1): as rb_raw_particles */
	const double *ctf_defU, *ctf_defV, *ctf_defAngle, *ctf_Bfac, *ctf_scale, *ctf_phase_shift;
	const int *optics_group;     /* [n] or NULL (0)                                                                  */
	const double *og_kV, *og_Cs, *og_Q0;
	double pixel_size;           /* Angstrom / pixel                                                                 */
	int ctf_premultiplied;
} rb_posed_raw;
int rb_backproject_posed_raw(rb_ctx *ctx, int iclass, const rb_posed_raw *raw);
/* The same scatter on a batch that is staged on the device once (roofline measurement without the PCIe copy). */
int rb_bp_posed_stage(rb_ctx *ctx, int img_size, int n, const float *F2D_complex, const float *Fctf, const float *eulers);
int rb_bp_posed_run(rb_ctx *ctx, int iclass);

#ifdef __cplusplus
}
#endif
#endif /* RELION_B200_H_ */
