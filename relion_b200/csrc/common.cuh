// relion_b200 — shared declarations for the sm_100a E-step kernels and the C-ABI glue.
// Product code: must not reference anything under oracle/.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <string>
#include <vector>
#include <map>

#include "relion_b200.h"

// ---------------------------------------------------------------------------------------------
// error handling (HANDLE_ERROR analogue, /root/reference/src/acc/cuda/cuda_settings.h:48-68)
// ---------------------------------------------------------------------------------------------
void rb_set_error(const char *fmt, ...);

#define RB_CUDA(call)                                                                          \
	do {                                                                                       \
		cudaError_t e__ = (call);                                                              \
		if (e__ != cudaSuccess) {                                                              \
			rb_set_error("CUDA error %s (%d) at %s:%d: %s", cudaGetErrorName(e__), (int) e__,  \
			             __FILE__, __LINE__, cudaGetErrorString(e__));                         \
			return RB_ERR_CUDA;                                                                \
		}                                                                                      \
	} while (0)

#define RB_CHECK(st)                                                                           \
	do { int s__ = (st); if (s__ != RB_OK) return s__; } while (0)

#define RB_ARG(cond, ...)                                                                      \
	do { if (!(cond)) { rb_set_error(__VA_ARGS__); return RB_ERR_ARG; } } while (0)

static const int RB_MAX_CLASSES = 64;
static const int RB_NUM_SLOTS = 2;
#define RB_LOWEST (-FLT_MAX)

// ---------------------------------------------------------------------------------------------
// device-side descriptors
// ---------------------------------------------------------------------------------------------

// Slot of block (bx, by, bz) in a block-rank table (RbProjector::blk, RbBackprojector::blk).  The table itself is stored in
// BRICKS of 4 x 4 x 2 blocks (32 entries = one 128-byte line per 16 x 16 x 8 voxels): the 32 lanes of a warp look up the
// blocks along an arc ~64 voxels long, which crosses 4-8 bricks but ~16 rows of a [bz][by][bx] table.
// nbx / nbxy: bricks along x, bricks per z-layer of bricks.
__host__ __device__ inline uint32_t rb_blk_slot(int nbx, int nbxy, int bx, int by, int bz)
{
	return ((uint32_t) ((bz >> 1) * nbxy + (by >> 2) * nbx + (bx >> 2)) << 5) | (uint32_t) (((bz & 1) << 4) | ((by & 3) << 2) | (bx & 3));
}

// AccProjectorKernel state (acc_projectorkernel_impl.h:19-69); volume is interleaved (re,im)
struct RbProjector {
	const float2 *mdl;
	// neighbourhood-expanded copy: per voxel its 2x2x2 trilinear cell (8 x (re,im) = 64 B, 64-B aligned), so a
	// sample at an arbitrary position costs exactly two 32-B sectors instead of the 4-8 a compact layout touches
	const float4 *mdl8;
	// x-pair copy: per voxel (v[x], v[x+1]) as one aligned 16-byte word (2x the memory): the coarse pass, whose working
	// set has to stay in L2, gathers a trilinear sample with four 16-byte loads instead of eight 8-byte ones
	const float4 *mdl2;
	int mdlX, mdlY, mdlZ;
	int mdlXY;
	int mdlInitY, mdlInitZ;
	int mdlMaxR;
	float padding_factor;
	// Geometry of mdl2.  The coarse pass only samples inside the coarse window's sphere; while a pool runs, mdl2 is a
	// CONTIGUOUS copy of that core (77 MB at 256 px / coarse window 106) instead of a sub-box strewn over every z-plane of the
	// full volume: it stays in L2 and inside the TLB's reach (128 entries x 2 MB; a core embedded in the 515 x 515 x 258
	// volume touches > 200 pages and every L1 miss walks the page table).  Stage entry points use the full copy.
	int c2X, c2XY, c2InitY, c2InitZ;
	// xy-quad copy of the same core (nullptr when not built): per voxel (v[y][x], v[y][x+1], v[y+1][x], v[y+1][x+1]) as one aligned
	// 32-byte word, c2 geometry — a trilinear sample is TWO 32-byte loads (LDG.E.256), each exactly one sector
	const float4 *quad;
	// Addressing of mdl8.  The cells live in 4 x 4 x 4 blocks (4 KB) ordered by the distance of the block centre from the
	// origin, blk[] = rank of block (bz, by, bx): a thin spherical shell - what the band-major kernels sweep - is then one
	// contiguous range of memory (<= 125 pages at the edge of a 256-px reference) instead of a slice through all 2190 pages
	// of the volume.  Measured (tools/shell_prefetch_bench.cu): 64-byte gathers in a shell 35 G/s in [z][y][x] order
	// (TLB-miss bound, L2-resident or not), 160 - 200 G/s in this order.
	const uint32_t *blk;
	int nbx, nbxy;
};

// AccBackprojector state (acc_backprojector.h:24-60); accumulator is float4 (re, im, weight, 0)
struct RbBackprojector {
	float4 *vol;
	int mdlX, mdlY, mdlZ;
	int mdlInitY, mdlInitZ;
	int maxR;
	float padding_factor;
	// Scatter target of the band-major kernels (nullptr: scatter into vol).  The voxels of every 4 x 4 x 4 block are kept
	// together with their +1 halo (5 x 5 x 5 -> 128 float4 = 2 KB, so the eight corners of any cell that starts in the block
	// are in it: y stride 5, z stride 25) and the blocks are ordered by radius like the expanded reference's (RbProjector::blk):
	// a band sweep reduces into one contiguous range instead of a slice through all ~520 pages of vol.  Measured
	// (tools/shell_prefetch_bench.cu): 8 x red.v4 per sample in a shell 10.7 G samples/s into [z][y][x], 28.8 G into this.
	// rb_bp_fold adds the blocks into vol (and clears them) before anything reads vol.
	float4 *blkvol;
	const uint32_t *blk;
	int nbx, nbxy;
};
// voxel index inside blkvol of the origin of cell (x0, yi, zi); its corners are at +1, +5, +25 (and sums)
__device__ __forceinline__ size_t rb_bp_blk_cell(const RbBackprojector &b, int x0, int yi, int zi)
{
	const uint32_t rank = __ldg(b.blk + rb_blk_slot(b.nbx, b.nbxy, x0 >> 2, yi >> 2, zi >> 2));
	return ((size_t) rank << 7) + (size_t) ((zi & 3) * 25 + (yi & 3) * 5 + (x0 & 3));
}

// Pixel list entry for one window size: packed (x:10 | (y+512):11 | (ires+1):11).
// The list holds only pixels with Mresol >= 0 (ires >= 0), i.e. inside the Nyquist circle and not
// on the redundant x=0,y<0 half column (src/ml_optimiser.cpp:5784-5811).  Every other pixel has
// Minvsigma2 == 0 (src/ml_optimiser.cpp:6868-6879) and therefore contributes exactly 0 to diff2,
// wavg shell sums and back-projection weights.
__host__ __device__ inline uint32_t rb_pack_pix(int x, int y, int ires) { return (uint32_t) x | ((uint32_t) (y + 512) << 10) | ((uint32_t) (ires + 1) << 21); }
__host__ __device__ inline int rb_pix_x(uint32_t v) { return (int) (v & 1023u); }
__host__ __device__ inline int rb_pix_y(uint32_t v) { return (int) ((v >> 10) & 2047u) - 512; }
__host__ __device__ inline int rb_pix_ires(uint32_t v) { return (int) (v >> 21) - 1; }

// One image row of a window: pixels x_lo..x_hi (inclusive) of FFTW row iy, whose signed frequency is y.
// The translation phase e^{i(x tx + y ty)} factorises into a per-column and a per-row factor, so kernels that walk
// a row keep sum_x Z(x,y) e^{i x tx} in registers and apply e^{i y ty} once per row (see kernels_diff2.cu).
struct RbRow { short iy, y, x_lo, x_hi; };

// per-particle metadata resident on the device for one pool slot
struct RbPartMeta {
	int nd, np;              // sp.nr_dir, sp.nr_psi of this particle
	int dir_off, psi_off;    // offsets into the pool's dir/psi lists (-1: identity lists)
	int group, og;
	float scale;             // scale_correction[group] (1 if !do_scale_correction)
	float part_scale;        // clamped scale used by wavg/BP (acc_ml_optimiser_impl.h:3059-3079)
	float xi2_half;          // (XFLOAT)(highres_Xi2 / 2)
	int bp_off;              // accumulator = class + bp_off (pseudo half-sets of gradient refinement)
	double oldx, oldy, prx, pry;
	long long coarse_off;    // offset of this particle's dense Mweight block
	long long prior_off;     // offset of its pdf_orientation block (K*nd*np)
};

// per-particle running results on the device
struct RbPartState {
	int min_diff2_bits;      // float bits, atomicMin (diff2 >= 0)
	float min_diff2;         // coarse
	float cmax_weight; long long cmax_index;
	float csum_weight, csig_weight;
	int nr_sig_coarse, n_nonzero;
	int n_so, n_pairs;       // significant coarse orientations / significant coarse (o,t) pairs
	long long so_base, pair_base, fo_base, fs_base;  // bases into pool-level lists
	int fmin_bits;           // fine min diff2 bits
	float fmin_diff2;
	float fmax_weight; long long fmax_sample;
	float fsum_weight, fsig_weight;
	double min_diff2_final;
	long long best_ihid;     // ihidden_over of the maximum-weight fine sample
	int status;
	int n_bp;                // fine orientations the store stage processed (>= 1 significant sample)
	// store-stage accumulators
	double wsum_norm, wsum_XA, wsum_AA, sumw, wsum_s2off;
};

// one fine (oversampled) orientation of one particle
struct RbFineOrient {
	int particle;
	int iclass;
	int iorient;             // dense orientation index within the class (idl*np + ipl)
	int iover_rot;
	int pair_off;            // offset (pool-level) of the list of significant coarse translations
	int n_t;                 // number of significant coarse translations
	long long sample_off;    // pool-level offset of its n_t*NOT fine samples
	float e[9];
};

// MBL / MBR of a pool: orientation matrices are inverse(L * A * R) (helper.cuh:713-840)
struct RbLR {
	int doL = 0, doR = 0;
	double L[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
};

struct RbSamplingDev {
	int n_dir, n_psi, n_over_rot, n_trans, n_over_trans;
	const float *coarse_eulers;       // [n_dir*n_psi][9]
	const double *over_rot, *over_tilt, *over_psi; // [n_dir*n_psi*n_over_rot] (or coarse angles when n_over_rot==1)
	const double *rot, *tilt, *psi;
	const float *ctx, *cty;           // coarse trans, radians/pixel
	const float *ftx, *fty;           // fine trans
	const double *trans_x, *trans_y;  // coarse, pixels
	const double *over_trans_x, *over_trans_y;
};

struct RbModelDev {
	int nr_classes, ori_size, coarse_size, current_size, nshell;
	int Npc, Npf;            // full half-image sizes n*(n/2+1)
	int nvc, nvf;            // valid-pixel list lengths
	const uint32_t *pix_c, *pix_f;
	const uint32_t *pix_store; int nv_store;   // pixel list of the store stage: pix_f, or with --no_map the full x = 0 column too
	int dead_maxR;                             // rb_model.ref_max_r in effect (0: none): rows |y| > dead_maxR keep x == dead_maxR only in
	                                           // the diff2 / wavg sums; the store list still holds them for the back-projection
	// band-major kernels (kernels_band.cu): the store-stage pixel set sorted by |r| (ties by angle).  Entries [0, nv_rs_d2) are
	// the diff2 set (Mresol >= 0), [nv_rs_d2, nv_rs_st) the extra x = 0 half column of --no_map; nv_rs_pad = row stride of the
	// band-ordered arrays (slices, particle images)
	const uint32_t *pix_rs; int nv_rs_d2, nv_rs_st, nv_rs_pad;
	// the same pixel sets as row runs (one entry per image row holding valid pixels) and dense shell maps
	int nrows_c, nrows_f;
	const RbRow *rows_c, *rows_f;
	const short *ires_c, *ires_f;   // [n][n/2+1], -1 = excluded
	// pixel sets the SIMT diff2 kernels walk: the sets above, or (do_cc) every pixel inside the window's circle
	int d2_nvc, d2_nrows_c, d2_nrows_f;
	const uint32_t *d2_pix_c;
	const RbRow *d2_rows_c, *d2_rows_f;
	const short *d2_ires_c, *d2_ires_f;
	const float *minvs2;     // [nr_optics_groups][nshell] 1/(fudge*sigma2), entry 0 kept (DC restored for store)
	const double *pdf_direction; // [K][n_dir]
	const double *pdf_class;
	const double *prior_offset_class; // [K][2] pixels (2D references) or nullptr: the particle's own prior
	const unsigned char *dvp_gt3; // [K][nshell] data_vs_prior_class > 3
	double pixel_size, s2off, adaptive_fraction;
	int maximum_significants;
	int do_ctf_correction, refs_are_ctf_corrected, do_scale_correction, do_map, ctf_premultiplied, bp_circle_bound;
	int do_skip_rotate;      // orientation prior = pdf_class (do_skip_align || do_skip_rotate, acc_ml_optimiser_impl.h:1966)
	int do_cc;               // first-iteration cross-correlation criterion (acc_ml_optimiser_impl.h:1164)
	int do_grad;             // SGD / VDAM: back-project the weighted residual (BP.cuh:406-656), every pixel (no circle bound)
	// blocks of pdf_offset per particle: [Kp][n_trans], Kp = K with per-class prior centres, else 1.  Block 0 serves the
	// coarse pass for every class, as in the reference
	__host__ __device__ int prior_classes() const { return prior_offset_class ? nr_classes : 1; }
};

// ---------------------------------------------------------------------------------------------
// host-side context
// ---------------------------------------------------------------------------------------------
struct DevBuf {
	void *p = nullptr; size_t bytes = 0;
	int ensure(size_t n);   // grows (never shrinks); returns status
	void release();
	template <typename T> T *as() const { return (T *) p; }
};

// cudaFuncSetAttribute applies to the current device only: launchers remember what they configured per device
// (several MlDeviceBundles, one per GPU, may live in one process)
static const int RB_MAX_DEVICES = 64;

struct PoolSlot {
	int P = 0;
	bool has_priors = false;
	int max_no = 0;                 // max over particles of nd*np
	long long total_coarse = 0;     // sum over particles of K*nd*np*T
	int max_bp_off = 0;             // largest RbPartMeta::bp_off of the pool
	RbLR lr;                        // rb_particles.mat_left / mat_right
	// local searches: the particles sorted by the first direction of their prior list.  The fused coarse kernel hands CTA column i
	// particle order[i], so that CTAs in flight together sample neighbouring central planes and share their part of the coarse core
	// in L2 (a pool in acquisition order scatters the planes over the whole core)
	DevBuf order; bool has_order = false;
	DevBuf pre_shift; bool has_pre_shift = false;   // rb_particles.pre_shift: [P][2] doubles, applied to Fimg / Fnomask once
	long long total_prior = 0;
	DevBuf Fimg, Fnomask, Fctf, meta, state, dir_idx, dir_prior, psi_idx, psi_prior;
	DevBuf Mweight, pdf_orient, pdf_orient_zero, pdf_offset, pdf_offset_zero;
	DevBuf so_list, pair_list, fo, fs_w, fs_ihid, counters, shells, out_pdf_dir, out_pdf_class;
	DevBuf fimg4, cimg4;            // prepared (corrected) images of the pool at the fine / coarse window
	DevBuf cc_corr;                 // do_cc: [2][P] 1 / sqrtXi2^2 of the coarse / fine window (buildCorrImage)
	DevBuf slices;                  // reference slices written by the fine pass, re-read by the store stage
	long long slice_capacity = 0;   // number of fine orientations whose slice fits the cache
	size_t cap_fo = 0, cap_fs = 0;  // capacity of this slot's fine-pass lists (fo / fs_w, fs_ihid)
	// band-major path (kernels_band.cu)
	DevBuf simg4, sst, sctf;        // band-ordered particle images: prepared image, (X, X0), ctf * scale
	DevBuf bp_cnt, bp_item_of, bp_items, bp_samp;   // fine orientations holding significant samples + their (phase, weight) tables
	int band_rounds = 0;
	std::vector<RbPartMeta> h_meta;
	std::vector<int> h_order;
	cudaEvent_t uploaded = nullptr;
	cudaEvent_t done = nullptr;     // recorded when the E-step of this slot has been enqueued completely; rb_estep_fetch waits on it
};

struct rb_ctx {
	int device = 0;
	int num_sms = 0;
	cudaStream_t stream = nullptr, copy_stream = nullptr, fetch_stream = nullptr;
	cudaStream_t aux_stream = nullptr;               // side stream of the band-ordered image copies (rbk_band_images_async)
	cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
	long long launches = 0;

	RbProjector proj[RB_MAX_CLASSES];
	DevBuf proj_buf[RB_MAX_CLASSES];
	DevBuf proj8_buf[RB_MAX_CLASSES];
	DevBuf proj2_buf[RB_MAX_CLASSES];
	DevBuf proj4c_buf[RB_MAX_CLASSES];               // xy-quad copy of the coarse-window core (RbProjector::quad)
	DevBuf proj2c_buf[RB_MAX_CLASSES];               // x-pair copy of the coarse-window core (RbProjector::c2*)
	long long core_stamp[RB_MAX_CLASSES];            // ref_version the core was built from (-1: none), and its half width
	int core_R[RB_MAX_CLASSES];
	// radius-sorted block tables of the expanded references (RbProjector::blk), one per volume geometry
	struct BlockTable { int geom[5]; DevBuf buf; };
	std::vector<BlockTable *> blk_tables;
	RbBackprojector bp[RB_MAX_CLASSES];
	DevBuf bp_buf[RB_MAX_CLASSES];
	DevBuf bp_blk_buf[RB_MAX_CLASSES];               // padded block accumulators (RbBackprojector::blkvol)
	bool bp_blk_dirty[RB_MAX_CLASSES] = {false};     // blkvol holds contributions that are not in vol yet (rb_bp_fold)
	bool has_proj[RB_MAX_CLASSES] = {false}, has_bp[RB_MAX_CLASSES] = {false};
	// 2D references / accumulators (2D classification) live in a two-plane volume [2][Y][X] whose second plane is zero:
	// with in-plane rotations zp == 0, so the trilinear code paths reduce exactly to project2Dmodel / backproject2D
	bool ref_2d[RB_MAX_CLASSES] = {false}, bp_2d[RB_MAX_CLASSES] = {false};

	bool has_sampling = false, has_model = false;
	rb_sampling h_samp{};            // scalar fields only (pointers are not kept)
	rb_model h_model{};
	RbSamplingDev d_samp{};
	RbModelDev d_model{};
	DevBuf s_coarse_eulers, s_over_rot, s_over_tilt, s_over_psi, s_rot, s_tilt, s_psi, s_ctx, s_cty, s_ftx, s_fty,
	       s_tx, s_ty, s_otx, s_oty;
	DevBuf m_pix_c, m_pix_f, m_minvs2, m_pdf_dir, m_pdf_class, m_dvp;
	DevBuf m_rows_c, m_rows_f, m_ires_c, m_ires_f;
	DevBuf m_cc[6];                  // do_cc: coarse pixel list, coarse rows / shell map, fine rows / shell map with the full x = 0 column;
	                                 // [5]: !do_map: fine pixel list of the store stage with the full x = 0 column
	DevBuf d_proj, d_bp;             // device copies of the projector / backprojector tables
	std::vector<double> h_scale_correction;

	PoolSlot slot[RB_NUM_SLOTS];
	DevBuf m_pix_rs;
	DevBuf comm_buf, comm_buf2;      // staging of the NCCL reductions (comm.cu): planar accumulator, fp64 weighted sums
	DevBuf band_slices;              // band-ordered slices of one round of fine orientations (shared by the slots: E-steps run one after the other)
	long long band_slice_capacity = 0;
	// phase tables of the band-major diff2 pass (kernels_band.cu): [n_trans][stride] and [n_over_trans][stride] float2
	DevBuf band_tabc, band_tabo, band_tabu;
	bool band_separable = false;     // over_trans[t * NOT + j] == trans[t] + d[j] (checked by rb_set_sampling)
	std::vector<double> h_band_u;    // turns per pixel: coarse x [T], coarse y [T], offsets x [NOT], offsets y [NOT]
	long long band_tab_model = -1, band_tab_samp = -1;

	// stage timing
	std::map<std::string, std::pair<cudaEvent_t, cudaEvent_t>> stage_ev;
	DevBuf scratch[8];
	DevBuf wc_buf[10];
	DevBuf grid_buf[5];              // iterative gridding (kernels_recon.cu): Fweight, Fnewweight, Fconv, real-space volume (double), blob table
	DevBuf recon_buf[3];             // reconstruction on the device (kernels_recon.cu): FFT input / output, radial sums
	DevBuf prep_buf[5];              // device image preparation (kernels_prep.cu): cuFFT input / output, background values, spectra
	int prep_plan_inv = 0, prep_plan_inv_n = 0, prep_plan_inv_batch = 0;   // batched 2D C2R plan of the noise-filled mask
	std::vector<double> h_sigma2_noise;   // [nr_optics_groups][nshell] (the noise image of the noise-filled mask follows this spectrum)
	int prep_plan = 0, prep_plan_n = 0, prep_plan_batch = 0;   // this context's batched 2D R2C cuFFT plan (cufftHandle is an int); 0 batch: none
	DevBuf prep_raw[RB_NUM_SLOTS][7];   // per slot: raw images, shifts, norm factors, CTF parameters (filled on the copy stream)
	DevBuf posed_buf[2][3];          // staged posed images (F2D, Fctf, matrices), two buffers for upload / compute overlap
	DevBuf posed_pix, posed_sorted;  // band-major posed back-projection: pixel list of the image size, band-ordered images of a chunk
	int posed_pix_n = 0, posed_pix_count = 0;
	int posed_n = 0, posed_count = 0;   // what rb_bp_posed_stage left in posed_buf[0]
	cudaEvent_t posed_ev[2] = {nullptr, nullptr};               // partials / compact list of the multi-CTA coarse weight conversion
	DevBuf gemm_buf[10];             // operands of the tensor-core coarse pass (kernels_gemm.cu)
	// orientation operands (A, |A|^2; TF32 hi / lo) depend on the reference and the sampling only: cached across pools
	DevBuf gemmA[RB_MAX_CLASSES][4];
	long long gemmA_stamp[RB_MAX_CLASSES];
	DevBuf gemmA_all[4];             // the classes' orientation operands stacked along M (few orientations per class: 2D classification)
	long long gemmA_all_stamp = -1;
	long long ref_version[RB_MAX_CLASSES] = {0}, samp_version = 0, model_version = 0;
	RbLR coarse_lr;                 // the MBL / MBR the coarse matrices were last built with
	int fine_dead_maxR = 0;         // rb_model.ref_max_r when the fine pixel sets were built without the rows beyond it
};

int rb_stage_begin(rb_ctx *ctx, const char *name);
int rb_stage_end(rb_ctx *ctx, const char *name);
int rb_sync_tables(rb_ctx *ctx);   // refresh d_proj / d_bp device tables
int rb_bp_fold(rb_ctx *ctx, int k);  // blkvol -> vol (no-op when clean); every reader of vol calls it first
int rbk_bp_fold(rb_ctx *ctx, const RbBackprojector &bp);

// ---------------------------------------------------------------------------------------------
// kernel launchers (one per .cu)
// ---------------------------------------------------------------------------------------------
// kernels_misc.cu
int rbk_make_coarse_eulers(rb_ctx *ctx, const double *d_rot, const double *d_tilt, const double *d_psi, int n_dir, int n_psi, const RbLR &lr,
                           float *d_eulers);
int rbk_convert_volume(rb_ctx *ctx, const double *d_in, float2 *d_out, size_t n);
int rbk_expand_volume(rb_ctx *ctx, const RbProjector &pj, float4 *d_out, float4 *d_out2);
int rbk_pre_shift(rb_ctx *ctx, PoolSlot &s, cudaStream_t stream);
int rbk_xyquad_core(rb_ctx *ctx, const RbProjector &pj, int cX, int cY, int cInitY, int cInitZ, float4 *d_out);
int rbk_xpair_core(rb_ctx *ctx, const RbProjector &pj, int cX, int cY, int cInitY, int cInitZ, float4 *d_out);
int rbk_bp_deinterleave(rb_ctx *ctx, const float4 *vol, float *re, float *im, float *w, size_t n);
int rbk_backproject_posed(rb_ctx *ctx, const RbBackprojector &bp, int n, int count, const float2 *d_F, const float *d_W, const float *d_eulers);
struct RbPosedBandLayout { const uint32_t *pix; int npix, stride; float4 *sF; int *queue; };
int rbk_posed_band_layout(rb_ctx *ctx, int n, int count, RbPosedBandLayout *L);
int rbk_posed_band_scatter(rb_ctx *ctx, const RbBackprojector &bp, int n, int count, const float *d_eulers, const RbPosedBandLayout &L);
// raw images -> FFT, centring sign, shift phase, CTF, DC = 0 -> band-ordered staging buffer -> scatter (kernels_prep.cu)
int rbk_backproject_posed_raw(rb_ctx *ctx, const RbBackprojector &bp, int n, int count, float *d_images, const double *d_shift,
                              const double *d_ctfpar, double xs_angstrom, int ctf_premultiplied, const float *d_eulers);
int rbk_backproject_posed_band(rb_ctx *ctx, const RbBackprojector &bp, int n, int count, const float2 *d_F, const float *d_W, const float *d_eulers);
int rbk_project(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers, int count, float2 *d_out);

// kernels_diff2.cu
int rbk_prep_priors(rb_ctx *ctx, PoolSlot &s);
int rbk_diff2_coarse_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_diff2_fine_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_diff2_coarse_stage(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers, int O,
                           const float *d_tx, const float *d_ty, int T, const float *d_re, const float *d_im,
                           const float *d_corr, float *d_out, int cc = 0);
int rbk_cc_corr_pool(rb_ctx *ctx, PoolSlot &s);   // do_cc: 1 / sqrtXi2^2 of the coarse and the fine window, per particle
int rbk_diff2_fine_stage(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers,
                         const float *d_tx, const float *d_ty, const float *d_re, const float *d_im,
                         const float *d_corr, float sum_init,
                         const unsigned long long *d_rot_idx, const unsigned long long *d_trans_idx,
                         const unsigned long long *d_job_idx, const unsigned long long *d_job_num, int n_jobs,
                         float *d_out, int cc = 0);

// kernels_gemm.cu: global-search coarse pass as a 3xTF32 tcgen05 contraction
bool rbk_coarse_gemm_applicable(rb_ctx *ctx, const PoolSlot &s);
int rbk_diff2_coarse_gemm_pool(rb_ctx *ctx, PoolSlot &s, const float4 *cimg4);
bool rbk_coarse_fused_applicable(rb_ctx *ctx, const PoolSlot &s);
int rbk_diff2_coarse_fused_pool(rb_ctx *ctx, PoolSlot &s, const float4 *cimg4);
int rbk_gemm_tf32x3_stage(rb_ctx *ctx, const float *dA, const float *dB, int M, int N, int K, float *dC);

// kernels_prep.cu: getFourierTransformsAndCtfs on the device, batched over the pool
void rbk_prepare_release(rb_ctx *ctx);
int rbk_prepare_pool(rb_ctx *ctx, PoolSlot &s, const float *d_raw, const int *d_shift, const float *d_norm, const double *d_ctfpar,
                     int n, float radius, float cosine_width, float *d_power, const long long *d_seed = nullptr, const float *d_spectrum = nullptr, const float2 *d_og_factor = nullptr,
                     cudaStream_t stream = nullptr);

// kernels_recon.cu: BackProjector::reconstruct (skip_gridding) + windowToOridimRealSpace + griddingCorrect on the device
int rbk_reconstruct(rb_ctx *ctx, const RbBackprojector &bp, int ori, const double *d_tau2, int n_tau2, double tau2_fudge, int minres_map,
                    float *d_vol_out, int max_iter_preweight = 0, double normalise = 1.);

int rbk_bp_symmetrise(rb_ctx *ctx, const RbBackprojector &bp, DevBuf &tmp, const float *d_R, int nsym, const float *d_hR, const float *d_hz,
                      int nhel);
int rbk_update_ssnr(rb_ctx *ctx, const RbBackprojector &bp, bool is_2d, int ori, double tau2_fudge, double *tau2_io, double *sigma2_out,
                    double *dvp_out, double *cov_out, const double *fsc, const double *avgctf2, bool update_with_fsc, bool whole);
int rbk_ftmap(rb_ctx *ctx, const float *d_vol, int ori, int r_max, float pf, float2 *d_data, int pad, double *h_power);

// kernels_weights.cu
int rbk_weights_coarse_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_fine_setup_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_weights_fine_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_convert_weights_stage(rb_ctx *ctx, float *d_w, long long n_orient, int n_trans,
                              const float *d_pdf_o, const unsigned char *d_pdf_oz,
                              const float *d_pdf_t, const unsigned char *d_pdf_tz,
                              double adaptive_fraction, int maxsig, int filter_zero,
                              unsigned char *d_sig, rb_weights_out *d_out);

// kernels_band.cu: band-major (L2-resident) fine pass and store stage
bool rbk_band_applicable(rb_ctx *ctx);
int rbk_band_fine_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_band_images_async(rb_ctx *ctx, PoolSlot &s);   // band-ordered particle images on a side stream, joined by rbk_band_fine_pool
int rbk_band_store_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_band_phase_tables(rb_ctx *ctx, bool &ok);

// kernels_store.cu
int rbk_collect_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_store_pool(rb_ctx *ctx, PoolSlot &s);
int rbk_wavg_stage(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers, int O,
                   const float *d_tx, const float *d_ty, int T, const float *d_re, const float *d_im,
                   const float *d_w, const float *d_ctf, float weight_norm, float sig_w,
                   float *d_parts, float *d_AA, float *d_XA);
int rbk_backproject_stage(rb_ctx *ctx, const RbBackprojector &bp, int n, const float *d_eulers, int O,
                          const float *d_tx, const float *d_ty, int T, const float *d_re, const float *d_im,
                          const float *d_w, const float *d_minvs2, const float *d_ctf,
                          float weight_norm, float sig_w, int circle_bound, int ctf_premultiplied);

#define RB_LAUNCH_CHECK(ctx)                                                                   \
	do { (ctx)->launches++; RB_CUDA(cudaGetLastError()); } while (0)
