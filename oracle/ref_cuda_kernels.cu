/*
 * TEST / BENCH INFRASTRUCTURE ONLY - builds oracle/_ref/librefcuda.so (git-ignored, travels to the GPU box).
 *
 * The reference's OWN CUDA kernels (its --gpu path), timed on the same B200 beside ours.  This translation unit contains
 * no kernel of its own: it #includes the reference's CUDA kernel headers from where they lie under /root/reference
 * (never copied into this repo)
 *     src/acc/cuda/cuda_kernels/diff2.cuh   cuda_kernel_diff2_coarse / cuda_kernel_diff2_fine
 *     src/acc/cuda/cuda_kernels/wavg.cuh    cuda_kernel_wavg
 *     src/acc/cuda/cuda_kernels/BP.cuh      cuda_kernel_backproject3D
 *     src/acc/acc_projectorkernel_impl.h    AccProjectorKernel::project3Dmodel (texture path, the reference's default)
 * compiled for sm_100, and launches them with the reference's grid shapes
 *     runDiff2KernelCoarse   src/acc/acc_helper_functions_impl.h:1139-1400, AccUtilities::diff2_coarse src/acc/utilities.h:1101-1114
 *     runDiff2KernelFine     :1813-1920, AccUtilities::diff2_fine utilities.h:1287-1308
 *     runWavgKernel          :316-503,  AccUtilities::kernel_wavg utilities.h:968-991
 *     runBackProjectKernel   :505-1092 (cuda_kernel_backproject3D<false,false><<<imageCount, BP_REF3D_BLOCK_SIZE>>>, :1008-1016)
 * behind the kernel table of oracle/oracle_kernels.h, so that the restated per-particle driver (oracle/estep_driver.cpp)
 * runs the reference's E-step on them exactly as it does on the ALTCPU kernels.  Every kernel is bracketed by CUDA events;
 * refcuda_timers() returns the summed device time per kernel family.  Host<->device copies around each launch are NOT
 * in those sums (the reference keeps its buffers on the device too).
 *
 * PROJECTOR_NO_TEXTURES does not build for CUDA in the reference (makeKernel passes one pointer to a two-pointer
 * constructor, acc_projectorkernel_impl.h:301-319), so this is the texture path: two float 3D cudaArrays with linear
 * filtering, configured like AccProjector::initMdl (src/acc/acc_projector_impl.h:46-105).  Texture interpolation uses
 * 8-bit fractions: this provider is a TIMING baseline, parity stays pinned on the ALTCPU kernels.
 *
 * Recipe: oracle/Makefile target `refcuda`:
 *   nvcc -gencode arch=compute_100,code=sm_100 -O3 -std=c++17 -D_CUDA_ENABLED -DACC_CUDA=2 -DACC_CPU=1 -DCUDA_NO_CUSTOM_ALLOCATION
 *        -I/root/reference -Ioracle/shim -Ioracle --shared -Xcompiler -fPIC
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "src/acc/settings.h"
#include "src/acc/acc_ptr.h"
#include "src/acc/acc_projector.h"
#include "src/acc/acc_backprojector.h"
#include "src/acc/acc_projectorkernel_impl.h"
#include "src/acc/cuda/cuda_settings.h"
#include "src/acc/cuda/cuda_kernels/cuda_device_utils.cuh"
#include "src/acc/cuda/cuda_kernels/helper.cuh"
#include "src/acc/cuda/cuda_kernels/diff2.cuh"
#include "src/acc/cuda/cuda_kernels/wavg.cuh"
#include "src/acc/cuda/cuda_kernels/BP.cuh"

#include "oracle_kernels.h"

// CPU pieces of the table that have no CUDA twin on this path (priors, weights, collect): the restated port
extern "C" const ok_kernel_table *portk_kernel_table(void);

#define RC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "refcuda: %s at %s:%d: %s\n", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); abort(); } } while (0)

namespace {

struct DevVol { cudaArray_t aR = nullptr, aI = nullptr; cudaTextureObject_t tR = 0, tI = 0; };
struct DevBP { float *re = nullptr, *im = nullptr, *w = nullptr; size_t n = 0; };
struct Buf {
	void *p = nullptr; size_t cap = 0;
	template <typename T> T *up(const T *src, size_t n)
	{
		const size_t bytes = (n ? n : 1) * sizeof(T);
		if (bytes > cap) { if (p) cudaFree(p); RC(cudaMalloc(&p, bytes)); cap = bytes; }
		if (src && n) RC(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
		else RC(cudaMemset(p, 0, bytes));
		return (T *) p;
	}
};

std::map<const float *, DevVol> g_vols;
std::map<const float *, DevBP> g_bps;
Buf g_buf[16];
cudaEvent_t g_e0 = nullptr, g_e1 = nullptr;
double g_ms[4] = {0., 0., 0., 0.};       // coarse, fine, wavg, backproject
long g_launches[4] = {0, 0, 0, 0};

void tic() { if (!g_e0) { RC(cudaEventCreate(&g_e0)); RC(cudaEventCreate(&g_e1)); } RC(cudaEventRecord(g_e0, 0)); }
void toc(int which, int launches)
{
	RC(cudaEventRecord(g_e1, 0));
	RC(cudaEventSynchronize(g_e1));
	RC(cudaGetLastError());
	float ms = 0.f;
	RC(cudaEventElapsedTime(&ms, g_e0, g_e1));
	g_ms[which] += ms; g_launches[which] += launches;
}

// AccProjector::initMdl, texture branch for a 3D reference (acc_projector_impl.h:46-105, 196-312): split into real / imag float arrays
const DevVol &volume(const ok_projector *p)
{
	auto it = g_vols.find(p->mdl);
	if (it != g_vols.end()) return it->second;
	DevVol v;
	const size_t n = (size_t) p->mdlX * p->mdlY * p->mdlZ;
	std::vector<float> re(n), im(n);
	for (size_t i = 0; i < n; i++) { re[i] = p->mdl[2 * i]; im[i] = p->mdl[2 * i + 1]; }
	cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
	cudaExtent ext = make_cudaExtent(p->mdlX, p->mdlY, p->mdlZ);
	RC(cudaMalloc3DArray(&v.aR, &desc, ext));
	RC(cudaMalloc3DArray(&v.aI, &desc, ext));
	for (int c = 0; c < 2; c++)
	{
		cudaMemcpy3DParms cp;
		memset(&cp, 0, sizeof(cp));
		cp.srcPtr = make_cudaPitchedPtr(c ? im.data() : re.data(), p->mdlX * sizeof(float), p->mdlX, p->mdlY);
		cp.dstArray = c ? v.aI : v.aR;
		cp.extent = ext;
		cp.kind = cudaMemcpyHostToDevice;
		RC(cudaMemcpy3D(&cp));
	}
	cudaTextureDesc td;
	memset(&td, 0, sizeof(td));
	td.filterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeElementType;
	td.normalizedCoords = false;
	for (int i = 0; i < 3; i++) td.addressMode[i] = cudaAddressModeClamp;
	for (int c = 0; c < 2; c++)
	{
		cudaResourceDesc rd;
		memset(&rd, 0, sizeof(rd));
		rd.resType = cudaResourceTypeArray;
		rd.res.array.array = c ? v.aI : v.aR;
		RC(cudaCreateTextureObject(c ? &v.tI : &v.tR, &rd, &td, nullptr));
	}
	return g_vols[p->mdl] = v;
}

AccProjectorKernel make_kernel(const ok_projector *p, int imgX, int imgY)
{
	// AccProjectorKernel::makeKernel (acc_projectorkernel_impl.h:301-319) with imgMaxR = imgX - 1 (acc_ml_optimiser_impl.h:1322-1327)
	const DevVol &v = volume(p);
	const int imgMaxR = imgX - 1;
	const int maxR = p->mdlMaxR >= imgMaxR ? imgMaxR : p->mdlMaxR;
	return AccProjectorKernel(p->mdlX, p->mdlY, p->mdlZ, imgX, imgY, 1, p->mdlInitY, p->mdlInitZ, p->padding_factor, maxR, v.tR, v.tI);
}

DevBP &accumulator(const ok_backprojector *bp)
{
	auto it = g_bps.find(bp->real);
	if (it != g_bps.end()) return it->second;
	DevBP d;
	d.n = (size_t) bp->mdlX * bp->mdlY * bp->mdlZ;
	RC(cudaMalloc(&d.re, d.n * 4)); RC(cudaMalloc(&d.im, d.n * 4)); RC(cudaMalloc(&d.w, d.n * 4));
	RC(cudaMemcpy(d.re, bp->real, d.n * 4, cudaMemcpyHostToDevice));
	RC(cudaMemcpy(d.im, bp->imag, d.n * 4, cudaMemcpyHostToDevice));
	RC(cudaMemcpy(d.w, bp->weight, d.n * 4, cudaMemcpyHostToDevice));
	return g_bps[bp->real] = d;
}

template <int BS>
void coarse_launch(const AccProjectorKernel &k, float *d_e, float *d_tx, float *d_ty, float *d_tz, float *d_re, float *d_im, float *d_corr,
                   float *d_out, unsigned long O, unsigned long T, unsigned long image_size, int &launches)
{
	// runDiff2KernelCoarse, CUDA branch for a 3D reference and 2D data: full blocks of D2C_EULERS_PER_BLOCK_REF3D orientations,
	// the O % D2C_BLOCK_SIZE_REF3D rest one orientation per block
	const unsigned long rest = O % D2C_BLOCK_SIZE_REF3D, even = O - rest;
	if (even)
	{
		cuda_kernel_diff2_coarse<true, false, BS, D2C_EULERS_PER_BLOCK_REF3D, 4><<<even / D2C_EULERS_PER_BLOCK_REF3D, BS>>>(
			d_e, d_tx, d_ty, d_tz, d_re, d_im, k, d_corr, d_out, (int) T, (int) image_size);
		launches++;
	}
	if (rest)
	{
		cuda_kernel_diff2_coarse<true, false, BS, 1, 4><<<rest, BS>>>(
			d_e + 9 * even, d_tx, d_ty, d_tz, d_re, d_im, k, d_corr, d_out + T * even, (int) T, (int) image_size);
		launches++;
	}
}

void rc_diff2_coarse(const ok_projector *p, int imgX, int imgY, const float *eulers, unsigned long O,
                     const float *trans_x, const float *trans_y, unsigned long T,
                     const float *img_re, const float *img_im, const float *corr, float *diff2s)
{
	const AccProjectorKernel k = make_kernel(p, imgX, imgY);
	const unsigned long image_size = (unsigned long) imgX * imgY;
	float *d_e = g_buf[0].up(eulers, O * 9), *d_tx = g_buf[1].up(trans_x, T), *d_ty = g_buf[2].up(trans_y, T), *d_tz = g_buf[3].up((const float *) nullptr, T);
	float *d_re = g_buf[4].up(img_re, image_size), *d_im = g_buf[5].up(img_im, image_size), *d_c = g_buf[6].up(corr, image_size);
	float *d_o = g_buf[7].up(diff2s, O * T);
	int launches = 0;
	tic();
	if (T <= D2C_BLOCK_SIZE_REF3D) coarse_launch<D2C_BLOCK_SIZE_REF3D>(k, d_e, d_tx, d_ty, d_tz, d_re, d_im, d_c, d_o, O, T, image_size, launches);
	else if (T <= D2C_BLOCK_SIZE_REF3D * 2) coarse_launch<D2C_BLOCK_SIZE_REF3D * 2>(k, d_e, d_tx, d_ty, d_tz, d_re, d_im, d_c, d_o, O, T, image_size, launches);
	else if (T <= D2C_BLOCK_SIZE_REF3D * 4) coarse_launch<D2C_BLOCK_SIZE_REF3D * 4>(k, d_e, d_tx, d_ty, d_tz, d_re, d_im, d_c, d_o, O, T, image_size, launches);
	else { fprintf(stderr, "refcuda: %lu translations exceed the reference's limit (ERR_TRANSLIM)\n", T); abort(); }
	toc(0, launches);
	RC(cudaMemcpy(diff2s, d_o, O * T * 4, cudaMemcpyDeviceToHost));
}

void rc_diff2_fine(const ok_projector *p, int imgX, int imgY, const float *eulers,
                   const float *trans_x, const float *trans_y,
                   const float *img_re, const float *img_im, const float *corr, float sum_init,
                   unsigned long orientation_num, unsigned long translation_num, unsigned long num_jobs,
                   const unsigned long *rot_idx, const unsigned long *trans_idx,
                   const unsigned long *job_idx, const unsigned long *job_num, float *diff2s)
{
	if (num_jobs == 0) return;
	const AccProjectorKernel k = make_kernel(p, imgX, imgY);
	const unsigned long image_size = (unsigned long) imgX * imgY;
	unsigned long nw = 0;
	for (unsigned long j = 0; j < num_jobs; j++) nw = nw > job_idx[j] + job_num[j] ? nw : job_idx[j] + job_num[j];
	float *d_e = g_buf[0].up(eulers, orientation_num * 9), *d_tx = g_buf[1].up(trans_x, translation_num), *d_ty = g_buf[2].up(trans_y, translation_num);
	float *d_tz = g_buf[3].up((const float *) nullptr, translation_num);
	float *d_re = g_buf[4].up(img_re, image_size), *d_im = g_buf[5].up(img_im, image_size), *d_c = g_buf[6].up(corr, image_size);
	float *d_o = g_buf[7].up(diff2s, nw);
	unsigned long *d_ri = g_buf[8].up(rot_idx, nw), *d_ti = g_buf[9].up(trans_idx, nw), *d_ji = g_buf[10].up(job_idx, num_jobs), *d_jn = g_buf[11].up(job_num, num_jobs);
	tic();
	cuda_kernel_diff2_fine<true, false, D2F_BLOCK_SIZE_REF3D, D2F_CHUNK_REF3D><<<num_jobs, D2F_BLOCK_SIZE_REF3D>>>(
		d_e, d_re, d_im, d_tx, d_ty, d_tz, k, d_c, d_o, (unsigned) image_size, sum_init, orientation_num, translation_num, num_jobs,
		d_ri, d_ti, d_ji, d_jn);
	toc(1, 1);
	RC(cudaMemcpy(diff2s, d_o, nw * 4, cudaMemcpyDeviceToHost));
}

void rc_wavg(const ok_projector *p, int imgX, int imgY, const float *eulers, unsigned long orientation_num,
             const float *img_re, const float *img_im, const float *trans_x, const float *trans_y,
             const float *weights, const float *ctfs, float *parts, float *AA, float *XA,
             unsigned long trans_num, float weight_norm, float significant_weight, float part_scale)
{
	if (orientation_num == 0) return;
	const AccProjectorKernel k = make_kernel(p, imgX, imgY);
	const unsigned long image_size = (unsigned long) imgX * imgY;
	float *d_e = g_buf[0].up(eulers, orientation_num * 9), *d_tx = g_buf[1].up(trans_x, trans_num), *d_ty = g_buf[2].up(trans_y, trans_num);
	float *d_tz = g_buf[3].up((const float *) nullptr, trans_num);
	float *d_re = g_buf[4].up(img_re, image_size), *d_im = g_buf[5].up(img_im, image_size), *d_c = g_buf[6].up(ctfs, image_size);
	float *d_w = g_buf[7].up(weights, orientation_num * trans_num);
	float *d_p = g_buf[8].up(parts, image_size), *d_a = g_buf[9].up(AA, image_size), *d_x = g_buf[10].up(XA, image_size);
	tic();
	cuda_kernel_wavg<true, true, false, WAVG_BLOCK_SIZE><<<orientation_num, WAVG_BLOCK_SIZE, (3 * WAVG_BLOCK_SIZE + 9) * sizeof(XFLOAT)>>>(
		d_e, k, (unsigned) image_size, orientation_num, d_re, d_im, d_tx, d_ty, d_tz, d_w, d_c, d_p, d_a, d_x, trans_num,
		weight_norm, significant_weight, part_scale);
	toc(2, 1);
	RC(cudaMemcpy(parts, d_p, image_size * 4, cudaMemcpyDeviceToHost));
	RC(cudaMemcpy(AA, d_a, image_size * 4, cudaMemcpyDeviceToHost));
	RC(cudaMemcpy(XA, d_x, image_size * 4, cudaMemcpyDeviceToHost));
}

void rc_backproject(const ok_backprojector *bp, int imgX, int imgY, const float *img_re, const float *img_im,
                    const float *trans_x, const float *trans_y, const float *weights, const float *Minvsigma2s, const float *ctfs,
                    unsigned long trans_num, float significant_weight, float weight_norm, const float *eulers, unsigned long image_count)
{
	if (image_count == 0) return;
	DevBP &d = accumulator(bp);
	const unsigned long image_size = (unsigned long) imgX * imgY;
	float *d_e = g_buf[0].up(eulers, image_count * 9), *d_tx = g_buf[1].up(trans_x, trans_num), *d_ty = g_buf[2].up(trans_y, trans_num);
	float *d_tz = g_buf[3].up((const float *) nullptr, trans_num);
	float *d_re = g_buf[4].up(img_re, image_size), *d_im = g_buf[5].up(img_im, image_size), *d_c = g_buf[6].up(ctfs, image_size);
	float *d_w = g_buf[7].up(weights, image_count * trans_num), *d_m = g_buf[8].up(Minvsigma2s, image_size);
	tic();
	cuda_kernel_backproject3D<false, false><<<image_count, BP_REF3D_BLOCK_SIZE>>>(
		d_re, d_im, d_tx, d_ty, d_tz, d_w, d_m, d_c, trans_num, significant_weight, weight_norm, d_e,
		d.re, d.im, d.w, bp->maxR, bp->maxR * bp->maxR, bp->padding_factor,
		(unsigned) imgX, (unsigned) imgY, 1u, (unsigned) image_size, (unsigned) bp->mdlX, (unsigned) bp->mdlY, bp->mdlInitY, bp->mdlInitZ);
	toc(3, 1);
}

void *rc_sync_alloc(int, int) { return nullptr; }
void rc_sync_free(void *) {}

ok_kernel_table g_table;
bool g_table_ready = false;

} // namespace

extern "C" {

const ok_kernel_table *refcuda_kernel_table(void)
{
	if (!g_table_ready)
	{
		g_table = *portk_kernel_table();
		g_table.kind = "reference-cuda";
		g_table.diff2_coarse = rc_diff2_coarse;
		g_table.diff2_fine = rc_diff2_fine;
		g_table.wavg = rc_wavg;
		g_table.backproject = rc_backproject;
		g_table.bp_sync_alloc = rc_sync_alloc;
		g_table.bp_sync_free = rc_sync_free;
		g_table_ready = true;
	}
	return &g_table;
}

// summed device time (ms) and launch counts: [coarse, fine, wavg, backproject]
void refcuda_timers(double *ms, long *launches, int reset)
{
	for (int i = 0; i < 4; i++) { if (ms) ms[i] = g_ms[i]; if (launches) launches[i] = g_launches[i]; }
	if (reset) for (int i = 0; i < 4; i++) { g_ms[i] = 0.; g_launches[i] = 0; }
}

// copy a device accumulator back into the host arrays of its ok_backprojector (after the last particle)
void refcuda_bp_download(const ok_backprojector *bp)
{
	auto it = g_bps.find(bp->real);
	if (it == g_bps.end()) return;
	RC(cudaMemcpy(bp->real, it->second.re, it->second.n * 4, cudaMemcpyDeviceToHost));
	RC(cudaMemcpy(bp->imag, it->second.im, it->second.n * 4, cudaMemcpyDeviceToHost));
	RC(cudaMemcpy(bp->weight, it->second.w, it->second.n * 4, cudaMemcpyDeviceToHost));
}

// free every device object (volumes, accumulators, scratch)
void refcuda_release(void)
{
	for (auto &kv : g_vols) { cudaDestroyTextureObject(kv.second.tR); cudaDestroyTextureObject(kv.second.tI); cudaFreeArray(kv.second.aR); cudaFreeArray(kv.second.aI); }
	g_vols.clear();
	for (auto &kv : g_bps) { cudaFree(kv.second.re); cudaFree(kv.second.im); cudaFree(kv.second.w); }
	g_bps.clear();
	for (auto &b : g_buf) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }
}

} // extern "C"
