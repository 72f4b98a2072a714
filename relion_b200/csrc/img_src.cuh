// relion_b200 — image source shared by the diff2 kernels: either reference-style pre-corrected SoA arrays
// (stage API) or the pool's raw Fimg / Fctf with the corrections (pixel_correction, corr_img —
// /root/reference/src/acc/acc_ml_optimiser_impl.h:1251-1268, acc_helper_functions_impl.h:164-196)
// computed in registers.
#pragma once
#include "device_utils.cuh"

struct ImgSrc {
	const float *re, *im, *corr;   // stage mode when re != nullptr
	const float2 *F;               // pool mode
	const float *ctf;
	const float *minvs2;           // [nshell] of the particle's optics group
	float inv_scale, scale2;
	int do_ctf_refs;               // do_ctf_correction && refs_are_ctf_corrected
	int do_scale;
	int n_array;                   // window size the arrays are stored at
	float cc_corr;                 // > 0: cross-correlation criterion, corr = 1 / sqrtXi2^2 on every pixel, DC included
	                               // (buildCorrImage, acc_helper_functions_impl.h:172-174)
};

// by array index + shell index (ires only matters in pool mode)
__device__ __forceinline__ void img_load_idx(const ImgSrc &s, int idx, int ires, float2 &X, float &corr)
{
	if (s.re)
	{
		X = make_float2(__ldg(s.re + idx), __ldg(s.im + idx));
		corr = __ldg(s.corr + idx);
	}
	else
	{
		float2 F = __ldg(s.F + idx);
		float pc = s.inv_scale;
		float c = ires > 0 ? __ldg(s.minvs2 + ires) : 0.f;       // DC excluded (src/ml_optimiser.cpp:6874-6879)
		if (s.cc_corr > 0.f) c = s.cc_corr;
		if (s.do_ctf_refs)
		{
			float ctf = __ldg(s.ctf + idx);
			if (fabsf(ctf) > 1e-8f) pc = pc / ctf;               // acc_ml_optimiser_impl.h:1254-1264
			c *= ctf * ctf;                                      // buildCorrImage
		}
		if (s.do_scale) c *= s.scale2;
		X = make_float2(F.x * pc, F.y * pc);
		corr = c;
	}
}

// by packed pixel-list entry
__device__ __forceinline__ void img_load(const ImgSrc &s, uint32_t pk, float2 &X, float &corr)
{
	img_load_idx(s, rb_src_index(rb_pix_x(pk), rb_pix_y(pk), s.n_array), rb_pix_ires(pk), X, corr);
}

__device__ __forceinline__ void img_src_pool(ImgSrc &src, const RbModelDev &M, const RbPartMeta &m,
                                             const float2 *Fimg, const float *Fctf, int p)
{
	src.re = nullptr;
	src.F = Fimg + (size_t) p * M.Npf; src.ctf = Fctf ? Fctf + (size_t) p * M.Npf : nullptr;
	src.minvs2 = M.minvs2 + (size_t) m.og * M.nshell;
	src.inv_scale = 1.0f / m.scale; src.scale2 = m.scale * m.scale;
	src.do_ctf_refs = M.do_ctf_correction && M.refs_are_ctf_corrected && src.ctf;
	src.do_scale = M.do_scale_correction;
	src.n_array = M.current_size;
	src.cc_corr = 0.f;
}

// prepared image: corrections applied once per particle instead of once per (orientation, pixel)
struct PrepArgs {
	ImgSrc src;                    // stage mode: src.re != nullptr
	const RbPartMeta *metas; const float2 *Fimg; const float *Fctf;   // pool mode
	const short *ires;             // pool: dense shell map (-1 = excluded); stage: nullptr
	const RbRow *rows; int nrows;  // valid runs
	int n;
	float4 *out;
	// cross-correlation criterion: the prepared weight is corr itself (the CC kernels do not halve it, diff2.h:712-713);
	// pool mode takes the per-particle 1 / sqrtXi2^2 from cc_corr[p]
	int cc; const float *cc_corr;
};

static __global__ void k_prep_img4(PrepArgs A, RbModelDev M)
{
	const int xs = A.n / 2 + 1;
	const int p = blockIdx.y;
	ImgSrc src = A.src;
	if (!src.re) { img_src_pool(src, M, A.metas[p], A.Fimg, A.Fctf, p); if (A.cc) src.cc_corr = A.cc_corr[p]; }
	float4 *out = A.out + (size_t) p * A.n * xs;
	// rows without any valid pixel keep zero weight: clear first, then fill the runs
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.nrows * xs; i += gridDim.x * blockDim.x)
	{
		const int r = i / xs, x = i - r * xs;
		const RbRow rd = A.rows[r];
		const int idx = rd.iy * xs + x;                       // position in the n-window
		float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
		const int ires = A.ires ? (int) A.ires[idx] : 0;
		if (x >= rd.x_lo && x <= rd.x_hi && ires >= 0)
		{
			float2 X; float corr;
			img_load_idx(src, rb_src_index(x, rd.y, src.n_array), ires, X, corr);   // windowFourierTransform (src/fftw.h:850-856)
			v = make_float4(X.x, X.y, A.cc ? corr : corr * 0.5f, 0.f);
		}
		out[idx] = v;
	}
}

