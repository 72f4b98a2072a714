// relion_b200 — coarse-pass squared differences (sm_100a).
//
// Replaces cuda_kernel_diff2_coarse (/root/reference/src/acc/cuda/cuda_kernels/diff2.cuh:24-189; ALTCPU twin
// src/acc/cpu/cpu_kernels/diff2.h:32-282) and mapAllWeightsToMweights (helper.cu:782-796) for a whole pool:
// one launch, no host sync, images corrected once per particle (k_prep_img4).
//
// diff2[o][t] = sum c (|A_o|^2 + |X|^2) - 2 Re sum Z_o e^{i phi_t},  Z_o = c conj(A_o) X,  c = corr/2:
// per (pixel, translation) one phase factor (two table lookups + a complex product) is shared by the CO_EO
// orientations of the CTA, each of which then costs two FMAs.  Lanes walk consecutive x of a row, so the gathers of
// neighbouring lanes fall into the same 128-byte lines of the (L2-resident, ~20 MB) low-resolution core of the
// compact reference.  A quad-per-row variant with separable phases (as in the fine pass) was measured slower here
// (16.2 ms vs 10.6 ms for the headline pool): with only ~30 translations the 4x redundant coordinate math dominates.
#include "img_src.cuh"
#include <cstdlib>

// phase tables for a chunk of translations: tab_x[t][x] = (cos, sin)(x*tx), tab_y[t][y+yoff] = (cos, sin)(y*ty)
// (computeSincosLookupTable2D, cpu_kernels/helper.h:622-660; negative y uses cos(-a)=cos a, sin(-a)=-sin a)
__device__ __forceinline__ void build_tables(float2 *tab_x, float2 *tab_y, int imgX, int ny, int yoff,
                                             const float *tx, const float *ty, int ntr)
{
	for (int i = threadIdx.x; i < ntr * imgX; i += blockDim.x)
	{
		int t = i / imgX, x = i - t * imgX;
		float s, c; sincosf(x * tx[t], &s, &c);
		tab_x[i] = make_float2(c, s);
	}
	for (int i = threadIdx.x; i < ntr * ny; i += blockDim.x)
	{
		int t = i / ny, yy = i - t * ny;
		int y = yy - yoff;
		float s, c; sincosf((y < 0 ? -y : y) * ty[t], &s, &c);
		tab_y[i] = make_float2(c, y < 0 ? -s : s);
	}
}

// Transposing butterfly: every lane holds N (multiple of 32) partial values; afterwards lane L holds,
// for each group g, the warp total of value g*32+L in v[g*32].  31 shuffles per 32 values instead of 160.
template <int N>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[N])
{
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int g = 0; g < N / 32; g++)
	{
#pragma unroll
		for (int s = 16; s >= 1; s >>= 1)
		{
			const bool up = (lane & s) != 0;
#pragma unroll
			for (int j = 0; j < s; j++)
			{
				float send = up ? v[g * 32 + j] : v[g * 32 + j + s];
				float keep = up ? v[g * 32 + j + s] : v[g * 32 + j];
				v[g * 32 + j] = keep + __shfl_xor_sync(RB_FULL_MASK, send, s);
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// coarse pass
// ---------------------------------------------------------------------------------------------
static const int CO_THREADS = 128;
// CO_EO orientations per CTA x CO_TT translations per register chunk = 96 accumulators per thread; the launcher
// picks (EO, TT) from {(8,12), (4,24), (3,32)} so that all translations fit one chunk whenever T <= 32 and every
// projected sample is gathered once.

struct CoarseArgs {
	// pool mode
	const RbPartMeta *metas; RbPartState *states;
	const int *dir_idx, *psi_idx;
	const unsigned char *pdf_orient_zero;
	float *Mweight;
	// stage mode (metas == nullptr)
	const float *st_eulers; float *st_out; int st_O; int st_class;
	const float4 *img4;            // prepared images at the coarse window [P][n][n/2+1]: (X'.re, X'.im, corr/2, 0)
	// common
	const RbProjector *projs;
	const uint32_t *pix; int npix; int n; // window
	const float *tx, *ty; int T;
	int ny, yoff;                  // extent / offset of the y phase table
	int tiles_per_class;           // CTAs per class (tiles never straddle classes)
	int cc;                        // cross-correlation criterion (cuda_kernel_diff2_CC_coarse, diff2.cuh:336-460; ALTCPU
	                               // cpu_kernels/diff2.h:611-742): value = -sum(corr Re(A conj X_t)) / sqrt(sum(corr |A|^2)),
	                               // img4.z holds corr (not corr/2), no Xi2 term, no running minimum (values are negative)
};

template <int CO_EO, int CO_TT, bool XP>
__global__ void __launch_bounds__(CO_THREADS)
k_diff2_coarse(CoarseArgs A, RbModelDev M, RbSamplingDev S)
{
	extern __shared__ float2 smem2[];
	__shared__ float s_e[CO_EO][6];
	__shared__ int s_valid[CO_EO];
	__shared__ float s_red[CO_THREADS / 32][CO_EO * CO_TT];
	__shared__ float s_min[32];
	__shared__ float s_base[CO_THREADS / 32][CO_EO];

	const int imgX = A.n / 2 + 1;
	const int ny = A.ny, yoff = A.yoff;
	float2 *tab_x = smem2;
	float2 *tab_y = smem2 + CO_TT * imgX;

	const bool stage = (A.metas == nullptr);
	const int p = blockIdx.y;
	const int cls = stage ? A.st_class : blockIdx.x / A.tiles_per_class;
	const int oi0 = (blockIdx.x - (stage ? 0 : cls * A.tiles_per_class)) * CO_EO;
	int no, np = 1;
	RbPartMeta m;
	if (stage) no = A.st_O;
	else { m = A.metas[p]; no = m.nd * m.np; np = m.np; }
	if (oi0 >= no) return;
	const int o0 = cls * no + oi0;   // dense orientation index (iorientclass) of the tile's first entry

	if (threadIdx.x < CO_EO)
	{
		const int e = threadIdx.x, oi = oi0 + e;
		int valid = oi < no;
		const float *eu = nullptr;
		if (valid)
		{
			if (stage) eu = A.st_eulers + (size_t) oi * 9;
			else
			{
				valid = !A.pdf_orient_zero[m.prior_off + o0 + e];
				int idl = oi / np, ipl = oi - idl * np;
				int gd = m.dir_off < 0 ? idl : A.dir_idx[m.dir_off + idl];
				int gp = m.psi_off < 0 ? ipl : A.psi_idx[m.psi_off + ipl];
				eu = S.coarse_eulers + ((size_t) gd * S.n_psi + gp) * 9;
			}
		}
		s_valid[e] = valid;
		if (valid) { s_e[e][0] = eu[0]; s_e[e][1] = eu[1]; s_e[e][2] = eu[3]; s_e[e][3] = eu[4]; s_e[e][4] = eu[6]; s_e[e][5] = eu[7]; }
	}
	__syncthreads();
	bool any = false;
#pragma unroll
	for (int e = 0; e < CO_EO; e++) any |= (s_valid[e] != 0);
	if (!any) return;   // Mweight stays lowest() (acc_ml_optimiser_impl.h:3849)

	const float4 *img = A.img4 + (size_t) p * A.n * imgX;
	const RbProjK pk = XP ? rb_make_projk2(A.projs[cls], imgX) : rb_make_projk(A.projs[cls], imgX);   // XP: geometry of the x-pair copy
	const float4 *mdl2 = A.projs[cls].mdl2;   // nullptr in stage mode with a caller-supplied projector copy

	float bmin = FLT_MAX;
	for (int t0 = 0; t0 < A.T; t0 += CO_TT)
	{
		const int ntr = min(CO_TT, A.T - t0);
		__syncthreads();
		build_tables(tab_x, tab_y, imgX, ny, yoff, A.tx + t0, A.ty + t0, ntr);
		__syncthreads();

		float acc[CO_EO * CO_TT];
#pragma unroll
		for (int i = 0; i < CO_EO * CO_TT; i++) acc[i] = 0.f;
		float base[CO_EO];
#pragma unroll
		for (int e = 0; e < CO_EO; e++) base[e] = 0.f;

		for (int ip = threadIdx.x; ip < A.npix; ip += CO_THREADS)
		{
			const uint32_t pkx = __ldg(A.pix + ip);
			const int x = rb_pix_x(pkx), y = rb_pix_y(pkx);
			const float4 im = __ldg(img + rb_src_index(x, y, A.n));
			const float hc = im.z;                              // s_corr = corr/2 (diff2.h:114)
			float zr[CO_EO], zi[CO_EO];
#pragma unroll
			for (int e = 0; e < CO_EO; e++)
			{
				float2 ref = make_float2(0.f, 0.f);
				if (s_valid[e])
					ref = XP ? rb_project3d_xp(pk, mdl2, x, y, s_e[e][0], s_e[e][1], s_e[e][2], s_e[e][3], s_e[e][4], s_e[e][5])
					         : rb_project3d(pk, x, y, s_e[e][0], s_e[e][1], s_e[e][2], s_e[e][3], s_e[e][4], s_e[e][5]);
				zr[e] = hc * (ref.x * im.x + ref.y * im.y);
				zi[e] = hc * (ref.x * im.y - ref.y * im.x);
				base[e] += hc * ((ref.x * ref.x + ref.y * ref.y) + (A.cc ? 0.f : (im.x * im.x + im.y * im.y)));
			}
			const float2 *txp = tab_x + x, *typ = tab_y + (y + yoff);
#pragma unroll
			for (int t = 0; t < CO_TT; t++)
			{
				if (t < ntr)
				{
					const float2 a = txp[t * imgX], b = typ[t * ny];
					const float ss = a.y * b.x + a.x * b.y;     // sin(x tx + y ty)
					const float cc = a.x * b.x - a.y * b.y;     // cos
#pragma unroll
					for (int e = 0; e < CO_EO; e++)
						acc[e * CO_TT + t] = fmaf(zr[e], cc, fmaf(-zi[e], ss, acc[e * CO_TT + t]));
				}
			}
		}
		// reduce over the CTA: butterfly inside each warp, then across the 4 warps through smem
		warp_transpose_reduce<CO_EO * CO_TT>(acc);
		const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
		for (int g = 0; g < CO_EO * CO_TT / 32; g++) s_red[wid][g * 32 + lane] = acc[g * 32];
#pragma unroll
		for (int e = 0; e < CO_EO; e++) { const float bsum = warp_sum(base[e]); if (lane == 0) s_base[wid][e] = bsum; }
		__syncthreads();
		if (threadIdx.x < CO_EO * CO_TT)
		{
			const int e = threadIdx.x / CO_TT, t = threadIdx.x - e * CO_TT;
			if (t < ntr && s_valid[e])
			{
				float cr = 0.f, bs = 0.f;
#pragma unroll
				for (int w = 0; w < CO_THREADS / 32; w++) { cr += s_red[w][threadIdx.x]; bs += s_base[w][e]; }
				float v = A.cc ? -(cr / sqrtf(bs)) : fmaxf(bs - 2.f * cr, 0.f);   // CC: diff2.h:729-735
				const int o = o0 + e;
				if (stage) A.st_out[(size_t) o * A.T + t0 + t] += v;          // += like the reference kernel
				else
				{
					if (!A.cc) v += m.xi2_half;                               // :1287-1297 (no Xi2 term with CC)
					A.Mweight[m.coarse_off + (long long) o * A.T + t0 + t] = v;
					bmin = fminf(bmin, v);
				}
			}
		}
	}
	if (!stage && !A.cc)
	{
		bmin = -block_max(-bmin, s_min);
		if (threadIdx.x == 0 && bmin < FLT_MAX) rb_atomic_min_pos(&A.states[p].min_diff2_bits, bmin);
	}
}

template <int EO, int TT, bool XP>
static int launch_coarse_cfg(rb_ctx *ctx, CoarseArgs &A, int no_max, int n_classes, int P)
{
	size_t sm = (size_t) TT * ((A.n / 2 + 1) + A.ny) * sizeof(float2);
	static size_t configured[RB_MAX_DEVICES] = {};
	size_t &cfg = configured[ctx->device % RB_MAX_DEVICES];
	if (sm > cfg)
	{
		RB_CUDA(cudaFuncSetAttribute(k_diff2_coarse<EO, TT, XP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
		cfg = sm;
	}
	A.tiles_per_class = (no_max + EO - 1) / EO;
	dim3 grid(A.tiles_per_class * n_classes, P);
	k_diff2_coarse<EO, TT, XP><<<grid, CO_THREADS, sm, ctx->stream>>>(A, ctx->d_model, ctx->d_samp);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

static int launch_coarse(rb_ctx *ctx, CoarseArgs &A, int no_max, int n_classes, int P)
{
	static int xp = -1;
	if (xp < 0) { const char *e = getenv("RB_COARSE_XP"); xp = e ? atoi(e) : 1; }
	if (xp)
	{
		if (A.T <= 12) return launch_coarse_cfg<8, 12, true>(ctx, A, no_max, n_classes, P);
		if (A.T <= 24) return launch_coarse_cfg<4, 24, true>(ctx, A, no_max, n_classes, P);
		return launch_coarse_cfg<3, 32, true>(ctx, A, no_max, n_classes, P);
	}
	if (A.T <= 12) return launch_coarse_cfg<8, 12, false>(ctx, A, no_max, n_classes, P);
	if (A.T <= 24) return launch_coarse_cfg<4, 24, false>(ctx, A, no_max, n_classes, P);
	return launch_coarse_cfg<3, 32, false>(ctx, A, no_max, n_classes, P);
}

__global__ void k_fill(float *p, float v, size_t n)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}

// exp_local_sqrtXi2 (src/ml_optimiser.cpp:6846-6856): sqrt of the power of ALL pixels of the masked transform windowed to
// the pass' size, accumulated in double; buildCorrImage (acc_helper_functions_impl.h:172-174) turns it into the uniform
// weight 1 / (sqrtXi2 * sqrtXi2).  One CTA per (particle, window): blockIdx.y = 0 coarse window, 1 fine window.
__global__ void __launch_bounds__(256)
k_cc_corr(const float2 *Fimg, int n_src, int n_coarse, int P, float *out)
{
	__shared__ double red[8];
	const int p = blockIdx.x, which = blockIdx.y;
	const int n = which ? n_src : n_coarse, xs = n / 2 + 1;
	const float2 *F = Fimg + (size_t) p * n_src * (n_src / 2 + 1);
	double acc = 0.;
	for (int i = threadIdx.x; i < n * xs; i += blockDim.x)
	{
		const int iy = i / xs, x = i - iy * xs;
		const int y = iy < xs ? iy : iy - n;                      // windowFourierTransform (src/fftw.h:850-856)
		const float2 v = __ldg(F + rb_src_index(x, y, n_src));
		acc += (double) v.x * (double) v.x + (double) v.y * (double) v.y;
	}
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(RB_FULL_MASK, acc, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		double t = 0.;
		for (int w = 0; w < 8; w++) t += red[w];
		const double sq = sqrt(t);
		out[(size_t) which * P + p] = (float) (1. / (sq * sq));
	}
}

int rbk_cc_corr_pool(rb_ctx *ctx, PoolSlot &s)
{
	const RbModelDev &M = ctx->d_model;
	RB_CHECK(s.cc_corr.ensure((size_t) 2 * s.P * sizeof(float)));
	k_cc_corr<<<dim3(s.P, 2), 256, 0, ctx->stream>>>(s.Fimg.as<float2>(), M.current_size, M.coarse_size, s.P, s.cc_corr.as<float>());
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

int rbk_diff2_coarse_pool(rb_ctx *ctx, PoolSlot &s)
{
	// Mweight <- lowest() (acc_ml_optimiser_impl.h:3849); the contraction's epilogue writes every entry itself
	const bool gemm_path = rbk_coarse_gemm_applicable(ctx, s);
	if (!gemm_path)
	{
		k_fill<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(s.Mweight.as<float>(), RB_LOWEST, (size_t) s.total_coarse);
		RB_LAUNCH_CHECK(ctx);
	}
	// images at the coarse window with all corrections applied, once per particle
	const RbModelDev &M = ctx->d_model;
	const int nc = M.coarse_size, xsc = nc / 2 + 1;
	RB_CHECK(s.cimg4.ensure((size_t) s.P * nc * xsc * sizeof(float4)));
	RB_CUDA(cudaMemsetAsync(s.cimg4.p, 0, (size_t) s.P * nc * xsc * sizeof(float4), ctx->stream));
	PrepArgs PA;
	memset(&PA, 0, sizeof(PA));
	PA.metas = s.meta.as<RbPartMeta>(); PA.Fimg = s.Fimg.as<float2>();
	PA.Fctf = ctx->h_model.do_ctf_correction ? s.Fctf.as<float>() : nullptr;
	PA.ires = M.d2_ires_c; PA.rows = M.d2_rows_c; PA.nrows = M.d2_nrows_c; PA.n = nc; PA.out = s.cimg4.as<float4>();
	if (M.do_cc)
	{
		// exp_local_sqrtXi2 of both windows (src/ml_optimiser.cpp:6846-6856) -> 1 / sqrtXi2^2 per particle
		RB_CHECK(rbk_cc_corr_pool(ctx, s));
		PA.cc = 1; PA.cc_corr = s.cc_corr.as<float>();
	}
	dim3 pg((M.d2_nrows_c * xsc + 255) / 256, s.P);
	k_prep_img4<<<pg, 256, 0, ctx->stream>>>(PA, M);
	RB_LAUNCH_CHECK(ctx);

	// global searches: the cross term is a dense contraction shared by the whole pool -> tensor cores (also with the
	// cross-correlation criterion: same cross and norm terms, different epilogue)
	if (gemm_path) return rbk_diff2_coarse_gemm_pool(ctx, s, s.cimg4.as<float4>());
	// local searches: projection fused with the contraction, per (particle, 128-orientation tile), for both criteria
	// (--always_cc late in a refinement: CC epilogue)
	if (rbk_coarse_fused_applicable(ctx, s)) return rbk_diff2_coarse_fused_pool(ctx, s, s.cimg4.as<float4>());

	CoarseArgs A;
	memset(&A, 0, sizeof(A));
	A.metas = s.meta.as<RbPartMeta>(); A.states = s.state.as<RbPartState>();
	A.img4 = s.cimg4.as<float4>();
	A.dir_idx = s.dir_idx.as<int>(); A.psi_idx = s.psi_idx.as<int>();
	A.pdf_orient_zero = s.pdf_orient_zero.as<unsigned char>();
	A.Mweight = s.Mweight.as<float>();
	A.projs = ctx->d_proj.as<RbProjector>();
	A.pix = ctx->d_model.d2_pix_c; A.npix = ctx->d_model.d2_nvc; A.n = ctx->d_model.coarse_size;
	A.tx = ctx->d_samp.ctx; A.ty = ctx->d_samp.cty; A.T = ctx->d_samp.n_trans;
	A.ny = A.n + 1; A.yoff = A.n / 2;
	A.cc = M.do_cc;
	return launch_coarse(ctx, A, s.max_no, ctx->d_model.nr_classes, s.P);
}

int rbk_diff2_coarse_stage(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers, int O,
                           const float *d_tx, const float *d_ty, int T, const float *d_re, const float *d_im,
                           const float *d_corr, float *d_out, int cc)
{
	// full pixel list with the coarse kernel's y-wrap: y > maxR -> y - imgY (diff2.cuh:89-90, diff2.h:109-110)
	const int imgX = n / 2 + 1;
	RbProjK pk = rb_make_projk(pj, imgX);
	std::vector<uint32_t> pix((size_t) n * imgX);
	for (int iy = 0; iy < n; iy++)
		for (int x = 0; x < imgX; x++)
		{
			int y = iy > pk.maxR ? iy - n : iy;
			pix[(size_t) iy * imgX + x] = rb_pack_pix(x, y, 0);
		}
	RB_CHECK(ctx->scratch[0].ensure(pix.size() * 4));
	RB_CHECK(ctx->scratch[1].ensure(sizeof(RbProjector)));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[0].p, pix.data(), pix.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[1].p, &pj, sizeof(RbProjector), cudaMemcpyHostToDevice, ctx->stream));
	std::vector<RbRow> rows;
	for (int iy = 0; iy < n; iy++)
		rows.push_back(RbRow{(short) iy, (short) (iy > pk.maxR ? iy - n : iy), (short) 0, (short) (imgX - 1)});
	RB_CHECK(ctx->scratch[7].ensure(rows.size() * sizeof(RbRow)));
	RB_CHECK(ctx->scratch[6].ensure((size_t) n * imgX * sizeof(float4)));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[7].p, rows.data(), rows.size() * sizeof(RbRow), cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));   // `pix` and `rows` are pageable temporaries
	PrepArgs PA;
	memset(&PA, 0, sizeof(PA));
	PA.src.re = d_re; PA.src.im = d_im; PA.src.corr = d_corr; PA.src.n_array = n;
	PA.rows = ctx->scratch[7].as<RbRow>(); PA.nrows = (int) rows.size(); PA.n = n; PA.out = ctx->scratch[6].as<float4>();
	PA.cc = cc;
	k_prep_img4<<<dim3((n * imgX + 255) / 256, 1), 256, 0, ctx->stream>>>(PA, ctx->d_model);
	RB_LAUNCH_CHECK(ctx);
	CoarseArgs A;
	memset(&A, 0, sizeof(A));
	A.st_eulers = d_eulers; A.st_out = d_out; A.st_O = O; A.st_class = 0;
	A.img4 = ctx->scratch[6].as<float4>();
	A.projs = ctx->scratch[1].as<RbProjector>();
	A.pix = ctx->scratch[0].as<uint32_t>(); A.npix = (int) pix.size(); A.n = n;
	A.tx = d_tx; A.ty = d_ty; A.T = T;
	A.ny = 2 * n + 1; A.yoff = n;   // un-wrapped rows can reach -n when the projector's r_max < n/2
	A.cc = cc;
	return launch_coarse(ctx, A, O, 1, 1);
}
