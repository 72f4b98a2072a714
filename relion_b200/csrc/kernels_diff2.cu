// relion_b200 — squared-difference kernels (coarse and fine pass), sm_100a.
//
// Replaces cuda_kernel_diff2_coarse / cuda_kernel_diff2_fine
// (/root/reference/src/acc/cuda/cuda_kernels/diff2.cuh:24-189, 193-332) and their ALTCPU twins
// (src/acc/cpu/cpu_kernels/diff2.h:32-430) with kernels batched over a whole pool of particles:
// no per-particle launch, no host sync, image corrections (pixel_correction, corr_img —
// acc_ml_optimiser_impl.h:1251-1268, acc_helper_functions_impl.h:164-196) applied on the fly.
#include "device_utils.cuh"

// ---------------------------------------------------------------------------------------------
// image source: either reference-style pre-corrected SoA arrays (stage API) or the pool's raw
// Fimg / Fctf with the corrections computed in registers
// ---------------------------------------------------------------------------------------------
struct ImgSrc {
	const float *re, *im, *corr;   // stage mode when re != nullptr
	const float2 *F;               // pool mode
	const float *ctf;
	const float *minvs2;           // [nshell] of the particle's optics group
	float inv_scale, scale2;
	int do_ctf_refs;               // do_ctf_correction && refs_are_ctf_corrected
	int do_scale;
	int n_array;                   // window size the arrays are stored at
};

__device__ __forceinline__ void img_load(const ImgSrc &s, uint32_t pk, float2 &X, float &corr)
{
	const int x = rb_pix_x(pk), y = rb_pix_y(pk);
	const int idx = rb_src_index(x, y, s.n_array);
	if (s.re)
	{
		X = make_float2(__ldg(s.re + idx), __ldg(s.im + idx));
		corr = __ldg(s.corr + idx);
	}
	else
	{
		const int ires = rb_pix_ires(pk);
		float2 F = __ldg(s.F + idx);
		float pc = s.inv_scale;
		float c = ires > 0 ? __ldg(s.minvs2 + ires) : 0.f;       // DC excluded (src/ml_optimiser.cpp:6874-6879)
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
// priors: pdf_orientation = log(pdf), zero flags (initOrientations, utilities_impl.h:656-668);
// pdf_offset (acc_ml_optimiser_impl.h:2094-2171)
// ---------------------------------------------------------------------------------------------
__global__ void k_prep_priors(const RbPartMeta *metas, RbModelDev M, RbSamplingDev S,
                              const int *dir_idx, const double *dir_prior, const int *psi_idx, const double *psi_prior,
                              float *pdf_orient, unsigned char *pdf_orient_zero,
                              float *pdf_offset, unsigned char *pdf_offset_zero, RbPartState *states)
{
	const int p = blockIdx.y;
	const RbPartMeta m = metas[p];
	const int no = m.nd * m.np;
	const int ndense = M.nr_classes * no;
	for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < ndense; o += gridDim.x * blockDim.x)
	{
		int k = o / no, oi = o - k * no, idl = oi / m.np, ipl = oi - idl * m.np;
		double pdf;
		if (m.dir_off < 0) pdf = M.pdf_direction[(size_t) k * S.n_dir + idl];
		else pdf = dir_prior[m.dir_off + idl] * psi_prior[m.psi_off + ipl];
		if (!(M.pdf_class[k] > 0.)) pdf = 0.;   // classes with zero pdf_class are never evaluated (:1069)
		pdf_orient_zero[m.prior_off + o] = (pdf == 0.);
		pdf_orient[m.prior_off + o] = (pdf == 0.) ? 0.f : (float) log(pdf);
	}
	if (blockIdx.x == 0)
	{
		for (int t = threadIdx.x; t < S.n_trans; t += blockDim.x)
		{
			double offx = m.oldx + S.trans_x[t], offy = m.oldy + S.trans_y[t];
			double tdiff2 = (offx - m.prx) * (offx - m.prx) / (-2. * M.s2off) + (offy - m.pry) * (offy - m.pry) / (-2. * M.s2off);
			tdiff2 *= M.pixel_size * M.pixel_size;
			double pdf; bool z;
			if (M.s2off < 0.0001) { z = tdiff2 > 0.; pdf = z ? 0. : 1.; }
			else { z = false; pdf = tdiff2; }
			pdf_offset_zero[(size_t) p * S.n_trans + t] = z;
			pdf_offset[(size_t) p * S.n_trans + t] = (float) pdf;
		}
		if (threadIdx.x == 0)
		{
			RbPartState st;
			memset(&st, 0, sizeof(st));
			st.min_diff2_bits = 0x7f7fffff; st.fmin_bits = 0x7f7fffff;
			states[p] = st;
		}
	}
}

int rbk_prep_priors(rb_ctx *ctx, PoolSlot &s)
{
	dim3 grid((s.max_no * ctx->d_model.nr_classes + 255) / 256, s.P);
	if (grid.x > 64) grid.x = 64;
	if (grid.x < 1) grid.x = 1;
	k_prep_priors<<<grid, 256, 0, ctx->stream>>>(s.meta.as<RbPartMeta>(), ctx->d_model, ctx->d_samp,
		s.dir_idx.as<int>(), s.dir_prior.as<double>(), s.psi_idx.as<int>(), s.psi_prior.as<double>(),
		s.pdf_orient.as<float>(), s.pdf_orient_zero.as<unsigned char>(),
		s.pdf_offset.as<float>(), s.pdf_offset_zero.as<unsigned char>(), s.state.as<RbPartState>());
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
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
	const float2 *Fimg; const float *Fctf;
	const int *dir_idx, *psi_idx;
	const unsigned char *pdf_orient_zero;
	float *Mweight;
	// stage mode (metas == nullptr)
	const float *st_eulers; const float *st_re, *st_im, *st_corr; float *st_out; int st_O; int st_class;
	// common
	const RbProjector *projs;
	const uint32_t *pix; int npix; int n; // window
	const float *tx, *ty; int T;
	int ny, yoff;                  // extent / offset of the y phase table
	int tiles_per_class;           // CTAs per class (tiles never straddle classes)
};

template <int CO_EO, int CO_TT>
__global__ void __launch_bounds__(CO_THREADS)
k_diff2_coarse(CoarseArgs A, RbModelDev M, RbSamplingDev S)
{
	extern __shared__ float2 smem2[];
	__shared__ float s_e[CO_EO][6];
	__shared__ int s_valid[CO_EO];
	__shared__ float s_red[CO_THREADS / 32][CO_EO * CO_TT];
	__shared__ float s_min[32];

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

	ImgSrc src;
	if (stage) { src.re = A.st_re; src.im = A.st_im; src.corr = A.st_corr; src.n_array = A.n; }
	else
	{
		src.re = nullptr;
		src.F = A.Fimg + (size_t) p * M.Npf; src.ctf = A.Fctf ? A.Fctf + (size_t) p * M.Npf : nullptr;
		src.minvs2 = M.minvs2 + (size_t) m.og * M.nshell;
		src.inv_scale = 1.0f / m.scale; src.scale2 = m.scale * m.scale;
		src.do_ctf_refs = M.do_ctf_correction && M.refs_are_ctf_corrected && src.ctf;
		src.do_scale = M.do_scale_correction;
		src.n_array = M.current_size;
	}
	const RbProjK pk = rb_make_projk(A.projs[cls], imgX);

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

		for (int ip = threadIdx.x; ip < A.npix; ip += CO_THREADS)
		{
			const uint32_t pkx = __ldg(A.pix + ip);
			const int x = rb_pix_x(pkx), y = rb_pix_y(pkx);
			float2 X; float corr;
			img_load(src, pkx, X, corr);
			const float hc = corr * 0.5f;                       // s_corr = corr/2 (diff2.h:114)
			float2 ref[CO_EO];
#pragma unroll
			for (int e = 0; e < CO_EO; e++)
				ref[e] = s_valid[e] ? rb_project3d(pk, x, y, s_e[e][0], s_e[e][1], s_e[e][2], s_e[e][3], s_e[e][4], s_e[e][5])
				                    : make_float2(0.f, 0.f);
			const float2 *txp = tab_x + x, *typ = tab_y + (y + yoff);
#pragma unroll
			for (int t = 0; t < CO_TT; t++)
			{
				if (t < ntr)
				{
					const float2 a = txp[t * imgX], b = typ[t * ny];
					const float ss = a.y * b.x + a.x * b.y;     // sin(x tx + y ty)
					const float cc = a.x * b.x - a.y * b.y;     // cos
					const float sr = cc * X.x - ss * X.y;
					const float si = cc * X.y + ss * X.x;
#pragma unroll
					for (int e = 0; e < CO_EO; e++)
					{
						const float dr = ref[e].x - sr, di = ref[e].y - si;
						acc[e * CO_TT + t] += (dr * dr + di * di) * hc;
					}
				}
			}
		}
		// reduce over the CTA: butterfly inside each warp, then across the 4 warps through smem
		warp_transpose_reduce<CO_EO * CO_TT>(acc);
		const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
		for (int g = 0; g < CO_EO * CO_TT / 32; g++) s_red[wid][g * 32 + lane] = acc[g * 32];
		__syncthreads();
		if (threadIdx.x < CO_EO * CO_TT)
		{
			const int e = threadIdx.x / CO_TT, t = threadIdx.x - e * CO_TT;
			if (t < ntr && s_valid[e])
			{
				float v = 0.f;
#pragma unroll
				for (int w = 0; w < CO_THREADS / 32; w++) v += s_red[w][threadIdx.x];
				const int o = o0 + e;
				if (stage) A.st_out[(size_t) o * A.T + t0 + t] += v;          // += like the reference kernel
				else
				{
					v += m.xi2_half;                                          // :1290-1296
					A.Mweight[m.coarse_off + (long long) o * A.T + t0 + t] = v;
					bmin = fminf(bmin, v);
				}
			}
		}
	}
	if (!stage)
	{
		bmin = -block_max(-bmin, s_min);
		if (threadIdx.x == 0 && bmin < FLT_MAX) rb_atomic_min_pos(&A.states[p].min_diff2_bits, bmin);
	}
}

template <int EO, int TT>
static int launch_coarse_cfg(rb_ctx *ctx, CoarseArgs &A, int no_max, int n_classes, int P)
{
	size_t sm = (size_t) TT * ((A.n / 2 + 1) + A.ny) * sizeof(float2);
	static size_t configured = 0;
	if (sm > configured)
	{
		RB_CUDA(cudaFuncSetAttribute(k_diff2_coarse<EO, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
		configured = sm;
	}
	A.tiles_per_class = (no_max + EO - 1) / EO;
	dim3 grid(A.tiles_per_class * n_classes, P);
	k_diff2_coarse<EO, TT><<<grid, CO_THREADS, sm, ctx->stream>>>(A, ctx->d_model, ctx->d_samp);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

static int launch_coarse(rb_ctx *ctx, CoarseArgs &A, int no_max, int n_classes, int P)
{
	if (A.T <= 12) return launch_coarse_cfg<8, 12>(ctx, A, no_max, n_classes, P);
	if (A.T <= 24) return launch_coarse_cfg<4, 24>(ctx, A, no_max, n_classes, P);
	return launch_coarse_cfg<3, 32>(ctx, A, no_max, n_classes, P);
}

__global__ void k_fill(float *p, float v, size_t n)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}

int rbk_diff2_coarse_pool(rb_ctx *ctx, PoolSlot &s)
{
	// Mweight <- lowest() (acc_ml_optimiser_impl.h:3849)
	k_fill<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(s.Mweight.as<float>(), RB_LOWEST, (size_t) s.total_coarse);
	RB_LAUNCH_CHECK(ctx);
	CoarseArgs A;
	memset(&A, 0, sizeof(A));
	A.metas = s.meta.as<RbPartMeta>(); A.states = s.state.as<RbPartState>();
	A.Fimg = s.Fimg.as<float2>(); A.Fctf = ctx->h_model.do_ctf_correction ? s.Fctf.as<float>() : nullptr;
	A.dir_idx = s.dir_idx.as<int>(); A.psi_idx = s.psi_idx.as<int>();
	A.pdf_orient_zero = s.pdf_orient_zero.as<unsigned char>();
	A.Mweight = s.Mweight.as<float>();
	A.projs = ctx->d_proj.as<RbProjector>();
	A.pix = ctx->d_model.pix_c; A.npix = ctx->d_model.nvc; A.n = ctx->d_model.coarse_size;
	A.tx = ctx->d_samp.ctx; A.ty = ctx->d_samp.cty; A.T = ctx->d_samp.n_trans;
	A.ny = A.n + 1; A.yoff = A.n / 2;
	return launch_coarse(ctx, A, s.max_no, ctx->d_model.nr_classes, s.P);
}

int rbk_diff2_coarse_stage(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers, int O,
                           const float *d_tx, const float *d_ty, int T, const float *d_re, const float *d_im,
                           const float *d_corr, float *d_out)
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
	RB_CUDA(cudaStreamSynchronize(ctx->stream));   // `pix` is a pageable temporary
	CoarseArgs A;
	memset(&A, 0, sizeof(A));
	A.st_eulers = d_eulers; A.st_re = d_re; A.st_im = d_im; A.st_corr = d_corr; A.st_out = d_out; A.st_O = O; A.st_class = 0;
	A.projs = ctx->scratch[1].as<RbProjector>();
	A.pix = ctx->scratch[0].as<uint32_t>(); A.npix = (int) pix.size(); A.n = n;
	A.tx = d_tx; A.ty = d_ty; A.T = T;
	A.ny = 2 * n + 1; A.yoff = n;   // un-wrapped rows can reach -n when the projector's r_max < n/2
	return launch_coarse(ctx, A, O, 1, 1);
}

// ---------------------------------------------------------------------------------------------
// fine pass: one CTA per oversampled orientation (persistent grid), each projected pixel is reused
// for all of that orientation's significant translations
// ---------------------------------------------------------------------------------------------
static const int FI_THREADS = 256;
static const int FI_TF = 32;   // fine translations accumulated per pass (8 significant coarse translations x 4)

struct FineArgs {
	// pool mode
	const RbPartMeta *metas; RbPartState *states;
	const float2 *Fimg; const float *Fctf;
	const RbFineOrient *fo; const int *pair_list; const int *counters; // counters[0] = number of fine orientations
	float *fs_w;
	// stage mode (fo == nullptr): the reference's job lists
	const float *st_eulers, *st_re, *st_im, *st_corr; float st_sum_init;
	const unsigned long long *st_rot_idx, *st_trans_idx, *st_job_idx, *st_job_num; int st_njobs;
	float *st_out;
	// common
	const RbProjector *projs;
	const uint32_t *pix; int npix; int n;
	const float *tx, *ty; int NOT;
};

struct FinePix {
	RbProjFetch pf;
	float2 X;
	float hc;
	int x, y;
};

__device__ __forceinline__ void fine_issue(const FineArgs &A, const ImgSrc &src, const RbProjK8 &pk, int ip,
                                           float e0, float e1, float e3, float e4, float e6, float e7, FinePix &f)
{
	const uint32_t pkx = __ldg(A.pix + ip);
	f.x = rb_pix_x(pkx); f.y = rb_pix_y(pkx);
	rb_proj_issue(pk, f.x, f.y, e0, e1, e3, e4, e6, e7, f.pf);
	float corr;
	img_load(src, pkx, f.X, corr);
	f.hc = corr * 0.5f;
}

// diff2[t] = sum_pix hc*|ref - S_t X|^2 evaluated as  sum hc*(|ref|^2 + |X|^2)  -  2 * sum Re(hc*conj(ref)*X * e^{i phi_t}):
// the cross term is the only part that depends on the translation (the contraction the north star names), so each
// (pixel, translation) costs one phase factor and two FMAs; accumulation stays fp32.
__global__ void __launch_bounds__(FI_THREADS, 2)
k_diff2_fine(FineArgs A, RbModelDev M)
{
	__shared__ float s_ux[FI_TF], s_uy[FI_TF];
	__shared__ float s_red[FI_THREADS / 32][FI_TF + 1];
	__shared__ float s_e[6];

	const int imgX = A.n / 2 + 1;
	const bool stage = (A.fo == nullptr);
	const int nwork = stage ? A.st_njobs : A.counters[0];

	for (int w = blockIdx.x; w < nwork; w += gridDim.x)
	{
		int nsamp, cls = 0, p = 0;
		long long out_off;
		const float *eu;
		RbFineOrient F;
		if (stage)
		{
			unsigned long long j0 = A.st_job_idx[w];
			nsamp = (int) A.st_job_num[w];
			eu = A.st_eulers + A.st_rot_idx[j0] * 9;
			out_off = (long long) j0;
		}
		else
		{
			F = A.fo[w];
			nsamp = F.n_t * A.NOT; cls = F.iclass; p = F.particle; out_off = F.sample_off;
			eu = A.fo[w].e;
		}
		__syncthreads();
		if (threadIdx.x < 6) s_e[threadIdx.x] = eu[threadIdx.x + threadIdx.x / 2];   // elements 0,1,3,4,6,7
		ImgSrc src;
		float xi2_half;
		if (stage) { src.re = A.st_re; src.im = A.st_im; src.corr = A.st_corr; src.n_array = A.n; xi2_half = A.st_sum_init; }
		else
		{
			const RbPartMeta m = A.metas[p];
			src.re = nullptr;
			src.F = A.Fimg + (size_t) p * M.Npf; src.ctf = A.Fctf ? A.Fctf + (size_t) p * M.Npf : nullptr;
			src.minvs2 = M.minvs2 + (size_t) m.og * M.nshell;
			src.inv_scale = 1.0f / m.scale; src.scale2 = m.scale * m.scale;
			src.do_ctf_refs = M.do_ctf_correction && M.refs_are_ctf_corrected && src.ctf;
			src.do_scale = M.do_scale_correction;
			src.n_array = M.current_size;
			xi2_half = m.xi2_half;
		}
		const RbProjK8 pk = rb_make_projk8(A.projs[cls], imgX);
		float bmin = FLT_MAX;

		for (int c0 = 0; c0 < nsamp; c0 += FI_TF)
		{
			const int ntr = min(FI_TF, nsamp - c0);
			__syncthreads();
			if (threadIdx.x < FI_TF)
			{
				float ux = 0.f, uy = 0.f;
				if (threadIdx.x < ntr)
				{
					int j = c0 + threadIdx.x, it;
					if (stage) it = (int) A.st_trans_idx[A.st_job_idx[w]] + j;                   // consecutive translations in a job
					else it = A.pair_list[F.pair_off + j / A.NOT] * A.NOT + (j % A.NOT);
					ux = A.tx[it] * 0.15915494309189535f; uy = A.ty[it] * 0.15915494309189535f;  // radians -> turns per pixel
				}
				s_ux[threadIdx.x] = ux; s_uy[threadIdx.x] = uy;
			}
			__syncthreads();
			const float e0 = s_e[0], e1 = s_e[1], e3 = s_e[2], e4 = s_e[3], e6 = s_e[4], e7 = s_e[5];

			float acc[FI_TF];
#pragma unroll
			for (int i = 0; i < FI_TF; i++) acc[i] = 0.f;
			float base = 0.f;

			// software-pipelined pixel loop: the next pixel's gathers are in flight while this one is accumulated
			int ip = threadIdx.x;
			bool have = ip < A.npix;
			FinePix cur;
			if (have) fine_issue(A, src, pk, ip, e0, e1, e3, e4, e6, e7, cur);
			while (have)
			{
				const int ipn = ip + FI_THREADS;
				const bool haven = ipn < A.npix;
				FinePix nxt;
				if (haven) fine_issue(A, src, pk, ipn, e0, e1, e3, e4, e6, e7, nxt);

				const float2 ref = (cur.pf.flags & 1) ? rb_proj_finish(cur.pf) : make_float2(0.f, 0.f);
				const float zr = cur.hc * (ref.x * cur.X.x + ref.y * cur.X.y);
				const float zi = cur.hc * (ref.x * cur.X.y - ref.y * cur.X.x);
				base += cur.hc * ((ref.x * ref.x + ref.y * ref.y) + (cur.X.x * cur.X.x + cur.X.y * cur.X.y));
#pragma unroll
				for (int t = 0; t < FI_TF; t++)
				{
					if (t < ntr)
					{
						const float2 ph = rb_phase(cur.x, cur.y, s_ux[t], s_uy[t]);
						acc[t] += zr * ph.x - zi * ph.y;
					}
				}
				if (haven) cur = nxt;
				ip = ipn; have = haven;
			}
			warp_transpose_reduce<FI_TF>(acc);
			base = warp_sum(base);
			const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
			s_red[wid][lane] = acc[0];
			if (lane == 0) s_red[wid][FI_TF] = base;
			__syncthreads();
			if (threadIdx.x < ntr)
			{
				float c = 0.f, b = 0.f;
#pragma unroll
				for (int ww = 0; ww < FI_THREADS / 32; ww++) { c += s_red[ww][threadIdx.x]; b += s_red[ww][FI_TF]; }
				float v = (b - 2.f * c) + xi2_half;
				v = fmaxf(v, 0.f);
				if (stage) A.st_out[out_off + c0 + threadIdx.x] += v;                         // diff2.h:424-428
				else { A.fs_w[out_off + c0 + threadIdx.x] = v; bmin = fminf(bmin, v); }
			}
		}
		if (!stage && threadIdx.x < FI_TF && bmin < FLT_MAX) rb_atomic_min_pos(&A.states[p].fmin_bits, bmin);
	}
}

static int launch_fine(rb_ctx *ctx, FineArgs &A, int grid)
{
	k_diff2_fine<<<grid, FI_THREADS, 0, ctx->stream>>>(A, ctx->d_model);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

int rbk_diff2_fine_pool(rb_ctx *ctx, PoolSlot &s)
{
	FineArgs A;
	memset(&A, 0, sizeof(A));
	A.metas = s.meta.as<RbPartMeta>(); A.states = s.state.as<RbPartState>();
	A.Fimg = s.Fimg.as<float2>(); A.Fctf = ctx->h_model.do_ctf_correction ? s.Fctf.as<float>() : nullptr;
	A.fo = s.fo.as<RbFineOrient>(); A.pair_list = s.pair_list.as<int>(); A.counters = s.counters.as<int>();
	A.fs_w = s.fs_w.as<float>();
	A.projs = ctx->d_proj.as<RbProjector>();
	A.pix = ctx->d_model.pix_f; A.npix = ctx->d_model.nvf; A.n = ctx->d_model.current_size;
	A.tx = ctx->d_samp.ftx; A.ty = ctx->d_samp.fty; A.NOT = ctx->d_samp.n_over_trans;
	return launch_fine(ctx, A, ctx->num_sms * 2);
}

int rbk_diff2_fine_stage(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers,
                         const float *d_tx, const float *d_ty, const float *d_re, const float *d_im,
                         const float *d_corr, float sum_init,
                         const unsigned long long *d_rot_idx, const unsigned long long *d_trans_idx,
                         const unsigned long long *d_job_idx, const unsigned long long *d_job_num, int n_jobs,
                         float *d_out)
{
	// pixel list with the fine kernels' row rule (diff2.cuh:268-274, diff2.h:344-355): rows in the dead
	// band maxR < iy < imgY-maxR contribute only the pixel x = maxR (which projects to zero)
	const int imgX = n / 2 + 1;
	RbProjK pk = rb_make_projk(pj, imgX);
	std::vector<uint32_t> pix;
	pix.reserve((size_t) n * imgX);
	for (int iy = 0; iy < n; iy++)
	{
		int xs = 0, xe = imgX, y = iy;
		if (iy > pk.maxR)
		{
			if (iy >= n - pk.maxR) y = iy - n;
			else { xs = pk.maxR; xe = xs + 1; }
		}
		for (int x = xs; x < xe; x++) pix.push_back(rb_pack_pix(x, y, 0));
	}
	RB_CHECK(ctx->scratch[0].ensure(pix.size() * 4));
	RB_CHECK(ctx->scratch[1].ensure(sizeof(RbProjector)));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[0].p, pix.data(), pix.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[1].p, &pj, sizeof(RbProjector), cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	FineArgs A;
	memset(&A, 0, sizeof(A));
	A.st_eulers = d_eulers; A.st_re = d_re; A.st_im = d_im; A.st_corr = d_corr; A.st_sum_init = sum_init;
	A.st_rot_idx = d_rot_idx; A.st_trans_idx = d_trans_idx; A.st_job_idx = d_job_idx; A.st_job_num = d_job_num;
	A.st_njobs = n_jobs; A.st_out = d_out;
	A.projs = ctx->scratch[1].as<RbProjector>();
	A.pix = ctx->scratch[0].as<uint32_t>(); A.npix = (int) pix.size(); A.n = n;
	A.tx = d_tx; A.ty = d_ty; A.NOT = 1;
	int grid = n_jobs < ctx->num_sms * 2 ? n_jobs : ctx->num_sms * 2;
	if (grid < 1) return RB_OK;
	return launch_fine(ctx, A, grid);
}
