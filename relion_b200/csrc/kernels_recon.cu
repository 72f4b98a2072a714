// relion_b200 — reconstruction of a map from the back-projection accumulators on the device (SURVEY.md §8f "next" row 2).
//
// BackProjector::reconstruct, default skip_gridding branch (/root/reference/src/backprojector.cpp:1379-1575: decenter,
// MAP regularisation of the weights :1463-1507, division by max(weight, radial average / 1000) :1513-1573), followed by
// windowToOridimRealSpace (:2530-2665: window the transform to pad*ori, CenterFFTbySign, inverse FFT, window to the box,
// normalise, softMaskOutsideMap src/mask.cpp:43-96) and Projector::griddingCorrect (src/projector.cpp:595-628).
// The accumulator never leaves the device; the 3D inverse FFT is cuFFT (library), the rest is fused into five kernels.
#include "device_utils.cuh"
#include <cufft.h>

struct ReconArgs {
	const float4 *acc;              // (re, im, weight, 0) [mdlZ][mdlY][mdlX], centred in y and z
	int mdlX, mdlY, mdlZ, initY, initZ;
	int pad;                        // pad_size of the accumulator (= mdlY)
	int r_max; float pf;
	long long max_r2, round_max_r2;
	const double *tau2; int n_tau2; double tau2_fudge, oversampling_correction; int minres_map;   // tau2 == nullptr: no MAP term
	double *radsum; double *radcnt; // [r_max]
	int padori, ori;
	const double *newweight;        // iterative gridding: Fnewweight on the pad^3 half transform (nullptr: skip_gridding branch)
	double normalise;
};

__device__ __forceinline__ int fftw_freq(int k, int n) { return k < n / 2 + 1 ? k : k - n; }

// regularised weight of FFTW-index voxel (kp, ip, jp) of the pad^3 transform (decenter + MAP term)
__device__ __forceinline__ double recon_weight(const ReconArgs &A, int kp, int ip, int jp, long long r2, float4 *val)
{
	float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
	if (r2 <= A.max_r2) v = __ldg(A.acc + ((size_t) (kp - A.initZ) * A.mdlY + (ip - A.initY)) * A.mdlX + jp);   // decenter, projector.h:249-260
	if (val) *val = v;
	double w = (double) v.z;
	if (A.tau2 && r2 < A.max_r2)
	{
		const int ires = (int) floor(sqrt((double) r2) / (double) A.pf + 0.5);
		const double t = A.tau2[ires < A.n_tau2 ? ires : A.n_tau2 - 1];
		double invtau2;
		if (t > 0.) invtau2 = 1. / (A.oversampling_correction * A.tau2_fudge * t);
		else invtau2 = w > 1e-20 ? 1. / (0.001 * w) : 0.;
		if (ires >= A.minres_map) w += invtau2;
	}
	return w;
}

// radial average of the (regularised) weights over r2 < round(r_max pf r_max pf), shells floor(r / pf)   (:1513-1540)
__global__ void __launch_bounds__(256)
k_recon_radavg(ReconArgs A)
{
	__shared__ double s_sum[1024], s_cnt[1024];
	for (int i = threadIdx.x; i < A.r_max; i += blockDim.x) { s_sum[i] = 0.; s_cnt[i] = 0.; }
	__syncthreads();
	const int xh = A.pad / 2 + 1;
	const size_t n = (size_t) A.pad * A.pad * xh;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int jp = (int) (i % xh), ii = (int) ((i / xh) % A.pad), kk = (int) (i / ((size_t) xh * A.pad));
		const int ip = fftw_freq(ii, A.pad), kp = fftw_freq(kk, A.pad);
		const long long r2 = (long long) kp * kp + (long long) ip * ip + (long long) jp * jp;
		if (r2 < A.round_max_r2)
		{
			const int ires = (int) floor(sqrt((double) r2) / (double) A.pf);
			if (ires < A.r_max)
			{
				atomicAdd(&s_sum[ires], recon_weight(A, kp, ip, jp, r2, nullptr));
				atomicAdd(&s_cnt[ires], 1.);
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < A.r_max; i += blockDim.x)
		if (s_cnt[i] > 0.) { atomicAdd(A.radsum + i, s_sum[i]); atomicAdd(A.radcnt + i, s_cnt[i]); }
}

// Fconv = data / max(weight, radavg / 1000), windowed to the pad*ori transform, sign-centred: the input of the inverse FFT
__global__ void __launch_bounds__(256)
k_recon_fin(ReconArgs A, float2 *Fin)
{
	const int xo = A.padori / 2 + 1;
	const size_t n = (size_t) A.padori * A.padori * xo;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int j = (int) (i % xo), ii = (int) ((i / xo) % A.padori), kk = (int) (i / ((size_t) xo * A.padori));
		const int ip = fftw_freq(ii, A.padori), kp = fftw_freq(kk, A.padori);        // windowFourierTransform: same (kp, ip, jp)
		float2 out = make_float2(0.f, 0.f);
		const int half = A.pad / 2;                                                    // frequencies held by the pad^3 transform
		if (j <= half && ip >= -half && ip <= half && kp >= -half && kp <= half)
		{
			const long long r2 = (long long) kp * kp + (long long) ip * ip + (long long) j * j;
			float4 v;
			const double w0 = recon_weight(A, kp, ip, j, r2, &v);
			int ires = (int) floor(sqrt((double) r2) / (double) A.pf);
			if (ires > A.r_max - 1) ires = A.r_max - 1;
			const double ra = A.radsum[ires] / (1000. * (A.radcnt[ires] > 0. ? A.radcnt[ires] : 1.));
			const double w = w0 > ra ? w0 : ra;                                         // :1547
			double re = (double) v.x, im = (double) v.y;
			if (A.newweight)
			{
				// gridding branch: the data times the iteratively determined weight (:1678-1690)
				const int kq = kp < 0 ? kp + A.pad : kp, iq = ip < 0 ? ip + A.pad : ip;
				const double nw = A.newweight[((size_t) kq * A.pad + iq) * (A.pad / 2 + 1) + j] / A.normalise;
				re *= nw; im *= nw;
			}
			else if (w != 0.) { re /= w; im /= w; }
			if ((kk ^ ii ^ j) & 1) { re = -re; im = -im; }                              // CenterFFTbySign, src/fftw.h:390-403
			out = make_float2((float) re, (float) im);
		}
		Fin[i] = out;
	}
}

// window to the box, normalise, and accumulate the background sums of softMaskOutsideMap (radius ori/2, width 3)
__global__ void __launch_bounds__(256)
k_recon_window(const float *real, float *vol, int padori, int ori, float inv_normfft, double *bg_sums)
{
	__shared__ double dred[32];
	const size_t n = (size_t) ori * ori * ori;
	const int o = padori / 2 - ori / 2;
	const float radius = (float) ori / 2.f, radius_p = radius + 3.f;
	double s = 0., sb = 0.;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (i % ori), y = (int) ((i / ori) % ori), z = (int) (i / ((size_t) ori * ori));
		const float v = real[((size_t) (z + o) * padori + (y + o)) * padori + (x + o)] * inv_normfft;
		vol[i] = v;
		const int cx = x - ori / 2, cy = y - ori / 2, cz = z - ori / 2;
		const float r = sqrtf((float) (cx * cx + cy * cy + cz * cz));
		if (r >= radius)
		{
			const float rc = r > radius_p ? 1.f : 0.5f + 0.5f * cosf((float) M_PI * (radius_p - r) / 3.f);
			s += (double) rc; sb += (double) (rc * v);
		}
	}
	s = block_sum(s, dred);
	sb = block_sum(sb, dred);
	if (threadIdx.x == 0) { atomicAdd(bg_sums, s); atomicAdd(bg_sums + 1, sb); }
}

// soft mask to the background value + gridding correction (divide by sinc^2(r / (ori pf)))
__global__ void __launch_bounds__(256)
k_recon_finish(float *vol, int ori, float pf, const double *bg_sums)
{
	const size_t n = (size_t) ori * ori * ori;
	const float radius = (float) ori / 2.f, radius_p = radius + 3.f;
	const float bg = (float) (bg_sums[1] / bg_sums[0]);
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (i % ori), y = (int) ((i / ori) % ori), z = (int) (i / ((size_t) ori * ori));
		const int cx = x - ori / 2, cy = y - ori / 2, cz = z - ori / 2;
		const float r = sqrtf((float) (cx * cx + cy * cy + cz * cz));
		float v = vol[i];
		if (r >= radius)
		{
			const float rc = r > radius_p ? 1.f : 0.5f + 0.5f * cosf((float) M_PI * (radius_p - r) / 3.f);
			v = (1.f - rc) * v + rc * bg;
		}
		if (r > 0.f)
		{
			const float rval = r / ((float) ori * pf);
			const float sinc = sinf((float) M_PI * rval) / ((float) M_PI * rval);
			v /= sinc * sinc;
		}
		vol[i] = v;
	}
}

// ---------------------------------------------------------------------------------------------
// Iterative gridding (--dont_skip_gridding): Eq. [14] of Pipe & Menon (1999) as BackProjector::reconstruct runs it
// (/root/reference/src/backprojector.cpp:1577-1700): Fnewweight = 1 inside the sphere; max_iter_preweight times
//   Fconv = Fnewweight * Fweight -> inverse FFT -> x FT of the Kaiser-Bessel blob (convoluteBlobRealSpace, :2483-2528)
//   -> forward FFT / N -> Fnewweight /= max(1e-6, |Fconv|)  for r2 < max_r2
// in double like the reference (Fnewweight "can become too large for a float"); the 3D transforms are cuFFT Z2D / D2Z.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_grid_init(ReconArgs A, double *Fw, double *Fnw)
{
	const int xh = A.pad / 2 + 1;
	const size_t n = (size_t) A.pad * A.pad * xh;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int jp = (int) (i % xh), ii = (int) ((i / xh) % A.pad), kk = (int) (i / ((size_t) xh * A.pad));
		const int ip = fftw_freq(ii, A.pad), kp = fftw_freq(kk, A.pad);
		const long long r2 = (long long) kp * kp + (long long) ip * ip + (long long) jp * jp;
		Fw[i] = recon_weight(A, kp, ip, jp, r2, nullptr) / A.normalise;                   // :1583-1587
		Fnw[i] = r2 < A.max_r2 ? 1. : 0.;                                                // :1596-1606
	}
}

__global__ void __launch_bounds__(256)
k_grid_mul(const double *Fw, const double *Fnw, double2 *Fc, size_t n)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
		Fc[i] = make_double2(Fnw[i] * Fw[i], 0.);
}

// real space: multiply with tab_ftblob(rval) / tab_ftblob(0), rval = |k| / (ori pf), k wrapped at pad / 2 (:2510-2524)
__global__ void __launch_bounds__(256)
k_grid_blob(double *M, int pad, const double *tab, int nr, double sampling, double ori_pf)
{
	const size_t n = (size_t) pad * pad * pad;
	const int h = pad / 2;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int j = (int) (i % pad), ii = (int) ((i / pad) % pad), k = (int) (i / ((size_t) pad * pad));
		const int jp = j < h ? j : j - pad, ip = ii < h ? ii : ii - pad, kp = k < h ? k : k - pad;
		const double rval = sqrt((double) ((long long) kp * kp + (long long) ip * ip + (long long) jp * jp)) / ori_pf;
		const int idx = (int) (rval / sampling);
		M[i] *= idx >= nr ? 0. : tab[idx];
	}
}

__global__ void __launch_bounds__(256)
k_grid_div(const double2 *Fc, double *Fnw, int pad, long long max_r2, double inv_n)
{
	const int xh = pad / 2 + 1;
	const size_t n = (size_t) pad * pad * xh;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int jp = (int) (i % xh), ii = (int) ((i / xh) % pad), kk = (int) (i / ((size_t) xh * pad));
		const int ip = fftw_freq(ii, pad), kp = fftw_freq(kk, pad);
		const long long r2 = (long long) kp * kp + (long long) ip * ip + (long long) jp * jp;
		if (r2 < max_r2)
		{
			const double2 c = Fc[i];
			const double w = fmax(1e-6, sqrt(c.x * c.x + c.y * c.y) * inv_n);             // :1636-1645
			Fnw[i] /= w;
		}
	}
}

// Fnewweight of accumulator `A` after max_iter iterations, left in grid_buf[1]
static int recon_gridding(rb_ctx *ctx, ReconArgs &A, int max_iter, double blob_radius, double blob_alpha)
{
	const int pad = A.pad, xh = pad / 2 + 1;
	const size_t nh = (size_t) pad * pad * xh, nr3 = (size_t) pad * pad * pad;
	DevBuf &bFw = ctx->grid_buf[0], &bFnw = ctx->grid_buf[1], &bFc = ctx->grid_buf[2], &bM = ctx->grid_buf[3], &bTab = ctx->grid_buf[4];
	RB_CHECK(bFw.ensure(nh * 8)); RB_CHECK(bFnw.ensure(nh * 8)); RB_CHECK(bFc.ensure(nh * 16)); RB_CHECK(bM.ensure(nr3 * 8));
	// tab_ftblob (src/tabfuncs.cpp:95-121, src/funcs.cpp:244-251): order 0, radius 2 * blob_radius, 10000 entries over [0, 0.5)
	const int nr = 10000;
	const double sampling = 0.5 / nr, a = 2. * blob_radius;
	std::vector<double> tab(nr);
	for (int i = 0; i < nr; i++)
	{
		const double arg = 2. * M_PI * a * (i * sampling);
		double sigma = sqrt(fabs(blob_alpha * blob_alpha - arg * arg));
		if (sigma == 0.) sigma = 1e-300;
		const double b = arg > blob_alpha ? sqrt(2. / (M_PI * sigma)) * (sin(sigma) / sigma - cos(sigma))
		                                  : sqrt(2. / (M_PI * sigma)) * (cosh(sigma) - sinh(sigma) / sigma);
		tab[i] = b / pow(sigma, 1.5);
	}
	for (int i = nr - 1; i >= 0; i--) tab[i] /= tab[0];
	RB_CHECK(bTab.ensure(nr * 8));
	RB_CUDA(cudaMemcpyAsync(bTab.p, tab.data(), nr * 8, cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	cufftHandle pinv, pfwd;
	if (cufftPlan3d(&pinv, pad, pad, pad, CUFFT_Z2D) != CUFFT_SUCCESS) { rb_set_error("cufftPlan3d(Z2D %d^3) failed", pad); return RB_ERR_CUDA; }
	if (cufftPlan3d(&pfwd, pad, pad, pad, CUFFT_D2Z) != CUFFT_SUCCESS) { cufftDestroy(pinv); rb_set_error("cufftPlan3d(D2Z %d^3) failed", pad); return RB_ERR_CUDA; }
	cufftSetStream(pinv, ctx->stream); cufftSetStream(pfwd, ctx->stream);
	const int g = ctx->num_sms * 8;
	k_grid_init<<<g, 256, 0, ctx->stream>>>(A, bFw.as<double>(), bFnw.as<double>());
	int rc = RB_OK;
	for (int it = 0; it < max_iter && rc == RB_OK; it++)
	{
		k_grid_mul<<<g, 256, 0, ctx->stream>>>(bFw.as<double>(), bFnw.as<double>(), bFc.as<double2>(), nh);
		if (cufftExecZ2D(pinv, bFc.as<cufftDoubleComplex>(), bM.as<double>()) != CUFFT_SUCCESS) { rb_set_error("cufftExecZ2D failed"); rc = RB_ERR_CUDA; break; }
		k_grid_blob<<<g, 256, 0, ctx->stream>>>(bM.as<double>(), pad, bTab.as<double>(), nr, sampling, (double) A.ori * (double) A.pf);
		if (cufftExecD2Z(pfwd, bM.as<double>(), bFc.as<cufftDoubleComplex>()) != CUFFT_SUCCESS) { rb_set_error("cufftExecD2Z failed"); rc = RB_ERR_CUDA; break; }
		k_grid_div<<<g, 256, 0, ctx->stream>>>(bFc.as<double2>(), bFnw.as<double>(), pad, A.max_r2, 1. / (double) nr3);
		ctx->launches += 5;
	}
	if (rc == RB_OK && cudaGetLastError() != cudaSuccess) { rb_set_error("gridding kernels failed"); rc = RB_ERR_CUDA; }
	cudaStreamSynchronize(ctx->stream);
	cufftDestroy(pinv); cufftDestroy(pfwd);
	bFw.release(); bFc.release(); bM.release();
	return rc;
}

// d_vol_out: [ori][ori][ori] device floats
int rbk_reconstruct(rb_ctx *ctx, const RbBackprojector &bp, int ori, const double *d_tau2, int n_tau2, double tau2_fudge, int minres_map,
                    float *d_vol_out, int max_iter_preweight, double normalise)
{
	ReconArgs A;
	memset(&A, 0, sizeof(A));
	A.acc = bp.vol; A.mdlX = bp.mdlX; A.mdlY = bp.mdlY; A.mdlZ = bp.mdlZ; A.initY = bp.mdlInitY; A.initZ = bp.mdlInitZ;
	A.pad = bp.mdlY; A.r_max = bp.maxR; A.pf = bp.padding_factor;
	const long long rr = (long long) floor((double) bp.maxR * (double) bp.padding_factor + 0.5);
	A.max_r2 = rr * rr;
	A.round_max_r2 = (long long) floor((double) bp.maxR * bp.padding_factor * bp.maxR * bp.padding_factor + 0.5);
	A.tau2 = d_tau2; A.n_tau2 = n_tau2; A.tau2_fudge = tau2_fudge; A.minres_map = minres_map;
	A.oversampling_correction = (double) bp.padding_factor * bp.padding_factor * bp.padding_factor;
	int padori = (int) floor((double) bp.padding_factor * ori + 0.5);
	padori += padori % 2;
	A.padori = padori; A.ori = ori;
	if (A.r_max > 1024) { rb_set_error("rb_reconstruct: r_max %d too large", A.r_max); return RB_ERR_ARG; }
	const size_t nfin = (size_t) padori * padori * (padori / 2 + 1), nreal = (size_t) padori * padori * padori;
	DevBuf &bFin = ctx->recon_buf[0], &bReal = ctx->recon_buf[1], &bRad = ctx->recon_buf[2];
	RB_CHECK(bFin.ensure(nfin * 8)); RB_CHECK(bReal.ensure(nreal * 4)); RB_CHECK(bRad.ensure((size_t) (2 * 1024 + 2) * 8));
	RB_CUDA(cudaMemsetAsync(bRad.p, 0, (size_t) (2 * 1024 + 2) * 8, ctx->stream));
	A.radsum = bRad.as<double>(); A.radcnt = A.radsum + 1024;
	double *bg_sums = A.radsum + 2048;
	A.normalise = normalise > 0. ? normalise : 1.;
	if (max_iter_preweight > 0)
	{
		RB_CHECK(recon_gridding(ctx, A, max_iter_preweight, 1.9, 15.));                 // BackProjector's default blob (src/backprojector.h:84-86)
		A.newweight = ctx->grid_buf[1].as<double>();
	}
	k_recon_radavg<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(A); RB_LAUNCH_CHECK(ctx);
	k_recon_fin<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(A, bFin.as<float2>()); RB_LAUNCH_CHECK(ctx);
	cufftHandle plan;
	if (cufftPlan3d(&plan, padori, padori, padori, CUFFT_C2R) != CUFFT_SUCCESS) { rb_set_error("cufftPlan3d(%d^3) failed", padori); return RB_ERR_CUDA; }
	cufftSetStream(plan, ctx->stream);
	const cufftResult r = cufftExecC2R(plan, bFin.as<cufftComplex>(), bReal.as<float>());
	ctx->launches++;
	if (r != CUFFT_SUCCESS) { cufftDestroy(plan); rb_set_error("cufftExecC2R failed (%d)", (int) r); return RB_ERR_CUDA; }
	const float inv_normfft = 1.f / (bp.padding_factor * bp.padding_factor * bp.padding_factor * (float) ori);   // ref_dim 3, data_dim 2 (:2583-2586)
	k_recon_window<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(bReal.as<float>(), d_vol_out, padori, ori, inv_normfft, bg_sums); RB_LAUNCH_CHECK(ctx);
	k_recon_finish<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(d_vol_out, ori, bp.padding_factor, bg_sums); RB_LAUNCH_CHECK(ctx);
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	cufftDestroy(plan);
	ctx->grid_buf[1].release();
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// BackProjector::updateSSNRarrays (/root/reference/src/backprojector.cpp:1041-1204): the spectra the M-step needs next to
// the map - sigma2 (shell average of the inverse noise power in the reconstruction), tau2 (from the FSC between the
// half-maps when asked), data_vs_prior and fourier_coverage.  Two passes over the weight plane of the accumulator, shell
// sums in fp64; the handful of per-shell formulas in between run on the host.
// ---------------------------------------------------------------------------------------------
struct SsnrArgs {
	const float4 *acc; int mdlX, mdlY, mdlZ, initY, initZ;
	long long max_r2; float pf; int nshell;
	double oc, tau2_fudge;
	const double *tau2, *avgctf2;   // pass 2
	double *sum_a, *sum_b, *cnt;    // pass 1: sum of oc * w, -, count; pass 2: sum of w / invtau2, coverage count, count
};

template <int PASS>
__global__ void __launch_bounds__(256)
k_ssnr_pass(SsnrArgs A)
{
	__shared__ double s_a[1024], s_b[1024], s_c[1024];
	for (int i = threadIdx.x; i < A.nshell; i += blockDim.x) { s_a[i] = 0.; s_b[i] = 0.; s_c[i] = 0.; }
	__syncthreads();
	const size_t n = (size_t) A.mdlZ * A.mdlY * A.mdlX;
	for (size_t idx = blockIdx.x * (size_t) blockDim.x + threadIdx.x; idx < n; idx += (size_t) gridDim.x * blockDim.x)
	{
		const int j = (int) (idx % A.mdlX), i = (int) ((idx / A.mdlX) % A.mdlY) + A.initY;
		const int k = A.mdlZ > 1 ? (int) (idx / ((size_t) A.mdlX * A.mdlY)) + A.initZ : 0;
		const long long r2 = (long long) k * k + (long long) i * i + (long long) j * j;
		if (r2 >= A.max_r2) continue;
		const int ires = (int) floor(sqrt((double) r2) / (double) A.pf + 0.5);
		if (ires >= A.nshell) continue;
		const double w = (double) __ldg(A.acc + idx).z;
		if (PASS == 1) atomicAdd(&s_a[ires], A.oc * w);
		else
		{
			const double t = A.tau2[ires];
			double invtau2;
			if (t > 0.)
			{
				invtau2 = 1. / (A.oc * A.tau2_fudge * t);
				if (A.avgctf2 && A.avgctf2[ires] > 0.) invtau2 *= 1. / A.avgctf2[ires];
			}
			else invtau2 = 1. / (0.001 * w);                  // tau2 == 0: "use small value instead" (:1153-1157)
			const double ratio = w / invtau2;                  // w == 0 with tau2 == 0: 0 / inf = 0
			if (ratio == ratio) atomicAdd(&s_a[ires], ratio);
			if (ratio >= 1.) atomicAdd(&s_b[ires], 1.);
		}
		atomicAdd(&s_c[ires], 1.);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < A.nshell; i += blockDim.x)
		if (s_c[i] > 0.) { atomicAdd(A.sum_a + i, s_a[i]); atomicAdd(A.sum_b + i, s_b[i]); atomicAdd(A.cnt + i, s_c[i]); }
}

int rbk_update_ssnr(rb_ctx *ctx, const RbBackprojector &bp, bool is_2d, int ori, double tau2_fudge, double *tau2_io, double *sigma2_out,
                    double *dvp_out, double *cov_out, const double *fsc, const double *avgctf2, bool update_with_fsc, bool whole)
{
	const int ns = ori / 2 + 1;
	if (ns > 1024) { rb_set_error("rb_update_ssnr: ori_size %d too large", ori); return RB_ERR_ARG; }
	if (update_with_fsc && !fsc) { rb_set_error("rb_update_ssnr: update_tau2_with_fsc needs an fsc spectrum"); return RB_ERR_ARG; }
	SsnrArgs A;
	memset(&A, 0, sizeof(A));
	A.acc = bp.vol; A.mdlX = bp.mdlX; A.mdlY = bp.mdlY; A.mdlZ = is_2d ? 1 : bp.mdlZ; A.initY = bp.mdlInitY; A.initZ = bp.mdlInitZ;
	const long long rr = (long long) floor((double) bp.maxR * (double) bp.padding_factor + 0.5);
	A.max_r2 = rr * rr; A.pf = bp.padding_factor; A.nshell = ns;
	A.oc = is_2d ? (double) bp.padding_factor * bp.padding_factor : (double) bp.padding_factor * bp.padding_factor * bp.padding_factor;
	A.tau2_fudge = tau2_fudge;
	DevBuf &b = ctx->recon_buf[2];
	RB_CHECK(b.ensure((size_t) 5 * 1024 * 8));
	double *d = b.as<double>();
	A.sum_a = d; A.sum_b = d + 1024; A.cnt = d + 2048;
	double *d_tau2 = d + 3072, *d_ctf2 = d + 4096;
	std::vector<double> h((size_t) 3 * 1024);
	RB_CUDA(cudaMemsetAsync(d, 0, (size_t) 3 * 1024 * 8, ctx->stream));
	k_ssnr_pass<1><<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(A); RB_LAUNCH_CHECK(ctx);
	RB_CUDA(cudaMemcpyAsync(h.data(), d, (size_t) 3 * 1024 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	std::vector<double> sigma2(ns), tau2(tau2_io, tau2_io + ns), dvp(ns, 0.), counter(h.begin() + 2048, h.begin() + 2048 + ns);
	for (int i = 0; i < ns; i++)
	{
		const double s = h[i];
		if (s > 1e-20) sigma2[i] = counter[i] / s;
		else if (s == 0.) sigma2[i] = 0.;
		else { rb_set_error("rb_update_ssnr: unexpectedly small, yet non-zero sigma2 value %g in shell %d", s, i); return RB_ERR_ARG; }
	}
	if (update_with_fsc)
		for (int i = 0; i < ns; i++)
		{
			double f = std::max(0.001, fsc[i]);
			if (whole) f = sqrt(2. * f / (f + 1.));
			f = std::min(0.999, f);
			const double ssnr = f / (1. - f) * tau2_fudge;
			tau2[i] = ssnr * sigma2[i];
			dvp[i] = ssnr;
		}
	for (int i = 0; i < ns; i++)
		if (tau2[i] < 0.) { rb_set_error("rb_update_ssnr: negative value %g in the tau2 spectrum (shell %d)", tau2[i], i); return RB_ERR_ARG; }
	RB_CUDA(cudaMemcpyAsync(d_tau2, tau2.data(), (size_t) ns * 8, cudaMemcpyHostToDevice, ctx->stream));
	if (avgctf2) RB_CUDA(cudaMemcpyAsync(d_ctf2, avgctf2, (size_t) ns * 8, cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaMemsetAsync(d, 0, (size_t) 3 * 1024 * 8, ctx->stream));
	A.tau2 = d_tau2; A.avgctf2 = avgctf2 ? d_ctf2 : nullptr;
	k_ssnr_pass<2><<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(A); RB_LAUNCH_CHECK(ctx);
	RB_CUDA(cudaMemcpyAsync(h.data(), d, (size_t) 3 * 1024 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	for (int i = 0; i < ns; i++)
	{
		const double c = h[2048 + i];
		if (!update_with_fsc) dvp[i] = i > bp.maxR ? 0. : (c < 0.001 ? 999. : h[i] / c);
		cov_out[i] = c > 0. ? h[1024 + i] / c : h[1024 + i];
		tau2_io[i] = tau2[i]; sigma2_out[i] = sigma2[i]; dvp_out[i] = dvp[i];
	}
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// The inverse direction (SURVEY.md §8f "next" row 3): Projector::computeFourierTransformMap
// (/root/reference/src/projector.cpp:116-592) for a 3D reference used with 2D images: gridding correction (divide by
// sinc^2, :595-628), zero-padding to pad*ori, forward FFT (normalised), CenterFFTbySign, window to the projector's
// (2 (round(pf r_max) + 1) + 1)^3 half volume with everything beyond round(pf r_max) zeroed, scaled by normfft = pf^3 ori
// (:147-163), radial power spectrum (:497-545).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ftmap_pad(const float *vol, float *real, int ori, int padori, float pf)
{
	const size_t n = (size_t) padori * padori * padori;
	const int o = padori / 2 - ori / 2;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (i % padori) - o, y = (int) ((i / padori) % padori) - o, z = (int) (i / ((size_t) padori * padori)) - o;
		float v = 0.f;
		if (x >= 0 && x < ori && y >= 0 && y < ori && z >= 0 && z < ori)
		{
			v = vol[((size_t) z * ori + y) * ori + x];
			const int cx = x - ori / 2, cy = y - ori / 2, cz = z - ori / 2;
			const float r = sqrtf((float) (cx * cx + cy * cy + cz * cz));
			if (r > 0.f)
			{
				const float rval = r / ((float) ori * pf);
				const float sinc = sinf((float) M_PI * rval) / ((float) M_PI * rval);
				v /= sinc * sinc;
			}
		}
		real[i] = v;
	}
}

__global__ void __launch_bounds__(256)
k_ftmap_window(const float2 *F, float2 *data, int padori, int pad, long long max_r2, float scale, float pf, double *pow_sum, double *pow_cnt, int nshell)
{
	__shared__ double s_sum[1024], s_cnt[1024];
	for (int i = threadIdx.x; i < nshell; i += blockDim.x) { s_sum[i] = 0.; s_cnt[i] = 0.; }
	__syncthreads();
	const int xd = pad / 2 + 1, h = (pad - 1) / 2, xf = padori / 2 + 1;
	const size_t n = (size_t) pad * pad * xd;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int jp = (int) (i % xd), ip = (int) ((i / xd) % pad) - h, kp = (int) (i / ((size_t) xd * pad)) - h;
		const long long r2 = (long long) kp * kp + (long long) ip * ip + (long long) jp * jp;
		float2 out = make_float2(0.f, 0.f);
		const bool in_fft = kp >= -(padori / 2 - 1) && kp <= padori / 2 && ip >= -(padori / 2 - 1) && ip <= padori / 2 && jp <= padori / 2;
		if (r2 <= max_r2 && in_fft)
		{
			const int k = kp < 0 ? kp + padori : kp, ii = ip < 0 ? ip + padori : ip;
			const float2 v = F[((size_t) k * padori + ii) * xf + jp];
			const float sg = ((k ^ ii ^ jp) & 1) ? -scale : scale;                       // CenterFFTbySign, 1/N and normfft folded in
			out = make_float2(v.x * sg, v.y * sg);
			const int ires = (int) floor(sqrt((double) r2) / (double) pf + 0.5);
			if (ires < nshell) { atomicAdd(&s_sum[ires], 0.5 * ((double) out.x * out.x + (double) out.y * out.y)); atomicAdd(&s_cnt[ires], 1.); }
		}
		data[i] = out;
	}
	__syncthreads();
	for (int i = threadIdx.x; i < nshell; i += blockDim.x)
		if (s_cnt[i] > 0.) { atomicAdd(pow_sum + i, s_sum[i]); atomicAdd(pow_cnt + i, s_cnt[i]); }
}

// d_vol: [ori]^3 device floats; d_data: the projector's compact (re, im) volume [pad][pad][pad/2+1]; h_power: [ori/2+1] host doubles or nullptr
int rbk_ftmap(rb_ctx *ctx, const float *d_vol, int ori, int r_max, float pf, float2 *d_data, int pad, double *h_power)
{
	int padori = (int) floor((double) pf * ori + 0.5);
	padori += padori % 2;
	const float pfe = (float) padori / (float) ori;                                   // re-calculated padding factor (:133)
	const size_t nreal = (size_t) padori * padori * padori, nF = (size_t) padori * padori * (padori / 2 + 1);
	DevBuf &bF = ctx->recon_buf[0], &bReal = ctx->recon_buf[1], &bRad = ctx->recon_buf[2];
	RB_CHECK(bF.ensure(nF * 8)); RB_CHECK(bReal.ensure(nreal * 4)); RB_CHECK(bRad.ensure((size_t) (2 * 1024 + 2) * 8));
	RB_CUDA(cudaMemsetAsync(bRad.p, 0, (size_t) (2 * 1024 + 2) * 8, ctx->stream));
	k_ftmap_pad<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(d_vol, bReal.as<float>(), ori, padori, pfe); RB_LAUNCH_CHECK(ctx);
	cufftHandle plan;
	if (cufftPlan3d(&plan, padori, padori, padori, CUFFT_R2C) != CUFFT_SUCCESS) { rb_set_error("cufftPlan3d(%d^3) failed", padori); return RB_ERR_CUDA; }
	cufftSetStream(plan, ctx->stream);
	const cufftResult r = cufftExecR2C(plan, bReal.as<float>(), bF.as<cufftComplex>());
	ctx->launches++;
	if (r != CUFFT_SUCCESS) { cufftDestroy(plan); rb_set_error("cufftExecR2C failed (%d)", (int) r); return RB_ERR_CUDA; }
	const long long rr = (long long) floor((double) r_max * pfe + 0.5);
	const float scale = (pfe * pfe * pfe * (float) ori) / ((float) padori * (float) padori * (float) padori);   // normfft / N
	const int nshell = ori / 2 + 1;
	double *pow_sum = bRad.as<double>(), *pow_cnt = pow_sum + 1024;
	k_ftmap_window<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(bF.as<float2>(), d_data, padori, pad, rr * rr, scale, pfe, pow_sum, pow_cnt, nshell);
	RB_LAUNCH_CHECK(ctx);
	if (h_power)
	{
		std::vector<double> ps(nshell), pc(nshell);
		RB_CUDA(cudaMemcpyAsync(ps.data(), pow_sum, nshell * 8, cudaMemcpyDeviceToHost, ctx->stream));
		RB_CUDA(cudaMemcpyAsync(pc.data(), pow_cnt, nshell * 8, cudaMemcpyDeviceToHost, ctx->stream));
		RB_CUDA(cudaStreamSynchronize(ctx->stream));
		for (int i = 0; i < nshell; i++) h_power[i] = pc[i] < 1. ? 0. : ps[i] / pc[i];                              // :563-568
	}
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	cufftDestroy(plan);
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// BackProjector::symmetrise (/root/reference/src/backprojector.cpp:2136-2146) on the interleaved accumulator:
// enforceHermitianSymmetry (:2148-2165) on the x = 0 plane, then applyPointGroupSymmetry (:2324-2480): every voxel inside
// round(r_max pf) receives the trilinearly interpolated values of its nsym symmetry mates (Hermitian fold for x < 0).
// applyHelicalSymmetry (:2167-2322) is the same sweep with rotations about Z and, per operator, a phase ramp along z
// (`zshift[m]`, cycles per voxel of z) applied to the interpolated data term (:2284-2296); it runs between the two.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bp_hermitian(float4 *acc, int mdlX, int mdlY, int mdlZ, int initY, int initZ)
{
	const int finY = mdlY - 1 + initY, finZ = mdlZ - 1 + initZ;
	const int n = mdlZ * mdlY;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const int iz = i / mdlY + initZ, iy = i % mdlY + initY;
		const int starty = iz < 0 ? 0 : 1;
		if (iy < starty || iy > finY || -iz < initZ || -iz > finZ || -iy < initY) continue;
		float4 *a = acc + ((size_t) (iz - initZ) * mdlY + (iy - initY)) * mdlX;
		float4 *b = acc + ((size_t) (-iz - initZ) * mdlY + (-iy - initY)) * mdlX;
		const float4 va = *a, vb = *b;
		const float sr = va.x + vb.x, si = va.y - vb.y, sw = va.z + vb.z;             // fsum = a + conj(b)
		*a = make_float4(sr, si, sw, 0.f);
		*b = make_float4(sr, -si, sw, 0.f);
	}
}

__global__ void __launch_bounds__(256)
k_bp_pointgroup(const float4 *acc, float4 *out, int mdlX, int mdlY, int mdlZ, int initY, int initZ, long long rmax2, const float *R, const float *zshift, int nsym)
{
	const size_t n = (size_t) mdlX * mdlY * mdlZ;
	const size_t sy = mdlX, sz = (size_t) mdlX * mdlY;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const int j = (int) (i % mdlX), iy = (int) ((i / mdlX) % mdlY) + initY, iz = (int) (i / sz) + initZ;
		float4 s = acc[i];
		const long long r2 = (long long) j * j + (long long) iy * iy + (long long) iz * iz;
		if (r2 <= rmax2)
		{
			const float x = (float) j, y = (float) iy, z = (float) iz;
			for (int m = 0; m < nsym; m++)
			{
				const float *r = R + 9 * m;
				float xp = x * r[0] + y * r[1] + z * r[2];
				float yp = x * r[3] + y * r[4] + z * r[5];
				float zp = x * r[6] + y * r[7] + z * r[8];
				const bool neg = xp < 0.f;
				if (neg) { xp = -xp; yp = -yp; zp = -zp; }
				const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
				const float fx = xp - fx0, fy = yp - fy0, fz = zp - fz0;
				const int x0 = (int) fx0, y0 = (int) fy0 - initY, z0 = (int) fz0 - initZ;
				if (x0 < 0 || y0 < 0 || z0 < 0 || x0 + 1 >= mdlX || y0 + 1 >= mdlY || z0 + 1 >= mdlZ) continue;
				const float4 *b = acc + (size_t) z0 * sz + (size_t) y0 * sy + x0;
				const float4 d000 = __ldg(b), d001 = __ldg(b + 1), d010 = __ldg(b + sy), d011 = __ldg(b + sy + 1);
				const float4 d100 = __ldg(b + sz), d101 = __ldg(b + sz + 1), d110 = __ldg(b + sz + sy), d111 = __ldg(b + sz + sy + 1);
#define RB_LERP3(c) ({ \
	const float dx00 = d000.c + (d001.c - d000.c) * fx, dx01 = d100.c + (d101.c - d100.c) * fx; \
	const float dx10 = d010.c + (d011.c - d010.c) * fx, dx11 = d110.c + (d111.c - d110.c) * fx; \
	const float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy; \
	dxy0 + (dxy1 - dxy0) * fz; })
				float vr = RB_LERP3(x), vi = RB_LERP3(y);
				const float vw = RB_LERP3(z);
#undef RB_LERP3
				if (neg) vi = -vi;
				if (zshift && zshift[m] != 0.f)
				{
					float sn, cs;
					sincospif(2.f * z * zshift[m], &sn, &cs);
					const float tr = cs * vr - sn * vi;
					vi = cs * vi + sn * vr;
					vr = tr;
				}
				s.x += vr; s.y += vi; s.z += vw;
			}
		}
		out[i] = s;
	}
}

int rbk_bp_symmetrise(rb_ctx *ctx, const RbBackprojector &bp, DevBuf &tmp, const float *d_R, int nsym, const float *d_hR, const float *d_hz, int nhel)
{
	k_bp_hermitian<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(bp.vol, bp.mdlX, bp.mdlY, bp.mdlZ, bp.mdlInitY, bp.mdlInitZ);
	RB_LAUNCH_CHECK(ctx);
	const size_t n = (size_t) bp.mdlX * bp.mdlY * bp.mdlZ;
	const long long rr = (long long) floor((double) bp.maxR * (double) bp.padding_factor + 0.5);
	// the helical operators first: the point group then acts on their sum (:2143-2145)
	for (int pass = 0; pass < 2; pass++)
	{
		const int nops = pass == 0 ? nhel : nsym;
		if (nops <= 0) continue;
		RB_CHECK(tmp.ensure(n * sizeof(float4)));
		k_bp_pointgroup<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(bp.vol, tmp.as<float4>(), bp.mdlX, bp.mdlY, bp.mdlZ, bp.mdlInitY, bp.mdlInitZ,
			rr * rr, pass == 0 ? d_hR : d_R, pass == 0 ? d_hz : nullptr, nops);
		RB_LAUNCH_CHECK(ctx);
		RB_CUDA(cudaMemcpyAsync(bp.vol, tmp.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
	}
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}
