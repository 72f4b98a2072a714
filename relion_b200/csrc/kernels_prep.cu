// relion_b200 — image preparation on the device (SURVEY.md §8f "next" row 1): what getFourierTransformsAndCtfs
// (/root/reference/src/acc/acc_ml_optimiser_impl.h:11-1010) does per particle on one host thread with one cuFFT plan,
// batched here over the whole pool.  Branch covered: 2D images, one body, zero-masking, no helix / tomo / beam tilt / MTF.
//
//   raw image --translate (rounded old offset), x norm factor--> t          TranslateAndNormCorrect, utilities_impl.h:374-436
//   t        --centre (shift n/2), FFT / n^2, window to current size--> Fimg_nomask      normalizeAndTransformImage :438-486
//   t        --soft circular mask to the background value of its edge--> m               helper.cpp:117-253
//   m        --centre, FFT / n^2, window--> Fimg;  full-size |F|^2 per shell --> power_img, highres_Xi2    helper.h:468-540
//   CTF parameters --> Fctf on the current-size window                                    CTF::getFftwImage, src/ctf.h:184-256
// The FFTs are cuFFT (library, batched R2C); everything around them is fused into four small kernels.
#include "device_utils.cuh"
#include <cufft.h>

// tile-major index of the band-ordered staging buffers (bd_at of kernels_band.cu: 128-pixel tiles)
__device__ __forceinline__ size_t bd_at_prep(int row, int ip, int nrows) { return ((size_t) (ip >> 7) * (size_t) nrows + (size_t) row) * 128 + (size_t) (ip & 127); }

struct PrepRaw {
	const float *raw;            // [P][n][n]
	const int *shift;            // [P][2] rounded old offsets (dx, dy)
	const float *norm;           // [P]
	const float *bg;             // [P] background value of the soft edge (masked pass)
	const float *noise;          // [P][n][n] noise images (centred like raw) or nullptr: fill value of the soft mask
	int n;
	float radius, radius_p, cosine_width;
};

// translated + norm-corrected pixel (y, x) of particle p (cpu_translate2D: targets outside the box are dropped, the rest is zero)
__device__ __forceinline__ float prep_translated(const PrepRaw &A, int p, int y, int x)
{
	const int sx = x - A.shift[2 * p], sy = y - A.shift[2 * p + 1];
	if (sx < 0 || sy < 0 || sx >= A.n || sy >= A.n) return 0.f;
	return __ldg(A.raw + ((size_t) p * A.n + sy) * A.n + sx) * A.norm[p];
}

// soft-edge weight of pixel (y, x): 0 inside radius, 1 beyond radius_p, raised cosine in between
__device__ __forceinline__ float prep_edge(const PrepRaw &A, int y, int x, bool &inside)
{
	const int cx = x - A.n / 2, cy = y - A.n / 2;
	const float r = sqrtf((float) (cx * cx + cy * cy));
	inside = r < A.radius;
	if (inside) return 0.f;
	if (r > A.radius_p) return 1.f;
	return 0.5f + 0.5f * cosf((A.radius_p - r) / A.cosine_width * (float) M_PI);
}

// background value = sum(edge * t) / sum(edge) over r >= radius (softMaskBackgroundValue + getSumOnDevice, :632-652)
__global__ void __launch_bounds__(256)
k_prep_mask_bg(PrepRaw A, float *bg)
{
	__shared__ double dred[32];
	const int p = blockIdx.x;
	double s = 0., sb = 0.;
	for (int i = threadIdx.x; i < A.n * A.n; i += blockDim.x)
	{
		const int y = i / A.n, x = i - y * A.n;
		bool inside;
		const float e = prep_edge(A, y, x, inside);
		if (!inside) { s += (double) e; sb += (double) (e * prep_translated(A, p, y, x)); }
	}
	s = block_sum(s, dred);
	sb = block_sum(sb, dred);
	if (threadIdx.x == 0) bg[p] = (float) (sb / s);
}

// cuFFT input: translated (and, in the masked pass, soft-masked) image with the origin moved to pixel (0, 0)
// (runCenterFFT(forward = false): circular shift by n/2)
template <bool MASKED>
__global__ void __launch_bounds__(256)
k_prep_real(PrepRaw A, float *real)
{
	const int p = blockIdx.y, n = A.n;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * n; i += gridDim.x * blockDim.x)
	{
		const int yc = i / n, xc = i - yc * n;
		const int y = (yc + n / 2) % n, x = (xc + n / 2) % n;
		float v = prep_translated(A, p, y, x);
		if (MASKED)
		{
			bool inside;
			const float e = prep_edge(A, y, x, inside);
			const float fill = A.noise ? __ldg(A.noise + ((size_t) p * n + y) * n + x) : A.bg[p];   // do_noise: defVal = noise[texel] (helper.cpp:231-232)
			if (!inside) v = (e == 1.f) ? fill : v * (1.f - e) + fill * e;              // cosineFilter, helper.cpp:234-247
		}
		real[(size_t) p * n * n + i] = v;
	}
}

// scale by 1/n^2 and window to the current size (windowFourierTransform2, shrinking branch); then the optics group's factor image
// (beam-tilt demodulation, MTF: obs_model.cpp:528-626 as applied at acc_ml_optimiser_impl.h:535-536) when there is one
__global__ void __launch_bounds__(256)
k_prep_window(const float2 *F, float2 *out, int n, int cs, float scale, const float2 *og_factor, const RbPartMeta *metas)
{
	const int p = blockIdx.y;
	const int xf = n / 2 + 1, xo = cs / 2 + 1;
	const float2 *fac = og_factor ? og_factor + (size_t) metas[p].og * cs * xo : nullptr;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cs * xo; i += gridDim.x * blockDim.x)
	{
		const int iy = i / xo, x = i - iy * xo;
		const int ip = iy < xo ? iy : iy - cs;
		const float2 v = F[((size_t) p * n + (ip < 0 ? ip + n : ip)) * xf + x];
		float2 o = make_float2(v.x * scale, v.y * scale);
		if (fac) { const float2 c = __ldg(fac + i); o = make_float2(o.x * c.x - o.y * c.y, o.x * c.y + o.y * c.x); }
		out[(size_t) p * cs * xo + i] = o;
	}
}

// power spectrum of the full-size transform and the power beyond the current size (powerClass); one CTA per particle
__global__ void __launch_bounds__(256)
k_prep_power(const float2 *F, int n, int cs, float scale, float *power, RbPartMeta *metas)
{
	__shared__ double s_spec[1024];
	__shared__ double dred[32];
	const int p = blockIdx.x;
	const int xf = n / 2 + 1;
	for (int i = threadIdx.x; i < xf; i += blockDim.x) s_spec[i] = 0.;
	__syncthreads();
	double xi2 = 0.;
	const int res_limit = cs / 2 + 1;
	for (int i = threadIdx.x; i < n * xf; i += blockDim.x)
	{
		const int iy = i / xf, x = i - iy * xf;
		const int y = iy < xf ? iy : iy - n;
		const int ires = (int) (sqrtf((float) (x * x + y * y)) + 0.5f);
		if (ires < xf && !(x == 0 && y < 0))
		{
			const float2 v = F[(size_t) p * n * xf + i];
			const float vr = v.x * scale, vi = v.y * scale;
			const double nf = (double) (vr * vr + vi * vi);
			atomicAdd(&s_spec[ires], nf);
			if (ires >= res_limit) xi2 += nf;
		}
	}
	xi2 = block_sum(xi2, dred);
	__syncthreads();
	if (power) for (int i = threadIdx.x; i < xf; i += blockDim.x) power[(size_t) p * xf + i] = (float) s_spec[i];
	if (threadIdx.x == 0) metas[p].xi2_half = (float) (xi2 / 2.);
}

// CTF::getCTF with damping, no flips (src/ctf.h:184-256) evaluated in double like the host code, on the current-size window
__global__ void __launch_bounds__(256)
k_prep_ctf(const double *par, float *Fctf, int cs, double xs_angstrom)
{
	const int p = blockIdx.y;
	const double *q = par + (size_t) p * 9;   // K1, K2, K3, K4, K5, Axx, Axy, Ayy, scale
	const int xo = cs / 2 + 1;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cs * xo; i += gridDim.x * blockDim.x)
	{
		const int iy = i / xo, jp = i - iy * xo;
		const int ip = iy < xo ? iy : iy - cs;
		const double X = (double) jp / xs_angstrom, Y = (double) ip / xs_angstrom;
		const double u2 = X * X + Y * Y;
		const double gamma = q[0] * (q[5] * X * X + 2.0 * q[6] * X * Y + q[7] * Y * Y) + q[1] * u2 * u2 - q[4] - q[2];
		double r = -sin(gamma) * exp(q[3] * u2) * q[8];
		if (fabs(r) < 1e-8) r = r < 0 ? -1e-8 : 1e-8;                                    // ctf.h:229-232
		Fctf[(size_t) p * cs * xo + i] = (float) r;
	}
}

// Noise image of the noise-filled mask, Fourier side (makeNoiseImage, utilities_impl.h:231-371 ->
// RNDnormalDitributionComplexWithPowerModulation2D, cuda_kernels/helper.cu:90-135): every pixel of the half transform an
// independent complex normal times spectrum[ires] (zero beyond the last shell).  Counter-based generator: splitmix64 of
// (seed, pixel) -> two uniforms -> Box-Muller, independent of the launch shape.
__device__ __forceinline__ unsigned long long prep_mix64(unsigned long long z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256)
k_prep_noise_fourier(const long long *seed, const float *spectrum, const RbPartMeta *metas, int nshell, int n, float2 *F)
{
	const int p = blockIdx.y, xf = n / 2 + 1;
	const float *spec = spectrum + (size_t) metas[p].og * nshell;
	const unsigned long long key = prep_mix64((unsigned long long) seed[p]);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * xf; i += gridDim.x * blockDim.x)
	{
		const int iy = i / xf, x = i - iy * xf;
		const int y = iy >= xf ? iy - n : iy;
		const int ires = (int) rintf(sqrtf((float) (x * x + y * y)));
		float2 v = make_float2(0.f, 0.f);
		if (ires < xf && ires < nshell)
		{
			const unsigned long long h = prep_mix64(key ^ ((unsigned long long) i * 0xD1342543DE82EF95ull));
			const float u1 = ((float) (unsigned) (h >> 40) + 0.5f) * (1.f / 16777216.f);            // (0, 1)
			const float u2 = ((float) (unsigned) ((h >> 8) & 0xFFFFFF) + 0.5f) * (1.f / 16777216.f);
			const float r = sqrtf(-2.f * logf(u1)) * spec[ires];
			float sn, cs;
			sincospif(2.f * u2, &sn, &cs);
			v = make_float2(r * cs, r * sn);
		}
		F[(size_t) p * n * xf + i] = v;
	}
}

// the inverse transform leaves the origin at pixel (0, 0); the mask works on images centred at (n/2, n/2) like the particle
__global__ void __launch_bounds__(256)
k_prep_noise_centre(const float *in, float *out, int n)
{
	const int p = blockIdx.y;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * n; i += gridDim.x * blockDim.x)
	{
		const int y = i / n, x = i - y * n;
		out[(size_t) p * n * n + i] = in[((size_t) p * n + (y + n / 2) % n) * n + (x + n / 2) % n];
	}
}

// One plan per context: a cuFFT plan belongs to the device that was current when it was made, and contexts run on their own host threads.
static int get_plan(rb_ctx *ctx, int n, int batch, cufftHandle *out, cudaStream_t stream)
{
	if (ctx->prep_plan_batch == 0 || ctx->prep_plan_n != n || ctx->prep_plan_batch != batch)
	{
		if (ctx->prep_plan_batch) cufftDestroy((cufftHandle) ctx->prep_plan);
		ctx->prep_plan_batch = 0;
		int dims[2] = {n, n};
		cufftHandle h;
		cufftResult r = cufftPlanMany(&h, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, batch);
		if (r != CUFFT_SUCCESS) { rb_set_error("cufftPlanMany(%d x %d, batch %d) failed (%d)", n, n, batch, (int) r); return RB_ERR_CUDA; }
		ctx->prep_plan = (int) h; ctx->prep_plan_n = n; ctx->prep_plan_batch = batch;
	}
	cufftResult r = cufftSetStream((cufftHandle) ctx->prep_plan, stream);
	if (r != CUFFT_SUCCESS) { rb_set_error("cufftSetStream failed (%d)", (int) r); return RB_ERR_CUDA; }
	*out = (cufftHandle) ctx->prep_plan;
	return RB_OK;
}

static int get_plan_inv(rb_ctx *ctx, int n, int batch, cufftHandle *out, cudaStream_t stream)
{
	if (ctx->prep_plan_inv_batch == 0 || ctx->prep_plan_inv_n != n || ctx->prep_plan_inv_batch != batch)
	{
		if (ctx->prep_plan_inv_batch) cufftDestroy((cufftHandle) ctx->prep_plan_inv);
		ctx->prep_plan_inv_batch = 0;
		int dims[2] = {n, n};
		cufftHandle h;
		cufftResult r = cufftPlanMany(&h, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, batch);
		if (r != CUFFT_SUCCESS) { rb_set_error("cufftPlanMany(C2R %d x %d, batch %d) failed (%d)", n, n, batch, (int) r); return RB_ERR_CUDA; }
		ctx->prep_plan_inv = (int) h; ctx->prep_plan_inv_n = n; ctx->prep_plan_inv_batch = batch;
	}
	cufftResult r = cufftSetStream((cufftHandle) ctx->prep_plan_inv, stream);
	if (r != CUFFT_SUCCESS) { rb_set_error("cufftSetStream failed (%d)", (int) r); return RB_ERR_CUDA; }
	*out = (cufftHandle) ctx->prep_plan_inv;
	return RB_OK;
}

void rbk_prepare_release(rb_ctx *ctx)
{
	if (ctx->prep_plan_batch) cufftDestroy((cufftHandle) ctx->prep_plan);
	ctx->prep_plan_batch = 0;
	if (ctx->prep_plan_inv_batch) cufftDestroy((cufftHandle) ctx->prep_plan_inv);
	ctx->prep_plan_inv_batch = 0;
}

// d_raw: [P][n][n] device, d_shift [P][2], d_norm [P], d_ctfpar [P][9] (nullptr: Fctf untouched), outputs into the slot buffers
// d_seed / d_spectrum: noise-filled mask (nullptr: zero mask); d_noise_out: optional copy of the noise images (tests)
int rbk_prepare_pool(rb_ctx *ctx, PoolSlot &s, const float *d_raw, const int *d_shift, const float *d_norm, const double *d_ctfpar,
                     int n, float radius, float cosine_width, float *d_power, const long long *d_seed, const float *d_spectrum, const float2 *d_og_factor,
                     cudaStream_t stream)
{
	if (!stream) stream = ctx->stream;
	const RbModelDev &M = ctx->d_model;
	const int P = s.P, cs = M.current_size;
	const int xf = n / 2 + 1, xo = cs / 2 + 1;
	DevBuf &bReal = ctx->prep_buf[0], &bF = ctx->prep_buf[1], &bBg = ctx->prep_buf[2];
	RB_CHECK(bReal.ensure((size_t) P * n * n * 4)); RB_CHECK(bF.ensure((size_t) P * n * xf * 8)); RB_CHECK(bBg.ensure((size_t) P * 4));
	cufftHandle plan;
	RB_CHECK(get_plan(ctx, n, P, &plan, stream));
	PrepRaw A;
	A.raw = d_raw; A.shift = d_shift; A.norm = d_norm; A.bg = bBg.as<float>(); A.n = n; A.noise = nullptr;
	A.radius = radius < 0.f ? (float) n / 2.f : radius; A.cosine_width = cosine_width; A.radius_p = A.radius + cosine_width;
	const float scale = 1.f / ((float) n * (float) n);
	dim3 gr((n * n + 255) / 256 > 64 ? 64 : (n * n + 255) / 256, P), gw((cs * xo + 255) / 256 > 64 ? 64 : (cs * xo + 255) / 256, P);
	// unmasked image -> Fimg_nomask
	k_prep_real<false><<<gr, 256, 0, stream>>>(A, bReal.as<float>()); RB_LAUNCH_CHECK(ctx);
	if (cufftExecR2C(plan, bReal.as<float>(), bF.as<cufftComplex>()) != CUFFT_SUCCESS) { rb_set_error("cufftExecR2C failed (%d x %d, batch %d)", n, n, P); return RB_ERR_CUDA; }
	ctx->launches++;
	k_prep_window<<<gw, 256, 0, stream>>>(bF.as<float2>(), s.Fnomask.as<float2>(), n, cs, scale, d_og_factor, s.meta.as<RbPartMeta>()); RB_LAUNCH_CHECK(ctx);
	// masked image -> Fimg, power spectrum, highres_Xi2
	if (d_seed)
	{
		// noise images: Fourier-space normals with the model's noise spectrum -> inverse FFT (unnormalised: the coefficients are in
		// RELION's 1/N-normalised convention) -> centred, kept in prep_buf[4] until the masked pass has read them
		DevBuf &bNoise = ctx->prep_buf[4];
		RB_CHECK(bNoise.ensure((size_t) P * n * n * 4));
		cufftHandle iplan;
		RB_CHECK(get_plan_inv(ctx, n, P, &iplan, stream));
		dim3 gf((n * xf + 255) / 256 > 64 ? 64 : (n * xf + 255) / 256, P);
		k_prep_noise_fourier<<<gf, 256, 0, stream>>>(d_seed, d_spectrum, s.meta.as<RbPartMeta>(), M.nshell, n, bF.as<float2>()); RB_LAUNCH_CHECK(ctx);
		if (cufftExecC2R(iplan, bF.as<cufftComplex>(), bReal.as<float>()) != CUFFT_SUCCESS) { rb_set_error("cufftExecC2R failed (%d x %d, batch %d)", n, n, P); return RB_ERR_CUDA; }
		ctx->launches++;
		k_prep_noise_centre<<<gr, 256, 0, stream>>>(bReal.as<float>(), bNoise.as<float>(), n); RB_LAUNCH_CHECK(ctx);
		A.noise = bNoise.as<float>();
	}
	else k_prep_mask_bg<<<P, 256, 0, stream>>>(A, bBg.as<float>()); RB_LAUNCH_CHECK(ctx);
	k_prep_real<true><<<gr, 256, 0, stream>>>(A, bReal.as<float>()); RB_LAUNCH_CHECK(ctx);
	if (cufftExecR2C(plan, bReal.as<float>(), bF.as<cufftComplex>()) != CUFFT_SUCCESS) { rb_set_error("cufftExecR2C failed (%d x %d, batch %d)", n, n, P); return RB_ERR_CUDA; }
	ctx->launches++;
	k_prep_window<<<gw, 256, 0, stream>>>(bF.as<float2>(), s.Fimg.as<float2>(), n, cs, scale, d_og_factor, s.meta.as<RbPartMeta>()); RB_LAUNCH_CHECK(ctx);
	if (cs < n)
	{
		k_prep_power<<<P, 256, 0, stream>>>(bF.as<float2>(), n, cs, scale, d_power, s.meta.as<RbPartMeta>()); RB_LAUNCH_CHECK(ctx);
	}
	else if (d_power) RB_CUDA(cudaMemsetAsync(d_power, 0, (size_t) P * xf * 4, stream));
	if (d_ctfpar)
	{
		k_prep_ctf<<<gw, 256, 0, stream>>>(d_ctfpar, s.Fctf.as<float>(), cs, (double) M.ori_size * M.pixel_size); RB_LAUNCH_CHECK(ctx);
	}
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// relion_reconstruct from raw images: what Reconstructor::backprojectOneParticle (/root/reference/src/reconstructor.cpp:428-745)
// does per particle before backproject2Dto3D, for a chunk of images on the device:
//   F2D = FFT(img) / n^2, CenterFFTbySign (x (-1)^(i + j), src/fftw.h:390-403), shiftImageInFourierTransform by the particle's
//   origin offset (x e^{-2 pi i (x tx + y ty) / n}, src/fftw.cpp:874-918, double), Fctf = CTF::getFftwImage (damping, no flips),
//   F2D *= Fctf (unless the data are CTF-premultiplied), Fctf = Fctf^2, F2D(0, 0) = 0                        (:563-745)
// written straight into the band-ordered staging buffer of the posed scatter (kernels_band.cu): the images cross PCIe as
// real-space pixels (4 bytes per pixel instead of 12 per Fourier pixel) and the prepared transforms never exist in [image][pixel] order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_posed_raw_to_band(const float2 *F, int n, int count, const double *shift, const double *ctfpar, double xs_angstrom, int premultiplied,
                    const uint32_t *pix, int npix, int stride, float4 *sF)
{
	const int img = blockIdx.y, xf = n / 2 + 1;
	const double *q = ctfpar ? ctfpar + (size_t) img * 9 : nullptr;
	const double sx = shift ? -shift[2 * img] / (double) n : 0., sy = shift ? -shift[2 * img + 1] / (double) n : 0.;
	const bool do_shift = fabs(sx) >= 1e-6 || fabs(sy) >= 1e-6;                              // XMIPP_EQUAL_ACCURACY
	const float scale = 1.f / ((float) n * (float) n);
	for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < stride; ip += gridDim.x * blockDim.x)
	{
		float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
		if (ip < npix)
		{
			const uint32_t pk = __ldg(pix + ip);
			const int x = rb_pix_x(pk), y = rb_pix_y(pk);
			const int iy = y < 0 ? y + n : y;
			float2 v = __ldg(F + ((size_t) img * n + iy) * xf + x);
			const float sg = ((iy ^ x) & 1) ? -scale : scale;
			double re = (double) (v.x * sg), im = (double) (v.y * sg);
			if (do_shift)
			{
				double sn, cs;
				sincospi(2. * ((double) x * sx + (double) y * sy), &sn, &cs);
				const double r2 = cs * re - sn * im, i2 = cs * im + sn * re;
				re = r2; im = i2;
			}
			double ctf = 1.;
			if (q)
			{
				const double X = (double) x / xs_angstrom, Y = (double) y / xs_angstrom;
				const double u2 = X * X + Y * Y;
				const double gamma = q[0] * (q[5] * X * X + 2.0 * q[6] * X * Y + q[7] * Y * Y) + q[1] * u2 * u2 - q[4] - q[2];
				ctf = -sin(gamma) * exp(q[3] * u2) * q[8];
				if (fabs(ctf) < 1e-8) ctf = ctf < 0 ? -1e-8 : 1e-8;                          // ctf.h:229-232
			}
			const float cf = (float) ctf;                                                    // Fctf is an RFLOAT image in the reference; fp32 here
			float fr = (float) re, fi = (float) im;
			if (!premultiplied) { fr *= cf; fi *= cf; }
			if (x == 0 && y == 0) { fr = 0.f; fi = 0.f; }                                    // DIRECT_A2D_ELEM(F2D, 0, 0) = 0 (:745)
			o = make_float4(fr, fi, cf * cf, 0.f);
		}
		sF[bd_at_prep(img, ip, count)] = o;
	}
}

int rbk_backproject_posed_raw(rb_ctx *ctx, const RbBackprojector &bp, int n, int count, float *d_images, const double *d_shift,
                              const double *d_ctfpar, double xs_angstrom, int ctf_premultiplied, const float *d_eulers)
{
	if (count < 1) return RB_OK;
	const int xf = n / 2 + 1;
	DevBuf &bF = ctx->prep_buf[1];
	RB_CHECK(bF.ensure((size_t) count * n * xf * 8));
	cufftHandle plan;
	RB_CHECK(get_plan(ctx, n, count, &plan, ctx->stream));
	if (cufftExecR2C(plan, d_images, bF.as<cufftComplex>()) != CUFFT_SUCCESS) { rb_set_error("cufftExecR2C failed (%d x %d, batch %d)", n, n, count); return RB_ERR_CUDA; }
	ctx->launches++;
	RbPosedBandLayout L;
	RB_CHECK(rbk_posed_band_layout(ctx, n, count, &L));
	dim3 g((unsigned) ((L.stride + 255) / 256), (unsigned) count);
	k_posed_raw_to_band<<<g, 256, 0, ctx->stream>>>(bF.as<float2>(), n, count, d_shift, d_ctfpar, xs_angstrom, ctf_premultiplied, L.pix, L.npix, L.stride, L.sF);
	RB_LAUNCH_CHECK(ctx);
	return rbk_posed_band_scatter(ctx, bp, n, count, d_eulers, L);
}

// ---------------------------------------------------------------------------------------------
// rb_particles.pre_shift: Fimg and Fimg_nomask of particle p times exp(-2 pi i (x dx_p + y dy_p) / ori_size), the phase ramp every
// translation kernel applies for a sampled shift (trans = -2 pi shift / n_full, acc_ml_optimiser_impl.h:1239-1241): the particle's
// own translation of --skip_align (src/ml_optimiser.cpp:4196-4225), applied once instead of being sampled.
// ---------------------------------------------------------------------------------------------
__global__ void k_pre_shift(float2 *Fimg, float2 *Fnomask, const double *shift, int n, int ori_size, int P)
{
	const int xs = n / 2 + 1;
	const size_t per = (size_t) n * xs, tot = per * P;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < tot; i += (size_t) gridDim.x * blockDim.x)
	{
		const int p = (int) (i / per);
		const int r = (int) (i - (size_t) p * per), iy = r / xs, x = r - iy * xs, y = iy < xs ? iy : iy - n;
		const double dx = shift[2 * p], dy = shift[2 * p + 1];
		if (dx == 0. && dy == 0.) continue;
		double sn, cs;
		sincospi(-2. * ((double) x * dx + (double) y * dy) / (double) ori_size, &sn, &cs);
		const float c = (float) cs, s = (float) sn;
		float2 a = Fimg[i], b = Fnomask[i];
		Fimg[i] = make_float2(c * a.x - s * a.y, c * a.y + s * a.x);
		Fnomask[i] = make_float2(c * b.x - s * b.y, c * b.y + s * b.x);
	}
}

int rbk_pre_shift(rb_ctx *ctx, PoolSlot &s, cudaStream_t stream)
{
	const RbModelDev &M = ctx->d_model;
	k_pre_shift<<<ctx->num_sms * 4, 256, 0, stream>>>(s.Fimg.as<float2>(), s.Fnomask.as<float2>(), s.pre_shift.as<double>(), M.current_size, M.ori_size, s.P);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}
