// relion_b200 — small kernels: Euler matrices, volume upload conversion, accumulator read-back,
// stand-alone projection, posed back-projection.
#include "device_utils.cuh"

// cuda_kernel_make_eulers_3D<invert=true,doL,doR> (helper.cuh:713-840): fp32 degrees -> radians -> sincosf -> ZYZ matrix A,
// B = L (A R) in fp32 when the pool carries MBL / MBR, written inverted: the transpose without L (also with R alone, as the
// reference does), the adjugate over the determinant with L ("this could have anisotropy, so inverse neq transpose").
struct CoarseLR { int doL, doR; float L[9], R[9]; };

__global__ void k_make_coarse_eulers(const double *rot, const double *tilt, const double *psi, int n_dir, int n_psi, CoarseLR lr, float *eulers)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_dir * n_psi) return;
	const int d = i / n_psi, q = i - d * n_psi;
	const float a = (float) rot[d] * (float) 3.14159265358979323846 / 180.0f;     // XFLOAT alphas / betas / gammas of AccProjectorPlan
	const float b = (float) tilt[d] * (float) 3.14159265358979323846 / 180.0f;
	const float g = (float) psi[q] * (float) 3.14159265358979323846 / 180.0f;
	float sa, ca, sb, cb, sg, cg;
	sincosf(a, &sa, &ca); sincosf(b, &sb, &cb); sincosf(g, &sg, &cg);
	const float cc = cb * ca, cs = cb * sa, sc = sb * ca, ss = sb * sa;
	float A[9], B[9];
	A[0] = cg * cc - sg * sa;  A[1] = cg * cs + sg * ca;  A[2] = -cg * sb;
	A[3] = -sg * cc - cg * sa; A[4] = -sg * cs + cg * ca; A[5] = sg * sb;
	A[6] = sc;                 A[7] = ss;                 A[8] = cb;
	if (lr.doR)
	{
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++)
			{
				float v = 0.f;
				for (int k = 0; k < 3; k++) v += A[r * 3 + k] * lr.R[k * 3 + c];
				B[r * 3 + c] = v;
			}
	}
	else
		for (int k = 0; k < 9; k++) B[k] = A[k];
	float *e = eulers + (size_t) i * 9;
	if (lr.doL)
	{
		for (int k = 0; k < 9; k++) A[k] = B[k];
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++)
			{
				float v = 0.f;
				for (int k = 0; k < 3; k++) v += lr.L[r * 3 + k] * A[k * 3 + c];
				B[r * 3 + c] = v;
			}
		const float det = B[0] * (B[4] * B[8] - B[7] * B[5]) - B[1] * (B[3] * B[8] - B[6] * B[5]) + B[2] * (B[3] * B[7] - B[6] * B[4]);
		e[0] = (B[4] * B[8] - B[7] * B[5]) / det; e[1] = (B[7] * B[2] - B[1] * B[8]) / det; e[2] = (B[1] * B[5] - B[4] * B[2]) / det;
		e[3] = (B[5] * B[6] - B[8] * B[3]) / det; e[4] = (B[8] * B[0] - B[2] * B[6]) / det; e[5] = (B[2] * B[3] - B[5] * B[0]) / det;
		e[6] = (B[3] * B[7] - B[6] * B[4]) / det; e[7] = (B[6] * B[1] - B[0] * B[7]) / det; e[8] = (B[0] * B[4] - B[3] * B[1]) / det;
	}
	else
	{
		e[0] = B[0]; e[1] = B[3]; e[2] = B[6];
		e[3] = B[1]; e[4] = B[4]; e[5] = B[7];
		e[6] = B[2]; e[7] = B[5]; e[8] = B[8];
	}
}

int rbk_make_coarse_eulers(rb_ctx *ctx, const double *d_rot, const double *d_tilt, const double *d_psi, int n_dir, int n_psi, const RbLR &lr,
                           float *d_eulers)
{
	const int n = n_dir * n_psi;
	CoarseLR c;
	c.doL = lr.doL; c.doR = lr.doR;
	for (int i = 0; i < 9; i++) { c.L[i] = (float) lr.L[i]; c.R[i] = (float) lr.R[i]; }
	k_make_coarse_eulers<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_rot, d_tilt, d_psi, n_dir, n_psi, c, d_eulers);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// fp64 -> fp32 cast of the reference volume (AccProjector::initMdl, acc_projector_impl.h:196-312)
__global__ void k_convert_volume(const double *in, float2 *out, size_t n)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
		out[i] = make_float2((float) in[2 * i], (float) in[2 * i + 1]);
}

int rbk_convert_volume(rb_ctx *ctx, const double *d_in, float2 *d_out, size_t n)
{
	k_convert_volume<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(d_in, d_out, n);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// Neighbourhood-expanded reference: out[4*v + {0,1,2,3}] = {(v000,v001), (v010,v011), (v100,v101), (v110,v111)} of voxel v,
// zeros beyond the array edge.  8x the memory (4.4 GB per class at 256 px, 16.6 GB at 400 px — sized for 180 GB HBM3e),
// in exchange every trilinear sample of the fine pass / store stage is one aligned 64-byte read.
__global__ void k_expand_volume(RbProjector pj, float4 *out, float4 *out2)
{
	const size_t n = (size_t) pj.mdlXY * pj.mdlZ;
	for (size_t v = blockIdx.x * (size_t) blockDim.x + threadIdx.x; v < n; v += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (v % pj.mdlX);
		const int y = (int) ((v / pj.mdlX) % pj.mdlY);
		const int z = (int) (v / pj.mdlXY);
		const bool hx = x + 1 < pj.mdlX, hy = y + 1 < pj.mdlY, hz = z + 1 < pj.mdlZ;
		const float2 zero = make_float2(0.f, 0.f);
		const float2 *b = pj.mdl + v;
		const float2 d000 = b[0], d001 = hx ? b[1] : zero;
		const float2 d010 = hy ? b[pj.mdlX] : zero, d011 = (hy && hx) ? b[pj.mdlX + 1] : zero;
		const float2 d100 = hz ? b[pj.mdlXY] : zero, d101 = (hz && hx) ? b[pj.mdlXY + 1] : zero;
		const float2 d110 = (hz && hy) ? b[pj.mdlXY + pj.mdlX] : zero, d111 = (hz && hy && hx) ? b[pj.mdlXY + pj.mdlX + 1] : zero;
		float4 *o = out + 4 * (size_t) rb_cell8(pj.blk, pj.nbx, pj.nbxy, x, y, z);      // radius-sorted 4 x 4 x 4 blocks (RbProjector::blk)
		o[0] = make_float4(d000.x, d000.y, d001.x, d001.y);
		o[1] = make_float4(d010.x, d010.y, d011.x, d011.y);
		o[2] = make_float4(d100.x, d100.y, d101.x, d101.y);
		o[3] = make_float4(d110.x, d110.y, d111.x, d111.y);
		out2[v] = make_float4(d000.x, d000.y, d001.x, d001.y);
	}
}

int rbk_expand_volume(rb_ctx *ctx, const RbProjector &pj, float4 *d_out, float4 *d_out2)
{
	k_expand_volume<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(pj, d_out, d_out2);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// x-pair copy of the box [0, cX) x [cInitY, cInitY + cY) x [cInitZ, cInitZ + cY) of the reference, contiguous (RbProjector::c2*);
// voxels outside the stored volume read as zero
__global__ void k_xpair_core(RbProjector pj, int cX, int cY, int cInitY, int cInitZ, float4 *out)
{
	const size_t n = (size_t) cX * cY * cY;
	for (size_t v = blockIdx.x * (size_t) blockDim.x + threadIdx.x; v < n; v += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (v % cX);
		const int y = (int) ((v / cX) % cY) + cInitY - pj.mdlInitY;
		const int z = (int) (v / ((size_t) cX * cY)) + cInitZ - pj.mdlInitZ;
		float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
		if (y >= 0 && y < pj.mdlY && z >= 0 && z < pj.mdlZ && x < pj.mdlX)
		{
			const float2 *b = pj.mdl + ((size_t) z * pj.mdlXY + (size_t) y * pj.mdlX + x);
			const float2 a = b[0], c = x + 1 < pj.mdlX ? b[1] : make_float2(0.f, 0.f);
			o = make_float4(a.x, a.y, c.x, c.y);
		}
		out[v] = o;
	}
}

// xy-quad copy of the same box: entry v holds the x-pairs of rows y and y + 1 (two float4)
__global__ void k_xyquad_core(RbProjector pj, int cX, int cY, int cInitY, int cInitZ, float4 *out)
{
	const size_t n = (size_t) cX * cY * cY;
	for (size_t v = blockIdx.x * (size_t) blockDim.x + threadIdx.x; v < n; v += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (v % cX);
		const int y = (int) ((v / cX) % cY) + cInitY - pj.mdlInitY;
		const int z = (int) (v / ((size_t) cX * cY)) + cInitZ - pj.mdlInitZ;
		float4 o[2];
		for (int dy = 0; dy < 2; dy++)
		{
			o[dy] = make_float4(0.f, 0.f, 0.f, 0.f);
			const int yy = y + dy;
			if (yy >= 0 && yy < pj.mdlY && z >= 0 && z < pj.mdlZ && x < pj.mdlX)
			{
				const float2 *b = pj.mdl + ((size_t) z * pj.mdlXY + (size_t) yy * pj.mdlX + x);
				const float2 a = b[0], c = x + 1 < pj.mdlX ? b[1] : make_float2(0.f, 0.f);
				o[dy] = make_float4(a.x, a.y, c.x, c.y);
			}
		}
		out[2 * v] = o[0]; out[2 * v + 1] = o[1];
	}
}

int rbk_xyquad_core(rb_ctx *ctx, const RbProjector &pj, int cX, int cY, int cInitY, int cInitZ, float4 *d_out)
{
	k_xyquad_core<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(pj, cX, cY, cInitY, cInitZ, d_out);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

int rbk_xpair_core(rb_ctx *ctx, const RbProjector &pj, int cX, int cY, int cInitY, int cInitZ, float4 *d_out)
{
	k_xpair_core<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(pj, cX, cY, cInitY, cInitZ, d_out);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// padded block accumulator -> canonical accumulator: voxel (x, y, z) is local (x & 3, ...) of its own block and, where a
// coordinate is a multiple of 4, local 4 of the block before it along that axis: up to 8 copies, each read (and cleared)
// by exactly one voxel, summed in a fixed order
__global__ void k_bp_fold(RbBackprojector bp)
{
	const size_t n = (size_t) bp.mdlX * bp.mdlY * bp.mdlZ;
	const int nbx = (bp.mdlX + 3) >> 2, nby = (bp.mdlY + 3) >> 2, nbz = (bp.mdlZ + 3) >> 2;
	for (size_t v = blockIdx.x * (size_t) blockDim.x + threadIdx.x; v < n; v += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (v % bp.mdlX);
		const int y = (int) ((v / bp.mdlX) % bp.mdlY);
		const int z = (int) (v / ((size_t) bp.mdlX * bp.mdlY));
		float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
		for (int cz = 0; cz < 2; cz++)
		{
			int bz = z >> 2, lz = z & 3;
			if (cz) { if (lz != 0 || bz == 0) continue; bz--; lz = 4; }
			if (bz >= nbz) continue;
			for (int cy = 0; cy < 2; cy++)
			{
				int by = y >> 2, ly = y & 3;
				if (cy) { if (ly != 0 || by == 0) continue; by--; ly = 4; }
				if (by >= nby) continue;
				for (int cx = 0; cx < 2; cx++)
				{
					int bx = x >> 2, lx = x & 3;
					if (cx) { if (lx != 0 || bx == 0) continue; bx--; lx = 4; }
					if (bx >= nbx) continue;
					const uint32_t rank = bp.blk[rb_blk_slot(bp.nbx, bp.nbxy, bx, by, bz)];
					float4 *p = bp.blkvol + ((size_t) rank << 7) + (lz * 25 + ly * 5 + lx);
					const float4 a = *p;
					if (a.x != 0.f || a.y != 0.f || a.z != 0.f)
					{
						acc.x += a.x; acc.y += a.y; acc.z += a.z;
						*p = make_float4(0.f, 0.f, 0.f, 0.f);
					}
				}
			}
		}
		if (acc.x != 0.f || acc.y != 0.f || acc.z != 0.f)
		{
			float4 c = bp.vol[v];
			c.x += acc.x; c.y += acc.y; c.z += acc.z;
			bp.vol[v] = c;
		}
	}
}

int rbk_bp_fold(rb_ctx *ctx, const RbBackprojector &bp)
{
	k_bp_fold<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(bp);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// AccBackprojector::getMdlData (acc_backprojector_impl.h:109-138): interleaved float4 -> three SoA arrays
__global__ void k_bp_deinterleave(const float4 *vol, float *re, float *im, float *w, size_t n)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const float4 v = vol[i];
		re[i] = v.x; im[i] = v.y; w[i] = v.z;
	}
}

int rbk_bp_deinterleave(rb_ctx *ctx, const float4 *vol, float *re, float *im, float *w, size_t n)
{
	k_bp_deinterleave<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(vol, re, im, w, n);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// Fourier-slice projection of `count` orientations (fine-pass row rule, see rbk_diff2_fine_stage)
__global__ void k_project(RbProjector pj, int n, const float *eulers, float2 *out)
{
	const int imgX = n / 2 + 1;
	const RbProjK pk = rb_make_projk(pj, imgX);
	const float *e = eulers + (size_t) blockIdx.y * 9;
	float2 *o = out + (size_t) blockIdx.y * n * imgX;
	for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n * imgX; pix += gridDim.x * blockDim.x)
	{
		int x = pix % imgX, iy = pix / imgX, y = iy;
		float2 v = make_float2(0.f, 0.f);
		bool skip = false;
		if (iy > pk.maxR)
		{
			if (iy >= n - pk.maxR) y = iy - n;
			else skip = (x != pk.maxR);
		}
		if (!skip) v = rb_project3d(pk, x, y, e[0], e[1], e[3], e[4], e[6], e[7]);
		o[pix] = v;
	}
}

int rbk_project(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers, int count, float2 *d_out)
{
	if (count < 1) return RB_OK;
	dim3 grid((n * (n / 2 + 1) + 255) / 256, count);
	k_project<<<grid, 256, 0, ctx->stream>>>(pj, n, d_eulers, d_out);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// relion_reconstruct-style posed back-projection (BASELINE config #2): BackProjector::backproject2Dto3D
// (/root/reference/src/backprojector.cpp:55-357, TRILINEAR, no Ewald sphere / magnification), one orientation and unit
// weight per image, images already multiplied by their CTF, weights Fctf = ctf^2 (src/reconstructor.cpp:632-716).
// Positions are computed in fp64 like the reference (RFLOAT = double there), so the set of pixels inside r_max and their
// cells are the reference's; the interpolation weights and the accumulation are fp32 (float4 (re, im, w, 0) voxels,
// one 16-byte vector reduction per corner, two lanes per pixel so that the x-neighbours of a corner pair share a sector).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4_misc(float4 *addr, float a, float b, float c)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(0.f) : "memory");
}

__global__ void __launch_bounds__(256)
k_backproject_posed(RbBackprojector bp, int n, int count, const float2 *F2D, const float *Fctf, const float *eulers)
{
	const int xs = n / 2 + 1;
	const double pf = (double) bp.padding_factor;
	const long long rr = (long long) floor((double) bp.maxR * pf + 0.5);
	const double max_r2 = (double) (rr * rr);
	const int npix = n * xs;
	const int nit = (npix + 255) / 256;
	const size_t sy = bp.mdlX, sz = (size_t) bp.mdlX * bp.mdlY;
	for (int img = blockIdx.x; img < count; img += gridDim.x)
	{
		const float *e = eulers + (size_t) img * 9;
		const double a00 = (double) e[0] * pf, a01 = (double) e[1] * pf, a10 = (double) e[3] * pf, a11 = (double) e[4] * pf,
		             a20 = (double) e[6] * pf, a21 = (double) e[7] * pf;
		const double AtA_xx = a00 * a00 + a10 * a10 + a20 * a20, AtA_xy = a00 * a01 + a10 * a11 + a20 * a21,
		             AtA_yy = a01 * a01 + a11 * a11 + a21 * a21;
		const float2 *F = F2D + (size_t) img * npix;
		const float *W = Fctf + (size_t) img * npix;
		for (int it = 0; it < nit; it++)                      // uniform trip count: the scatter below is warp-cooperative
		{
			const int pix = it * 256 + threadIdx.x;
			long long cell = -1;
			float sfx = 0.f, sfy = 0.f, sfz = 0.f, vr = 0.f, vi = 0.f, vw = 0.f;
			if (pix < npix)
			{
				const int i = pix / xs, x = pix - i * xs;
				const int y = i < xs ? i : i - n;
				const int first_allowed_x = i < xs ? 0 : 1;
				const double discr = AtA_xy * AtA_xy * y * y - AtA_xx * (AtA_yy * y * y - max_r2);   // :118-128
				if (discr >= 0.)
				{
					const double d = sqrt(discr) / AtA_xx, q = -AtA_xy * y / AtA_xx;
					int first_x = (int) ceil(q - d), last_x = (int) floor(q + d);
					if (first_x < first_allowed_x) first_x = first_allowed_x;
					if (last_x > xs - 1) last_x = xs - 1;
					const float w = __ldg(W + pix);
					if (x >= first_x && x <= last_x && w > 0.f)
					{
						double xp = a00 * x + a01 * y, yp = a10 * x + a11 * y, zp = a20 * x + a21 * y;
						if (xp * xp + yp * yp + zp * zp <= max_r2)
						{
							float2 v = __ldg(F + pix);
							if (xp < 0.) { xp = -xp; yp = -yp; zp = -zp; v.y = -v.y; }
							const double fx0 = floor(xp), fy0 = floor(yp), fz0 = floor(zp);
							const int x0 = (int) fx0, y0 = (int) fy0 - bp.mdlInitY, z0 = (int) fz0 - bp.mdlInitZ;
							if (x0 >= 0 && x0 + 1 < bp.mdlX && y0 >= 0 && y0 + 1 < bp.mdlY && z0 >= 0 && z0 + 1 < bp.mdlZ)   // :213-218
							{
								sfx = (float) (xp - fx0); sfy = (float) (yp - fy0); sfz = (float) (zp - fz0);
								vr = v.x; vi = v.y; vw = w;
								cell = ((long long) z0 * bp.mdlY + y0) * bp.mdlX + x0;
							}
						}
					}
				}
			}
#pragma unroll
			for (int h = 0; h < 2; h++)
			{
				const int src = 16 * h + ((threadIdx.x & 31) >> 1);
				const long long c = __shfl_sync(0xffffffffu, cell, src);
				const float fx = __shfl_sync(0xffffffffu, sfx, src), fy = __shfl_sync(0xffffffffu, sfy, src), fz = __shfl_sync(0xffffffffu, sfz, src);
				const float r = __shfl_sync(0xffffffffu, vr, src), im = __shfl_sync(0xffffffffu, vi, src), w = __shfl_sync(0xffffffffu, vw, src);
				if (c >= 0)
				{
					const int px = threadIdx.x & 1;
					const float wx = px ? fx : 1.f - fx, mfy = 1.f - fy, mfz = 1.f - fz;
					float4 *b = bp.vol + (size_t) c + px;
					float dd;
					dd = mfz * mfy * wx; red_add_v4_misc(b, dd * r, dd * im, dd * w);
					dd = mfz * fy * wx;  red_add_v4_misc(b + sy, dd * r, dd * im, dd * w);
					dd = fz * mfy * wx;  red_add_v4_misc(b + sz, dd * r, dd * im, dd * w);
					dd = fz * fy * wx;   red_add_v4_misc(b + sz + sy, dd * r, dd * im, dd * w);
				}
			}
		}
	}
}

int rbk_backproject_posed(rb_ctx *ctx, const RbBackprojector &bp, int n, int count, const float2 *d_F, const float *d_W, const float *d_eulers)
{
	if (count < 1) return RB_OK;
	// band-major order (kernels_band.cu) unless RB_POSED_BAND=0 or the batch is too small to fill the GPU shell by shell
	static int band = -1;
	if (band < 0) { const char *e = getenv("RB_POSED_BAND"); band = e ? atoi(e) : 1; }
	if (band && (count >= 64 || band == 2)) return rbk_backproject_posed_band(ctx, bp, n, count, d_F, d_W, d_eulers);
	const int grid = count < ctx->num_sms * 8 ? count : ctx->num_sms * 8;
	k_backproject_posed<<<grid, 256, 0, ctx->stream>>>(bp, n, count, d_F, d_W, d_eulers);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}
