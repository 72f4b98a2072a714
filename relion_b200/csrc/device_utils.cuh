// relion_b200 — device helpers shared by the kernels (sm_100a).
#pragma once
#include "common.cuh"

#define RB_FULL_MASK 0xffffffffu

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RB_FULL_MASK, v, o);
	return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RB_FULL_MASK, v, o);
	return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(RB_FULL_MASK, v, o));
	return v;
}
__device__ __forceinline__ long long warp_sum(long long v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RB_FULL_MASK, v, o);
	return v;
}
__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RB_FULL_MASK, v, o);
	return v;
}

// Block-wide sum; every thread gets the result.  `sm` needs 32 elements.  Deterministic tree.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *sm)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	v = warp_sum(v);
	__syncthreads();
	if (lane == 0) sm[wid] = v;
	__syncthreads();
	T r = (lane < nw) ? sm[lane] : (T) 0;
	r = warp_sum(r);
	return r;
}
__device__ __forceinline__ float block_max(float v, float *sm)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	v = warp_max(v);
	__syncthreads();
	if (lane == 0) sm[wid] = v;
	__syncthreads();
	float r = (lane < nw) ? sm[lane] : RB_LOWEST;
	r = warp_max(r);
	return r;
}

// maxR / maxR2_padded exactly as AccProjectorKernel's ctor + makeKernel compute them
// (acc_projectorkernel_impl.h:51-69, 301-319): imgMaxR = imgX-1, int*int*float*float -> int.
struct RbProjK {
	const float2 *mdl;
	int mdlX, mdlXY, mdlInitY, mdlInitZ, maxR, maxR2_padded;
	float pf;
};
__host__ __device__ inline RbProjK rb_make_projk(const RbProjector &p, int imgX)
{
	RbProjK k;
	int imgMaxR = imgX - 1;
	k.maxR = p.mdlMaxR >= imgMaxR ? imgMaxR : p.mdlMaxR;
	k.mdl = p.mdl; k.mdlX = p.mdlX; k.mdlXY = p.mdlXY; k.mdlInitY = p.mdlInitY; k.mdlInitZ = p.mdlInitZ;
	k.pf = p.padding_factor;
	k.maxR2_padded = (int) (k.maxR * k.maxR * k.pf * k.pf);
	return k;
}

// AccProjectorKernel::project3Dmodel, 2D-image overload, exact fp32 lerps (PROJECTOR_NO_TEXTURES /
// CpuKernels::complex3D semantics: acc_projectorkernel_impl.h:161-231, cpu_kernels/cpu_utils.h:159-205).
// One 8-byte load per tap (re,im interleaved) instead of the reference's two separate textures.
__device__ __forceinline__ float2 rb_project3d(const RbProjK &k, int x, int y,
                                               float e0, float e1, float e3, float e4, float e6, float e7)
{
	float xp = (e0 * x + e1 * y) * k.pf;
	float yp = (e3 * x + e4 * y) * k.pf;
	float zp = (e6 * x + e7 * y) * k.pf;
	int r2 = (int) (xp * xp + yp * yp + zp * zp);
	if (r2 > k.maxR2_padded) return make_float2(0.f, 0.f);
	const bool inv = xp < 0.f;
	if (inv) { xp = -xp; yp = -yp; zp = -zp; }
	const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
	const float fx = xp - fx0, fy = yp - fy0, fz = zp - fz0;
	const int x0 = (int) fx0, y0 = (int) fy0, z0 = (int) fz0;
	const float2 *b = k.mdl + ((size_t) (z0 - k.mdlInitZ) * (size_t) k.mdlXY + (size_t) (y0 - k.mdlInitY) * (size_t) k.mdlX + (size_t) x0);
	const float2 d000 = __ldg(b), d001 = __ldg(b + 1);
	const float2 d010 = __ldg(b + k.mdlX), d011 = __ldg(b + k.mdlX + 1);
	const float2 d100 = __ldg(b + k.mdlXY), d101 = __ldg(b + k.mdlXY + 1);
	const float2 d110 = __ldg(b + k.mdlXY + k.mdlX), d111 = __ldg(b + k.mdlXY + k.mdlX + 1);
	float2 r;
	{
		float dx00 = d000.x + (d001.x - d000.x) * fx, dx10 = d010.x + (d011.x - d010.x) * fx;
		float dx01 = d100.x + (d101.x - d100.x) * fx, dx11 = d110.x + (d111.x - d110.x) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.x = dxy0 + (dxy1 - dxy0) * fz;
	}
	{
		float dx00 = d000.y + (d001.y - d000.y) * fx, dx10 = d010.y + (d011.y - d010.y) * fx;
		float dx01 = d100.y + (d101.y - d100.y) * fx, dx11 = d110.y + (d111.y - d110.y) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.y = dxy0 + (dxy1 - dxy0) * fz;
	}
	if (inv) r.y = -r.y;
	return r;
}

// index of pixel (x, y) of an n-window inside an array stored at window nfull (windowFourierTransform,
// src/fftw.h:850-856): same (x, y), row = y<0 ? y+nfull : y
__device__ __forceinline__ int rb_src_index(int x, int y, int nfull)
{
	int row = y < 0 ? y + nfull : y;
	return row * (nfull / 2 + 1) + x;
}

__device__ __forceinline__ void rb_atomic_min_pos(int *addr, float v) { atomicMin(addr, __float_as_int(v)); }

// ZYZ Euler matrix, inverted (= transposed), in fp64 then cast: generateEulerMatrices(inverse=true)
// (acc_helper_functions_impl.h:198-262)
__device__ inline void rb_euler_fine(double rot, double tilt, double psi, float *e)
{
	const double d2r = 3.14159265358979323846 / 180.0;
	double sa, ca, sb, cb, sg, cg;
	sincos(rot * d2r, &sa, &ca); sincos(tilt * d2r, &sb, &cb); sincos(psi * d2r, &sg, &cg);
	double cc = cb * ca, cs = cb * sa, sc = sb * ca, ss = sb * sa;
	e[0] = (float) (cg * cc - sg * sa); e[3] = (float) (cg * cs + sg * ca); e[6] = (float) (-cg * sb);
	e[1] = (float) (-sg * cc - cg * sa); e[4] = (float) (-sg * cs + cg * ca); e[7] = (float) (sg * sb);
	e[2] = (float) sc; e[5] = (float) ss; e[8] = (float) cb;
}
