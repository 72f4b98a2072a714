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

// geometry of the x-pair copy (RbProjector::mdl2, the coarse-window core while a pool runs)
__host__ __device__ inline RbProjK rb_make_projk2(const RbProjector &p, int imgX)
{
	RbProjK k = rb_make_projk(p, imgX);
	k.mdlX = p.c2X; k.mdlXY = p.c2XY; k.mdlInitY = p.c2InitY; k.mdlInitZ = p.c2InitZ;
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

// Same sample from the x-pair copy (RbProjector::mdl2): four aligned 16-byte loads, one per (z, y) corner row.
// Arithmetic identical to rb_project3d.
__device__ __forceinline__ float2 rb_project3d_xp(const RbProjK &k, const float4 *mdl2, int x, int y,
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
	const float4 *b = mdl2 + ((size_t) (z0 - k.mdlInitZ) * (size_t) k.mdlXY + (size_t) (y0 - k.mdlInitY) * (size_t) k.mdlX + (size_t) x0);
	const float4 q0 = __ldg(b), q1 = __ldg(b + k.mdlX), q2 = __ldg(b + k.mdlXY), q3 = __ldg(b + k.mdlXY + k.mdlX);
	float2 r;
	{
		float dx00 = q0.x + (q0.z - q0.x) * fx, dx10 = q1.x + (q1.z - q1.x) * fx;
		float dx01 = q2.x + (q2.z - q2.x) * fx, dx11 = q3.x + (q3.z - q3.x) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.x = dxy0 + (dxy1 - dxy0) * fz;
	}
	{
		float dx00 = q0.y + (q0.w - q0.y) * fx, dx10 = q1.y + (q1.w - q1.y) * fx;
		float dx01 = q2.y + (q2.w - q2.y) * fx, dx11 = q3.y + (q3.w - q3.y) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.y = dxy0 + (dxy1 - dxy0) * fz;
	}
	if (inv) r.y = -r.y;
	return r;
}

// Same sample from the neighbourhood-expanded volume (RbProjector::mdl8): four aligned 16-byte loads that cover
// exactly two 32-byte sectors.  Arithmetic identical to rb_project3d.
struct RbProjK8 {
	const float4 *mdl8;
	const uint32_t *blk; int nbx, nbxy;
	int mdlX, mdlXY, mdlInitY, mdlInitZ, maxR, maxR2_padded;
	float pf;
};
// index (in 64-byte cells) of the cell whose origin is voxel (x0, yi, zi), yi / zi counted from the array edge: rank of its
// 4 x 4 x 4 block in the radius-sorted block order, then (z, y, x) inside the block (see RbProjector::blk)
__device__ __forceinline__ int rb_cell8(const uint32_t *blk, int nbx, int nbxy, int x0, int yi, int zi)
{
	const uint32_t rank = __ldg(blk + rb_blk_slot(nbx, nbxy, x0 >> 2, yi >> 2, zi >> 2));
	return (int) ((rank << 6) | (uint32_t) (((zi & 3) << 4) | ((yi & 3) << 2) | (x0 & 3)));
}
__host__ __device__ inline RbProjK8 rb_make_projk8(const RbProjector &p, int imgX)
{
	RbProjK8 k;
	int imgMaxR = imgX - 1;
	k.maxR = p.mdlMaxR >= imgMaxR ? imgMaxR : p.mdlMaxR;
	k.mdl8 = p.mdl8; k.mdlX = p.mdlX; k.mdlXY = p.mdlXY; k.mdlInitY = p.mdlInitY; k.mdlInitZ = p.mdlInitZ;
	k.blk = p.blk; k.nbx = p.nbx; k.nbxy = p.nbxy;
	k.pf = p.padding_factor;
	k.maxR2_padded = (int) (k.maxR * k.maxR * k.pf * k.pf);
	return k;
}
__device__ __forceinline__ float2 rb_project3d_x8(const RbProjK8 &k, int x, int y,
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
	const float4 *b = k.mdl8 + 4 * (size_t) rb_cell8(k.blk, k.nbx, k.nbxy, x0, y0 - k.mdlInitY, z0 - k.mdlInitZ);
	const float4 q0 = __ldg(b), q1 = __ldg(b + 1), q2 = __ldg(b + 2), q3 = __ldg(b + 3);
	float2 r;
	{
		float dx00 = q0.x + (q0.z - q0.x) * fx, dx10 = q1.x + (q1.z - q1.x) * fx;
		float dx01 = q2.x + (q2.z - q2.x) * fx, dx11 = q3.x + (q3.z - q3.x) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.x = dxy0 + (dxy1 - dxy0) * fz;
	}
	{
		float dx00 = q0.y + (q0.w - q0.y) * fx, dx10 = q1.y + (q1.w - q1.y) * fx;
		float dx01 = q2.y + (q2.w - q2.y) * fx, dx11 = q3.y + (q3.w - q3.y) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.y = dxy0 + (dxy1 - dxy0) * fz;
	}
	if (inv) r.y = -r.y;
	return r;
}

// Same sample from the expanded volume with TWO 32-byte loads (LDG.E.256, sm_100): each half of the 64-byte cell is one
// 32-byte sector, so a warp-wide request touches 32 sectors instead of the 4 x 32 of the x-pair gather, and the L1's tag stage
// sees two wavefronts per sample instead of four.  Arithmetic identical to rb_project3d.
__device__ __forceinline__ void rb_ldg256(const float4 *p, float4 &a, float4 &b)
{
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ float2 rb_project3d_c256(const RbProjK8 &k, int x, int y,
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
	const float4 *b = k.mdl8 + 4 * (size_t) rb_cell8(k.blk, k.nbx, k.nbxy, x0, y0 - k.mdlInitY, z0 - k.mdlInitZ);
	float4 q0, q1, q2, q3;
	rb_ldg256(b, q0, q1);
	rb_ldg256(b + 2, q2, q3);
	float2 r;
	{
		float dx00 = q0.x + (q0.z - q0.x) * fx, dx10 = q1.x + (q1.z - q1.x) * fx;
		float dx01 = q2.x + (q2.z - q2.x) * fx, dx11 = q3.x + (q3.z - q3.x) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.x = dxy0 + (dxy1 - dxy0) * fz;
	}
	{
		float dx00 = q0.y + (q0.w - q0.y) * fx, dx10 = q1.y + (q1.w - q1.y) * fx;
		float dx01 = q2.y + (q2.w - q2.y) * fx, dx11 = q3.y + (q3.w - q3.y) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.y = dxy0 + (dxy1 - dxy0) * fz;
	}
	if (inv) r.y = -r.y;
	return r;
}

// Same sample from the xy-quad copy of the coarse core (RbProjector::quad, geometry of k = rb_make_projk2): the rows y0 / y0 + 1 of
// plane z0 in one 32-byte load, those of plane z0 + 1 in another.  Arithmetic identical to rb_project3d_xp.
__device__ __forceinline__ float2 rb_project3d_q256(const RbProjK &k, const float4 *quad, int x, int y,
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
	const float4 *b = quad + 2 * ((size_t) (z0 - k.mdlInitZ) * (size_t) k.mdlXY + (size_t) (y0 - k.mdlInitY) * (size_t) k.mdlX + (size_t) x0);
	float4 q0, q1, q2, q3;
	rb_ldg256(b, q0, q1);
	rb_ldg256(b + 2 * (size_t) k.mdlXY, q2, q3);
	float2 r;
	{
		float dx00 = q0.x + (q0.z - q0.x) * fx, dx10 = q1.x + (q1.z - q1.x) * fx;
		float dx01 = q2.x + (q2.z - q2.x) * fx, dx11 = q3.x + (q3.z - q3.x) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.x = dxy0 + (dxy1 - dxy0) * fz;
	}
	{
		float dx00 = q0.y + (q0.w - q0.y) * fx, dx10 = q1.y + (q1.w - q1.y) * fx;
		float dx01 = q2.y + (q2.w - q2.y) * fx, dx11 = q3.y + (q3.w - q3.y) * fx;
		float dxy0 = dx00 + (dx10 - dx00) * fy, dxy1 = dx01 + (dx11 - dx01) * fy;
		r.y = dxy0 + (dxy1 - dxy0) * fz;
	}
	if (inv) r.y = -r.y;
	return r;
}

// Split, branch-free form of rb_project3d_q256: rb_quad_issue computes the position and issues the two loads unconditionally (a sample
// outside r_max, or one the caller masks out, reads the first word of the copy and is zeroed in rb_quad_finish), so that a thread can
// have the loads of SEVERAL samples in flight — inside an `if (inside)` region the compiler must finish one sample before it may
// issue the next one's loads.
struct RbQuadFetch {
	float4 q0, q1, q2, q3;
	float fx, fy, fz;
	int flags;   // bit0: sample is live (inside r_max and not masked), bit1: Hermitian mate (conjugate)
};
__device__ __forceinline__ void rb_quad_issue(const RbProjK &k, const float4 *quad, bool live, int x, int y,
                                              float e0, float e1, float e3, float e4, float e6, float e7, RbQuadFetch &f)
{
	float xp = (e0 * x + e1 * y) * k.pf;
	float yp = (e3 * x + e4 * y) * k.pf;
	float zp = (e6 * x + e7 * y) * k.pf;
	const int r2 = (int) (xp * xp + yp * yp + zp * zp);
	live = live && r2 <= k.maxR2_padded;
	const bool inv = xp < 0.f;
	if (inv) { xp = -xp; yp = -yp; zp = -zp; }
	const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
	f.fx = xp - fx0; f.fy = yp - fy0; f.fz = zp - fz0;
	f.flags = (live ? 1 : 0) | (inv ? 2 : 0);
	const size_t off = live ? 2 * ((size_t) ((int) fz0 - k.mdlInitZ) * (size_t) k.mdlXY + (size_t) ((int) fy0 - k.mdlInitY) * (size_t) k.mdlX + (size_t) (int) fx0) : 0;
	rb_ldg256(quad + off, f.q0, f.q1);
	rb_ldg256(quad + off + (live ? 2 * (size_t) k.mdlXY : 0), f.q2, f.q3);
}
__device__ __forceinline__ float2 rb_quad_finish(const RbQuadFetch &f)
{
	float2 r;
	{
		float dx00 = f.q0.x + (f.q0.z - f.q0.x) * f.fx, dx10 = f.q1.x + (f.q1.z - f.q1.x) * f.fx;
		float dx01 = f.q2.x + (f.q2.z - f.q2.x) * f.fx, dx11 = f.q3.x + (f.q3.z - f.q3.x) * f.fx;
		float dxy0 = dx00 + (dx10 - dx00) * f.fy, dxy1 = dx01 + (dx11 - dx01) * f.fy;
		r.x = dxy0 + (dxy1 - dxy0) * f.fz;
	}
	{
		float dx00 = f.q0.y + (f.q0.w - f.q0.y) * f.fx, dx10 = f.q1.y + (f.q1.w - f.q1.y) * f.fx;
		float dx01 = f.q2.y + (f.q2.w - f.q2.y) * f.fx, dx11 = f.q3.y + (f.q3.w - f.q3.y) * f.fx;
		float dxy0 = dx00 + (dx10 - dx00) * f.fy, dxy1 = dx01 + (dx11 - dx01) * f.fy;
		r.y = dxy0 + (dxy1 - dxy0) * f.fz;
	}
	if (f.flags & 2) r.y = -r.y;
	if (!(f.flags & 1)) r = make_float2(0.f, 0.f);
	return r;
}

// Split form of rb_project3d_x8 for software pipelining: rb_proj_issue computes the sample position and
// issues the four 16-byte loads; rb_proj_finish does the lerps once the data is needed.
struct RbProjFetch {
	float4 q0, q1, q2, q3;
	float fx, fy, fz;
	int flags;   // bit0: inside r_max (else the sample is zero), bit1: Hermitian mate (conjugate)
};
__device__ __forceinline__ void rb_proj_issue(const RbProjK8 &k, int x, int y,
                                              float e0, float e1, float e3, float e4, float e6, float e7, RbProjFetch &f)
{
	float xp = (e0 * x + e1 * y) * k.pf;
	float yp = (e3 * x + e4 * y) * k.pf;
	float zp = (e6 * x + e7 * y) * k.pf;
	const int r2 = (int) (xp * xp + yp * yp + zp * zp);
	const bool inside = r2 <= k.maxR2_padded;
	const bool inv = xp < 0.f;
	if (inv) { xp = -xp; yp = -yp; zp = -zp; }
	const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
	f.fx = xp - fx0; f.fy = yp - fy0; f.fz = zp - fz0;
	f.flags = (inside ? 1 : 0) | (inv ? 2 : 0);
	if (inside)
	{
		const float4 *b = k.mdl8 + 4 * (size_t) rb_cell8(k.blk, k.nbx, k.nbxy, (int) fx0, (int) fy0 - k.mdlInitY, (int) fz0 - k.mdlInitZ);
		f.q0 = __ldg(b); f.q1 = __ldg(b + 1); f.q2 = __ldg(b + 2); f.q3 = __ldg(b + 3);
	}
	else f.q0 = f.q1 = f.q2 = f.q3 = make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ float2 rb_proj_finish(const RbProjFetch &f)
{
	float2 r;
	{
		float dx00 = f.q0.x + (f.q0.z - f.q0.x) * f.fx, dx10 = f.q1.x + (f.q1.z - f.q1.x) * f.fx;
		float dx01 = f.q2.x + (f.q2.z - f.q2.x) * f.fx, dx11 = f.q3.x + (f.q3.z - f.q3.x) * f.fx;
		float dxy0 = dx00 + (dx10 - dx00) * f.fy, dxy1 = dx01 + (dx11 - dx01) * f.fy;
		r.x = dxy0 + (dxy1 - dxy0) * f.fz;
	}
	{
		float dx00 = f.q0.y + (f.q0.w - f.q0.y) * f.fx, dx10 = f.q1.y + (f.q1.w - f.q1.y) * f.fx;
		float dx01 = f.q2.y + (f.q2.w - f.q2.y) * f.fx, dx11 = f.q3.y + (f.q3.w - f.q3.y) * f.fx;
		float dxy0 = dx00 + (dx10 - dx00) * f.fy, dxy1 = dx01 + (dx11 - dx01) * f.fy;
		r.y = dxy0 + (dxy1 - dxy0) * f.fz;
	}
	if (f.flags & 2) r.y = -r.y;
	return r;
}

// Phase factor (cos, sin)(x*tx + y*ty) of a translation given in TURNS per pixel (ux = tx / 2pi): the turn count is
// reduced exactly to [-0.5, 0.5] before the fast hardware sincos, so the absolute error stays ~5e-7 for any shift
// (translatePixel, cuda_device_utils.cuh:110-160, uses sincosf of the same argument).
__device__ __forceinline__ float2 rb_phase(int x, int y, float ux, float uy)
{
	float u = fmaf((float) x, ux, (float) y * uy);
	u -= rintf(u);
	float s, c;
	__sincosf(6.283185307179586f * u, &s, &c);
	return make_float2(c, s);
}

// index of pixel (x, y) of an n-window inside an array stored at window nfull (windowFourierTransform,
// src/fftw.h:850-856): same (x, y), row = y<0 ? y+nfull : y
__device__ __forceinline__ int rb_src_index(int x, int y, int nfull)
{
	int row = y < 0 ? y + nfull : y;
	return row * (nfull / 2 + 1) + x;
}

// Persistent CTAs pull work items from a shared queue instead of striding over them: items differ several-fold in cost
// (fine orientations without significant samples, partial tiles), and a static round-robin left 6 - 23 % of the SM time idle
// at the end of the fine / store kernels (ncu sm__cycles_active.avg vs sm__cycles_elapsed.max).  The first item of a CTA is
// its own index, later ones come from the counter (which starts at zero); queue == nullptr keeps the static stride.
// Every thread of the CTA must call this (it synchronises).
__device__ __forceinline__ int rb_next_work(int *queue, int *s_slot, int prev, bool first)
{
	if (first) return (int) blockIdx.x;
	if (!queue) return prev + (int) gridDim.x;
	__syncthreads();
	if (threadIdx.x == 0) *s_slot = (int) gridDim.x + atomicAdd(queue, 1);
	__syncthreads();
	return *s_slot;
}

__device__ __forceinline__ void rb_atomic_min_pos(int *addr, float v) { atomicMin(addr, __float_as_int(v)); }

// ZYZ Euler matrix, inverted (= transposed), in fp64 then cast: generateEulerMatrices(inverse=true)
// (acc_helper_functions_impl.h:198-262).  With MBL / MBR: A = L * A; A = A * R; A = A.inv() in fp64 (:248-255).
__device__ inline void rb_euler_fine(double rot, double tilt, double psi, const RbLR &lr, float *e)
{
	const double d2r = 3.14159265358979323846 / 180.0;
	double sa, ca, sb, cb, sg, cg;
	sincos(rot * d2r, &sa, &ca); sincos(tilt * d2r, &sb, &cb); sincos(psi * d2r, &sg, &cg);
	double cc = cb * ca, cs = cb * sa, sc = sb * ca, ss = sb * sa;
	if (!lr.doL && !lr.doR)
	{
		e[0] = (float) (cg * cc - sg * sa); e[3] = (float) (cg * cs + sg * ca); e[6] = (float) (-cg * sb);
		e[1] = (float) (-sg * cc - cg * sa); e[4] = (float) (-sg * cs + cg * ca); e[7] = (float) (sg * sb);
		e[2] = (float) sc; e[5] = (float) ss; e[8] = (float) cb;
		return;
	}
	double A[9] = { cg * cc - sg * sa, cg * cs + sg * ca, -cg * sb, -sg * cc - cg * sa, -sg * cs + cg * ca, sg * sb, sc, ss, cb }, B[9];
	if (lr.doL)
	{
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) B[r * 3 + c] = lr.L[r * 3] * A[c] + lr.L[r * 3 + 1] * A[3 + c] + lr.L[r * 3 + 2] * A[6 + c];
		for (int k = 0; k < 9; k++) A[k] = B[k];
	}
	if (lr.doR)
	{
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) B[r * 3 + c] = A[r * 3] * lr.R[c] + A[r * 3 + 1] * lr.R[3 + c] + A[r * 3 + 2] * lr.R[6 + c];
		for (int k = 0; k < 9; k++) A[k] = B[k];
	}
	const double det = A[0] * (A[4] * A[8] - A[7] * A[5]) - A[1] * (A[3] * A[8] - A[6] * A[5]) + A[2] * (A[3] * A[7] - A[6] * A[4]);
	e[0] = (float) ((A[4] * A[8] - A[7] * A[5]) / det); e[1] = (float) ((A[7] * A[2] - A[1] * A[8]) / det); e[2] = (float) ((A[1] * A[5] - A[4] * A[2]) / det);
	e[3] = (float) ((A[5] * A[6] - A[8] * A[3]) / det); e[4] = (float) ((A[8] * A[0] - A[2] * A[6]) / det); e[5] = (float) ((A[2] * A[3] - A[5] * A[0]) / det);
	e[6] = (float) ((A[3] * A[7] - A[6] * A[4]) / det); e[7] = (float) ((A[6] * A[1] - A[0] * A[7]) / det); e[8] = (float) ((A[0] * A[4] - A[3] * A[1]) / det);
}
