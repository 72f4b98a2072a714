// relion_b200 — fine-pass squared differences (sm_100a).
//
// Replaces cuda_kernel_diff2_fine (/root/reference/src/acc/cuda/cuda_kernels/diff2.cuh:193-332; ALTCPU twin
// src/acc/cpu/cpu_kernels/diff2.h:284-430) for a whole pool of particles in one launch.
//
// Formulation.  With Z(x,y) = (corr/2) * conj(A(x,y)) * X(x,y)  (A = reference slice, X = corrected image):
//     diff2[t] = sum (corr/2) (|A|^2 + |X|^2)  -  2 Re sum_y e^{i y ty} ( sum_x Z(x,y) e^{i x tx} )
// The translation phase factorises per column/row, so whoever walks one image row keeps
// R_t = sum_x Z e^{i x tx} in registers (4 FMA per pixel and translation, column factors from a shared-memory
// table) and applies the row factor once per row.  The reference evaluates sincos(x tx + y ty) for every
// (pixel, translation).
//
// Mapping.  A QUAD of 4 lanes walks one row.  Per pixel the quad issues ONE 64-byte gather: lane k loads the
// k-th 16-byte quarter of the voxel's 2x2x2 cell from the neighbourhood-expanded volume (one L1 wavefront per
// pixel instead of 32 divergent ones per warp-load), the trilinear lerp is finished with two shuffle rounds, and
// lane k owns 8 of the (up to) 32 translations of the pass, i.e. 16 accumulator registers instead of 64.
// Rows are dealt q, q+64, q+128, ... so that every quad gets the same number of pixels (+-2 %).
#include "img_src.cuh"
#include <cstdlib>

static const int FI_THREADS = 256;
static const int FI_QUADS = FI_THREADS / 4;
static const int FI_TF = 32;   // translations per pass = 8 significant coarse translations x 4

struct FineArgs {
	// pool mode
	const RbPartMeta *metas; RbPartState *states;
	const RbFineOrient *fo; const int *pair_list; const int *counters; // counters[0] = number of fine orientations
	float *fs_w;
	// stage mode (fo == nullptr): the reference's job lists
	const float *st_eulers; float st_sum_init;
	const unsigned long long *st_rot_idx, *st_trans_idx, *st_job_idx, *st_job_num; int st_njobs;
	float *st_out;
	// common
	const float4 *img4;            // [P][n][n/2+1] (X'.re, X'.im, corr/2, 0); zero weight outside the valid runs
	float2 *slices; long long slice_capacity;   // pool mode: cache of the projected slices for the store stage
	const RbProjector *projs;
	const RbRow *rows; int nrows; int n;
	const float *tx, *ty; int NOT;
	int cc;                        // cross-correlation criterion (cuda_kernel_diff2_CC_fine, diff2.cuh:464-640; ALTCPU
	                               // cpu_kernels/diff2.h:904-1050): img4.z holds corr, value = -cross / sqrt(sum corr |A|^2)
	int *queue;                    // pool mode: work queue counter (zero at launch), see rb_next_work
};

struct FineFetch {
	float4 q;      // this lane's quarter of the 2x2x2 cell
	float4 img;    // (X.re, X.im, corr/2, -)
	float fx, fy, fz;
	int flags;     // bit0 inside r_max, bit1 Hermitian mate
};

__device__ __forceinline__ void fine_issue(const RbProjK8 &pk, const float4 *img_row, int k, int x, int y,
                                           float e0, float e1, float e3, float e4, float e6, float e7, FineFetch &f)
{
	float xp = (e0 * x + e1 * y) * pk.pf;
	float yp = (e3 * x + e4 * y) * pk.pf;
	float zp = (e6 * x + e7 * y) * pk.pf;
	const int r2 = (int) (xp * xp + yp * yp + zp * zp);
	const bool inside = r2 <= pk.maxR2_padded;
	const bool inv = xp < 0.f;
	if (inv) { xp = -xp; yp = -yp; zp = -zp; }
	const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
	f.fx = xp - fx0; f.fy = yp - fy0; f.fz = zp - fz0;
	f.flags = (inside ? 1 : 0) | (inv ? 2 : 0);
	f.q = make_float4(0.f, 0.f, 0.f, 0.f);
	if (inside)
	{
		const size_t cell = (size_t) rb_cell8(pk.blk, pk.nbx, pk.nbxy, (int) fx0, (int) fy0 - pk.mdlInitY, (int) fz0 - pk.mdlInitZ);
		f.q = __ldg(pk.mdl8 + 4 * cell + k);
	}
	f.img = __ldg(img_row + x);
}

// finish the trilinear interpolation inside the quad: lane k holds (z,y) corner pair k = 2*dz + dy
__device__ __forceinline__ float2 fine_finish(const FineFetch &f, int k, unsigned qmask)
{
	float dxr = f.q.x + (f.q.z - f.q.x) * f.fx;
	float dxi = f.q.y + (f.q.w - f.q.y) * f.fx;
	// quads of one warp walk rows of different length, so only the quad's own 4 lanes take part in the shuffles
	const float or1 = __shfl_xor_sync(qmask, dxr, 1), oi1 = __shfl_xor_sync(qmask, dxi, 1);
	// lanes with dy == 0 hold dx_y0, partner holds dx_y1:  dxy = dx_y0 + (dx_y1 - dx_y0) * fy
	const bool y1 = (k & 1) != 0;
	const float a_r = y1 ? or1 : dxr, b_r = y1 ? dxr : or1;
	const float a_i = y1 ? oi1 : dxi, b_i = y1 ? dxi : oi1;
	float dxyr = a_r + (b_r - a_r) * f.fy;
	float dxyi = a_i + (b_i - a_i) * f.fy;
	const float or2 = __shfl_xor_sync(qmask, dxyr, 2), oi2 = __shfl_xor_sync(qmask, dxyi, 2);
	const bool z1 = (k & 2) != 0;
	const float c_r = z1 ? or2 : dxyr, d_r = z1 ? dxyr : or2;
	const float c_i = z1 ? oi2 : dxyi, d_i = z1 ? dxyi : oi2;
	float2 r;
	r.x = c_r + (d_r - c_r) * f.fz;
	r.y = c_i + (d_i - c_i) * f.fz;
	if (f.flags & 2) r.y = -r.y;
	return r;
}

template <bool CC>
__global__ void __launch_bounds__(FI_THREADS, 3)
k_diff2_fine(FineArgs A, RbModelDev M)
{
	extern __shared__ float4 s_px[];                 // [xs][16]: entry f = j*4 + k holds (cos, sin) of translations 2f, 2f+1
	__shared__ float s_ux[FI_TF], s_uy[FI_TF];
	__shared__ float s_red[FI_THREADS / 32][FI_TF + 1];
	__shared__ float s_e[6];

	const int xs = A.n / 2 + 1;
	const bool stage = (A.fo == nullptr);
	const int nwork = stage ? A.st_njobs : A.counters[0];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int k = threadIdx.x & 3, qd = threadIdx.x >> 2;
	const unsigned qmask = 0xFu << (lane & ~3);

	__shared__ int s_next;
	for (int w = rb_next_work(A.queue, &s_next, 0, true); w < nwork; w = rb_next_work(A.queue, &s_next, w, false))
	{
		int nsamp, cls = 0, p = 0;
		long long out_off;
		const float *eu;
		RbFineOrient F;
		float xi2_half;
		if (stage)
		{
			unsigned long long j0 = A.st_job_idx[w];
			nsamp = (int) A.st_job_num[w];
			eu = A.st_eulers + A.st_rot_idx[j0] * 9;
			out_off = (long long) j0;
			xi2_half = A.st_sum_init;
		}
		else
		{
			F = A.fo[w];
			nsamp = F.n_t * A.NOT; cls = F.iclass; p = F.particle; out_off = F.sample_off;
			eu = A.fo[w].e;
			xi2_half = A.metas[p].xi2_half;
		}
		__syncthreads();
		if (threadIdx.x < 6) s_e[threadIdx.x] = eu[threadIdx.x + threadIdx.x / 2];   // elements 0,1,3,4,6,7
		const float4 *img = A.img4 + (size_t) p * A.n * xs;
		float2 *slice = (!stage && A.slices && w < A.slice_capacity) ? A.slices + (size_t) w * A.n * xs : nullptr;
		const RbProjK8 pk = rb_make_projk8(A.projs[cls], xs);
		float bmin = FLT_MAX;

		for (int c0 = 0; c0 < nsamp; c0 += FI_TF)
		{
			const int ntr = min(FI_TF, nsamp - c0);
			const int nj = (ntr + 7) >> 3;            // groups of 8 translations in use
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
			for (int i = threadIdx.x; i < xs * 16; i += FI_THREADS)
			{
				const int x = i >> 4, f = i & 15;
				if (f < nj * 4)
				{
					const float2 p0 = rb_phase(x, 0, s_ux[2 * f], 0.f), p1 = rb_phase(x, 0, s_ux[2 * f + 1], 0.f);
					s_px[i] = make_float4(p0.x, p0.y, p1.x, p1.y);
				}
			}
			__syncthreads();
			const float e0 = s_e[0], e1 = s_e[1], e3 = s_e[2], e4 = s_e[3], e6 = s_e[4], e7 = s_e[5];

			float tot[8];
#pragma unroll
			for (int i = 0; i < 8; i++) tot[i] = 0.f;
			float base = 0.f;

			for (int r = qd; r < A.nrows; r += FI_QUADS)
			{
				const RbRow rd = A.rows[r];
				const float4 *img_row = img + (size_t) rd.iy * xs;
				float2 *slice_row = (slice && c0 == 0) ? slice + (size_t) rd.iy * xs : nullptr;
				float accr[8], acci[8];
#pragma unroll
				for (int i = 0; i < 8; i++) { accr[i] = 0.f; acci[i] = 0.f; }
				int x = 0;
				FineFetch cur;
				fine_issue(pk, img_row, k, x, rd.y, e0, e1, e3, e4, e6, e7, cur);
				while (true)
				{
					const bool haven = x < rd.x_hi;
					FineFetch nxt;
					if (haven) fine_issue(pk, img_row, k, x + 1, rd.y, e0, e1, e3, e4, e6, e7, nxt);

					const float2 ref = fine_finish(cur, k, qmask);
					if (slice_row && k == 0) slice_row[x] = ref;   // the store stage streams this instead of gathering again
					const float hc = cur.img.z;
					const float zr = hc * (ref.x * cur.img.x + ref.y * cur.img.y);
					const float zi = hc * (ref.x * cur.img.y - ref.y * cur.img.x);
					if (k == 0) base += hc * ((ref.x * ref.x + ref.y * ref.y) + (CC ? 0.f : (cur.img.x * cur.img.x + cur.img.y * cur.img.y)));
					const float4 *pp = s_px + (x << 4) + k;
#pragma unroll
					for (int j = 0; j < 4; j++)
					{
						if (j < nj)
						{
							const float4 cs = pp[j * 4];
							accr[2 * j] = fmaf(zr, cs.x, fmaf(-zi, cs.y, accr[2 * j]));
							acci[2 * j] = fmaf(zr, cs.y, fmaf(zi, cs.x, acci[2 * j]));
							accr[2 * j + 1] = fmaf(zr, cs.z, fmaf(-zi, cs.w, accr[2 * j + 1]));
							acci[2 * j + 1] = fmaf(zr, cs.w, fmaf(zi, cs.z, acci[2 * j + 1]));
						}
					}
					if (!haven) break;
					cur = nxt;
					x++;
				}
				// row factor e^{i y ty}
#pragma unroll
				for (int j = 0; j < 4; j++)
#pragma unroll
					for (int h = 0; h < 2; h++)
					{
						const int t = 2 * (j * 4 + k) + h;
						if (t < ntr)
						{
							const float2 py = rb_phase(0, rd.y, 0.f, s_uy[t]);
							tot[2 * j + h] += accr[2 * j + h] * py.x - acci[2 * j + h] * py.y;
						}
					}
			}
			// sum over the 8 quads of the warp (lanes with equal k), then over the 8 warps; fixed order
#pragma unroll
			for (int i = 0; i < 8; i++)
			{
				float v = tot[i];
				v += __shfl_xor_sync(RB_FULL_MASK, v, 4);
				v += __shfl_xor_sync(RB_FULL_MASK, v, 8);
				v += __shfl_xor_sync(RB_FULL_MASK, v, 16);
				if (lane < 4) s_red[wid][2 * ((i >> 1) * 4 + k) + (i & 1)] = v;
			}
			base = warp_sum(base);
			if (lane == 0) s_red[wid][FI_TF] = base;
			__syncthreads();
			if (threadIdx.x < ntr)
			{
				float c = 0.f, b = 0.f;
#pragma unroll
				for (int ww = 0; ww < FI_THREADS / 32; ww++) { c += s_red[ww][threadIdx.x]; b += s_red[ww][FI_TF]; }
				float v = CC ? -c / sqrtf(b) : fmaxf((b - 2.f * c) + xi2_half, 0.f);          // CC: diff2.h:1041-1046
				if (stage) A.st_out[out_off + c0 + threadIdx.x] += v;                         // diff2.h:424-428
				else { A.fs_w[out_off + c0 + threadIdx.x] = v; bmin = fminf(bmin, v); }
			}
		}
		if (!stage && !CC && threadIdx.x < FI_TF && bmin < FLT_MAX) rb_atomic_min_pos(&A.states[p].fmin_bits, bmin);
	}
}

// ---------------------------------------------------------------------------------------------
// Variant with the gathers staged through shared memory by cp.async: a lane keeps FA_DEPTH 16-byte copies of its cell
// quarter in flight without holding them in registers (the register-prefetch kernel above is bound by latency x bytes in
// flight: 768 lanes x 16 B per SM).  Per pixel a quad issues: the cell quarter (16 B, lane k -> quarter k), component k of
// the prepared image value (4 B) and stores component k of (fx, fy, fz, flags); the consumer reads its own quarter and the
// quad's four components back.  Same arithmetic, same summation order as k_diff2_fine.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int FA_DEPTH, int FA_MINB, bool CC>
__global__ void __launch_bounds__(FI_THREADS, FA_MINB)
k_diff2_fine_async(FineArgs A, RbModelDev M)
{
	extern __shared__ float4 s_dyn[];
	float4 *s_px = s_dyn;                                   // [xs][16] phase table
	const int xs = A.n / 2 + 1;
	float4 *s_cell = s_px + xs * 16;                        // [FA_DEPTH][FI_THREADS]
	float *s_frac = (float *) (s_cell + FA_DEPTH * FI_THREADS);   // [FA_DEPTH][FI_THREADS]
	float *s_img = s_frac + FA_DEPTH * FI_THREADS;          // [FA_DEPTH][FI_THREADS]
	__shared__ float s_ux[FI_TF], s_uy[FI_TF];
	__shared__ float s_red[FI_THREADS / 32][FI_TF + 1];
	__shared__ float s_e[6];

	const int nwork = A.counters[0];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int k = threadIdx.x & 3, qd = threadIdx.x >> 2, qbase = threadIdx.x & ~3;
	const unsigned qmask = 0xFu << (lane & ~3);

	__shared__ int s_next;
	for (int w = rb_next_work(A.queue, &s_next, 0, true); w < nwork; w = rb_next_work(A.queue, &s_next, w, false))
	{
		const RbFineOrient F = A.fo[w];
		const int nsamp = F.n_t * A.NOT, cls = F.iclass, p = F.particle;
		const long long out_off = F.sample_off;
		const float xi2_half = A.metas[p].xi2_half;
		__syncthreads();
		if (threadIdx.x < 6) s_e[threadIdx.x] = A.fo[w].e[threadIdx.x + threadIdx.x / 2];
		const float4 *img = A.img4 + (size_t) p * A.n * xs;
		float2 *slice = (A.slices && w < A.slice_capacity) ? A.slices + (size_t) w * A.n * xs : nullptr;
		const RbProjK8 pk = rb_make_projk8(A.projs[cls], xs);
		float bmin = FLT_MAX;

		for (int c0 = 0; c0 < nsamp; c0 += FI_TF)
		{
			const int ntr = min(FI_TF, nsamp - c0);
			const int nj = (ntr + 7) >> 3;
			__syncthreads();
			if (threadIdx.x < FI_TF)
			{
				float ux = 0.f, uy = 0.f;
				if (threadIdx.x < ntr)
				{
					const int j = c0 + threadIdx.x;
					const int it = A.pair_list[F.pair_off + j / A.NOT] * A.NOT + (j % A.NOT);
					ux = A.tx[it] * 0.15915494309189535f; uy = A.ty[it] * 0.15915494309189535f;
				}
				s_ux[threadIdx.x] = ux; s_uy[threadIdx.x] = uy;
			}
			__syncthreads();
			for (int i = threadIdx.x; i < xs * 16; i += FI_THREADS)
			{
				const int x = i >> 4, f = i & 15;
				if (f < nj * 4)
				{
					const float2 p0 = rb_phase(x, 0, s_ux[2 * f], 0.f), p1 = rb_phase(x, 0, s_ux[2 * f + 1], 0.f);
					s_px[i] = make_float4(p0.x, p0.y, p1.x, p1.y);
				}
			}
			__syncthreads();
			const float e0 = s_e[0], e1 = s_e[1], e3 = s_e[2], e4 = s_e[3], e6 = s_e[4], e7 = s_e[5];

			float tot[8];
#pragma unroll
			for (int i = 0; i < 8; i++) tot[i] = 0.f;
			float base = 0.f;

			for (int r = qd; r < A.nrows; r += FI_QUADS)
			{
				const RbRow rd = A.rows[r];
				const float4 *img_row = img + (size_t) rd.iy * xs;
				float2 *slice_row = (slice && c0 == 0) ? slice + (size_t) rd.iy * xs : nullptr;
				float accr[8], acci[8];
#pragma unroll
				for (int i = 0; i < 8; i++) { accr[i] = 0.f; acci[i] = 0.f; }

				// issue pixel x into ring slot (x % FA_DEPTH)
				auto issue = [&](int x)
				{
					const int slot = x % FA_DEPTH;
					float xp = (e0 * x + e1 * rd.y) * pk.pf;
					float yp = (e3 * x + e4 * rd.y) * pk.pf;
					float zp = (e6 * x + e7 * rd.y) * pk.pf;
					const int r2 = (int) (xp * xp + yp * yp + zp * zp);
					const bool inside = r2 <= pk.maxR2_padded;
					const bool inv = xp < 0.f;
					if (inv) { xp = -xp; yp = -yp; zp = -zp; }
					const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
					const float comp = k == 0 ? xp - fx0 : (k == 1 ? yp - fy0 : (k == 2 ? zp - fz0 : __int_as_float((inside ? 1 : 0) | (inv ? 2 : 0))));
					s_frac[slot * FI_THREADS + threadIdx.x] = comp;
					if (inside)
					{
						const size_t cell = (size_t) rb_cell8(pk.blk, pk.nbx, pk.nbxy, (int) fx0, (int) fy0 - pk.mdlInitY, (int) fz0 - pk.mdlInitZ);
						cp_async16(s_cell + slot * FI_THREADS + threadIdx.x, pk.mdl8 + 4 * cell + k);
					}
					else s_cell[slot * FI_THREADS + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
					cp_async4(s_img + slot * FI_THREADS + threadIdx.x, (const float *) (img_row + x) + k);
					cp_async_commit();
				};
#pragma unroll
				for (int d = 0; d < FA_DEPTH; d++) { if (d <= rd.x_hi) issue(d); else cp_async_commit(); }
				for (int x = 0; x <= rd.x_hi; x++)
				{
					const int slot = x % FA_DEPTH;
					cp_async_wait<FA_DEPTH - 1>();
					__syncwarp(qmask);                                  // the quad's four copies of this pixel are visible
					const float4 q = s_cell[slot * FI_THREADS + threadIdx.x];
					const float4 fr = *(const float4 *) (s_frac + slot * FI_THREADS + qbase);
					const float4 im = *(const float4 *) (s_img + slot * FI_THREADS + qbase);
					__syncwarp(qmask);                                  // everyone has read the slot before it is refilled
					if (x + FA_DEPTH <= rd.x_hi) issue(x + FA_DEPTH); else cp_async_commit();
					FineFetch f;
					f.q = q; f.fx = fr.x; f.fy = fr.y; f.fz = fr.z; f.flags = __float_as_int(fr.w);
					const float2 ref = fine_finish(f, k, qmask);
					if (slice_row && k == 0) slice_row[x] = ref;
					const float hc = im.z;
					const float zr = hc * (ref.x * im.x + ref.y * im.y);
					const float zi = hc * (ref.x * im.y - ref.y * im.x);
					if (k == 0) base += hc * ((ref.x * ref.x + ref.y * ref.y) + (CC ? 0.f : (im.x * im.x + im.y * im.y)));
					const float4 *pp = s_px + (x << 4) + k;
#pragma unroll
					for (int j = 0; j < 4; j++)
					{
						if (j < nj)
						{
							const float4 cs = pp[j * 4];
							accr[2 * j] = fmaf(zr, cs.x, fmaf(-zi, cs.y, accr[2 * j]));
							acci[2 * j] = fmaf(zr, cs.y, fmaf(zi, cs.x, acci[2 * j]));
							accr[2 * j + 1] = fmaf(zr, cs.z, fmaf(-zi, cs.w, accr[2 * j + 1]));
							acci[2 * j + 1] = fmaf(zr, cs.w, fmaf(zi, cs.z, acci[2 * j + 1]));
						}
					}
				}
				cp_async_wait<0>();
#pragma unroll
				for (int j = 0; j < 4; j++)
#pragma unroll
					for (int h = 0; h < 2; h++)
					{
						const int t = 2 * (j * 4 + k) + h;
						if (t < ntr)
						{
							const float2 py = rb_phase(0, rd.y, 0.f, s_uy[t]);
							tot[2 * j + h] += accr[2 * j + h] * py.x - acci[2 * j + h] * py.y;
						}
					}
			}
#pragma unroll
			for (int i = 0; i < 8; i++)
			{
				float v = tot[i];
				v += __shfl_xor_sync(RB_FULL_MASK, v, 4);
				v += __shfl_xor_sync(RB_FULL_MASK, v, 8);
				v += __shfl_xor_sync(RB_FULL_MASK, v, 16);
				if (lane < 4) s_red[wid][2 * ((i >> 1) * 4 + k) + (i & 1)] = v;
			}
			base = warp_sum(base);
			if (lane == 0) s_red[wid][FI_TF] = base;
			__syncthreads();
			if (threadIdx.x < ntr)
			{
				float c = 0.f, b = 0.f;
#pragma unroll
				for (int ww = 0; ww < FI_THREADS / 32; ww++) { c += s_red[ww][threadIdx.x]; b += s_red[ww][FI_TF]; }
				const float v = CC ? -c / sqrtf(b) : fmaxf((b - 2.f * c) + xi2_half, 0.f);
				A.fs_w[out_off + c0 + threadIdx.x] = v; bmin = fminf(bmin, v);
			}
		}
		if (!CC && threadIdx.x < FI_TF && bmin < FLT_MAX) rb_atomic_min_pos(&A.states[p].fmin_bits, bmin);   // CC values are negative: the minimum is taken by k_weights_cc_fine
	}
}

static int launch_fine(rb_ctx *ctx, FineArgs &A, int grid)
{
	const int xs = A.n / 2 + 1;
	size_t sm = (size_t) xs * 16 * sizeof(float4);
	static size_t configured[RB_MAX_DEVICES][2] = {};
	size_t &cfg = configured[ctx->device % RB_MAX_DEVICES][A.cc ? 1 : 0];
	void (*kern)(FineArgs, RbModelDev) = A.cc ? k_diff2_fine<true> : k_diff2_fine<false>;
	if (sm > cfg)
	{
		RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
		cfg = sm;
	}
	kern<<<grid, FI_THREADS, sm, ctx->stream>>>(A, ctx->d_model);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

int rbk_diff2_fine_pool(rb_ctx *ctx, PoolSlot &s)
{
	const RbModelDev &M = ctx->d_model;
	const int n = M.current_size, xs = n / 2 + 1;
	RB_CHECK(s.fimg4.ensure((size_t) s.P * n * xs * sizeof(float4)));
	RB_CUDA(cudaMemsetAsync(s.fimg4.p, 0, (size_t) s.P * n * xs * sizeof(float4), ctx->stream));
	PrepArgs PA;
	memset(&PA, 0, sizeof(PA));
	PA.metas = s.meta.as<RbPartMeta>(); PA.Fimg = s.Fimg.as<float2>();
	PA.Fctf = ctx->h_model.do_ctf_correction ? s.Fctf.as<float>() : nullptr;
	PA.ires = M.d2_ires_f; PA.rows = M.d2_rows_f; PA.nrows = M.d2_nrows_f; PA.n = n; PA.out = s.fimg4.as<float4>();
	if (M.do_cc) { PA.cc = 1; PA.cc_corr = s.cc_corr.as<float>() + s.P; }   // [1][P]: the fine window's 1 / sqrtXi2^2
	dim3 pg((M.d2_nrows_f * xs + 255) / 256, s.P);
	k_prep_img4<<<pg, 256, 0, ctx->stream>>>(PA, M);
	RB_LAUNCH_CHECK(ctx);

	FineArgs A;
	memset(&A, 0, sizeof(A));
	A.metas = s.meta.as<RbPartMeta>(); A.states = s.state.as<RbPartState>();
	A.fo = s.fo.as<RbFineOrient>(); A.pair_list = s.pair_list.as<int>(); A.counters = s.counters.as<int>();
	A.fs_w = s.fs_w.as<float>();
	A.img4 = s.fimg4.as<float4>();
	A.slices = s.slices.as<float2>(); A.slice_capacity = s.slice_capacity;
	A.projs = ctx->d_proj.as<RbProjector>();
	A.rows = M.d2_rows_f; A.nrows = M.d2_nrows_f; A.n = n;
	A.tx = ctx->d_samp.ftx; A.ty = ctx->d_samp.fty; A.NOT = ctx->d_samp.n_over_trans;
	A.cc = M.do_cc;
	A.queue = s.counters.as<int>() + 8;
	// cp.async-staged variant whenever three CTAs per SM still fit next to the phase table (measured at 256 px: 6.44 ms vs
	// 6.97 ms; with two CTAs per SM it loses: 9.9 vs 9.2 ms at 400 px with a 6-deep ring)
	static int use_async = -1;
	if (use_async < 0) { const char *e = getenv("RB_FINE_ASYNC"); use_async = e ? atoi(e) : 1; }
	if (use_async)
	{
		const size_t table = (size_t) xs * 16 * sizeof(float4);
		const size_t ring1 = (size_t) FI_THREADS * (sizeof(float4) + 2 * sizeof(float));
		const size_t limit = 74 * 1024;
		// ring depth 2 (measured: depth 2 / 3 / 4 within 1 % at 256 px, depth 2 best at 400 px where it keeps 3 CTAs per SM);
		// RB_FINE_ASYNC=4 tries four CTAs per SM (64 registers)
		const int mode = (use_async > 1) ? use_async : (table + 2 * ring1 <= limit ? 2 : 0);
		if (mode == 2 || mode == 4)
		{
			const size_t sm = table + 2 * ring1;
			static size_t configured_dev[RB_MAX_DEVICES][4] = {};
			size_t *configured = configured_dev[ctx->device % RB_MAX_DEVICES];
			void (*kern)(FineArgs, RbModelDev) = A.cc ? (mode == 2 ? k_diff2_fine_async<2, 3, true> : k_diff2_fine_async<2, 4, true>)
			                                          : (mode == 2 ? k_diff2_fine_async<2, 3, false> : k_diff2_fine_async<2, 4, false>);
			const int ci = (mode == 4 ? 1 : 0) + (A.cc ? 2 : 0);
			if (sm > configured[ci])
			{
				RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
				configured[ci] = sm;
			}
			kern<<<ctx->num_sms * (mode == 4 ? 4 : 3), FI_THREADS, sm, ctx->stream>>>(A, ctx->d_model);
			RB_LAUNCH_CHECK(ctx);
			return RB_OK;
		}
	}
	return launch_fine(ctx, A, ctx->num_sms * 3);
}

int rbk_diff2_fine_stage(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers,
                         const float *d_tx, const float *d_ty, const float *d_re, const float *d_im,
                         const float *d_corr, float sum_init,
                         const unsigned long long *d_rot_idx, const unsigned long long *d_trans_idx,
                         const unsigned long long *d_job_idx, const unsigned long long *d_job_num, int n_jobs,
                         float *d_out, int cc)
{
	// rows with the fine kernels' rule (diff2.cuh:268-274, diff2.h:344-355): rows in the dead band
	// maxR < iy < imgY-maxR contribute only the pixel x = maxR (which projects to zero)
	const int xs = n / 2 + 1;
	RbProjK pk = rb_make_projk(pj, xs);
	std::vector<RbRow> rows;
	for (int iy = 0; iy < n; iy++)
	{
		int lo = 0, hi = xs - 1, y = iy;
		if (iy > pk.maxR)
		{
			if (iy >= n - pk.maxR) y = iy - n;
			else { lo = pk.maxR; hi = pk.maxR; }
		}
		rows.push_back(RbRow{(short) iy, (short) y, (short) lo, (short) hi});
	}
	RB_CHECK(ctx->scratch[0].ensure(rows.size() * sizeof(RbRow)));
	RB_CHECK(ctx->scratch[1].ensure(sizeof(RbProjector)));
	RB_CHECK(ctx->scratch[6].ensure((size_t) n * xs * sizeof(float4)));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[0].p, rows.data(), rows.size() * sizeof(RbRow), cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[1].p, &pj, sizeof(RbProjector), cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));   // `rows` is a pageable temporary
	PrepArgs PA;
	memset(&PA, 0, sizeof(PA));
	PA.src.re = d_re; PA.src.im = d_im; PA.src.corr = d_corr; PA.src.n_array = n;
	PA.rows = ctx->scratch[0].as<RbRow>(); PA.nrows = (int) rows.size(); PA.n = n; PA.out = ctx->scratch[6].as<float4>();
	PA.cc = cc;
	k_prep_img4<<<dim3((n * xs + 255) / 256, 1), 256, 0, ctx->stream>>>(PA, ctx->d_model);
	RB_LAUNCH_CHECK(ctx);
	FineArgs A;
	memset(&A, 0, sizeof(A));
	A.st_eulers = d_eulers; A.st_sum_init = sum_init;
	A.st_rot_idx = d_rot_idx; A.st_trans_idx = d_trans_idx; A.st_job_idx = d_job_idx; A.st_job_num = d_job_num;
	A.st_njobs = n_jobs; A.st_out = d_out;
	A.img4 = ctx->scratch[6].as<float4>();
	A.projs = ctx->scratch[1].as<RbProjector>();
	A.rows = ctx->scratch[0].as<RbRow>(); A.nrows = (int) rows.size(); A.n = n;
	A.tx = d_tx; A.ty = d_ty; A.NOT = 1;
	A.cc = cc;
	int grid = n_jobs < ctx->num_sms * 3 ? n_jobs : ctx->num_sms * 3;
	if (grid < 1) return RB_OK;
	return launch_fine(ctx, A, grid);
}
