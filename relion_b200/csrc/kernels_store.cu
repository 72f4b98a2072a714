// relion_b200 — weighted sums and back-projection (storeWeightedSums), sm_100a.
//
// Replaces cuda_kernel_collect2jobs (helper.cuh:69-165), cuda_kernel_wavg (wavg.cuh:13-152) and
// cuda_kernel_backproject3D (BP.cuh:174-403) of /root/reference/src/acc/cuda/cuda_kernels and their
// ALTCPU twins (cpu_kernels/helper.h:65-153, wavg.h:22-199, BP.h:497-753); orchestration being
// replaced: storeWeightedSums (acc_ml_optimiser_impl.h:2553-3667).
//
// B200-first choices:
//  * wavg and back-projection are ONE kernel per pool: both walk (fine orientation, pixel,
//    significant translation); the reference slice, the phase factors and the weights are computed
//    once and feed both the sigma2/XA/AA sums and the scatter.
//  * shell sums are reduced in shared memory per orientation (the reference writes per-pixel arrays
//    with atomics and sums the shells on the host, acc_ml_optimiser_impl.h:3466-3494).
//  * the accumulator volume is float4 (re, im, weight, 0): each trilinear corner is one 16-byte vector
//    reduction (red.global.add.v4.f32) instead of three scalar atomics into three arrays; the two x
//    neighbours of a corner pair share a 32-byte sector.
#include "device_utils.cuh"
#include <cstdlib>

static const int ST_THREADS = 256;
static const int ST_MAXSAMP = 2048;

__device__ __forceinline__ void red_add_v4(float4 *addr, float a, float b, float c)
{
	// 16-byte vector reduction, sm_90+ (PTX ISA 8.1: red.global.add.v4.f32)
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(0.f) : "memory");
}

// 8-corner trilinear scatter with Hermitian fold (BP.cuh:301-401 / BP.h:632-748)
__device__ __forceinline__ void bp_scatter(const RbBackprojector &bp, int max_r2_vol, int x, int y,
                                           float e0, float e1, float e3, float e4, float e6, float e7,
                                           float real, float imag, float Fweight)
{
	float xp = (e0 * x + e1 * y) * bp.padding_factor;
	float yp = (e3 * x + e4 * y) * bp.padding_factor;
	float zp = (e6 * x + e7 * y) * bp.padding_factor;
	if (xp * xp + yp * yp + zp * zp > (float) max_r2_vol) return;
	if (xp < 0.f) { xp = -xp; yp = -yp; zp = -zp; imag = -imag; }
	const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
	const float fx = xp - fx0, fy = yp - fy0, fz = zp - fz0;
	const int x0 = (int) fx0, y0 = (int) fy0 - bp.mdlInitY, z0 = (int) fz0 - bp.mdlInitZ;
	const float mfx = 1.f - fx, mfy = 1.f - fy, mfz = 1.f - fz;
	float4 *b = bp.vol + ((size_t) z0 * bp.mdlY + y0) * (size_t) bp.mdlX + x0;
	const size_t sy = bp.mdlX, sz = (size_t) bp.mdlX * bp.mdlY;
	float d;
	d = mfz * mfy * mfx; red_add_v4(b, d * real, d * imag, d * Fweight);
	d = mfz * mfy * fx;  red_add_v4(b + 1, d * real, d * imag, d * Fweight);
	d = mfz * fy * mfx;  red_add_v4(b + sy, d * real, d * imag, d * Fweight);
	d = mfz * fy * fx;   red_add_v4(b + sy + 1, d * real, d * imag, d * Fweight);
	d = fz * mfy * mfx;  red_add_v4(b + sz, d * real, d * imag, d * Fweight);
	d = fz * mfy * fx;   red_add_v4(b + sz + 1, d * real, d * imag, d * Fweight);
	d = fz * fy * mfx;   red_add_v4(b + sz + sy, d * real, d * imag, d * Fweight);
	d = fz * fy * fx;    red_add_v4(b + sz + sy + 1, d * real, d * imag, d * Fweight);
}

__device__ __forceinline__ void build_tables_st(float2 *tab_x, float2 *tab_y, int imgX, int ny, int yoff,
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
		int t = i / ny, yy = i - t * ny, y = yy - yoff;
		float s, c; sincosf((y < 0 ? -y : y) * ty[t], &s, &c);
		tab_y[i] = make_float2(c, y < 0 ? -s : s);
	}
}

// ---------------------------------------------------------------------------------------------
// collect (collect2jobs + host loop acc_ml_optimiser_impl.h:2821-2854): one thread per fine orientation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_collect(const RbPartMeta *metas, RbPartState *states, const RbFineOrient *fo, const int *pair_list,
          const float *fs_w, const long long *fs_ihid, const int *dir_idx, RbModelDev M, RbSamplingDev S,
          double *out_pdf_dir, double *out_pdf_class, double *out_prior_class, const int *counters)
{
	__shared__ double dred[32];
	if (counters[2]) return;
	const int p = blockIdx.x;
	const RbPartMeta m = metas[p];
	RbPartState *st = states + p;
	if (st->status != 0) return;
	const int NOR = S.n_over_rot, NOT = S.n_over_trans;
	const int nfo = st->n_so * NOR;
	const float sig = st->fsig_weight, sumw = st->fsum_weight;
	double a_w = 0., a_s2 = 0.;
	for (int i = threadIdx.x; i < nfo; i += blockDim.x)
	{
		const RbFineOrient F = fo[st->fo_base + i];
		float sw = 0.f, ss2 = 0.f, spx = 0.f, spy = 0.f;
		// centre of the translation prior: the class' own for 2D references (:2673-2677), else the particle's
		const double prx = M.prior_offset_class ? M.prior_offset_class[2 * F.iclass] : m.prx;
		const double pry = M.prior_offset_class ? M.prior_offset_class[2 * F.iclass + 1] : m.pry;
		for (int j = 0; j < F.n_t * NOT; j++)
		{
			float w = fs_w[F.sample_off + j];
			w = (w >= sig) ? w / sumw : 0.f;                                          // helper.cuh:118-127
			const int it = pair_list[F.pair_off + j / NOT] * NOT + (j % NOT);
			const double xs = m.oldx + S.over_trans_x[it], ys = m.oldy + S.over_trans_y[it];
			const double dx = prx - xs, dy = pry - ys;
			const float o2 = (float) (dx * dx + dy * dy);                             // :2728-2736
			sw += w; ss2 += w * o2;
			spx += w * (float) xs; spy += w * (float) ys;                             // helper.cuh:129-131
		}
		const int idl = F.iorient / m.np;
		const int mydir = m.dir_off < 0 ? idl : dir_idx[m.dir_off + idl];             // :2833-2837
		if (sw != 0.f)
		{
			atomicAdd(out_pdf_dir + (size_t) F.iclass * S.n_dir + mydir, (double) sw);
			atomicAdd(out_pdf_class + F.iclass, (double) sw);
			if (M.prior_offset_class)                                                 // :2847-2851
			{
				atomicAdd(out_prior_class + 2 * F.iclass, M.pixel_size * (double) spx);
				atomicAdd(out_prior_class + 2 * F.iclass + 1, M.pixel_size * (double) spy);
			}
		}
		a_w += (double) sw;
		a_s2 += M.pixel_size * M.pixel_size * (double) ss2;                           // :2845
	}
	a_w = block_sum(a_w, dred);
	a_s2 = block_sum(a_s2, dred);
	if (threadIdx.x == 0)
	{
		st->sumw = a_w; st->wsum_s2off = a_s2;
		st->best_ihid = fs_ihid[st->fs_base + st->fmax_sample];
	}
}

int rbk_collect_pool(rb_ctx *ctx, PoolSlot &s)
{
	k_collect<<<s.P, 256, 0, ctx->stream>>>(s.meta.as<RbPartMeta>(), s.state.as<RbPartState>(), s.fo.as<RbFineOrient>(),
		s.pair_list.as<int>(), s.fs_w.as<float>(), s.fs_ihid.as<long long>(), s.dir_idx.as<int>(),
		ctx->d_model, ctx->d_samp, s.out_pdf_dir.as<double>(), s.out_pdf_class.as<double>(),
		s.out_pdf_class.as<double>() + ctx->d_model.nr_classes, s.counters.as<int>());
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// fused wavg + back-projection over the pool
// ---------------------------------------------------------------------------------------------
struct StoreArgs {
	const RbPartMeta *metas; RbPartState *states;
	const float2 *Fimg, *Fnomask; const float *Fctf;
	const RbFineOrient *fo; const int *pair_list; const int *counters;
	const float *fs_w;
	float *shells;                 // [P][nshell]
	const RbProjector *projs; const RbBackprojector *bps;
	const uint32_t *pix; int npix; int n;
	const float *tx, *ty; int NOT;
	const float2 *slices; long long slice_capacity;   // slices cached by the fine pass ([fine orientation][n][n/2+1])
	int *queue;                    // two work queue counters (sliced launch, gather launch), zero at launch: rb_next_work
};

static const int ST_CHUNK = 256;   // significant samples held in shared memory at a time

// What one lane keeps in flight per pixel.  SLICED: the reference sample comes from the slice the fine pass cached
// (one streaming 8-byte read); otherwise it is gathered from the expanded reference again (64 bytes + lerp state).
template <bool SLICED> struct StoreRef;
template <> struct StoreRef<true> { float2 r; };
template <> struct StoreRef<false> { RbProjFetch pf; };

template <bool SLICED>
struct StorePix {
	StoreRef<SLICED> ref;
	float2 X, X0;
	float ctf;
	uint32_t pkx;
};

template <bool SLICED>
__device__ __forceinline__ void store_issue(const StoreArgs &A, const RbProjK8 &pk, const float2 *X, const float2 *X0,
                                            const float *C, const float2 *slice, float part_scale, int ip,
                                            float e0, float e1, float e3, float e4, float e6, float e7, StorePix<SLICED> &f)
{
	f.pkx = __ldg(A.pix + ip);
	const int x = rb_pix_x(f.pkx), y = rb_pix_y(f.pkx);
	const int idx = rb_src_index(x, y, A.n);
	if constexpr (SLICED) f.ref.r = __ldg(slice + idx);
	else rb_proj_issue(pk, x, y, e0, e1, e3, e4, e6, e7, f.ref.pf);
	f.X = __ldg(X + idx); f.X0 = __ldg(X0 + idx);
	f.ctf = C ? __ldg(C + idx) * part_scale : part_scale;                                     // :3087-3096
}

// Per pixel everything the store stage needs from the translations is  Phi = sum_t wn_t * e^{i phi_t}  and
// W = sum_t wn_t  (wn_t = weight_t / sum_weight over the significant samples of this orientation):
//   XA    = Re(conj(ref) * X * Phi)                      (wavg.cuh:131-133 summed over t)
//   AA    = W * |ref|^2
//   wdiff = W * (|ref|^2 + |X|^2) - 2 * XA               (= sum_t wn_t |ref - S_t X|^2)
//   F     = g * X0 * Phi,  Fweight = W * g * ctf,  g = ctf * Minvsigma2   (BP.cuh:276-299)
// so the per-sample work is one phase factor and two FMAs, and no sample count limit applies.
//
// SLICED launches cover the fine orientations whose slice sits in the cache (w < slice_capacity), the gather variant
// the rest; DEPTH pixels are kept in flight per lane (the stage is latency-bound: 8 scattered reductions per pixel).
template <bool SLICED, int DEPTH, int MINB>
__global__ void __launch_bounds__(ST_THREADS, MINB)
k_store(StoreArgs A, RbModelDev M)
{
	__shared__ float s_ux[ST_CHUNK], s_uy[ST_CHUNK], s_wn[ST_CHUNK];
	__shared__ float s_e[6];
	__shared__ int s_nsig;
	__shared__ float s_W;
	__shared__ int s_sigidx[ST_MAXSAMP];
	__shared__ double dred[32];
	__shared__ float s_shell[1024];

	const int imgX = A.n / 2 + 1;
	const int nfo = A.counters[0];
	const long long cap = A.slices ? A.slice_capacity : 0;
	const int w_begin = SLICED ? 0 : (int) (cap < nfo ? cap : nfo);
	const int w_end = SLICED ? (int) (cap < nfo ? cap : nfo) : nfo;
	const int half = A.n / 2;

	__shared__ int s_next;
	int *queue = A.queue ? A.queue + (SLICED ? 0 : 1) : nullptr;
	for (int wi = rb_next_work(queue, &s_next, 0, true); w_begin + wi < w_end; wi = rb_next_work(queue, &s_next, wi, false))
	{
		const int w = w_begin + wi;
		const RbFineOrient F = A.fo[w];
		const int p = F.particle;
		const RbPartState *st = A.states + p;
		if (st->status != 0) continue;
		const float sig = st->fsig_weight, sumw = st->fsum_weight;
		const int nsamp = F.n_t * A.NOT;
		// ordered list of significant samples (warp 0)
		__syncthreads();
		if (threadIdx.x < 32)
		{
			int cnt = 0;
			for (int j0 = 0; j0 < nsamp; j0 += 32)
			{
				const int j = j0 + threadIdx.x;
				const bool s = j < nsamp && A.fs_w[F.sample_off + j] >= sig;          // wavg.cuh:106 / BP.cuh:278
				const unsigned b = __ballot_sync(RB_FULL_MASK, s);
				if (s) { int pos = cnt + __popc(b & ((1u << threadIdx.x) - 1)); if (pos < ST_MAXSAMP) s_sigidx[pos] = j; }
				cnt += __popc(b);
			}
			if (threadIdx.x == 0) s_nsig = min(cnt, ST_MAXSAMP);
		}
		if (threadIdx.x < 6) s_e[threadIdx.x] = A.fo[w].e[threadIdx.x + threadIdx.x / 2];   // elements 0,1,3,4,6,7
		for (int i = threadIdx.x; i < M.nshell; i += ST_THREADS) s_shell[i] = 0.f;
		__syncthreads();
		const int nsig = s_nsig;
		if (nsig == 0) continue;
		if (threadIdx.x == 0) atomicAdd(&A.states[p].n_bp, 1);

		const RbPartMeta m = A.metas[p];
		const float2 *X = A.Fimg + (size_t) p * M.Npf, *X0 = A.Fnomask + (size_t) p * M.Npf;
		const float *C = A.Fctf ? A.Fctf + (size_t) p * M.Npf : nullptr;
		const float2 *slice = SLICED ? A.slices + (size_t) w * M.Npf : nullptr;
		const float *mtab = M.minvs2 + (size_t) m.og * M.nshell;
		const unsigned char *dvp = M.dvp_gt3 + (size_t) F.iclass * M.nshell;
		const RbProjK8 pk = rb_make_projk8(A.projs[F.iclass], imgX);
		const RbBackprojector bp = A.bps[F.iclass + m.bp_off];
		const int max_r2_vol = (int) (bp.maxR * bp.maxR * bp.padding_factor * bp.padding_factor);   // BP.cuh:209
		const float wni = 1.0f / sumw;
		double aXA = 0., aAA = 0.;

		for (int c0 = 0; c0 < nsig; c0 += ST_CHUNK)
		{
			const int ntr = min(ST_CHUNK, nsig - c0);
			__syncthreads();
			for (int i = threadIdx.x; i < ntr; i += ST_THREADS)
			{
				const int j = s_sigidx[c0 + i];
				const int it = A.pair_list[F.pair_off + j / A.NOT] * A.NOT + (j % A.NOT);
				s_ux[i] = A.tx[it] * 0.15915494309189535f; s_uy[i] = A.ty[it] * 0.15915494309189535f;
				s_wn[i] = A.fs_w[F.sample_off + j] * wni;                              // weight * weight_norm_inverse (wavg.h:138)
			}
			__syncthreads();
			if (threadIdx.x == 0) { float Wt = 0.f; for (int i = 0; i < ntr; i++) Wt += s_wn[i]; s_W = Wt; }
			__syncthreads();
			const float W = s_W;
			const float e0 = s_e[0], e1 = s_e[1], e3 = s_e[2], e4 = s_e[3], e6 = s_e[4], e7 = s_e[5];

			// the trip count is uniform per warp (lanes past the end idle) because the scatter below is cooperative
			const int wbase = threadIdx.x & ~31;
			const int nit = wbase < A.npix ? (A.npix - wbase + ST_THREADS - 1) / ST_THREADS : 0;
			StorePix<SLICED> f[DEPTH];
#pragma unroll
			for (int d = 0; d < DEPTH; d++)
			{
				const int ip = threadIdx.x + d * ST_THREADS;
				if (d < nit && ip < A.npix) store_issue<SLICED>(A, pk, X, X0, C, slice, m.part_scale, ip, e0, e1, e3, e4, e6, e7, f[d]);
			}
			for (int it0 = 0; it0 < nit; it0 += DEPTH)
			{
#pragma unroll
				for (int d = 0; d < DEPTH; d++)
				{
					const int it = it0 + d;
					if (it >= nit) break;
					const int ip = threadIdx.x + it * ST_THREADS;
					const bool have = ip < A.npix;
					const StorePix<SLICED> cur = f[d];
					{
						const int ipn = ip + DEPTH * ST_THREADS;
						if (it + DEPTH < nit && ipn < A.npix)
							store_issue<SLICED>(A, pk, X, X0, C, slice, m.part_scale, ipn, e0, e1, e3, e4, e6, e7, f[d]);
					}

					int cell = -1;           // accumulator voxel of corner (0,0,0), -1: nothing to scatter
					float sfx = 0.f, sfy = 0.f, sfz = 0.f, Fr = 0.f, Fi = 0.f, Fw = 0.f;
					if (have)
					{
						const int x = rb_pix_x(cur.pkx), y = rb_pix_y(cur.pkx), ires = rb_pix_ires(cur.pkx);
						float2 ref;
						if constexpr (SLICED) ref = cur.ref.r;
						else ref = (cur.ref.pf.flags & 1) ? rb_proj_finish(cur.ref.pf) : make_float2(0.f, 0.f);
						const float ctf = cur.ctf;
						const float2 ref_ctf = make_float2(ref.x * ctf, ref.y * ctf);                      // BP.cuh:520-521 (SGD)
						if (M.refs_are_ctf_corrected) { ref.x *= ctf; ref.y *= ctf; }                     // wavg.cuh:96-104
						else { ref.x *= m.part_scale; ref.y *= m.part_scale; }
						float phr = 0.f, phi = 0.f;
						for (int t = 0; t < ntr; t++)
						{
							const float2 ph = rb_phase(x, y, s_ux[t], s_uy[t]);
							const float wn = s_wn[t];
							phr = fmaf(wn, ph.x, phr); phi = fmaf(wn, ph.y, phi);
						}
						const float refn = ref.x * ref.x + ref.y * ref.y;
						const float Xn = cur.X.x * cur.X.x + cur.X.y * cur.X.y;
						const float xa = (ref.x * cur.X.x + ref.y * cur.X.y) * phr - (ref.x * cur.X.y - ref.y * cur.X.x) * phi;
						const float aa = W * refn;
						const float wd = fmaxf(W * (refn + Xn) - 2.f * xa, 0.f);
						// (x = 0, y < 0) is only in the list with --no_map: Mresol excludes it from the shell sums (:3466-3494)
						// the rows beyond references that end inside the window: skipped by wavg (wavg.h:74-82), walked by the back-projection
						const bool in_mresol = !(x == 0 && y < 0) && !(M.dead_maxR > 0 && abs(y) > M.dead_maxR && x != M.dead_maxR);
						if (in_mresol) atomicAdd(&s_shell[ires], wd);
						if (in_mresol && dvp[ires] && M.do_scale_correction) { aXA += (double) xa; aAA += (double) aa; }   // :3473-3479
						// back-projection
						const float minvs2 = M.do_map ? __ldg(mtab + ires) : 1.f;                          // :2586, :3110-3115
						const float g = M.ctf_premultiplied ? minvs2 : ctf * minvs2;                       // BP.cuh:280-289
						Fw = W * g * ctf;
						bool do_bp = Fw > 0.f;
						if (M.bp_circle_bound && !M.do_grad)                                                 // the SGD kernel walks every pixel (BP.h:757-1047)
						{
							const int xmax = (int) sqrtf((float) (half * half - y * y));                  // BP.h:565
							do_bp = do_bp && (x < xmax);
						}
						if (do_bp)
						{
							Fr = (cur.X0.x * phr - cur.X0.y * phi) * g;
							Fi = (cur.X0.x * phi + cur.X0.y * phr) * g;
							if (M.do_grad) { Fr -= ref_ctf.x * (W * g); Fi -= ref_ctf.y * (W * g); }             // sum_t w_t (X_t - CTF A), BP.cuh:540-541
							// position in the accumulator (BP.cuh:301-347)
							float xp = (e0 * x + e1 * y) * bp.padding_factor;
							float yp = (e3 * x + e4 * y) * bp.padding_factor;
							float zp = (e6 * x + e7 * y) * bp.padding_factor;
							if (xp * xp + yp * yp + zp * zp <= (float) max_r2_vol)
							{
								if (xp < 0.f) { xp = -xp; yp = -yp; zp = -zp; Fi = -Fi; }
								const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
								sfx = xp - fx0; sfy = yp - fy0; sfz = zp - fz0;
								cell = (((int) fz0 - bp.mdlInitZ) * bp.mdlY + ((int) fy0 - bp.mdlInitY)) * bp.mdlX + (int) fx0;
							}
						}
					}
					// Cooperative scatter: two lanes per pixel, one per x-neighbour, so that the two 16-byte reductions of a
					// corner pair sit in the same instruction and share one 32-byte sector / L1 wavefront.
#pragma unroll
					for (int h = 0; h < 2; h++)
					{
						const int src = 16 * h + ((threadIdx.x & 31) >> 1);
						const int c = __shfl_sync(RB_FULL_MASK, cell, src);
						const float fx = __shfl_sync(RB_FULL_MASK, sfx, src), fy = __shfl_sync(RB_FULL_MASK, sfy, src), fz = __shfl_sync(RB_FULL_MASK, sfz, src);
						const float vr = __shfl_sync(RB_FULL_MASK, Fr, src), vi = __shfl_sync(RB_FULL_MASK, Fi, src), vw = __shfl_sync(RB_FULL_MASK, Fw, src);
						if (c >= 0)
						{
							const int px = threadIdx.x & 1;
							const float wx = px ? fx : 1.f - fx;
							const float mfy = 1.f - fy, mfz = 1.f - fz;
							float4 *b = bp.vol + (size_t) c + px;
							const size_t sy = bp.mdlX, sz = (size_t) bp.mdlX * bp.mdlY;
							float d2;
							d2 = mfz * mfy * wx; red_add_v4(b, d2 * vr, d2 * vi, d2 * vw);
							d2 = mfz * fy * wx;  red_add_v4(b + sy, d2 * vr, d2 * vi, d2 * vw);
							d2 = fz * mfy * wx;  red_add_v4(b + sz, d2 * vr, d2 * vi, d2 * vw);
							d2 = fz * fy * wx;   red_add_v4(b + sz + sy, d2 * vr, d2 * vi, d2 * vw);
						}
					}
				}
			}
		}
		__syncthreads();
		for (int i = threadIdx.x; i < M.nshell; i += ST_THREADS)
		{
			const float v = s_shell[i];
			if (v != 0.f) atomicAdd(A.shells + (size_t) p * M.nshell + i, v);
		}
		aXA = block_sum(aXA, dred);
		aAA = block_sum(aAA, dred);
		if (threadIdx.x == 0 && (aXA != 0. || aAA != 0.))
		{
			atomicAdd(&A.states[p].wsum_XA, aXA);
			atomicAdd(&A.states[p].wsum_AA, aAA);
		}
	}
}

int rbk_store_pool(rb_ctx *ctx, PoolSlot &s)
{
	StoreArgs A;
	memset(&A, 0, sizeof(A));
	A.metas = s.meta.as<RbPartMeta>(); A.states = s.state.as<RbPartState>();
	A.Fimg = s.Fimg.as<float2>(); A.Fnomask = s.Fnomask.as<float2>();
	A.Fctf = ctx->h_model.do_ctf_correction ? s.Fctf.as<float>() : nullptr;
	A.fo = s.fo.as<RbFineOrient>(); A.pair_list = s.pair_list.as<int>(); A.counters = s.counters.as<int>();
	A.fs_w = s.fs_w.as<float>(); A.shells = s.shells.as<float>();
	A.projs = ctx->d_proj.as<RbProjector>(); A.bps = ctx->d_bp.as<RbBackprojector>();
	A.pix = ctx->d_model.pix_store; A.npix = ctx->d_model.nv_store; A.n = ctx->d_model.current_size;
	A.tx = ctx->d_samp.ftx; A.ty = ctx->d_samp.fty; A.NOT = ctx->d_samp.n_over_trans;
	A.slices = s.slices.as<float2>(); A.slice_capacity = s.slice_capacity;
	A.queue = s.counters.as<int>() + 9;
	// measured on the headline pool (tools/sweep_variants.sh): depth 2 / 2 CTAs per SM 5.66 ms, depth 3 / 3 CTAs 5.93 ms,
	// depth 4 / 3 CTAs 6.10 ms: the stage is bound by the reductions' L2 round trips, not by load latency
	if (A.slices && A.slice_capacity > 0)
	{
		k_store<true, 2, 2><<<ctx->num_sms * 2, ST_THREADS, 0, ctx->stream>>>(A, ctx->d_model);
		RB_LAUNCH_CHECK(ctx);
	}
	// fine orientations beyond the slice cache (none at the default budget unless the pool is very large)
	k_store<false, 2, 2><<<ctx->num_sms * 2, ST_THREADS, 0, ctx->stream>>>(A, ctx->d_model);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// stage entry points: reference-style dense inputs (one CTA per orientation, all translations)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_wavg_stage(RbProjector pj, int n, const float *eulers, const float *tx, const float *ty, int T,
             const float *re, const float *im, const float *weights, const float *ctfs,
             float weight_norm, float sig_w, float *parts, float *AA, float *XA)
{
	const int imgX = n / 2 + 1;
	const RbProjK pk = rb_make_projk(pj, imgX);
	const int o = blockIdx.x;
	const float *e = eulers + (size_t) o * 9;
	const float wni = 1.0f / weight_norm;
	for (int pix = threadIdx.x; pix < n * imgX; pix += blockDim.x)
	{
		int x = pix % imgX, iy = pix / imgX, y = iy;
		if (iy > pk.maxR)                                                                    // wavg.cuh:81-87
		{
			if (iy >= n - pk.maxR) y = iy - n;
			else if (x != pk.maxR) continue;   // ALTCPU visits only x = maxR in the dead band (wavg.h:74-82)
		}
		float2 ref = rb_project3d(pk, x, y, e[0], e[1], e[3], e[4], e[6], e[7]);
		const float ctf = ctfs[pix];
		ref.x *= ctf; ref.y *= ctf;
		const float ir = re[pix], ii = im[pix];
		float wd = 0.f, xa = 0.f, aa = 0.f;
		for (int t = 0; t < T; t++)
		{
			float w = weights[(size_t) o * T + t];
			if (w < sig_w) continue;
			w *= wni;
			float s, c; sincosf(x * tx[t] + y * ty[t], &s, &c);                              // translatePixel
			const float tr = c * ir - s * ii, ti = c * ii + s * ir;
			const float dr = ref.x - tr, di = ref.y - ti;
			wd += w * (dr * dr + di * di);
			xa += w * (ref.x * tr + ref.y * ti);
			aa += w * (ref.x * ref.x + ref.y * ref.y);
		}
		atomicAdd(parts + pix, wd); atomicAdd(XA + pix, xa); atomicAdd(AA + pix, aa);
	}
}

int rbk_wavg_stage(rb_ctx *ctx, const RbProjector &pj, int n, const float *d_eulers, int O,
                   const float *d_tx, const float *d_ty, int T, const float *d_re, const float *d_im,
                   const float *d_w, const float *d_ctf, float weight_norm, float sig_w,
                   float *d_parts, float *d_AA, float *d_XA)
{
	if (O < 1) return RB_OK;
	k_wavg_stage<<<O, 256, 0, ctx->stream>>>(pj, n, d_eulers, d_tx, d_ty, T, d_re, d_im, d_w, d_ctf, weight_norm, sig_w, d_parts, d_AA, d_XA);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

__global__ void __launch_bounds__(128)
k_backproject_stage(RbBackprojector bp, int n, const float *eulers, const float *tx, const float *ty, int T,
                    const float *re, const float *im, const float *weights, const float *minvs2s, const float *ctfs,
                    float weight_norm, float sig_w, int circle_bound, int ctf_premultiplied)
{
	const int imgX = n / 2 + 1, half = n / 2;
	const int o = blockIdx.x;
	const float *e = eulers + (size_t) o * 9;
	const int max_r2_vol = (int) (bp.maxR * bp.maxR * bp.padding_factor * bp.padding_factor);
	const float wni = 1.0f / weight_norm;
	for (int pix = threadIdx.x; pix < n * imgX; pix += blockDim.x)
	{
		int x = pix % imgX, iy = pix / imgX, y = iy > half ? iy - n : iy;                    // BP.cuh:258-261
		if (circle_bound) { int xmax = (int) sqrtf((float) (half * half - y * y)); if (x >= xmax) continue; }
		const float minvs2 = minvs2s[pix], ctf = ctfs[pix], ir = re[pix], ii = im[pix];
		float Fr = 0.f, Fi = 0.f, Fw = 0.f;
		for (int t = 0; t < T; t++)
		{
			const float w = weights[(size_t) o * T + t];
			if (w < sig_w) continue;
			float myw = ctf_premultiplied ? w * (wni * minvs2) : w * (wni * ctf * minvs2);
			Fw += myw * ctf;
			float s, c; sincosf(x * tx[t] + y * ty[t], &s, &c);
			Fr += (c * ir - s * ii) * myw;
			Fi += (c * ii + s * ir) * myw;
		}
		if (Fw > 0.f) bp_scatter(bp, max_r2_vol, x, y, e[0], e[1], e[3], e[4], e[6], e[7], Fr, Fi, Fw);
	}
}

int rbk_backproject_stage(rb_ctx *ctx, const RbBackprojector &bp, int n, const float *d_eulers, int O,
                          const float *d_tx, const float *d_ty, int T, const float *d_re, const float *d_im,
                          const float *d_w, const float *d_minvs2, const float *d_ctf,
                          float weight_norm, float sig_w, int circle_bound, int ctf_premultiplied)
{
	if (O < 1) return RB_OK;
	k_backproject_stage<<<O, 128, 0, ctx->stream>>>(bp, n, d_eulers, d_tx, d_ty, T, d_re, d_im, d_w, d_minvs2, d_ctf,
		weight_norm, sig_w, circle_bound, ctf_premultiplied);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}
