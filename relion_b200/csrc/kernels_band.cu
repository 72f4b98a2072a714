// relion_b200 — band-major fine pass and store stage (sm_100a): the volume traffic of the two HBM-bound stages made
// L2-resident.
//
// Replaces, for a whole pool: cuda_kernel_diff2_fine (/root/reference/src/acc/cuda/cuda_kernels/diff2.cuh:193-332),
// cuda_kernel_wavg (wavg.cuh:13-152) and cuda_kernel_backproject3D (BP.cuh:174-403); ALTCPU twins
// cpu_kernels/diff2.h:284-430, wavg.h:22-199, BP.h:497-753.
//
// Why.  A 512-particle pool at 256 px makes 1.8e8 trilinear samples into a half sphere of 3.5e7 voxels: every 64-byte cell
// of the expanded reference is needed ~5 times per pool, every accumulator voxel is hit ~10 times.  When a CTA walks one
// slice (orientation-major order), concurrent CTAs touch unrelated parts of Fourier space, nothing is reused in the 126 MB
// L2 and every sample costs a random 64-byte DRAM read (3.1 TB/s ceiling, 1.55x the algorithmic bytes) or a DRAM
// read-modify-write.  Here the work is ordered RADIAL-BAND-MAJOR: the pixel list is sorted by |r|, cut into tiles of
// BD_TP pixels (a ring 0.3 - 0.6 pixels thick at the edge of a 256-px image), and the work queue runs over
// (tile, chunk of orientations) with the tile as the slow index.  All resident CTAs therefore work inside one thin
// spherical shell at a time: ~45 MB of reference cells, ~15 MB of accumulator - the shell is read from HBM once, hit in
// L2 by every orientation of the pool, and (store stage) written back once.
//
//   k_prep_sorted     particle images -> band-ordered arrays (prepared image for diff2; X, X0, ctf for the store stage)
//   k_project_band    fine orientations x tile -> slices in band order.  One lane owns one sample (address arithmetic once
//                     per sample), the 64-byte cell is fetched by the lane's QUAD with four 16-byte loads (one L1 line visit
//                     per sample) and handed over through a conflict-free shared-memory tile; one sample ahead in flight.
//   k_diff2_slices*   streams slice + prepared image per fine orientation: diff2 of all its fine translations; phases from
//                     L2-resident tables when the oversampled translations factorise (no transcendental in the loop)
//   k_bp_*            compact list of the fine orientations holding >= 1 significant sample, with (phase, weight) tables
//   k_store_band      (tile, chunk of those orientations): wavg sums + trilinear scatter (red.global.add.v4.f32) into the
//                     L2-resident shell of the accumulator
#include "img_src.cuh"
#include <cstdlib>
#include <algorithm>

static const int BD_THREADS = 256;
static const int BD_TP = 128;                    // pixels per tile
static const int BD_WPT = BD_TP / 32;            // warps per tile row
static const int BD_NPH = BD_THREADS / BD_TP;    // orientation phases per CTA
static const int BD_MAXCHUNK = 128;

// Band-major arrays are TILE-major: element (row, pixel ip) of an array with `nrows` rows (fine orientations of the round,
// particles) sits at ((ip / BD_TP) * nrows + row) * BD_TP + ip % BD_TP.  What the resident CTAs of a band sweep touch at one
// time - one tile of every row - is then a contiguous range (a few 2 MB pages) instead of one 1 KB piece out of every row of
// a 1.4 GB array (700 pages: every access a TLB miss that also evicts the pages of the reference shell).
__device__ __forceinline__ size_t bd_at(int row, int ip, int nrows) { return ((size_t) (ip >> 7) * (size_t) nrows + (size_t) row) * BD_TP + (size_t) (ip & (BD_TP - 1)); }

__device__ __forceinline__ void bd_red_add_v4(float4 *addr, float a, float b, float c)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(0.f) : "memory");
}

// ---------------------------------------------------------------------------------------------
// band-ordered particle images
// ---------------------------------------------------------------------------------------------
struct BandPrepArgs {
	const RbPartMeta *metas; const float2 *Fimg, *Fnomask; const float *Fctf;
	const uint32_t *pix; int nd2, nst, stride;
	float4 *simg4;      // [P][stride] (X'.re, X'.im, corr / 2, 0): diff2 passes
	float4 *sst;        // tile-major [stride / BD_TP][P][BD_TP] (X.re, X.im, X0.re, X0.im): store stage
	float *sctf;        // tile-major, ctf * part_scale
};

// WHICH: 0 = the diff2 pass' prepared image (fine stage), 1 = the store stage's X, X0, ctf
template <int WHICH>
static __global__ void __launch_bounds__(256)
k_prep_sorted(BandPrepArgs A, RbModelDev M)
{
	const int p = blockIdx.y;
	const RbPartMeta m = A.metas[p];
	ImgSrc src;
	img_src_pool(src, M, m, A.Fimg, A.Fctf, p);
	const float2 *X = A.Fimg + (size_t) p * M.Npf, *X0 = A.Fnomask + (size_t) p * M.Npf;
	const float *C = A.Fctf ? A.Fctf + (size_t) p * M.Npf : nullptr;
	for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < A.stride; ip += gridDim.x * blockDim.x)
	{
		float4 d = make_float4(0.f, 0.f, 0.f, 0.f), s = d;
		float c = 0.f;
		if (ip < A.nst)
		{
			const uint32_t pk = __ldg(A.pix + ip);
			const int x = rb_pix_x(pk), y = rb_pix_y(pk), ires = rb_pix_ires(pk);
			const int idx = rb_src_index(x, y, M.current_size);
			if (WHICH == 0 && ip < A.nd2)
			{
				float2 Xc; float corr;
				img_load_idx(src, idx, ires, Xc, corr);
				d = make_float4(Xc.x, Xc.y, corr * 0.5f, 0.f);
			}
			if (WHICH == 1)
			{
				const float2 a = __ldg(X + idx), b = __ldg(X0 + idx);
				s = make_float4(a.x, a.y, b.x, b.y);
				c = C ? __ldg(C + idx) * m.part_scale : m.part_scale;                              // acc_ml_optimiser_impl.h:3087-3096
			}
		}
		if (WHICH == 0) A.simg4[(size_t) p * A.stride + ip] = d;
		else
		{
			A.sst[bd_at(p, ip, (int) gridDim.y)] = s;
			A.sctf[bd_at(p, ip, (int) gridDim.y)] = c;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// projection, band-major
// ---------------------------------------------------------------------------------------------
struct BandProjArgs {
	const RbFineOrient *fo;
	const int *indir;            // nullptr: orientation j of the round is fo[begin + j]; else fo[indir[begin + j].w] (RbBpItem list)
	const int *count_ptr;        // number of list entries (device): counters[0] or the BP list length
	int begin, capacity;         // this round covers list entries [begin, min(count, begin + capacity))
	const uint32_t *pix; int npix, stride;
	float2 *slices;              // tile-major [stride / BD_TP][n][BD_TP], n = orientations of the round
	const RbProjector *projs; int imgX; int nr_classes;
	int *queue;
	int chunk_min;
	const int *nfo_ptr; int only_if_nfo_above;   // re-projection rounds of the store stage: run only when the fine pass needed several rounds
};

struct RbBpItem { int w; int samp_off; int nsig; float W; };

static const int BD_STAGE_ROW = 33;           // float2 per piece row of a warp's exchange tile: piece k of sample s sits at k * 33 + s, which makes
                                                // both the quad-wise 8-byte writes and the lane-wise 8-byte reads conflict free (2 wavefronts per 256 bytes)

struct BandProjSmem {
	float2 cell[BD_THREADS / 32][4 * BD_STAGE_ROW];
	float e[BD_MAXCHUNK][6];
	int cls[BD_MAXCHUNK];
	int next;
};

// what a lane keeps about a sample in flight: trilinear fractions and flags (bit0 inside r_max, bit1 Hermitian mate),
// and the four 16-byte pieces it fetched for its QUAD: piece r is quarter (lane & 3) of the cell of the quad's sample r
struct BandFrac { float fx, fy, fz; int flags; };
struct BandLoad { float4 q[4]; BandFrac f; };

// Gathers go through plain 16-byte loads (LDG.128): per instruction the 8 quads of a warp touch 8 cells, one L1 line visit
// per sample.  (cp.async into shared memory was measured 2.6x slower here: with 8 distinct lines per instruction every line
// becomes its own shared-memory write transaction, ~40 cycles per instruction against ~16 for the register path.)
// Two steps, so that the block-table look-up of sample j + 2 (RbProjector::blk) is in flight while the cell of sample j + 1 is
// being fetched and sample j is interpolated.
struct BandAddr { BandFrac f; uint32_t rank; int sub; };

template <bool MULTI>
__device__ __forceinline__ void band_addr(const BandProjArgs &A, const BandProjSmem &S, const RbProjK8 &pk0, int j, int x, int y, bool have, BandAddr &a)
{
	const float2 ea = *(const float2 *) &S.e[j][0], eb = *(const float2 *) &S.e[j][2], ec = *(const float2 *) &S.e[j][4];
	RbProjK8 pk = pk0;
	if (MULTI) pk = rb_make_projk8(A.projs[S.cls[j]], A.imgX);
	float xp = (ea.x * x + ea.y * y) * pk.pf;
	float yp = (eb.x * x + eb.y * y) * pk.pf;
	float zp = (ec.x * x + ec.y * y) * pk.pf;
	const int r2 = (int) (xp * xp + yp * yp + zp * zp);
	const bool inside = have && r2 <= pk.maxR2_padded;
	const bool inv = xp < 0.f;
	if (inv) { xp = -xp; yp = -yp; zp = -zp; }
	const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
	a.f.fx = xp - fx0; a.f.fy = yp - fy0; a.f.fz = zp - fz0;
	a.f.flags = (inside ? 1 : 0) | (inv ? 2 : 0);
	const int x0 = (int) fx0, yi = (int) fy0 - pk.mdlInitY, zi = (int) fz0 - pk.mdlInitZ;
	a.sub = ((zi & 3) << 4) | ((yi & 3) << 2) | (x0 & 3);
	a.rank = inside ? __ldg(pk.blk + rb_blk_slot(pk.nbx, pk.nbxy, x0 >> 2, yi >> 2, zi >> 2)) : 0u;
}

template <bool MULTI>
__device__ __forceinline__ void band_fetch(const BandProjArgs &A, const BandProjSmem &S, const RbProjK8 &pk0, int j, const BandAddr &a, int lane, BandLoad &L)
{
	const float4 *mdl8 = pk0.mdl8;
	if (MULTI) mdl8 = A.projs[S.cls[j]].mdl8;
	L.f = a.f;
	const int cell = (a.f.flags & 1) ? (int) ((a.rank << 6) | (uint32_t) a.sub) : -1;
	const int k = lane & 3, qbase = lane & ~3;
#pragma unroll
	for (int r = 0; r < 4; r++)
	{
		const int cs = __shfl_sync(RB_FULL_MASK, cell, qbase + r);             // cell of the quad's sample r
		L.q[r] = cs >= 0 ? __ldcg(mdl8 + 4 * (size_t) cs + k) : make_float4(0.f, 0.f, 0.f, 0.f);
	}
}

// The lane holds piece k = (lane & 3) of the cells of its quad's four samples: (z, y) corner pair k, voxels x and x + 1.  It
// does the x-interpolation of those four pieces itself (the owner's fx comes by shuffle; same arithmetic, so the result is
// bit-identical to interpolating on the owner), which halves what has to cross the quad: 8-byte (re, im) values go through
// the warp's exchange tile, and the owner finishes with the y- and z-interpolation of its own sample.
__device__ __forceinline__ float2 band_consume(float2 *tile, int lane, const BandLoad &L)
{
	const int k = lane & 3, qbase = lane & ~3;
#pragma unroll
	for (int r = 0; r < 4; r++)
	{
		const float fxr = __shfl_sync(RB_FULL_MASK, L.f.fx, qbase + r);
		const float4 q = L.q[r];
		tile[k * BD_STAGE_ROW + qbase + r] = make_float2(q.x + (q.z - q.x) * fxr, q.y + (q.w - q.y) * fxr);
	}
	__syncwarp();
	const float2 d00 = tile[lane], d10 = tile[BD_STAGE_ROW + lane], d01 = tile[2 * BD_STAGE_ROW + lane], d11 = tile[3 * BD_STAGE_ROW + lane];
	__syncwarp();
	const BandFrac &f = L.f;
	float2 ref;
	{
		const float dxy0 = d00.x + (d10.x - d00.x) * f.fy, dxy1 = d01.x + (d11.x - d01.x) * f.fy;
		ref.x = dxy0 + (dxy1 - dxy0) * f.fz;
	}
	{
		const float dxy0 = d00.y + (d10.y - d00.y) * f.fy, dxy1 = d01.y + (d11.y - d01.y) * f.fy;
		ref.y = dxy0 + (dxy1 - dxy0) * f.fz;
	}
	if (f.flags & 2) ref.y = -ref.y;
	if (!(f.flags & 1)) ref = make_float2(0.f, 0.f);
	return ref;
}

template <bool MULTI, int MINB>
static __global__ void __launch_bounds__(BD_THREADS, MINB)
k_project_band(BandProjArgs A)
{
	__shared__ BandProjSmem S;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int total = *A.count_ptr;
	const int n = min(A.capacity, total - A.begin);
	if (n <= 0) return;
	if (A.nfo_ptr && *A.nfo_ptr <= A.only_if_nfo_above) return;
	const int ntiles = (A.npix + BD_TP - 1) / BD_TP;
	// chunk: about one tile's worth of items per resident wave, so that the wave stays inside one band
	int chunk = (n + (int) gridDim.x - 1) / (int) gridDim.x;
	chunk = max(A.chunk_min, min(BD_MAXCHUNK, chunk));
	const int nchunks = (n + chunk - 1) / chunk;
	const long long nitems = (long long) ntiles * nchunks;
	const RbProjK8 pk0 = rb_make_projk8(A.projs[0], A.imgX);
	const int ph = wid / BD_WPT;
	float2 *tile_buf = &S.cell[wid][0];

	long long item = blockIdx.x;
	while (item < nitems)
	{
		const int tile = (int) (item / nchunks), c = (int) (item - (long long) tile * nchunks);
		const int o0 = c * chunk, no = min(chunk, n - o0);
		// orientation matrices of the chunk
		for (int i = threadIdx.x; i < no * 6; i += BD_THREADS)
		{
			const int j = i / 6, q = i - j * 6;
			const int li = A.begin + o0 + j;
			const int w = A.indir ? A.indir[4 * li] : li;
			S.e[j][q] = A.fo[w].e[q + q / 2];                                   // elements 0,1,3,4,6,7
			if (MULTI && q == 0) S.cls[j] = A.fo[w].iclass;
		}
		// the next item is requested now and looked at after this one is done: the atomic's round trip is hidden
		int pending = 0;
		if (threadIdx.x == 0) pending = atomicAdd(A.queue, 1);
		__syncthreads();
		const int ip = tile * BD_TP + (wid % BD_WPT) * 32 + lane;
		const bool have = ip < A.npix;
		int x = 0, y = 0;
		if (have) { const uint32_t pkx = __ldg(A.pix + ip); x = rb_pix_x(pkx); y = rb_pix_y(pkx); }
		const int nmy = (no - ph + BD_NPH - 1) / BD_NPH;                        // orientations ph, ph + BD_NPH, ...

		// two orientations ahead: table look-up of sample jj + 2, cell loads of sample jj + 1, interpolation of sample jj
		BandAddr a1, a2;
		BandLoad cur, nxt;
		if (nmy > 0) { band_addr<MULTI>(A, S, pk0, ph, x, y, have, a1); band_fetch<MULTI>(A, S, pk0, ph, a1, lane, cur); }
		if (nmy > 1) band_addr<MULTI>(A, S, pk0, ph + BD_NPH, x, y, have, a1);
		for (int jj = 0; jj < nmy; jj++)
		{
			if (jj + 2 < nmy) band_addr<MULTI>(A, S, pk0, ph + (jj + 2) * BD_NPH, x, y, have, a2);
			if (jj + 1 < nmy) band_fetch<MULTI>(A, S, pk0, ph + (jj + 1) * BD_NPH, a1, lane, nxt);
			const float2 ref = band_consume(tile_buf, lane, cur);
			if (have) __stcs(A.slices + bd_at(o0 + ph + jj * BD_NPH, ip, n), ref);
			cur = nxt; a1 = a2;
		}
		if (threadIdx.x == 0) S.next = (int) gridDim.x + pending;
		__syncthreads();
		item = S.next;
	}
}

// ---------------------------------------------------------------------------------------------
// diff2 of every fine sample from the band-ordered slices (streaming)
// ---------------------------------------------------------------------------------------------
struct BandDiffArgs {
	const RbPartMeta *metas; RbPartState *states;
	const RbFineOrient *fo; const int *pair_list; const int *counters;
	float *fs_w;
	const float4 *simg4; const float2 *slices; const uint32_t *pix; int nd2, stride;
	int begin, capacity;
	const float *tx, *ty; int NOT;
	int *queue;
};

template <int NT>
__device__ __forceinline__ void band_diff_pass(const float2 *__restrict__ slices, int row, int nrows, const float4 *__restrict__ img, const uint32_t *__restrict__ pix,
                                               int nd2, const float *s_ux, const float *s_uy, float (&acc)[NT], float &base)
{
#pragma unroll
	for (int t = 0; t < NT; t++) acc[t] = 0.f;
	base = 0.f;
	for (int ip = threadIdx.x; ip < nd2; ip += BD_THREADS)
	{
		const uint32_t pk = __ldg(pix + ip);
		const int x = rb_pix_x(pk), y = rb_pix_y(pk);
		const float2 ref = __ldcs(slices + bd_at(row, ip, nrows));
		const float4 im = __ldg(img + ip);
		const float hc = im.z;
		const float zr = hc * (ref.x * im.x + ref.y * im.y);
		const float zi = hc * (ref.x * im.y - ref.y * im.x);
		base += hc * ((ref.x * ref.x + ref.y * ref.y) + (im.x * im.x + im.y * im.y));
		const float fx = (float) x, fy = (float) y;
#pragma unroll
		for (int t = 0; t < NT; t++)
		{
			float u = fmaf(fx, s_ux[t], fy * s_uy[t]);
			u -= rintf(u);
			float s, c;
			__sincosf(6.283185307179586f * u, &s, &c);
			acc[t] = fmaf(zr, c, fmaf(-zi, s, acc[t]));
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Phase tables.  RELION's oversampled translations are the coarse ones plus a fixed set of offsets
// (HealpixSampling::getTranslationsInPixel, src/healpix_sampling.cpp:1724-1790): over[t * NOT + j] = trans[t] + d[j], so the
// phase of fine translation (t, j) at a pixel factorises: e^{i phi_t} e^{i phi_j}.  Both factors are tabulated once per
// (model, sampling) in band order, [t][pixel] and [j][pixel] (fp64 sincos): 6 MB + 0.8 MB at 256 px / 29 translations,
// L2-resident, read coalesced.  The diff2 pass then needs no transcendental at all: per (pixel, coarse translation) one
// 8-byte load and a complex product, per fine sample two FMAs.  rb_set_sampling checks the factorisation; samplings that do
// not factorise (or NOT not in {1, 4}) take the generic kernel, which evaluates sincos per (pixel, sample).
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
k_band_tables(const uint32_t *pix, int npix, int stride, const double *ucx, const double *ucy, int T, const double *uox, const double *uoy, int NOT,
              float2 *tabc, float2 *tabo)
{
	const int ip = blockIdx.x * blockDim.x + threadIdx.x;
	if (ip >= stride) return;
	const int t = blockIdx.y;
	double x = 0., y = 0.;
	if (ip < npix) { const uint32_t pk = pix[ip]; x = rb_pix_x(pk); y = rb_pix_y(pk); }
	double u = (t < T) ? x * ucx[t] + y * ucy[t] : x * uox[t - T] + y * uoy[t - T];   // turns
	u -= rint(u);
	double sn, cs;
	sincospi(2. * u, &sn, &cs);
	float2 *dst = (t < T) ? tabc + (size_t) t * stride : tabo + (size_t) (t - T) * stride;
	dst[ip] = make_float2((float) cs, (float) sn);
}

template <int NC, int NOT>
__device__ __forceinline__ void band_diff_pass_sep(const float2 *__restrict__ slices, int row, int nrows, const float4 *__restrict__ img, int nd2, int stride,
                                                   const float2 *__restrict__ tabc, const float2 *__restrict__ tabo, const int *s_tc,
                                                   float (&acc)[NC * NOT], float &base)
{
#pragma unroll
	for (int t = 0; t < NC * NOT; t++) acc[t] = 0.f;
	base = 0.f;
	// two neighbouring pixels per thread and step: slice, phase tables and image come as 16-byte loads (the band arrays are
	// padded to whole tiles and 16-byte aligned at even pixels), half the load instructions of a pixel-per-thread loop
	for (int ip = 2 * threadIdx.x; ip < nd2; ip += 2 * BD_THREADS)
	{
		const bool two = ip + 1 < nd2;
		float4 rr = __ldcs((const float4 *) (slices + bd_at(row, ip, nrows)));
		const float4 ia = __ldg(img + ip);
		float4 ib = __ldg(img + ip + 1);
		if (!two) { rr.z = 0.f; rr.w = 0.f; ib = make_float4(0.f, 0.f, 0.f, 0.f); }          // the pad entry of the slice is not written
		float4 po[NOT];
#pragma unroll
		for (int j = 0; j < NOT; j++) po[j] = NOT > 1 ? __ldg((const float4 *) (tabo + (size_t) j * stride + ip)) : make_float4(1.f, 0.f, 1.f, 0.f);
		float4 pc[NC];
#pragma unroll
		for (int c = 0; c < NC; c++) pc[c] = __ldg((const float4 *) (tabc + (size_t) s_tc[c] * stride + ip));
		const float zra = ia.z * (rr.x * ia.x + rr.y * ia.y), zia = ia.z * (rr.x * ia.y - rr.y * ia.x);
		const float zrb = ib.z * (rr.z * ib.x + rr.w * ib.y), zib = ib.z * (rr.z * ib.y - rr.w * ib.x);
		base += ia.z * ((rr.x * rr.x + rr.y * rr.y) + (ia.x * ia.x + ia.y * ia.y));
		base += ib.z * ((rr.z * rr.z + rr.w * rr.w) + (ib.x * ib.x + ib.y * ib.y));
#pragma unroll
		for (int c = 0; c < NC; c++)
		{
			const float cra = zra * pc[c].x - zia * pc[c].y, cia = zra * pc[c].y + zia * pc[c].x;     // Z e^{i phi_c}
			const float crb = zrb * pc[c].z - zib * pc[c].w, cib = zrb * pc[c].w + zib * pc[c].z;
#pragma unroll
			for (int j = 0; j < NOT; j++)
			{
				float a = acc[c * NOT + j];
				a = fmaf(cra, po[j].x, fmaf(-cia, po[j].y, a));                                          // Re(Z e^{i phi_c} e^{i phi_j})
				a = fmaf(crb, po[j].z, fmaf(-cib, po[j].w, a));
				acc[c * NOT + j] = a;
			}
		}
	}
}

static const int BD_NCMAX = 8;   // coarse translations per pass of the factorised kernel

template <int NOT>
static __global__ void __launch_bounds__(BD_THREADS, 2)
k_diff2_slices_sep(BandDiffArgs A, const float2 *tabc, const float2 *tabo)
{
	__shared__ int s_tc[BD_NCMAX];
	__shared__ float s_red[BD_THREADS / 32][BD_NCMAX * NOT + 1];
	__shared__ int s_next;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int total = A.counters[0];
	const int n = min(A.capacity, total - A.begin);
	for (int wi = rb_next_work(A.queue, &s_next, 0, true); wi < n; wi = rb_next_work(A.queue, &s_next, wi, false))
	{
		const int w = A.begin + wi;
		const RbFineOrient F = A.fo[w];
		const int p = F.particle;
		const float xi2_half = A.metas[p].xi2_half;
		const float4 *img = A.simg4 + (size_t) p * A.stride;
		float bmin = FLT_MAX;
		for (int c0 = 0; c0 < F.n_t; c0 += BD_NCMAX)
		{
			const int nc = min(BD_NCMAX, F.n_t - c0);
			__syncthreads();
			if (threadIdx.x < BD_NCMAX) s_tc[threadIdx.x] = A.pair_list[F.pair_off + c0 + min((int) threadIdx.x, nc - 1)];   // padding repeats the last one
			__syncthreads();
			float acc[BD_NCMAX * NOT], base;
			if (nc == 1)
			{
				float a[NOT]; band_diff_pass_sep<1, NOT>(A.slices, wi, n, img, A.nd2, A.stride, tabc, tabo, s_tc, a, base);
#pragma unroll
				for (int t = 0; t < BD_NCMAX * NOT; t++) acc[t] = t < NOT ? a[t % NOT] : 0.f;
			}
			else if (nc == 2)
			{
				float a[2 * NOT]; band_diff_pass_sep<2, NOT>(A.slices, wi, n, img, A.nd2, A.stride, tabc, tabo, s_tc, a, base);
#pragma unroll
				for (int t = 0; t < BD_NCMAX * NOT; t++) acc[t] = t < 2 * NOT ? a[t % (2 * NOT)] : 0.f;
			}
			else if (nc <= 4)
			{
				float a[4 * NOT]; band_diff_pass_sep<4, NOT>(A.slices, wi, n, img, A.nd2, A.stride, tabc, tabo, s_tc, a, base);
#pragma unroll
				for (int t = 0; t < BD_NCMAX * NOT; t++) acc[t] = t < 4 * NOT ? a[t % (4 * NOT)] : 0.f;
			}
			else band_diff_pass_sep<BD_NCMAX, NOT>(A.slices, wi, n, img, A.nd2, A.stride, tabc, tabo, s_tc, acc, base);
			const int ntr = nc * NOT;
			// fixed-order reduction: lanes, then warps
#pragma unroll
			for (int t = 0; t < BD_NCMAX * NOT; t++)
			{
				if (t < ntr)
				{
					const float v = warp_sum(acc[t]);
					if (lane == 0) s_red[wid][t] = v;
				}
			}
			base = warp_sum(base);
			if (lane == 0) s_red[wid][BD_NCMAX * NOT] = base;
			__syncthreads();
			if (threadIdx.x < ntr)
			{
				float c = 0.f, b = 0.f;
#pragma unroll
				for (int ww = 0; ww < BD_THREADS / 32; ww++) { c += s_red[ww][threadIdx.x]; b += s_red[ww][BD_NCMAX * NOT]; }
				const float v = fmaxf((b - 2.f * c) + xi2_half, 0.f);
				A.fs_w[F.sample_off + (long long) c0 * NOT + threadIdx.x] = v; bmin = fminf(bmin, v);
			}
		}
		if (threadIdx.x < BD_NCMAX * NOT && bmin < FLT_MAX) rb_atomic_min_pos(&A.states[p].fmin_bits, bmin);
	}
}

static const int BD_TF = 32;   // translations per pass

static __global__ void __launch_bounds__(BD_THREADS, 3)
k_diff2_slices(BandDiffArgs A)
{
	__shared__ float s_ux[BD_TF], s_uy[BD_TF];
	__shared__ float s_red[BD_THREADS / 32][BD_TF + 1];
	__shared__ int s_next;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int total = A.counters[0];
	const int n = min(A.capacity, total - A.begin);
	for (int wi = rb_next_work(A.queue, &s_next, 0, true); wi < n; wi = rb_next_work(A.queue, &s_next, wi, false))
	{
		const int w = A.begin + wi;
		const RbFineOrient F = A.fo[w];
		const int nsamp = F.n_t * A.NOT, p = F.particle;
		const float xi2_half = A.metas[p].xi2_half;
		const float4 *img = A.simg4 + (size_t) p * A.stride;
		float bmin = FLT_MAX;
		for (int c0 = 0; c0 < nsamp; c0 += BD_TF)
		{
			const int ntr = min(BD_TF, nsamp - c0);
			__syncthreads();
			if (threadIdx.x < BD_TF)
			{
				float ux = 0.f, uy = 0.f;
				if (threadIdx.x < ntr)
				{
					const int j = c0 + threadIdx.x;
					const int it = A.pair_list[F.pair_off + j / A.NOT] * A.NOT + (j % A.NOT);
					ux = A.tx[it] * 0.15915494309189535f; uy = A.ty[it] * 0.15915494309189535f;   // radians -> turns per pixel
				}
				s_ux[threadIdx.x] = ux; s_uy[threadIdx.x] = uy;
			}
			__syncthreads();
			float acc[BD_TF], base;
			if (ntr <= 4)
			{
				float a4[4]; band_diff_pass<4>(A.slices, wi, n, img, A.pix, A.nd2, s_ux, s_uy, a4, base);
#pragma unroll
				for (int t = 0; t < BD_TF; t++) acc[t] = t < 4 ? a4[t & 3] : 0.f;
			}
			else if (ntr <= 8)
			{
				float a8[8]; band_diff_pass<8>(A.slices, wi, n, img, A.pix, A.nd2, s_ux, s_uy, a8, base);
#pragma unroll
				for (int t = 0; t < BD_TF; t++) acc[t] = t < 8 ? a8[t & 7] : 0.f;
			}
			else if (ntr <= 16)
			{
				float a16[16]; band_diff_pass<16>(A.slices, wi, n, img, A.pix, A.nd2, s_ux, s_uy, a16, base);
#pragma unroll
				for (int t = 0; t < BD_TF; t++) acc[t] = t < 16 ? a16[t & 15] : 0.f;
			}
			else band_diff_pass<BD_TF>(A.slices, wi, n, img, A.pix, A.nd2, s_ux, s_uy, acc, base);
			// fixed-order reduction: lanes, then warps
#pragma unroll
			for (int t = 0; t < BD_TF; t++)
			{
				if (t < ntr)
				{
					const float v = warp_sum(acc[t]);
					if (lane == 0) s_red[wid][t] = v;
				}
			}
			base = warp_sum(base);
			if (lane == 0) s_red[wid][BD_TF] = base;
			__syncthreads();
			if (threadIdx.x < ntr)
			{
				float c = 0.f, b = 0.f;
#pragma unroll
				for (int ww = 0; ww < BD_THREADS / 32; ww++) { c += s_red[ww][threadIdx.x]; b += s_red[ww][BD_TF]; }
				const float v = fmaxf((b - 2.f * c) + xi2_half, 0.f);
				A.fs_w[F.sample_off + c0 + threadIdx.x] = v; bmin = fminf(bmin, v);
			}
		}
		if (threadIdx.x < BD_TF && bmin < FLT_MAX) rb_atomic_min_pos(&A.states[p].fmin_bits, bmin);
	}
}

// ---------------------------------------------------------------------------------------------
// list of the fine orientations that hold significant samples (the only ones wavg / back-projection work on)
// ---------------------------------------------------------------------------------------------
struct BpListArgs {
	const RbPartState *states_c; RbPartState *states;
	const RbFineOrient *fo; const int *pair_list; int *counters;   // counters[0] fine orientations, [10] BP items, [11] BP samples, [2] overflow
	const float *fs_w;
	int *cnt;                    // [cap_fo] significant samples per fine orientation, then exclusive prefix
	int *item_of;                // [cap_fo] BP item index of a fine orientation (or -1)
	RbBpItem *items; float4 *samp; long long samp_cap;
	const float *tx, *ty; int NOT;
};

static __global__ void __launch_bounds__(256)
k_bp_count(BpListArgs A)
{
	if (A.counters[2]) return;
	const int nfo = A.counters[0];
	const int lane = threadIdx.x & 31;
	for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nfo; w += gridDim.x * (blockDim.x >> 5))
	{
		const RbFineOrient F = A.fo[w];
		const RbPartState *st = A.states_c + F.particle;
		int cnt = 0;
		if (st->status == 0)
		{
			const float sig = st->fsig_weight;
			const int nsamp = F.n_t * A.NOT;
			for (int j = lane; j < nsamp; j += 32) cnt += A.fs_w[F.sample_off + j] >= sig ? 1 : 0;   // wavg.cuh:106 / BP.cuh:278
			cnt = warp_sum(cnt);
		}
		if (lane == 0) A.cnt[w] = cnt;
	}
}

// single CTA: exclusive prefixes of (has samples, number of samples) over the fine orientations
static __global__ void __launch_bounds__(1024)
k_bp_scan(BpListArgs A)
{
	__shared__ long long s_a[1024], s_b[1024];
	if (A.counters[2]) { if (threadIdx.x == 0) { A.counters[10] = 0; A.counters[11] = 0; } return; }
	const int nfo = A.counters[0];
	const int per = (nfo + 1023) / 1024;
	const int i0 = min(nfo, threadIdx.x * per), i1 = min(nfo, i0 + per);
	long long a = 0, b = 0;
	for (int i = i0; i < i1; i++) { const int c = A.cnt[i]; a += c > 0; b += c; }
	s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
	__syncthreads();
	for (int off = 1; off < 1024; off <<= 1)
	{
		long long ta = 0, tb = 0;
		if (threadIdx.x >= off) { ta = s_a[threadIdx.x - off]; tb = s_b[threadIdx.x - off]; }
		__syncthreads();
		s_a[threadIdx.x] += ta; s_b[threadIdx.x] += tb;
		__syncthreads();
	}
	long long ea = s_a[threadIdx.x] - a, eb = s_b[threadIdx.x] - b;
	const bool overflow = s_b[1023] > A.samp_cap;
	for (int i = i0; i < i1; i++)
	{
		const int c = A.cnt[i];
		A.item_of[i] = (c > 0 && !overflow) ? (int) ea : -1;
		A.cnt[i] = (int) eb;
		ea += c > 0; eb += c;
	}
	if (threadIdx.x == 1023)
	{
		A.counters[10] = overflow ? 0 : (int) s_a[1023];
		A.counters[11] = overflow ? 0 : (int) s_b[1023];
		if (overflow) { A.counters[3] = 1; ((long long *) A.counters)[6] = s_b[1023]; }   // counters[12..13]: samples needed
	}
}

static __global__ void __launch_bounds__(256)
k_bp_fill(BpListArgs A)
{
	if (A.counters[2] || A.counters[3]) return;
	const int nfo = A.counters[0];
	const int lane = threadIdx.x & 31;
	for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nfo; w += gridDim.x * (blockDim.x >> 5))
	{
		const int item = A.item_of[w];
		if (item < 0) continue;
		const RbFineOrient F = A.fo[w];
		const RbPartState *st = A.states_c + F.particle;
		const float sig = st->fsig_weight, wni = 1.0f / st->fsum_weight;
		const int nsamp = F.n_t * A.NOT, soff = A.cnt[w];
		int run = 0;
		for (int j0 = 0; j0 < nsamp; j0 += 32)
		{
			const int j = j0 + lane;
			const float wv = j < nsamp ? A.fs_w[F.sample_off + j] : -1.f;
			const bool s = j < nsamp && wv >= sig;
			const unsigned b = __ballot_sync(RB_FULL_MASK, s);
			if (s)
			{
				const int pos = run + __popc(b & ((1u << lane) - 1));
				const int it = A.pair_list[F.pair_off + j / A.NOT] * A.NOT + (j % A.NOT);
				A.samp[soff + pos] = make_float4(A.tx[it] * 0.15915494309189535f, A.ty[it] * 0.15915494309189535f, wv * wni, 0.f);   // weight * weight_norm_inverse (wavg.h:138)
			}
			run += __popc(b);
		}
		__syncwarp();
		if (lane == 0)
		{
			float W = 0.f;
			for (int i = 0; i < run; i++) W += A.samp[soff + i].z;
			RbBpItem it; it.w = w; it.samp_off = soff; it.nsig = run; it.W = W;
			A.items[item] = it;
			atomicAdd(&A.states[F.particle].n_bp, 1);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// wavg + back-projection, band-major
// ---------------------------------------------------------------------------------------------
struct BandStoreArgs {
	const RbPartMeta *metas; RbPartState *states;
	const RbFineOrient *fo; const RbBpItem *items; const float4 *samp; const int *counters;
	int begin, capacity;         // this round covers BP items [begin, min(nbp, begin + capacity))
	int slice_by_item;           // 1: slices are indexed by (item - begin) (re-projected rounds, run only when the fine pass took
	                             // several rounds: counters[0] > fits); 0: by fine orientation index (only when counters[0] <= fits)
	int fits;                    // slice buffer capacity in fine orientations
	const float4 *sst; const float *sctf; const float2 *slices; const uint32_t *pix; int nst, stride;
	float *shells;               // [P][nshell]
	const RbBackprojector *bps;
	int n; int *queue; int chunk_min;
	int nr_classes; int P;
};

struct BandStoreSmem {
	RbBpItem item[BD_MAXCHUNK];
	float e[BD_MAXCHUNK][6];
	int particle[BD_MAXCHUNK], cls[BD_MAXCHUNK], og[BD_MAXCHUNK], bpi[BD_MAXCHUNK];
	float part_scale[BD_MAXCHUNK];
	int next;
};

struct BandStoreIn { float2 ref; float4 XX; float ctf; };

template <bool MULTI>
static __global__ void __launch_bounds__(BD_THREADS, 3)
k_store_band(BandStoreArgs A, RbModelDev M)
{
	__shared__ BandStoreSmem S;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (A.counters[2] || A.counters[3]) return;
	if ((A.counters[0] > A.fits) != (A.slice_by_item != 0)) return;
	const int total = A.counters[10];
	const int n = min(A.capacity, total - A.begin);
	if (n <= 0) return;
	const int nrows = A.slice_by_item ? n : A.counters[0];     // rows of the tile-major slice array: what the producing round projected
	const int ntiles = (A.nst + BD_TP - 1) / BD_TP;
	int chunk = (n + (int) gridDim.x - 1) / (int) gridDim.x;
	chunk = max(A.chunk_min, min(BD_MAXCHUNK, chunk));
	const int nchunks = (n + chunk - 1) / chunk;
	const long long nitems = (long long) ntiles * nchunks;
	const int half = A.n / 2;
	const RbBackprojector bp0 = A.bps[0];
	const int ph = wid / BD_WPT;

	long long item = blockIdx.x;
	while (item < nitems)
	{
		const int tile = (int) (item / nchunks), c = (int) (item - (long long) tile * nchunks);
		const int o0 = c * chunk, no = min(chunk, n - o0);
		for (int j = threadIdx.x; j < no; j += BD_THREADS)
		{
			const RbBpItem it = A.items[A.begin + o0 + j];
			const int p = A.fo[it.w].particle;
			S.item[j] = it;
			S.particle[j] = p; S.cls[j] = A.fo[it.w].iclass;
			S.og[j] = A.metas[p].og; S.part_scale[j] = A.metas[p].part_scale; S.bpi[j] = A.fo[it.w].iclass + A.metas[p].bp_off;
		}
		for (int i = threadIdx.x; i < no * 6; i += BD_THREADS)
		{
			const int j = i / 6, q = i - j * 6;
			S.e[j][q] = A.fo[A.items[A.begin + o0 + j].w].e[q + q / 2];
		}
		int pending = 0;
		if (threadIdx.x == 0) pending = atomicAdd(A.queue, 1);      // next item, looked at after this one
		__syncthreads();
		const int ip = tile * BD_TP + (wid % BD_WPT) * 32 + lane;
		const bool have = ip < A.nst;
		int x = 0, y = 0, ires = 0;
		if (have) { const uint32_t pkx = __ldg(A.pix + ip); x = rb_pix_x(pkx); y = rb_pix_y(pkx); ires = rb_pix_ires(pkx); }
		// (x = 0, y < 0) is only in the list with --no_map: Mresol excludes it from the shell sums (:3466-3494)
		// ... and so are the rows the reference's wavg kernel skips when the references end inside the window (wavg.h:74-82); its
		// back-projection kernel still walks them
		const bool in_mresol = have && !(x == 0 && y < 0) && !(M.dead_maxR > 0 && abs(y) > M.dead_maxR && x != M.dead_maxR);
		bool circle_ok = true;
		if (M.bp_circle_bound && !M.do_grad) { const int xmax = (int) sqrtf((float) (half * half - y * y)); circle_ok = x < xmax; }   // BP.h:565 (not in the SGD kernel, BP.h:757-1047)
		const float fxp = (float) x, fyp = (float) y;

		// the inputs of the next orientation are requested before the current one is worked on
		auto fetch = [&](int j, BandStoreIn &in)
		{
			in.ref = make_float2(0.f, 0.f); in.XX = make_float4(0.f, 0.f, 0.f, 0.f); in.ctf = 0.f;
			if (have)
			{
				const size_t so = bd_at(A.slice_by_item ? (o0 + j) : S.item[j].w, ip, nrows);
				const size_t po = bd_at(S.particle[j], ip, A.P);
				in.ref = __ldcs(A.slices + so); in.XX = __ldg(A.sst + po); in.ctf = __ldg(A.sctf + po);
			}
		};
		BandStoreIn cur, nxt;
		if (ph < no) fetch(ph, cur);
		for (int j = ph; j < no; j += BD_NPH)
		{
			if (j + BD_NPH < no) fetch(j + BD_NPH, nxt);
			const RbBpItem it = S.item[j];
			const int p = S.particle[j], cls = MULTI ? S.cls[j] : 0;
			float2 ref = cur.ref; const float4 XX = cur.XX; const float ctf = cur.ctf;
			const float2 ref_ctf = make_float2(ref.x * ctf, ref.y * ctf);                                 // BP.cuh:520-521 (SGD)
			const float part_scale = S.part_scale[j];
			if (M.refs_are_ctf_corrected) { ref.x *= ctf; ref.y *= ctf; }                                  // wavg.cuh:96-104
			else { ref.x *= part_scale; ref.y *= part_scale; }
			float phr = 0.f, phi = 0.f;
			const float4 *sp = A.samp + it.samp_off;
			for (int t = 0; t < it.nsig; t++)
			{
				const float4 sv = __ldg(sp + t);
				float u = fmaf(fxp, sv.x, fyp * sv.y);
				u -= rintf(u);
				float sn, cs;
				__sincosf(6.283185307179586f * u, &sn, &cs);
				phr = fmaf(sv.z, cs, phr); phi = fmaf(sv.z, sn, phi);
			}
			const float W = it.W;
			const float refn = ref.x * ref.x + ref.y * ref.y;
			const float Xn = XX.x * XX.x + XX.y * XX.y;
			const float xa = (ref.x * XX.x + ref.y * XX.y) * phr - (ref.x * XX.y - ref.y * XX.x) * phi;
			const float aa = W * refn;
			float wd = in_mresol ? fmaxf(W * (refn + Xn) - 2.f * xa, 0.f) : 0.f;
			// shell sums: the 32 pixels of a warp are neighbours in |r|, i.e. they sit in one or two shells
			{
				unsigned todo = __ballot_sync(RB_FULL_MASK, in_mresol);
				while (todo)
				{
					const int leader = __ffs(todo) - 1;
					const int curs = __shfl_sync(RB_FULL_MASK, ires, leader);
					const bool mine = in_mresol && ires == curs;
					const float v = warp_sum(mine ? wd : 0.f);
					if (lane == leader && v != 0.f) atomicAdd(A.shells + (size_t) p * M.nshell + curs, v);
					todo &= ~__ballot_sync(RB_FULL_MASK, mine);
				}
			}
			if (M.do_scale_correction)                                                                        // :3473-3479
			{
				const bool use = in_mresol && M.dvp_gt3[(size_t) cls * M.nshell + ires];
				const double sxa = warp_sum(use ? (double) xa : 0.), saa = warp_sum(use ? (double) aa : 0.);
				if (lane == 0 && (sxa != 0. || saa != 0.))
				{
					atomicAdd(&A.states[p].wsum_XA, sxa);
					atomicAdd(&A.states[p].wsum_AA, saa);
				}
			}
			// back-projection
			RbBackprojector bp = bp0;
			if (MULTI) bp = A.bps[S.bpi[j]];
			const int max_r2_vol = (int) (bp.maxR * bp.maxR * bp.padding_factor * bp.padding_factor);   // BP.cuh:209
			int cell = -1;
			float sfx = 0.f, sfy = 0.f, sfz = 0.f, Fr = 0.f, Fi = 0.f, Fw = 0.f;
			if (have)
			{
				const float minvs2 = M.do_map ? __ldg(M.minvs2 + (size_t) S.og[j] * M.nshell + ires) : 1.f;   // :2586, :3110-3115
				const float g = M.ctf_premultiplied ? minvs2 : ctf * minvs2;                                // BP.cuh:280-289
				Fw = W * g * ctf;
				if (Fw > 0.f && circle_ok)
				{
					Fr = (XX.z * phr - XX.w * phi) * g;
					Fi = (XX.z * phi + XX.w * phr) * g;
					if (M.do_grad) { Fr -= ref_ctf.x * (W * g); Fi -= ref_ctf.y * (W * g); }                  // sum_t w_t (X_t - CTF A), BP.cuh:540-541
					const float e0 = S.e[j][0], e1 = S.e[j][1], e3 = S.e[j][2], e4 = S.e[j][3], e6 = S.e[j][4], e7 = S.e[j][5];
					float xp = (e0 * x + e1 * y) * bp.padding_factor;                                          // BP.cuh:301-347
					float yp = (e3 * x + e4 * y) * bp.padding_factor;
					float zp = (e6 * x + e7 * y) * bp.padding_factor;
					if (xp * xp + yp * yp + zp * zp <= (float) max_r2_vol)
					{
						if (xp < 0.f) { xp = -xp; yp = -yp; zp = -zp; Fi = -Fi; }
						const float fx0 = floorf(xp), fy0 = floorf(yp), fz0 = floorf(zp);
						sfx = xp - fx0; sfy = yp - fy0; sfz = zp - fz0;
						cell = bp.blkvol ? (int) rb_bp_blk_cell(bp, (int) fx0, (int) fy0 - bp.mdlInitY, (int) fz0 - bp.mdlInitZ)
						                 : (((int) fz0 - bp.mdlInitZ) * bp.mdlY + ((int) fy0 - bp.mdlInitY)) * bp.mdlX + (int) fx0;
					}
				}
			}
			// two lanes per pixel, one per x-neighbour: the two 16-byte reductions of a corner pair share a 32-byte sector
#pragma unroll
			for (int h = 0; h < 2; h++)
			{
				const int src = 16 * h + (lane >> 1);
				const int cc = __shfl_sync(RB_FULL_MASK, cell, src);
				const float fx = __shfl_sync(RB_FULL_MASK, sfx, src), fy = __shfl_sync(RB_FULL_MASK, sfy, src), fz = __shfl_sync(RB_FULL_MASK, sfz, src);
				const float vr = __shfl_sync(RB_FULL_MASK, Fr, src), vi = __shfl_sync(RB_FULL_MASK, Fi, src), vw = __shfl_sync(RB_FULL_MASK, Fw, src);
				if (cc >= 0)
				{
					const int px = lane & 1;
					const float wx = px ? fx : 1.f - fx;
					const float mfy = 1.f - fy, mfz = 1.f - fz;
					float4 *b = (bp.blkvol ? bp.blkvol : bp.vol) + (size_t) cc + px;
					const size_t sy = bp.blkvol ? 5 : bp.mdlX, sz = bp.blkvol ? 25 : (size_t) bp.mdlX * bp.mdlY;
					float d2;
					d2 = mfz * mfy * wx; bd_red_add_v4(b, d2 * vr, d2 * vi, d2 * vw);
					d2 = mfz * fy * wx;  bd_red_add_v4(b + sy, d2 * vr, d2 * vi, d2 * vw);
					d2 = fz * mfy * wx;  bd_red_add_v4(b + sz, d2 * vr, d2 * vi, d2 * vw);
					d2 = fz * fy * wx;   bd_red_add_v4(b + sz + sy, d2 * vr, d2 * vi, d2 * vw);
				}
			}
			cur = nxt;
		}
		if (threadIdx.x == 0) S.next = (int) gridDim.x + pending;
		__syncthreads();
		item = S.next;
	}
}

// ---------------------------------------------------------------------------------------------
// Posed back-projection (relion_reconstruct, BASELINE config #2), band-major.
// BackProjector::backproject2Dto3D (/root/reference/src/backprojector.cpp:55-357) for a batch of images: the per-sample
// arithmetic is k_backproject_posed's (kernels_misc.cu: fp64 positions, the reference's pixel set), the ORDER is the store
// stage's: the images are first copied into band order (tile-major float4 (re, im, ctf^2, 0)), then work items (tile,
// chunk of images) sweep Fourier space shell by shell, so the accumulator shell being reduced into stays in L2 instead
// of every image dragging its own great circle through the 1.1 GB accumulator.
// ---------------------------------------------------------------------------------------------
struct PosedBandArgs {
	RbBackprojector bp; int n, count;
	const float2 *F2D; const float *Fctf; const float *eulers;
	const uint32_t *pix; int npix, stride;
	float4 *sF;                  // tile-major [stride / BD_TP][count][BD_TP]
	int *queue; int chunk_min;
};

static __global__ void __launch_bounds__(256)
k_posed_sort(PosedBandArgs A)
{
	const int img = blockIdx.y, xs = A.n / 2 + 1;
	const float2 *F = A.F2D + (size_t) img * A.n * xs;
	const float *W = A.Fctf + (size_t) img * A.n * xs;
	for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < A.stride; ip += gridDim.x * blockDim.x)
	{
		float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
		if (ip < A.npix)
		{
			const uint32_t pk = __ldg(A.pix + ip);
			const int x = rb_pix_x(pk), y = rb_pix_y(pk);
			const int idx = (y < 0 ? y + A.n : y) * xs + x;
			const float2 v = __ldg(F + idx);
			o = make_float4(v.x, v.y, __ldg(W + idx), 0.f);
		}
		A.sF[bd_at(img, ip, A.count)] = o;
	}
}

struct PosedBandSmem { double a[BD_MAXCHUNK][6]; double ata[BD_MAXCHUNK][3]; int next; };

static __global__ void __launch_bounds__(BD_THREADS, 3)
k_posed_band(PosedBandArgs A)
{
	__shared__ PosedBandSmem S;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const RbBackprojector bp = A.bp;
	const double pf = (double) bp.padding_factor;
	const long long rr = (long long) floor((double) bp.maxR * pf + 0.5);
	const double max_r2 = (double) (rr * rr);
	const size_t sy = bp.blkvol ? 5 : bp.mdlX, sz = bp.blkvol ? 25 : (size_t) bp.mdlX * bp.mdlY;
	const int n = A.count;
	const int ntiles = (A.npix + BD_TP - 1) / BD_TP;
	int chunk = (n + (int) gridDim.x - 1) / (int) gridDim.x;
	chunk = max(A.chunk_min, min(BD_MAXCHUNK, chunk));
	const int nchunks = (n + chunk - 1) / chunk;
	const long long nitems = (long long) ntiles * nchunks;
	const int ph = wid / BD_WPT;

	long long item = blockIdx.x;
	while (item < nitems)
	{
		const int tile = (int) (item / nchunks), c = (int) (item - (long long) tile * nchunks);
		const int o0 = c * chunk, no = min(chunk, n - o0);
		for (int j = threadIdx.x; j < no; j += BD_THREADS)
		{
			const float *e = A.eulers + (size_t) (o0 + j) * 9;
			const double a00 = (double) e[0] * pf, a01 = (double) e[1] * pf, a10 = (double) e[3] * pf, a11 = (double) e[4] * pf,
			             a20 = (double) e[6] * pf, a21 = (double) e[7] * pf;
			S.a[j][0] = a00; S.a[j][1] = a01; S.a[j][2] = a10; S.a[j][3] = a11; S.a[j][4] = a20; S.a[j][5] = a21;
			S.ata[j][0] = a00 * a00 + a10 * a10 + a20 * a20; S.ata[j][1] = a00 * a01 + a10 * a11 + a20 * a21;
			S.ata[j][2] = a01 * a01 + a11 * a11 + a21 * a21;
		}
		int pending = 0;
		if (threadIdx.x == 0) pending = atomicAdd(A.queue, 1);
		__syncthreads();
		const int ip = tile * BD_TP + (wid % BD_WPT) * 32 + lane;
		const bool have = ip < A.npix;
		int x = 0, y = 0;
		if (have) { const uint32_t pkx = __ldg(A.pix + ip); x = rb_pix_x(pkx); y = rb_pix_y(pkx); }
		float4 cur = make_float4(0.f, 0.f, 0.f, 0.f), nxt = cur;
		if (have && ph < no) cur = __ldcs(A.sF + bd_at(o0 + ph, ip, n));
		for (int j = ph; j < no; j += BD_NPH)
		{
			if (have && j + BD_NPH < no) nxt = __ldcs(A.sF + bd_at(o0 + j + BD_NPH, ip, n));
			long long cell = -1;
			float sfx = 0.f, sfy = 0.f, sfz = 0.f, vr = 0.f, vi = 0.f, vw = 0.f;
			if (have && cur.z > 0.f)
			{
				// The reference first solves |A (x, y)|^2 <= max_r2 for the x-range of the row (:118-128, a double sqrt and two
				// divisions) and then tests the same quadratic form per pixel (:150): the two can only disagree when a pixel sits on
				// the sphere to rounding, so the row solve is evaluated for those pixels alone.
				double xp = S.a[j][0] * x + S.a[j][1] * y, yp = S.a[j][2] * x + S.a[j][3] * y, zp = S.a[j][4] * x + S.a[j][5] * y;
				const double r2 = xp * xp + yp * yp + zp * zp;
				bool ok = r2 <= max_r2;
				if (ok && max_r2 - r2 < 1e-9 * max_r2)
				{
					const double AtA_xx = S.ata[j][0], AtA_xy = S.ata[j][1], AtA_yy = S.ata[j][2];
					const double discr = AtA_xy * AtA_xy * y * y - AtA_xx * (AtA_yy * y * y - max_r2);
					ok = discr >= 0.;
					if (ok)
					{
						const double d = sqrt(discr) / AtA_xx, q = -AtA_xy * y / AtA_xx;
						ok = x >= (int) ceil(q - d) && x <= (int) floor(q + d);
					}
				}
				if (ok)
				{
					float2 v = make_float2(cur.x, cur.y);
					if (xp < 0.) { xp = -xp; yp = -yp; zp = -zp; v.y = -v.y; }
					const double fx0 = floor(xp), fy0 = floor(yp), fz0 = floor(zp);
					const int x0 = (int) fx0, y0 = (int) fy0 - bp.mdlInitY, z0 = (int) fz0 - bp.mdlInitZ;
					if (x0 >= 0 && x0 + 1 < bp.mdlX && y0 >= 0 && y0 + 1 < bp.mdlY && z0 >= 0 && z0 + 1 < bp.mdlZ)   // :213-218
					{
						sfx = (float) (xp - fx0); sfy = (float) (yp - fy0); sfz = (float) (zp - fz0);
						vr = v.x; vi = v.y; vw = cur.z;
						cell = bp.blkvol ? (long long) rb_bp_blk_cell(bp, x0, y0, z0) : ((long long) z0 * bp.mdlY + y0) * bp.mdlX + x0;
					}
				}
			}
#pragma unroll
			for (int h = 0; h < 2; h++)
			{
				const int src = 16 * h + (lane >> 1);
				const long long cc = __shfl_sync(RB_FULL_MASK, cell, src);
				const float fx = __shfl_sync(RB_FULL_MASK, sfx, src), fy = __shfl_sync(RB_FULL_MASK, sfy, src), fz = __shfl_sync(RB_FULL_MASK, sfz, src);
				const float r = __shfl_sync(RB_FULL_MASK, vr, src), im = __shfl_sync(RB_FULL_MASK, vi, src), w = __shfl_sync(RB_FULL_MASK, vw, src);
				if (cc >= 0)
				{
					const int px = lane & 1;
					const float wx = px ? fx : 1.f - fx, mfy = 1.f - fy, mfz = 1.f - fz;
					float4 *b = (bp.blkvol ? bp.blkvol : bp.vol) + (size_t) cc + px;
					float dd;
					dd = mfz * mfy * wx; bd_red_add_v4(b, dd * r, dd * im, dd * w);
					dd = mfz * fy * wx;  bd_red_add_v4(b + sy, dd * r, dd * im, dd * w);
					dd = fz * mfy * wx;  bd_red_add_v4(b + sz, dd * r, dd * im, dd * w);
					dd = fz * fy * wx;   bd_red_add_v4(b + sz + sy, dd * r, dd * im, dd * w);
				}
			}
			cur = nxt;
		}
		if (threadIdx.x == 0) S.next = (int) gridDim.x + pending;
		__syncthreads();
		item = S.next;
	}
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt)
{
	const char *v = getenv(name);
	return (v && *v) ? atoi(v) : dflt;
}

bool rbk_band_applicable(rb_ctx *ctx)
{
	// decided by pool_setup (api.cu): not with the cross-correlation criterion (it stays on k_diff2_fine / k_store), not with
	// RB_BAND=0 or without room for the band-ordered slices
	return !ctx->d_model.do_cc && ctx->d_model.pix_rs && ctx->band_slice_capacity > 0;
}

// phase tables of the current (model, sampling), rebuilt when either changed; ok == false: the sampling does not factorise
int rbk_band_phase_tables(rb_ctx *ctx, bool &ok)
{
	ok = false;
	static int on = -1;
	if (on < 0) on = env_int("RB_BAND_TABLES", 1);
	const int NOT = ctx->d_samp.n_over_trans, T = ctx->d_samp.n_trans;
	if (!on || !ctx->band_separable || !(NOT == 1 || NOT == 4)) return RB_OK;
	ok = true;
	if (ctx->band_tab_model == ctx->model_version && ctx->band_tab_samp == ctx->samp_version) return RB_OK;
	const RbModelDev &M = ctx->d_model;
	const size_t stride = (size_t) M.nv_rs_pad;
	RB_CHECK(ctx->band_tabc.ensure((size_t) T * stride * sizeof(float2)));
	RB_CHECK(ctx->band_tabo.ensure((size_t) NOT * stride * sizeof(float2)));
	RB_CHECK(ctx->band_tabu.ensure((size_t) 2 * (T + NOT) * sizeof(double)));
	RB_CUDA(cudaMemcpyAsync(ctx->band_tabu.p, ctx->h_band_u.data(), (size_t) 2 * (T + NOT) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	const double *u = ctx->band_tabu.as<double>();
	dim3 g((unsigned) ((stride + 255) / 256), T + NOT);
	k_band_tables<<<g, 256, 0, ctx->stream>>>(M.pix_rs, M.nv_rs_st, (int) stride, u, u + T, T, u + 2 * T, u + 2 * T + NOT, NOT,
	                                          ctx->band_tabc.as<float2>(), ctx->band_tabo.as<float2>());
	RB_LAUNCH_CHECK(ctx);
	ctx->band_tab_model = ctx->model_version; ctx->band_tab_samp = ctx->samp_version;
	return RB_OK;
}

// per-pool buffers and the band-ordered images (once per E-step of a slot): which = 0 before the fine pass, 1 before the store stage
int rbk_band_prepare_pool(rb_ctx *ctx, PoolSlot &s, int which, cudaStream_t stream)
{
	const RbModelDev &M = ctx->d_model;
	const size_t stride = (size_t) M.nv_rs_pad;
	if (which == 0) RB_CHECK(s.simg4.ensure((size_t) s.P * stride * sizeof(float4)));
	else { RB_CHECK(s.sst.ensure((size_t) s.P * stride * sizeof(float4))); RB_CHECK(s.sctf.ensure((size_t) s.P * stride * sizeof(float))); }
	BandPrepArgs A;
	memset(&A, 0, sizeof(A));
	A.metas = s.meta.as<RbPartMeta>(); A.Fimg = s.Fimg.as<float2>(); A.Fnomask = s.Fnomask.as<float2>();
	A.Fctf = ctx->h_model.do_ctf_correction ? s.Fctf.as<float>() : nullptr;
	A.pix = M.pix_rs; A.nd2 = M.nv_rs_d2; A.nst = M.nv_rs_st; A.stride = (int) stride;
	A.simg4 = s.simg4.as<float4>(); A.sst = s.sst.as<float4>(); A.sctf = s.sctf.as<float>();
	dim3 g((unsigned) ((stride + 255) / 256), s.P);
	if (which == 0) k_prep_sorted<0><<<g, 256, 0, stream>>>(A, M);
	else k_prep_sorted<1><<<g, 256, 0, stream>>>(A, M);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

static int launch_project_band(rb_ctx *ctx, PoolSlot &s, const int *indir, const int *count_ptr, int begin, int capacity, int *queue,
                               const int *nfo_ptr = nullptr, int only_if_nfo_above = 0)
{
	const RbModelDev &M = ctx->d_model;
	BandProjArgs A;
	memset(&A, 0, sizeof(A));
	A.fo = s.fo.as<RbFineOrient>(); A.indir = indir; A.count_ptr = count_ptr; A.begin = begin; A.capacity = capacity;
	A.pix = M.pix_rs; A.npix = M.nv_rs_st; A.stride = M.nv_rs_pad;
	A.slices = ctx->band_slices.as<float2>();
	A.projs = ctx->d_proj.as<RbProjector>(); A.imgX = M.current_size / 2 + 1; A.nr_classes = M.nr_classes;
	A.queue = queue;
	A.chunk_min = std::max(BD_NPH, env_int("RB_BAND_CHUNK_MIN", 48));   // measured (256 px, 6768 fine orientations): 16: 2.54 ms, 32: 2.36, 48: 2.30, 64: 2.31, 96: 2.44, 128: 2.56
	A.nfo_ptr = nfo_ptr; A.only_if_nfo_above = only_if_nfo_above;
	RB_CUDA(cudaMemsetAsync(queue, 0, 4, ctx->stream));
	const int ctas = env_int("RB_BAND_PROJ_CTAS", 3);
	const int grid = ctx->num_sms * ctas;
	if (M.nr_classes > 1) { if (ctas >= 3) k_project_band<true, 3><<<grid, BD_THREADS, 0, ctx->stream>>>(A); else k_project_band<true, 2><<<grid, BD_THREADS, 0, ctx->stream>>>(A); }
	else { if (ctas >= 3) k_project_band<false, 3><<<grid, BD_THREADS, 0, ctx->stream>>>(A); else k_project_band<false, 2><<<grid, BD_THREADS, 0, ctx->stream>>>(A); }
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// fine pass: projection of every fine orientation (band-major), then the streaming diff2 pass; in rounds when the slices of
// all fine orientations the pool could produce do not fit the slice buffer (rounds beyond the actual count exit at once)
// The band-ordered copies of the particle images depend on the uploaded pool only: they are made on a side stream while the
// coarse pass (L1-bound, HBM idle) runs, and joined before the fine pass.
int rbk_band_images_async(rb_ctx *ctx, PoolSlot &s)
{
	if (!ctx->aux_stream)
	{
		RB_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
		RB_CUDA(cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
		RB_CUDA(cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming));
	}
	RB_CUDA(cudaEventRecord(ctx->aux_fork, ctx->stream));              // after the upload and every earlier reader of the buffers
	RB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
	RB_CHECK(rbk_band_prepare_pool(ctx, s, 0, ctx->aux_stream));
	RB_CHECK(rbk_band_prepare_pool(ctx, s, 1, ctx->aux_stream));
	RB_CUDA(cudaEventRecord(ctx->aux_join, ctx->aux_stream));
	return RB_OK;
}

int rbk_band_fine_pool(rb_ctx *ctx, PoolSlot &s)
{
	const RbModelDev &M = ctx->d_model;
	RB_CHECK(rb_stage_begin(ctx, "fine_prep"));
	RB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_join, 0));       // band-ordered images (rbk_band_images_async)
	bool tables = false;
	RB_CHECK(rbk_band_phase_tables(ctx, tables));
	RB_CHECK(rb_stage_end(ctx, "fine_prep"));
	const long long cap = ctx->band_slice_capacity;
	const int rounds = s.band_rounds;               // pool_setup: ceil(cap_fo / cap), at most RB_BAND_ROUNDS
	int *queue = s.counters.as<int>() + 8;
	for (int r = 0; r < rounds; r++)
	{
		const int begin = (int) (r * cap);
		if (r == 0) RB_CHECK(rb_stage_begin(ctx, "fine_project"));
		RB_CHECK(launch_project_band(ctx, s, nullptr, s.counters.as<int>(), begin, (int) cap, queue));
		if (r == 0) { RB_CHECK(rb_stage_end(ctx, "fine_project")); RB_CHECK(rb_stage_begin(ctx, "fine_diff2")); }
		BandDiffArgs D;
		memset(&D, 0, sizeof(D));
		D.metas = s.meta.as<RbPartMeta>(); D.states = s.state.as<RbPartState>();
		D.fo = s.fo.as<RbFineOrient>(); D.pair_list = s.pair_list.as<int>(); D.counters = s.counters.as<int>();
		D.fs_w = s.fs_w.as<float>();
		D.simg4 = s.simg4.as<float4>(); D.slices = ctx->band_slices.as<float2>(); D.pix = M.pix_rs; D.nd2 = M.nv_rs_d2; D.stride = M.nv_rs_pad;
		D.begin = begin; D.capacity = (int) cap;
		D.tx = ctx->d_samp.ftx; D.ty = ctx->d_samp.fty; D.NOT = ctx->d_samp.n_over_trans;
		D.queue = queue + 1;
		RB_CUDA(cudaMemsetAsync(queue + 1, 0, 4, ctx->stream));
		const int NOT = ctx->d_samp.n_over_trans;
		const int dctas = env_int("RB_BAND_DIFF_CTAS", 2);
	if (tables && NOT == 4) k_diff2_slices_sep<4><<<ctx->num_sms * dctas, BD_THREADS, 0, ctx->stream>>>(D, ctx->band_tabc.as<float2>(), ctx->band_tabo.as<float2>());
		else if (tables && NOT == 1) k_diff2_slices_sep<1><<<ctx->num_sms * dctas, BD_THREADS, 0, ctx->stream>>>(D, ctx->band_tabc.as<float2>(), ctx->band_tabo.as<float2>());
		else k_diff2_slices<<<ctx->num_sms * 3, BD_THREADS, 0, ctx->stream>>>(D);
		RB_LAUNCH_CHECK(ctx);
		if (r == 0) RB_CHECK(rb_stage_end(ctx, "fine_diff2"));
	}
	return RB_OK;
}

int rbk_band_store_pool(rb_ctx *ctx, PoolSlot &s)
{
	const RbModelDev &M = ctx->d_model;
	const long long cap_fo = (long long) s.cap_fo;
	RB_CHECK(s.bp_cnt.ensure((size_t) cap_fo * 4)); RB_CHECK(s.bp_item_of.ensure((size_t) cap_fo * 4));
	RB_CHECK(s.bp_items.ensure((size_t) cap_fo * sizeof(RbBpItem)));
	const long long samp_cap = std::min<long long>((long long) s.cap_fs, (long long) env_int("RB_BP_SAMPLE_CAP", 1 << 22));
	RB_CHECK(s.bp_samp.ensure((size_t) samp_cap * sizeof(float4)));
	BpListArgs L;
	memset(&L, 0, sizeof(L));
	L.states_c = s.state.as<RbPartState>(); L.states = s.state.as<RbPartState>();
	L.fo = s.fo.as<RbFineOrient>(); L.pair_list = s.pair_list.as<int>(); L.counters = s.counters.as<int>();
	L.fs_w = s.fs_w.as<float>(); L.cnt = s.bp_cnt.as<int>(); L.item_of = s.bp_item_of.as<int>();
	L.items = s.bp_items.as<RbBpItem>(); L.samp = s.bp_samp.as<float4>(); L.samp_cap = samp_cap;
	L.tx = ctx->d_samp.ftx; L.ty = ctx->d_samp.fty; L.NOT = ctx->d_samp.n_over_trans;
	RB_CHECK(rb_stage_begin(ctx, "store_list"));
	k_bp_count<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(L);
	RB_LAUNCH_CHECK(ctx);
	k_bp_scan<<<1, 1024, 0, ctx->stream>>>(L);
	RB_LAUNCH_CHECK(ctx);
	k_bp_fill<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(L);
	RB_LAUNCH_CHECK(ctx);
	RB_CHECK(rb_stage_end(ctx, "store_list"));

	BandStoreArgs A;
	memset(&A, 0, sizeof(A));
	A.metas = s.meta.as<RbPartMeta>(); A.states = s.state.as<RbPartState>();
	A.fo = s.fo.as<RbFineOrient>(); A.items = s.bp_items.as<RbBpItem>(); A.samp = s.bp_samp.as<float4>(); A.counters = s.counters.as<int>();
	A.sst = s.sst.as<float4>(); A.sctf = s.sctf.as<float>(); A.slices = ctx->band_slices.as<float2>();
	A.pix = M.pix_rs; A.nst = M.nv_rs_st; A.stride = M.nv_rs_pad;
	A.shells = s.shells.as<float>(); A.bps = ctx->d_bp.as<RbBackprojector>();
	A.n = M.current_size; A.chunk_min = std::max(BD_NPH, env_int("RB_BAND_STORE_CHUNK_MIN", 16));
	A.nr_classes = M.nr_classes; A.P = s.P;
	int *queue = s.counters.as<int>() + 14;
	const int grid = ctx->num_sms * env_int("RB_BAND_STORE_CTAS", 3);
	const bool multi = M.nr_classes > 1 || s.max_bp_off > 0;
	const long long cap = ctx->band_slice_capacity;
	A.fits = (int) std::min<long long>(cap, 0x7fffffff);
	// (a) the fine pass fitted one round (decided on the device: counters[0] <= cap): its slices are still in the buffer
	A.begin = 0; A.capacity = 0x7fffffff; A.slice_by_item = 0; A.queue = queue;
	RB_CUDA(cudaMemsetAsync(queue, 0, 4, ctx->stream));
	RB_CHECK(rb_stage_begin(ctx, "store_band"));
	if (multi) k_store_band<true><<<grid, BD_THREADS, 0, ctx->stream>>>(A, M);
	else k_store_band<false><<<grid, BD_THREADS, 0, ctx->stream>>>(A, M);
	RB_LAUNCH_CHECK(ctx);
	RB_CHECK(rb_stage_end(ctx, "store_band"));
	// (b) it took several rounds and the buffer was reused: project the listed orientations again, round by round (these
	// launches exit at once in case (a))
	for (int r = 0; r < s.band_rounds && s.band_rounds > 1; r++)
	{
		const int begin = (int) (r * cap);
		RB_CHECK(launch_project_band(ctx, s, (const int *) s.bp_items.p, s.counters.as<int>() + 10, begin, (int) cap, queue + 1, s.counters.as<int>(), A.fits));
		A.begin = begin; A.capacity = (int) cap; A.slice_by_item = 1; A.queue = queue;
		RB_CUDA(cudaMemsetAsync(queue, 0, 4, ctx->stream));
		if (multi) k_store_band<true><<<grid, BD_THREADS, 0, ctx->stream>>>(A, M);
		else k_store_band<false><<<grid, BD_THREADS, 0, ctx->stream>>>(A, M);
		RB_LAUNCH_CHECK(ctx);
	}
	return RB_OK;
}

// band-ordered pixel list of an n x (n/2+1) half transform: every pixel backproject2Dto3D can use (x = 0 only for y >= 0)
static int posed_pixlist(rb_ctx *ctx, int n)
{
	if (ctx->posed_pix_n == n) return RB_OK;
	const int xs = n / 2 + 1;
	std::vector<uint32_t> v;
	v.reserve((size_t) n * xs);
	for (int i = 0; i < n; i++)
		for (int x = 0; x < xs; x++)
		{
			const int y = i < xs ? i : i - n;
			if (y < 0 && x == 0) continue;                                       // first_allowed_x (backprojector.cpp:107-116)
			v.push_back(rb_pack_pix(x, y, 0));
		}
	std::sort(v.begin(), v.end(), [](uint32_t a, uint32_t b) {
		const int xa = rb_pix_x(a), ya = rb_pix_y(a), xb = rb_pix_x(b), yb = rb_pix_y(b);
		const int ra = xa * xa + ya * ya, rb2 = xb * xb + yb * yb;
		if (ra != rb2) return ra < rb2;
		const double ta = atan2((double) ya, (double) xa), tb = atan2((double) yb, (double) xb);
		if (ta != tb) return ta < tb;
		return a < b;
	});
	ctx->posed_pix_count = (int) v.size();
	v.resize((v.size() + BD_TP - 1) / BD_TP * BD_TP, rb_pack_pix(0, 0, 0));
	RB_CHECK(ctx->posed_pix.ensure(v.size() * 4));
	RB_CUDA(cudaMemcpyAsync(ctx->posed_pix.p, v.data(), v.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->posed_pix_n = n;
	return RB_OK;
}

// band-ordered staging of a chunk of images: pixel list of the size, tile-major float4 buffer (re, im, weight, 0)
int rbk_posed_band_layout(rb_ctx *ctx, int n, int count, RbPosedBandLayout *L)
{
	RB_CHECK(posed_pixlist(ctx, n));
	L->pix = ctx->posed_pix.as<uint32_t>(); L->npix = ctx->posed_pix_count; L->stride = (L->npix + BD_TP - 1) / BD_TP * BD_TP;
	RB_CHECK(ctx->posed_sorted.ensure((size_t) count * L->stride * sizeof(float4) + 64));
	L->sF = ctx->posed_sorted.as<float4>();
	L->queue = (int *) ((char *) ctx->posed_sorted.p + (size_t) count * L->stride * sizeof(float4));
	return RB_OK;
}

// scatter of a chunk whose band-ordered staging buffer has been filled (k_posed_sort, or the raw-image preparation of
// kernels_prep.cu, which writes it directly)
int rbk_posed_band_scatter(rb_ctx *ctx, const RbBackprojector &bp, int n, int count, const float *d_eulers, const RbPosedBandLayout &L)
{
	PosedBandArgs A;
	memset(&A, 0, sizeof(A));
	A.bp = bp; A.n = n; A.count = count; A.eulers = d_eulers;
	A.pix = L.pix; A.npix = L.npix; A.stride = L.stride; A.sF = L.sF; A.queue = L.queue;
	A.chunk_min = std::max(BD_NPH, env_int("RB_POSED_CHUNK_MIN", 8));
	RB_CUDA(cudaMemsetAsync(A.queue, 0, 4, ctx->stream));
	k_posed_band<<<ctx->num_sms * 3, BD_THREADS, 0, ctx->stream>>>(A);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

int rbk_backproject_posed_band(rb_ctx *ctx, const RbBackprojector &bp, int n, int count, const float2 *d_F, const float *d_W, const float *d_eulers)
{
	if (count < 1) return RB_OK;
	RbPosedBandLayout L;
	RB_CHECK(rbk_posed_band_layout(ctx, n, count, &L));
	PosedBandArgs A;
	memset(&A, 0, sizeof(A));
	A.bp = bp; A.n = n; A.count = count; A.F2D = d_F; A.Fctf = d_W; A.eulers = d_eulers;
	A.pix = L.pix; A.npix = L.npix; A.stride = L.stride; A.sF = L.sF; A.queue = L.queue;
	dim3 g((unsigned) ((A.stride + 255) / 256), (unsigned) count);
	k_posed_sort<<<g, 256, 0, ctx->stream>>>(A);
	RB_LAUNCH_CHECK(ctx);
	return rbk_posed_band_scatter(ctx, bp, n, count, d_eulers, L);
}
