// relion_b200 — diff2 -> posterior weights, sort-free significance threshold, fine-pass job construction.
//
// Replaces, per pool and without host round trips:
//   cuda_kernel_weights_exponent_coarse / cuda_kernel_exponentiate (helper.cuh:16-67),
//   cuda_kernel_exponentiate_weights_fine (helper.cu:36-77),
//   CUB filter>0 + radix sort + inclusive scan + cuda_kernel_find_threshold_idx_in_cumulative +
//   cuda_kernel_array_over_threshold (cuda_utils_cub.cuh:30-359, cuda_device_utils.cuh:191-220),
//   generateProjectionSetupFine / makeJobsForDiff2Fine host loops (acc_helper_functions_impl.h:26-102, 265-314)
// of /root/reference/src/acc.  Orchestration being replaced: convertAllSquaredDifferencesToWeights
// (acc_ml_optimiser_impl.h:1895-2548).
//
// Significance rule (acc_ml_optimiser_impl.h:2240-2345): with the non-zero weights sorted ascending and
// cum their running sum, thresholdIdx = first i with cum[i] > (1-adaptive_fraction)*sum; significant_weight =
// sorted[thresholdIdx].  We find the same element without sorting: an 8-pass, 4-bit radix descent on the
// fp32 bit pattern (weights >= 0, so bit order == value order) that carries the exact (fp64) mass and count of
// everything below the current prefix.  Sums are fixed-tree fp64 reductions, hence deterministic.
#include "device_utils.cuh"
#include <cstdlib>
#include <algorithm>

static const int WT_THREADS = 512;   // 16 fp64 bins + 16 counters per thread: keep 128 registers available

struct SelResult {
	double total;        // fp64 sum of all weights
	float sum_f;         // (float) total  == op.sum_weight
	float sig_w;         // sorted[thresholdIdx]
	long long n_nonzero; // filteredSize
	long long thr_idx;   // thresholdIdx among the non-zero weights
};

struct SelSmem {
	double ws[32][16];
	int wc[32][16];
	double bs[16];
	long long bc[16];
	double dred[32];
	long long lred[32];
	unsigned prefix; double base; long long cbelow; long long cequal; int found;
};

// reduce per-thread 16-bin (sum,count) histograms over the CTA into sm.bs / sm.bc
__device__ __forceinline__ void reduce_bins(double (&s)[16], int (&c)[16], SelSmem &sm)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
	for (int b = 0; b < 16; b++)
	{
		double vs = warp_sum(s[b]);
		int vc = warp_sum(c[b]);
		if (lane == 0) { sm.ws[wid][b] = vs; sm.wc[wid][b] = vc; }
	}
	__syncthreads();
	if (threadIdx.x < 16)
	{
		double a = 0.; long long n = 0;
		for (int w = 0; w < nw; w++) { a += sm.ws[w][threadIdx.x]; n += sm.wc[w][threadIdx.x]; }
		sm.bs[threadIdx.x] = a; sm.bc[threadIdx.x] = n;
	}
	__syncthreads();
}

// Radix descent.  by_rank == false: smallest value v with (mass of weights <= v) > thr.
//                 by_rank == true : value of ascending rank `rank` among the non-zero weights.
__device__ void radix_descend(const float *w, long long n, bool by_rank, double thr, long long rank, SelSmem &sm)
{
	if (threadIdx.x == 0) { sm.prefix = 0u; sm.base = 0.; sm.cbelow = 0; sm.cequal = 0; }
	__syncthreads();
	for (int pass = 0; pass < 8; pass++)
	{
		const int shift = 28 - 4 * pass;
		const unsigned prefix = sm.prefix;
		double s[16]; int c[16];
#pragma unroll
		for (int b = 0; b < 16; b++) { s[b] = 0.; c[b] = 0; }
		for (long long i = threadIdx.x; i < n; i += blockDim.x)
		{
			const float v = w[i];
			const unsigned bits = __float_as_uint(v);
			if (v > 0.f && (pass == 0 || (bits >> (shift + 4)) == (prefix >> (shift + 4))))
			{
				const int bin = (bits >> shift) & 15;
#pragma unroll
				for (int b = 0; b < 16; b++) { const bool h = (bin == b); s[b] += h ? (double) v : 0.; c[b] += h ? 1 : 0; }
			}
		}
		reduce_bins(s, c, sm);
		if (threadIdx.x == 0)
		{
			double cum = sm.base; long long cc = sm.cbelow; int sel = -1;
			for (int b = 0; b < 16; b++)
			{
				if (sm.bc[b] > 0)
				{
					bool hit = by_rank ? (cc + sm.bc[b] > rank) : (cum + sm.bs[b] > thr);
					if (hit) { sel = b; break; }
				}
				cum += sm.bs[b]; cc += sm.bc[b];
			}
			if (sel < 0)
			{
				// nothing crosses the threshold (thr >= total): the reference's search leaves idx = 0
				sm.found = 0;
			}
			else
			{
				sm.found = 1;
				sm.prefix = prefix | ((unsigned) sel << shift);
				sm.base = cum; sm.cbelow = cc; sm.cequal = sm.bc[sel];
			}
		}
		__syncthreads();
		if (!sm.found) return;
	}
}

// Whole significance computation for one array; all threads of the CTA must call.
__device__ SelResult block_significance(const float *w, long long n, double adaptive_fraction, int maxsig, SelSmem &sm)
{
	SelResult r;
	// total mass and number of non-zero weights
	double t = 0.; long long cnt = 0;
	for (long long i = threadIdx.x; i < n; i += blockDim.x) { float v = w[i]; if (v > 0.f) { t += (double) v; cnt++; } }
	r.total = block_sum(t, sm.dred);
	r.n_nonzero = block_sum(cnt, sm.lred);
	r.sum_f = (float) r.total;
	r.sig_w = 0.f; r.thr_idx = 0;
	if (r.n_nonzero == 0) return r;
	const double thr = (double) (float) ((1. - adaptive_fraction) * (double) r.sum_f);   // (XFLOAT) threshold, :2277-2278
	radix_descend(w, n, false, thr, 0, sm);
	__syncthreads();
	if (!sm.found)
	{
		r.thr_idx = 0;
		radix_descend(w, n, true, 0., 0, sm);   // sorted[0]
		__syncthreads();
		r.sig_w = __uint_as_float(sm.prefix);
	}
	else
	{
		const float v = __uint_as_float(sm.prefix);
		// k-th copy of v is the first whose running sum exceeds thr
		double kk = floor((thr - sm.base) / (double) v) + 1.;
		long long k = kk < 1. ? 1 : (kk > (double) sm.cequal ? sm.cequal : (long long) kk);
		r.thr_idx = sm.cbelow + k - 1;
		r.sig_w = v;
	}
	__syncthreads();
	if (maxsig > 0 && r.n_nonzero - r.thr_idx > maxsig)                                   // :2301-2306
	{
		r.thr_idx = r.n_nonzero - maxsig;
		radix_descend(w, n, true, 0., r.thr_idx, sm);
		__syncthreads();
		r.sig_w = __uint_as_float(sm.prefix);
		__syncthreads();
	}
	return r;
}

struct ArgMaxSmem { float v[32]; long long i[32]; };

// first index holding the maximum value
__device__ void block_argmax(float v, long long idx, ArgMaxSmem &sm, float &out_v, long long &out_i)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		float ov = __shfl_xor_sync(RB_FULL_MASK, v, o);
		long long oi = __shfl_xor_sync(RB_FULL_MASK, idx, o);
		if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
	}
	__syncthreads();
	if (lane == 0) { sm.v[wid] = v; sm.i[wid] = idx; }
	__syncthreads();
	v = lane < nw ? sm.v[lane] : RB_LOWEST; idx = lane < nw ? sm.i[lane] : 0x7fffffffffffffffLL;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		float ov = __shfl_xor_sync(RB_FULL_MASK, v, o);
		long long oi = __shfl_xor_sync(RB_FULL_MASK, idx, o);
		if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
	}
	out_v = v; out_i = idx;
}

struct DenseOut { float min_diff2, wmax, max_weight; long long max_index; SelResult sel; };

// dense [n_orient][T] conversion in place (weights_exponent_coarse + exponentiate + significance + argmax)
__device__ DenseOut dense_convert(float *w, long long n, int T, const float *po, const unsigned char *pz,
                                  const float *pt, const unsigned char *tz, float min_diff2,
                                  double adaptive_fraction, int maxsig, SelSmem &sm, ArgMaxSmem &am, float *fred)
{
	DenseOut o;
	o.min_diff2 = min_diff2;
	float mx = RB_LOWEST;
	for (long long i = threadIdx.x; i < n; i += blockDim.x)
	{
		const long long io = i / T; const int it = (int) (i - io * T);
		const float d = w[i];
		float l;
		if (d < min_diff2 || pz[io] || tz[it]) l = RB_LOWEST;                      // helper.cuh:39-42
		else l = po[io] + pt[it] + min_diff2 - d;
		w[i] = l;
		mx = fmaxf(mx, l);
	}
	o.wmax = block_max(mx, fred);
	const float add = 50.f - o.wmax;                                               // acc_ml_optimiser_impl.h:2201-2207
	float bv = RB_LOWEST; long long bi = 0x7fffffffffffffffLL;
	__syncthreads();
	for (long long i = threadIdx.x; i < n; i += blockDim.x)
	{
		const float a = w[i] + add;
		const float e = (a < -88.f) ? 0.f : expf(a);                               // helper.cuh:57-66
		w[i] = e;
		if (e > bv) { bv = e; bi = i; }
	}
	__syncthreads();
	block_argmax(bv, bi, am, o.max_weight, o.max_index);
	o.sel = block_significance(w, n, adaptive_fraction, maxsig, sm);
	return o;
}

// ---------------------------------------------------------------------------------------------
// pool kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WT_THREADS)
k_weights_coarse(const RbPartMeta *metas, RbPartState *states, float *Mweight,
                 const float *pdf_orient, const unsigned char *pdf_orient_zero,
                 const float *pdf_offset, const unsigned char *pdf_offset_zero, RbModelDev M, int T)
{
	__shared__ SelSmem sm;
	__shared__ ArgMaxSmem am;
	__shared__ float fred[32];
	const int p = blockIdx.x;
	const RbPartMeta m = metas[p];
	RbPartState *st = states + p;
	const long long n = (long long) M.nr_classes * m.nd * m.np * T;
	float *w = Mweight + m.coarse_off;
	const float min_diff2 = __int_as_float(st->min_diff2_bits);
	DenseOut o = dense_convert(w, n, T, pdf_orient + m.prior_off, pdf_orient_zero + m.prior_off,
	                           pdf_offset + (size_t) p * M.prior_classes() * T, pdf_offset_zero + (size_t) p * M.prior_classes() * T, min_diff2,
	                           M.adaptive_fraction, M.maximum_significants, sm, am, fred);
	if (threadIdx.x == 0)
	{
		st->min_diff2 = min_diff2;
		st->cmax_weight = o.max_weight; st->cmax_index = o.max_index;
		st->csum_weight = o.sel.sum_f; st->n_nonzero = (int) o.sel.n_nonzero;
		if (n == 1) { st->csig_weight = 0.f; st->nr_sig_coarse = 1; }                  // :2347-2350
		else
		{
			st->csig_weight = o.sel.sig_w;
			st->nr_sig_coarse = (int) (o.sel.n_nonzero - o.sel.thr_idx);
			if (o.sel.n_nonzero == 0 || st->nr_sig_coarse == 0) st->status = RB_ERR_NO_SIGNIFICANT;   // :2242, :2282
		}
	}
}

// ---- the same conversion for LARGE dense arrays (global searches: K * n_dir * n_psi * T ~ 10^6..10^7 per particle) ----
// One CTA per particle would stream ~13 passes over megabytes with 512 threads (65 ms for a 256-particle pool at
// HEALPix order 3, where at low SNR most weights are non-zero).  Here every phase is spread over (chunk, particle) CTAs:
//   k_wc_max     max of the log-weights                                         (read)
//   k_wc_exp     weights in place + per-particle histogram over the top 13 bits of the fp32 pattern: (mass, count) per
//                bin.  A bin holds values of ONE binary exponent e, so its mass is 2^(e-23) times the INTEGER sum of the
//                24-bit significands: native 64-bit integer atomics, exact and independent of their order.  (read + write)
//   k_wc_pick    one CTA per particle: total mass, threshold, the bin in which the cumulative mass crosses it
//   k_wc_gather  compaction of the elements of that bin (a tiny fraction)                                (read)
//   k_wc_finish  the radix descent of block_significance on the compact list, seeded with the mass / count below the bin
static const int WC_THREADS = 256;
static const int WC_CHUNK = 1 << 15;            // elements per CTA
static const int WC_BINS = 4096;                // top 13 bits of a positive float: 8 exponent + 4 mantissa bits
static const int WC_SHIFT = 19;

struct WcPick {
	int mode;                 // 0: by mass in bin `bin`; 1: by rank (`rank`) in bin `bin`; -1: no non-zero weight
	int bin; double base; long long cbelow;
	int bin_r; long long cbelow_r, rank_r;   // maxsig candidate (rank mode), bin_r < 0: none
	double total; long long n_nonzero; double thr;
	long long off_a, off_r;   // compact list bases are fixed (particle block); counts:
	int cnt_a, cnt_r;
};

struct WcArgs {
	const RbPartMeta *metas; RbPartState *states; float *Mweight;
	const float *pdf_orient; const unsigned char *pdf_orient_zero;
	const float *pdf_offset; const unsigned char *pdf_offset_zero;
	int T, Tp, nchunk, P; long long n;           // n: dense elements per particle (identical for all particles); Tp: floats of pdf_offset per particle
	float *pmax; float *pav; long long *pai;
	unsigned long long *hsum; int *hcnt;         // [P][WC_BINS]: sum of the 24-bit significands of the bin's values (exact), count
	WcPick *pick;                                // [P]
	long long *gtot;                             // [P][2] number of gathered elements (picked bin, maxsig bin)
	long long *ntot;                             // [P][2] number of non-zero weights, number of them below the picked bin
	int counts;                                  // per-bin counts are kept (maxsig)
	float *compact_a, *compact_r;                // [P][cap]
	long long cap;
};

// Both streaming kernels evaluate the log-weight of dense element (io, it) like cuda_kernel_weights_exponent_coarse
// (helper.cuh:16-46), four consecutive elements per thread (one 16-byte load when the particle block is 16-byte aligned),
// with the translation priors of the particle in shared memory.
__device__ __forceinline__ void wc_load4(const float *w, long long i, long long i1, bool vec, float (&v)[4])
{
	if (vec && i + 3 < i1) { const float4 q = *(const float4 *) (w + i); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
	else
#pragma unroll
		for (int j = 0; j < 4; j++) v[j] = (i + j < i1) ? w[i + j] : 0.f;
}

// Each CTA walks its chunk in steps of WC_STEP = 8 * WC_THREADS elements: two 16-byte loads per thread in flight.  A thread's
// run of four consecutive elements (io, it), (io, it + 1), ... needs the orientation prior only when `it` wraps, so the
// cursor carries it: RB_LOWEST stands for "prior is zero" (helper.cuh:39-42), for orientations and translations alike.
static const int WC_STEP = 8 * WC_THREADS;
struct WcCursor {
	int io, it, dio, dit, T;
	const float *po; const unsigned char *pz;
	__device__ __forceinline__ WcCursor(long long i, int T_, int stride, const float *po_, const unsigned char *pz_) : T(T_), po(po_), pz(pz_)
	{
		io = (int) (i / T_); it = (int) (i - (long long) io * T_);
		dio = stride / T_; dit = stride - dio * T_;
	}
	__device__ __forceinline__ void advance() { io += dio; it += dit; if (it >= T) { it -= T; io++; } }
	__device__ __forceinline__ float prior(int o) const { return pz[o] ? RB_LOWEST : po[o]; }
};

// log-weights of the four elements starting at the cursor (elements at or beyond i1 get RB_LOWEST when CHECK).
// s_pt holds the translation priors twice over (index it + j needs no wrap), RB_LOWEST where the prior is zero: with a
// lowest orientation or translation prior the sum stays at (or below) RB_LOWEST, so one max() replaces the tests of
// cuda_kernel_weights_exponent_coarse (helper.cuh:39-42) and valid elements see exactly its arithmetic.
template <bool CHECK>
__device__ __forceinline__ void wc_logw4(const WcCursor &cur, const float *s_pt, float min_diff2, const float (&v)[4], long long i, long long i1,
                                         float (&l)[4])
{
	if (cur.T >= 4)
	{
		// at most one orientation boundary inside the run: elements j < r belong to io, the others to io + 1
		const int r = cur.T - cur.it;
		const float po0 = (!CHECK || i < i1) ? cur.prior(cur.io) : RB_LOWEST;
		const float po1 = (r < 4 && (!CHECK || i + r < i1)) ? cur.prior(cur.io + 1) : RB_LOWEST;
#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			const float po = j < r ? po0 : po1;
			const float lw = fmaxf(po + s_pt[cur.it + j] + min_diff2 - v[j], RB_LOWEST);
			l[j] = (v[j] < min_diff2 || (CHECK && i + j >= i1)) ? RB_LOWEST : lw;
		}
		return;
	}
	int io = cur.io, it = cur.it;
	float po = (!CHECK || i < i1) ? cur.prior(io) : RB_LOWEST;
#pragma unroll
	for (int j = 0; j < 4; j++)
	{
		const float lw = fmaxf(po + s_pt[it] + min_diff2 - v[j], RB_LOWEST);
		l[j] = (v[j] < min_diff2 || (CHECK && i + j >= i1)) ? RB_LOWEST : lw;
		if (j < 3 && ++it == cur.T) { it = 0; io++; po = (!CHECK || i + j + 1 < i1) ? cur.prior(io) : RB_LOWEST; }
	}
}

// translation priors of particle p, twice over: s_pt[k] = prior[k mod T] for k < 2T (T <= 64)
__device__ __forceinline__ void wc_stage_priors(const WcArgs &A, int p, float *s_pt)
{
	if (threadIdx.x < 2 * A.T)
	{
		const int t = threadIdx.x < A.T ? threadIdx.x : threadIdx.x - A.T;
		s_pt[threadIdx.x] = A.pdf_offset_zero[(size_t) p * A.Tp + t] ? RB_LOWEST : A.pdf_offset[(size_t) p * A.Tp + t];   // block 0 (:2187-2196)
	}
}

__global__ void __launch_bounds__(WC_THREADS)
k_wc_max(WcArgs A)
{
	__shared__ float fred[32];
	__shared__ float s_pt[128];
	const int c = blockIdx.x, p = blockIdx.y;
	const RbPartMeta m = A.metas[p];
	const float min_diff2 = __int_as_float(A.states[p].min_diff2_bits);
	const float *w = A.Mweight + m.coarse_off;
	const bool vec = (m.coarse_off & 3) == 0;
	wc_stage_priors(A, p, s_pt);
	__syncthreads();
	const long long i0 = (long long) c * WC_CHUNK, i1 = min(A.n, i0 + WC_CHUNK);
	float mx = RB_LOWEST;
	WcCursor cur(i0 + 4 * threadIdx.x, A.T, 4 * WC_THREADS, A.pdf_orient + m.prior_off, A.pdf_orient_zero + m.prior_off);
	for (long long base = i0; base < i1; base += WC_STEP)
	{
		float v[2][4], l[4];
		const long long ia = base + 4 * threadIdx.x, ib = ia + 4 * WC_THREADS;
		const bool full = vec && base + WC_STEP <= i1;
		wc_load4(w, ia, i1, vec, v[0]);
		wc_load4(w, ib, i1, vec, v[1]);
#pragma unroll
		for (int h = 0; h < 2; h++)
		{
			if (full) wc_logw4<false>(cur, s_pt, min_diff2, v[h], h ? ib : ia, i1, l);
			else wc_logw4<true>(cur, s_pt, min_diff2, v[h], h ? ib : ia, i1, l);
			mx = fmaxf(fmaxf(mx, fmaxf(l[0], l[1])), fmaxf(l[2], l[3]));
			cur.advance();
		}
	}
	mx = block_max(mx, fred);
	if (threadIdx.x == 0) A.pmax[(size_t) p * A.nchunk + c] = mx;
}

// Histogram words per bin in shared memory: 32-bit counters only (a 64-bit shared atomic add compiles to a compare-and-swap
// spin loop).  Lanes of a warp that hit the same bin are combined first (match.any + redux), then one lane adds the combined
// significand sum to a 32-bit word and counts its (rare) carries in a second one.
// COUNTS: also the number of elements per bin (only the --maxsig rank search needs it; otherwise the total number of
// non-zero weights comes from the ballots here and the count below the picked bin from k_wc_gather).
template <bool COUNTS>
__global__ void __launch_bounds__(WC_THREADS)
k_wc_exp(WcArgs A)
{
	extern __shared__ unsigned char wc_smem[];
	unsigned *h_lo = (unsigned *) wc_smem;                       // [WC_BINS] low word of the significand sum
	unsigned *h_hi = h_lo + WC_BINS;                             // [WC_BINS] its carries
	unsigned *h_c = h_hi + WC_BINS;                              // [WC_BINS] count (COUNTS)
	__shared__ ArgMaxSmem am;
	__shared__ float s_wmax;
	__shared__ float s_pt[128];
	__shared__ int s_nz;
	const int c = blockIdx.x, p = blockIdx.y;
	const int lane = threadIdx.x & 31;
	const RbPartMeta m = A.metas[p];
	const float min_diff2 = __int_as_float(A.states[p].min_diff2_bits);
	float *w = A.Mweight + m.coarse_off;
	const bool vec = (m.coarse_off & 3) == 0;
	for (int i = threadIdx.x; i < (COUNTS ? 3 : 2) * WC_BINS; i += WC_THREADS) h_lo[i] = 0u;
	wc_stage_priors(A, p, s_pt);
	if (threadIdx.x == 0) s_nz = 0;
	if (threadIdx.x < 32)
	{
		float mx = RB_LOWEST;
		for (int k = threadIdx.x; k < A.nchunk; k += 32) mx = fmaxf(mx, A.pmax[(size_t) p * A.nchunk + k]);
		mx = warp_max(mx);
		if (threadIdx.x == 0) s_wmax = mx;
	}
	__syncthreads();
	const float add = 50.f - s_wmax;                                               // acc_ml_optimiser_impl.h:2201-2207
	const long long i0 = (long long) c * WC_CHUNK, i1 = min(A.n, i0 + WC_CHUNK);
	float bv = RB_LOWEST; int brel = 0x7fffffff;                                    // maximum and its index relative to i0
	int nz_count = 0;                                                              // per thread
	WcCursor cur(i0 + 4 * threadIdx.x, A.T, 4 * WC_THREADS, A.pdf_orient + m.prior_off, A.pdf_orient_zero + m.prior_off);
	for (long long base = i0; base < i1; base += WC_STEP)                          // warp-uniform trip count
	{
		float v[2][4];
		const long long ia = base + 4 * threadIdx.x, ib = ia + 4 * WC_THREADS;
		const bool full = vec && base + WC_STEP <= i1;
		wc_load4(w, ia, i1, vec, v[0]);
		wc_load4(w, ib, i1, vec, v[1]);
#pragma unroll
		for (int h = 0; h < 2; h++)
		{
			const long long i = h ? ib : ia;
			float l[4], e[4];
			if (full) wc_logw4<false>(cur, s_pt, min_diff2, v[h], i, i1, l);
			else wc_logw4<true>(cur, s_pt, min_diff2, v[h], i, i1, l);
#pragma unroll
			for (int j = 0; j < 4; j++)
			{
				const float a = l[j] + add;
				e[j] = (a < -88.f) ? 0.f : expf(a);                                // helper.cuh:57-66 (elements beyond i1: l = lowest -> 0)
				const int rel = (int) (i - i0) + j;
				if (e[j] > bv && (full || i + j < i1)) { bv = e[j]; brel = rel; }
				// zero weights take part as (bin 0, significand 0): no divergence around the warp-wide match
				const unsigned bits = __float_as_uint(e[j]);
				const int bin = (int) (bits >> WC_SHIFT);
				// significand as an integer (normal: implicit one; denormal: exponent field 0)
				const unsigned sig = (bits >> 23) ? ((bits & 0x7fffffu) | 0x800000u) : (bits & 0x7fffffu);
				nz_count += bits != 0u;
				const unsigned peers = __match_any_sync(0xffffffffu, bin);
				const unsigned ssum = __reduce_add_sync(peers, sig);               // <= 32 * 2^24
				if (ssum && lane == __ffs(peers) - 1)
				{
					const unsigned old = atomicAdd(h_lo + bin, ssum);              // wraps at most 2^7 times per chunk
					if (old + ssum < old) atomicAdd(h_hi + bin, 1u);               // ... and the carries are counted
					if (COUNTS) atomicAdd(h_c + bin, (unsigned) __popc(peers));
				}
			}
			if (full || (vec && i + 3 < i1)) *(float4 *) (w + i) = make_float4(e[0], e[1], e[2], e[3]);
			else
#pragma unroll
				for (int j = 0; j < 4; j++) if (i + j < i1) w[i + j] = e[j];
			cur.advance();
		}
	}
	float ov; long long oi;
	block_argmax(bv, brel == 0x7fffffff ? 0x7fffffffffffffffLL : i0 + brel, am, ov, oi);
	if (threadIdx.x == 0) { A.pav[(size_t) p * A.nchunk + c] = ov; A.pai[(size_t) p * A.nchunk + c] = oi; }
	nz_count = (int) __reduce_add_sync(0xffffffffu, (unsigned) nz_count);
	if (lane == 0 && nz_count) atomicAdd(&s_nz, nz_count);
	__syncthreads();
	if (threadIdx.x == 0 && s_nz) atomicAdd((unsigned long long *) (A.ntot + 2 * p), (unsigned long long) s_nz);
	for (int i = threadIdx.x; i < WC_BINS; i += WC_THREADS)
		if (h_lo[i] | h_hi[i])
		{
			atomicAdd(A.hsum + (size_t) p * WC_BINS + i, ((unsigned long long) h_hi[i] << 32) + h_lo[i]);
			if (COUNTS) atomicAdd(A.hcnt + (size_t) p * WC_BINS + i, (int) h_c[i]);
		}
}

// exact mass of histogram bin b: integer significand sum times 2^(exponent - 23)  (denormal bins: exponent field 0 -> 2^-149)
__device__ __forceinline__ double wc_bin_mass(unsigned long long sig_sum, int bin)
{
	const int ef = bin >> 4;                          // the 8 exponent bits of the pattern
	return ldexp((double) sig_sum, (ef ? ef : 1) - 127 - 23);
}

// one CTA per particle: argmax over chunks, total mass, threshold, bin selection
__global__ void __launch_bounds__(WC_THREADS)
k_wc_pick(WcArgs A, RbModelDev M)
{
	__shared__ double dred[32];
	__shared__ long long lred[32];
	const int p = blockIdx.x;
	__shared__ double hs[WC_BINS];
	const int *hc = A.hcnt + (size_t) p * WC_BINS;
	double t = 0.; long long cnt = 0;
	for (int i = threadIdx.x; i < WC_BINS; i += WC_THREADS)
	{
		hs[i] = wc_bin_mass(A.hsum[(size_t) p * WC_BINS + i], i); t += hs[i];      // a non-empty bin has a positive mass
		if (A.counts) cnt += hc[i];
	}
	t = block_sum(t, dred);
	cnt = block_sum(cnt, lred);
	if (threadIdx.x != 0) return;
	if (!A.counts) cnt = A.ntot[2 * p];
	RbPartState *st = A.states + p;
	{
		float bv = RB_LOWEST; long long bi = 0x7fffffffffffffffLL;
		for (int c = 0; c < A.nchunk; c++)
		{
			const size_t k = (size_t) p * A.nchunk + c;
			if (A.pav[k] > bv) { bv = A.pav[k]; bi = A.pai[k]; }                   // chunks in order: first index of the maximum
		}
		st->cmax_weight = bv; st->cmax_index = bi;
	}
	WcPick pk;
	memset(&pk, 0, sizeof(pk));
	pk.total = t; pk.n_nonzero = cnt; pk.bin_r = -1; pk.mode = -1;
	if (cnt > 0)
	{
		const float sum_f = (float) t;
		pk.thr = (double) (float) ((1. - M.adaptive_fraction) * (double) sum_f);   // (XFLOAT) threshold, :2277-2278
		double cum = 0.; long long cc = 0; int sel = -1;
		for (int b = 0; b < WC_BINS; b++)
		{
			if (hs[b] > 0. && cum + hs[b] > pk.thr) { sel = b; break; }
			cum += hs[b]; if (A.counts) cc += hc[b];
		}
		if (sel >= 0) { pk.mode = 0; pk.bin = sel; pk.base = cum; pk.cbelow = cc; }
		else
		{
			// nothing crosses the threshold: the reference's search leaves idx = 0 -> sorted[0], the smallest weight
			int b = 0; while (hs[b] <= 0.) b++;
			pk.mode = 1; pk.bin = b; pk.base = 0.; pk.cbelow = 0;
		}
		if (M.maximum_significants > 0 && cnt > M.maximum_significants)            // :2301-2306 candidate
		{
			const long long r = cnt - M.maximum_significants;
			long long c2 = 0; int b = 0;
			for (; b < WC_BINS; b++) { if (c2 + hc[b] > r) break; c2 += hc[b]; }
			pk.bin_r = b; pk.cbelow_r = c2; pk.rank_r = r;
		}
	}
	A.pick[p] = pk;
}

// compaction of the elements falling in the picked bin(s).  Space is reserved per tile with one atomic, so the order of the
// compact list varies from run to run, but its elements share one binary exponent: every fp64 partial sum over them is exact and
// the selection that follows does not depend on the order.
__global__ void __launch_bounds__(WC_THREADS)
k_wc_gather(WcArgs A)
{
	__shared__ int s_scan[2][WC_THREADS];
	__shared__ long long s_base[2];
	const int c = blockIdx.x, p = blockIdx.y;
	const WcPick pk = A.pick[p];
	if (pk.mode < 0) return;
	const RbPartMeta m = A.metas[p];
	const float *w = A.Mweight + m.coarse_off;
	const long long i0 = (long long) c * WC_CHUNK, i1 = min(A.n, i0 + WC_CHUNK);
	const unsigned ba = (unsigned) pk.bin, br = pk.bin_r >= 0 ? (unsigned) pk.bin_r : 0xffffffffu;
	const int RUN = 8;
	int kb = 0;                                                   // non-zero weights below the picked bin (when no per-bin counts are kept)
	for (long long t0 = i0; t0 < i1; t0 += (long long) WC_THREADS * RUN)
	{
		float v[RUN]; int ka = 0, kr = 0;
		const long long b = t0 + (long long) threadIdx.x * RUN;
#pragma unroll
		for (int j = 0; j < RUN; j++)
		{
			v[j] = (b + j < i1) ? w[b + j] : 0.f;
			const unsigned bin = __float_as_uint(v[j]) >> WC_SHIFT;
			ka += (v[j] > 0.f && bin == ba); kr += (v[j] > 0.f && bin == br);
			kb += (v[j] > 0.f && bin < ba);
		}
		if (!__syncthreads_or(ka | kr)) continue;
		s_scan[0][threadIdx.x] = ka; s_scan[1][threadIdx.x] = kr;
		__syncthreads();
		for (int off = 1; off < WC_THREADS; off <<= 1)
		{
			const int ta = threadIdx.x >= off ? s_scan[0][threadIdx.x - off] : 0, tr = threadIdx.x >= off ? s_scan[1][threadIdx.x - off] : 0;
			__syncthreads();
			s_scan[0][threadIdx.x] += ta; s_scan[1][threadIdx.x] += tr;
			__syncthreads();
		}
		if (threadIdx.x == WC_THREADS - 1)
		{
			s_base[0] = s_scan[0][threadIdx.x] ? (long long) atomicAdd((unsigned long long *) (A.gtot + 2 * p), (unsigned long long) s_scan[0][threadIdx.x]) : 0;
			s_base[1] = s_scan[1][threadIdx.x] ? (long long) atomicAdd((unsigned long long *) (A.gtot + 2 * p + 1), (unsigned long long) s_scan[1][threadIdx.x]) : 0;
		}
		__syncthreads();
		long long pa = s_base[0] + s_scan[0][threadIdx.x] - ka, pr = s_base[1] + s_scan[1][threadIdx.x] - kr;
#pragma unroll
		for (int j = 0; j < RUN; j++)
		{
			const unsigned bin = __float_as_uint(v[j]) >> WC_SHIFT;
			if (v[j] > 0.f && bin == ba) { if (pa < A.cap) A.compact_a[(size_t) p * A.cap + pa] = v[j]; pa++; }
			if (v[j] > 0.f && bin == br) { if (pr < A.cap) A.compact_r[(size_t) p * A.cap + pr] = v[j]; pr++; }
		}
		__syncthreads();
	}
	if (!A.counts)
	{
		__shared__ int s_red[32];
		kb = block_sum(kb, s_red);
		if (threadIdx.x == 0 && kb) atomicAdd((unsigned long long *) (A.ntot + 2 * p + 1), (unsigned long long) kb);
	}
}

// Radix descent over a compact list whose elements all share the top 13 bits `bin`; seeded with the mass / count below.
__device__ void radix_descend_seeded(const float *w, long long n, bool by_rank, double thr, long long rank,
                                     double base0, long long cbelow0, SelSmem &sm)
{
	if (threadIdx.x == 0) { sm.prefix = 0u; sm.base = base0; sm.cbelow = cbelow0; sm.cequal = 0; }
	__syncthreads();
	for (int pass = 0; pass < 8; pass++)
	{
		const int shift = 28 - 4 * pass;
		const unsigned prefix = sm.prefix;
		double s[16]; int c[16];
#pragma unroll
		for (int b = 0; b < 16; b++) { s[b] = 0.; c[b] = 0; }
		for (long long i = threadIdx.x; i < n; i += blockDim.x)
		{
			const float v = w[i];
			const unsigned bits = __float_as_uint(v);
			if (v > 0.f && (pass == 0 || (bits >> (shift + 4)) == (prefix >> (shift + 4))))
			{
				const int bin = (bits >> shift) & 15;
#pragma unroll
				for (int b = 0; b < 16; b++) { const bool h = (bin == b); s[b] += h ? (double) v : 0.; c[b] += h ? 1 : 0; }
			}
		}
		reduce_bins(s, c, sm);
		if (threadIdx.x == 0)
		{
			double cum = sm.base; long long cc = sm.cbelow; int sel = -1;
			for (int b = 0; b < 16; b++)
			{
				if (sm.bc[b] > 0)
				{
					bool hit = by_rank ? (cc + sm.bc[b] > rank) : (cum + sm.bs[b] > thr);
					if (hit) { sel = b; break; }
				}
				cum += sm.bs[b]; cc += sm.bc[b];
			}
			if (sel < 0) sm.found = 0;
			else
			{
				sm.found = 1;
				sm.prefix = prefix | ((unsigned) sel << shift);
				sm.base = cum; sm.cbelow = cc; sm.cequal = sm.bc[sel];
			}
		}
		__syncthreads();
		if (!sm.found) return;
	}
}

__global__ void __launch_bounds__(WT_THREADS)
k_wc_finish(WcArgs A, RbModelDev M)
{
	__shared__ SelSmem sm;
	__shared__ int s_over;
	const int p = blockIdx.x;
	RbPartState *st = A.states + p;
	const WcPick pk = A.pick[p];
	if (pk.mode < 0)
	{
		if (threadIdx.x == 0)
		{
			st->min_diff2 = __int_as_float(st->min_diff2_bits);
			st->csum_weight = 0.f; st->n_nonzero = 0; st->csig_weight = 0.f; st->nr_sig_coarse = 0; st->status = RB_ERR_NO_SIGNIFICANT;
		}
		return;
	}
	const long long na = A.gtot[2 * p], nr = A.gtot[2 * p + 1];
	if (threadIdx.x == 0) s_over = (na > A.cap || nr > A.cap);
	__syncthreads();
	if (s_over)
	{
		if (threadIdx.x == 0) st->status = RB_ERR_CAPACITY;    // a single 1/16-octave bin larger than the compact buffer
		return;
	}
	float sig_w; long long thr_idx;
	const long long cbelow = A.counts ? pk.cbelow : (pk.mode == 1 ? 0 : A.ntot[2 * p + 1]);
	radix_descend_seeded(A.compact_a + (size_t) p * A.cap, na, pk.mode == 1, pk.thr, 0, pk.base, cbelow, sm);
	__syncthreads();
	if (pk.mode == 1) { thr_idx = 0; sig_w = __uint_as_float(sm.prefix); }
	else
	{
		const float v = __uint_as_float(sm.prefix);
		double kk = floor((pk.thr - sm.base) / (double) v) + 1.;                   // k-th copy of v crosses the threshold
		long long k = kk < 1. ? 1 : (kk > (double) sm.cequal ? sm.cequal : (long long) kk);
		thr_idx = sm.cbelow + k - 1;
		sig_w = v;
	}
	__syncthreads();
	if (pk.bin_r >= 0 && pk.n_nonzero - thr_idx > M.maximum_significants)          // :2301-2306
	{
		thr_idx = pk.rank_r;
		radix_descend_seeded(A.compact_r + (size_t) p * A.cap, nr, true, 0., pk.rank_r, 0., pk.cbelow_r, sm);
		__syncthreads();
		sig_w = __uint_as_float(sm.prefix);
	}
	if (threadIdx.x == 0)
	{
		st->min_diff2 = __int_as_float(st->min_diff2_bits);
		st->csum_weight = (float) pk.total; st->n_nonzero = (int) pk.n_nonzero;
		st->csig_weight = sig_w;
		st->nr_sig_coarse = (int) (pk.n_nonzero - thr_idx);
		if (pk.n_nonzero == 0 || st->nr_sig_coarse == 0) st->status = RB_ERR_NO_SIGNIFICANT;   // :2242, :2282
	}
}

static int weights_coarse_large(rb_ctx *ctx, PoolSlot &s, long long n)
{
	WcArgs A;
	memset(&A, 0, sizeof(A));
	const int P = s.P;
	A.metas = s.meta.as<RbPartMeta>(); A.states = s.state.as<RbPartState>(); A.Mweight = s.Mweight.as<float>();
	A.pdf_orient = s.pdf_orient.as<float>(); A.pdf_orient_zero = s.pdf_orient_zero.as<unsigned char>();
	A.pdf_offset = s.pdf_offset.as<float>(); A.pdf_offset_zero = s.pdf_offset_zero.as<unsigned char>();
	A.T = ctx->d_samp.n_trans; A.Tp = ctx->d_model.prior_classes() * A.T; A.n = n; A.nchunk = (int) ((n + WC_CHUNK - 1) / WC_CHUNK); A.P = P;
	const size_t np = (size_t) P * A.nchunk;
	// one 1/16-octave bin of one particle; n / 8 is generous (the weights span ~200 octaves), overflow -> RB_ERR_CAPACITY
	A.cap = std::max<long long>(1024, n / 8);
	const bool maxsig = ctx->d_model.maximum_significants > 0;
	RB_CHECK(ctx->wc_buf[0].ensure(np * 4)); RB_CHECK(ctx->wc_buf[1].ensure(np * 4)); RB_CHECK(ctx->wc_buf[2].ensure(np * 8));
	RB_CHECK(ctx->wc_buf[3].ensure((size_t) P * WC_BINS * 8)); RB_CHECK(ctx->wc_buf[4].ensure((size_t) P * WC_BINS * 4));
	RB_CHECK(ctx->wc_buf[5].ensure((size_t) P * sizeof(WcPick)));
	RB_CHECK(ctx->wc_buf[6].ensure((size_t) P * 16)); RB_CHECK(ctx->wc_buf[7].ensure((size_t) P * 16));
	RB_CHECK(ctx->wc_buf[8].ensure((size_t) P * A.cap * 4));
	RB_CHECK(ctx->wc_buf[9].ensure(maxsig ? (size_t) P * A.cap * 4 : 16));
	A.pmax = ctx->wc_buf[0].as<float>(); A.pav = ctx->wc_buf[1].as<float>(); A.pai = ctx->wc_buf[2].as<long long>();
	A.hsum = ctx->wc_buf[3].as<unsigned long long>(); A.hcnt = ctx->wc_buf[4].as<int>(); A.pick = ctx->wc_buf[5].as<WcPick>();
	A.gtot = ctx->wc_buf[6].as<long long>(); A.ntot = ctx->wc_buf[7].as<long long>(); A.counts = maxsig ? 1 : 0;
	A.compact_a = ctx->wc_buf[8].as<float>(); A.compact_r = ctx->wc_buf[9].as<float>();
	RB_CUDA(cudaMemsetAsync(A.hsum, 0, (size_t) P * WC_BINS * 8, ctx->stream));
	RB_CUDA(cudaMemsetAsync(A.hcnt, 0, (size_t) P * WC_BINS * 4, ctx->stream));
	RB_CUDA(cudaMemsetAsync(A.gtot, 0, (size_t) P * 16, ctx->stream));
	RB_CUDA(cudaMemsetAsync(A.ntot, 0, (size_t) P * 16, ctx->stream));
	dim3 grid(A.nchunk, P);
	static bool configured_dev[RB_MAX_DEVICES] = {};
	bool &configured = configured_dev[ctx->device % RB_MAX_DEVICES];
	if (!configured)
	{
		RB_CUDA(cudaFuncSetAttribute(k_wc_exp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WC_BINS * 12));
		RB_CUDA(cudaFuncSetAttribute(k_wc_exp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WC_BINS * 8));
		configured = true;
	}
	k_wc_max<<<grid, WC_THREADS, 0, ctx->stream>>>(A); RB_LAUNCH_CHECK(ctx);
	if (maxsig) k_wc_exp<true><<<grid, WC_THREADS, WC_BINS * 12, ctx->stream>>>(A);
	else k_wc_exp<false><<<grid, WC_THREADS, WC_BINS * 8, ctx->stream>>>(A);
	RB_LAUNCH_CHECK(ctx);
	k_wc_pick<<<P, WC_THREADS, 0, ctx->stream>>>(A, ctx->d_model); RB_LAUNCH_CHECK(ctx);
	k_wc_gather<<<grid, WC_THREADS, 0, ctx->stream>>>(A); RB_LAUNCH_CHECK(ctx);
	k_wc_finish<<<P, WT_THREADS, 0, ctx->stream>>>(A, ctx->d_model); RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// ---- cross-correlation criterion (first iteration, --firstiter_cc / --always_cc) --------------------------------------
// convertAllSquaredDifferencesToWeights, acc_ml_optimiser_impl.h:2012-2071: no priors, no exponentials.  The pose with the
// smallest value gets weight one, every other weight is zero; significant_weight = 0.999, NR_SIGN = 1, sum_weight keeps its
// initial value of one (:1986).  Ties go to the first index (getArgMinOnDevice).  Entries that were never computed
// (lowest(), orientations without prior) are skipped.
__global__ void __launch_bounds__(WT_THREADS)
k_weights_cc_coarse(const RbPartMeta *metas, RbPartState *states, float *Mweight, RbModelDev M, int T)
{
	__shared__ ArgMaxSmem am;
	const int p = blockIdx.x;
	const RbPartMeta m = metas[p];
	RbPartState *st = states + p;
	const long long n = (long long) M.nr_classes * m.nd * m.np * T;
	float *w = Mweight + m.coarse_off;
	float bv = RB_LOWEST; long long bi = 0x7fffffffffffffffLL;
	for (long long i = threadIdx.x; i < n; i += blockDim.x)
	{
		const float d = w[i];
		if (d > RB_LOWEST && -d > bv) { bv = -d; bi = i; }
	}
	float best; long long arg;
	block_argmax(bv, bi, am, best, arg);
	const bool none = arg == 0x7fffffffffffffffLL;
	__syncthreads();
	for (long long i = threadIdx.x; i < n; i += blockDim.x) w[i] = (i == arg) ? 1.f : 0.f;
	if (threadIdx.x == 0)
	{
		st->min_diff2 = none ? 0.f : -best;
		st->cmax_weight = 1.f; st->cmax_index = none ? 0 : arg;
		st->csum_weight = 1.f; st->n_nonzero = none ? 0 : 1;
		st->csig_weight = 0.999f; st->nr_sig_coarse = none ? 0 : 1;
		if (none) st->status = RB_ERR_NO_SIGNIFICANT;
	}
}

__global__ void __launch_bounds__(WT_THREADS)
k_weights_cc_fine(RbPartState *states, float *fs_w, int NOR, int NOT, const int *counters)
{
	__shared__ ArgMaxSmem am;
	if (counters[2]) return;
	const int p = blockIdx.x;
	RbPartState *st = states + p;
	const long long n = (long long) st->n_pairs * NOR * NOT;
	if (n == 0) { if (threadIdx.x == 0 && st->status == 0) st->status = RB_ERR_NO_SIGNIFICANT; return; }
	float *w = fs_w + st->fs_base;
	float bv = RB_LOWEST; long long bi = 0x7fffffffffffffffLL;
	for (long long i = threadIdx.x; i < n; i += blockDim.x)
	{
		const float d = w[i];
		if (-d > bv) { bv = -d; bi = i; }
	}
	float best; long long arg;
	block_argmax(bv, bi, am, best, arg);
	if (arg == 0x7fffffffffffffffLL) arg = 0;                     // only NaNs: keep the reference's first element
	__syncthreads();
	for (long long i = threadIdx.x; i < n; i += blockDim.x) w[i] = (i == arg) ? 1.f : 0.f;
	if (threadIdx.x == 0)
	{
		st->fmin_diff2 = -best;                                                             // :1881
		st->min_diff2_final = (double) -best;                                               // no "+ 50 - max" (:2444 is the other branch)
		st->fmax_weight = 1.f; st->fmax_sample = arg;
		st->fsum_weight = 1.f; st->fsig_weight = 0.999f;
	}
}

int rbk_weights_coarse_pool(rb_ctx *ctx, PoolSlot &s)
{
	if (ctx->d_model.do_cc)
	{
		k_weights_cc_coarse<<<s.P, WT_THREADS, 0, ctx->stream>>>(s.meta.as<RbPartMeta>(), s.state.as<RbPartState>(),
			s.Mweight.as<float>(), ctx->d_model, ctx->d_samp.n_trans);
		RB_LAUNCH_CHECK(ctx);
		return RB_OK;
	}
	if (!s.has_priors)
	{
		// global search: every particle has the same dense size
		const long long n = (long long) ctx->d_model.nr_classes * ctx->d_samp.n_dir * ctx->d_samp.n_psi * ctx->d_samp.n_trans;
		const char *e = getenv("RB_WEIGHTS_LARGE");
		// measured against the one-CTA-per-particle kernel: 0.85 vs 1.12 ms at 96 768 elements per particle (HEALPix order 2,
		// 21 translations, 256 particles), 0.89 vs 1.01 ms at 12 600 (2D classification, K = 10, 2000 particles)
		const long long min_n = e ? atoll(e) : (1 << 13);
		if (n >= min_n && n > 1 && ctx->d_samp.n_trans <= 64) return weights_coarse_large(ctx, s, n);
	}
	k_weights_coarse<<<s.P, WT_THREADS, 0, ctx->stream>>>(s.meta.as<RbPartMeta>(), s.state.as<RbPartState>(),
		s.Mweight.as<float>(), s.pdf_orient.as<float>(), s.pdf_orient_zero.as<unsigned char>(),
		s.pdf_offset.as<float>(), s.pdf_offset_zero.as<unsigned char>(), ctx->d_model, ctx->d_samp.n_trans);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// ---- fine-pass setup -------------------------------------------------------------------------
static const int FS_THREADS = 256;

// Orientations are handled in chunks of FS_CHUNK so that large dense arrays (global searches: 10^5 orientations per particle)
// spread over many CTAs: count per (chunk, particle), prefix over the chunks of a particle, ordered fill per chunk.
static const int FS_CHUNK = 2048;

__global__ void __launch_bounds__(FS_THREADS)
k_fine_count(const RbPartMeta *metas, const RbPartState *states, const float *Mweight, RbModelDev M, int T, int nchunk, int *part_so, int *part_pair)
{
	__shared__ int red[32];
	const int c = blockIdx.x, p = blockIdx.y;
	const RbPartMeta m = metas[p];
	const int ndense = M.nr_classes * m.nd * m.np;
	const float *w = Mweight + m.coarse_off;
	const float sig = states[p].csig_weight;
	int nso = 0, npair = 0;
	const int o1 = min(ndense, (c + 1) * FS_CHUNK);
	for (int o = c * FS_CHUNK + threadIdx.x; o < o1; o += FS_THREADS)
	{
		int cnt = 0;
		for (int t = 0; t < T; t++) cnt += (w[(long long) o * T + t] >= sig) ? 1 : 0;   // arrayOverThreshold
		nso += cnt > 0; npair += cnt;
	}
	nso = block_sum(nso, red);
	npair = block_sum(npair, red);
	if (threadIdx.x == 0) { part_so[(size_t) p * nchunk + c] = nso; part_pair[(size_t) p * nchunk + c] = npair; }
}

// per particle: totals and exclusive prefix over its chunks (in place)
__global__ void k_fine_chunkscan(RbPartState *states, int nchunk, int *part_so, int *part_pair)
{
	const int p = blockIdx.x;
	if (threadIdx.x != 0) return;
	int a = 0, b = 0;
	for (int c = 0; c < nchunk; c++)
	{
		const size_t k = (size_t) p * nchunk + c;
		const int va = part_so[k], vb = part_pair[k];
		part_so[k] = a; part_pair[k] = b;
		a += va; b += vb;
	}
	states[p].n_so = a; states[p].n_pairs = b;
}

// exclusive scan over the particles of the pool (single CTA)
__global__ void __launch_bounds__(1024)
k_fine_scan(RbPartState *states, int P, int NOR, int NOT, long long cap_fo, long long cap_fs, int *counters)
{
	__shared__ long long s_a[1024], s_b[1024];
	const int per = (P + 1023) / 1024;
	const int i0 = threadIdx.x * per, i1 = min(P, i0 + per);
	long long a = 0, b = 0;
	for (int i = i0; i < i1; i++) { a += states[i].n_so; b += states[i].n_pairs; }
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
	long long ea = s_a[threadIdx.x] - a, eb = s_b[threadIdx.x] - b;   // exclusive prefix of this thread's chunk
	for (int i = i0; i < i1; i++)
	{
		states[i].so_base = ea; states[i].pair_base = eb;
		states[i].fo_base = ea * NOR; states[i].fs_base = eb * NOR * NOT;
		ea += states[i].n_so; eb += states[i].n_pairs;
	}
	if (threadIdx.x == 1023)
	{
		long long tfo = s_a[1023] * NOR, tfs = s_b[1023] * NOR * NOT;
		int overflow = (tfo > cap_fo) || (tfs > cap_fs) || tfo > 0x7fffffffLL;
		counters[0] = overflow ? 0 : (int) tfo;
		counters[2] = overflow;
		((long long *) counters)[2] = tfo;   // counters[4..5]
		((long long *) counters)[3] = tfs;   // counters[6..7]
	}
}

__global__ void __launch_bounds__(FS_THREADS)
k_fine_fill(const RbPartMeta *metas, const RbPartState *states, const float *Mweight,
            const int *dir_idx, const int *psi_idx, RbModelDev M, RbSamplingDev S, RbLR lr,
            int *pair_list, RbFineOrient *fo, long long *fs_ihid, const int *counters, int nchunk, const int *part_so, const int *part_pair)
{
	__shared__ int s_scan_f[FS_THREADS], s_scan_c[FS_THREADS];
	__shared__ int s_run_f, s_run_c;
	if (counters[2]) return;   // capacity overflow: host reports RB_ERR_CAPACITY
	const int chunk = blockIdx.x, p = blockIdx.y;
	const RbPartMeta m = metas[p];
	const RbPartState st = states[p];
	const int T = S.n_trans, NOR = S.n_over_rot, NOT = S.n_over_trans;
	const int no = m.nd * m.np, ndense = M.nr_classes * no;
	const float *w = Mweight + m.coarse_off;
	const float sig = st.csig_weight;
	if (chunk * FS_CHUNK >= ndense) return;
	if (threadIdx.x == 0) { s_run_f = part_so[(size_t) p * nchunk + chunk]; s_run_c = part_pair[(size_t) p * nchunk + chunk]; }
	__syncthreads();
	const int c_end = min(ndense, (chunk + 1) * FS_CHUNK);
	for (int c0 = chunk * FS_CHUNK; c0 < c_end; c0 += FS_THREADS)
	{
		const int o = c0 + threadIdx.x;
		int cnt = 0;
		if (o < c_end) for (int t = 0; t < T; t++) cnt += (w[(long long) o * T + t] >= sig) ? 1 : 0;
		const int flag = cnt > 0;
		s_scan_f[threadIdx.x] = flag; s_scan_c[threadIdx.x] = cnt;
		__syncthreads();
		for (int off = 1; off < FS_THREADS; off <<= 1)
		{
			int tf = 0, tc = 0;
			if (threadIdx.x >= off) { tf = s_scan_f[threadIdx.x - off]; tc = s_scan_c[threadIdx.x - off]; }
			__syncthreads();
			s_scan_f[threadIdx.x] += tf; s_scan_c[threadIdx.x] += tc;
			__syncthreads();
		}
		const int sidx = s_run_f + s_scan_f[threadIdx.x] - flag;     // index among significant orientations
		const int poff = s_run_c + s_scan_c[threadIdx.x] - cnt;      // pairs before this orientation
		if (flag)
		{
			const int k = o / no, oi = o - k * no, idl = oi / m.np, ipl = oi - idl * m.np;
			const int gd = m.dir_off < 0 ? idl : dir_idx[m.dir_off + idl];
			const int gp = m.psi_off < 0 ? ipl : psi_idx[m.psi_off + ipl];
			const long long pair_off = st.pair_base + poff;
			int j = 0;
			for (int t = 0; t < T; t++) if (w[(long long) o * T + t] >= sig) pair_list[pair_off + j++] = t;
			for (int io = 0; io < NOR; io++)
			{
				RbFineOrient F;
				F.particle = p; F.iclass = k; F.iorient = oi; F.iover_rot = io;
				F.pair_off = (int) pair_off; F.n_t = cnt;
				F.sample_off = st.fs_base + ((long long) poff * NOR + (long long) io * cnt) * NOT;
				double rot, tilt, psi;
				if (S.over_rot)
				{
					const size_t g = ((size_t) gd * S.n_psi + gp) * NOR + io;
					rot = S.over_rot[g]; tilt = S.over_tilt[g]; psi = S.over_psi[g];
				}
				else { rot = S.rot[gd]; tilt = S.tilt[gd]; psi = S.psi[gp]; }
				rb_euler_fine(rot, tilt, psi, lr, F.e);
				fo[st.fo_base + (long long) sidx * NOR + io] = F;
				// ihidden_over (acc_helper_functions_impl.h:63)
				j = 0;
				for (int t = 0; t < T; t++)
					if (w[(long long) o * T + t] >= sig)
					{
						const long long ihidden = (long long) o * T + t;
						for (int iot = 0; iot < NOT; iot++)
							fs_ihid[F.sample_off + (long long) j * NOT + iot] = (ihidden * NOR + io) * NOT + iot;
						j++;
					}
			}
		}
		__syncthreads();
		if (threadIdx.x == FS_THREADS - 1) { s_run_f += s_scan_f[threadIdx.x]; s_run_c += s_scan_c[threadIdx.x]; }
		__syncthreads();
	}
}

int rbk_fine_setup_pool(rb_ctx *ctx, PoolSlot &s)
{
	const int T = ctx->d_samp.n_trans;
	const int nchunk = std::max(1, (s.max_no * ctx->d_model.nr_classes + FS_CHUNK - 1) / FS_CHUNK);
	RB_CHECK(ctx->wc_buf[7].ensure((size_t) 2 * s.P * nchunk * 4));
	int *part_so = ctx->wc_buf[7].as<int>(), *part_pair = part_so + (size_t) s.P * nchunk;
	dim3 grid(nchunk, s.P);
	k_fine_count<<<grid, FS_THREADS, 0, ctx->stream>>>(s.meta.as<RbPartMeta>(), s.state.as<RbPartState>(),
		s.Mweight.as<float>(), ctx->d_model, T, nchunk, part_so, part_pair);
	RB_LAUNCH_CHECK(ctx);
	k_fine_chunkscan<<<s.P, 32, 0, ctx->stream>>>(s.state.as<RbPartState>(), nchunk, part_so, part_pair);
	RB_LAUNCH_CHECK(ctx);
	k_fine_scan<<<1, 1024, 0, ctx->stream>>>(s.state.as<RbPartState>(), s.P, ctx->d_samp.n_over_rot, ctx->d_samp.n_over_trans,
		(long long) s.cap_fo, (long long) s.cap_fs, s.counters.as<int>());
	RB_LAUNCH_CHECK(ctx);
	k_fine_fill<<<grid, FS_THREADS, 0, ctx->stream>>>(s.meta.as<RbPartMeta>(), s.state.as<RbPartState>(),
		s.Mweight.as<float>(), s.dir_idx.as<int>(), s.psi_idx.as<int>(), ctx->d_model, ctx->d_samp, s.lr,
		s.pair_list.as<int>(), s.fo.as<RbFineOrient>(), s.fs_ihid.as<long long>(), s.counters.as<int>(), nchunk, part_so, part_pair);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// ---- fine-pass weights -----------------------------------------------------------------------
__global__ void __launch_bounds__(WT_THREADS)
k_weights_fine(const RbPartMeta *metas, RbPartState *states, float *fs_w, const long long *fs_ihid,
               const float *pdf_orient, const unsigned char *pdf_orient_zero,
               const float *pdf_offset, const unsigned char *pdf_offset_zero,
               RbModelDev M, int T, int NOR, int NOT, const int *counters)
{
	__shared__ SelSmem sm;
	__shared__ ArgMaxSmem am;
	__shared__ float fred[32];
	if (counters[2]) return;
	const int p = blockIdx.x;
	const RbPartMeta m = metas[p];
	RbPartState *st = states + p;
	const long long n = (long long) st->n_pairs * NOR * NOT;
	if (n == 0) { if (threadIdx.x == 0 && st->status == 0) st->status = RB_ERR_NO_SIGNIFICANT; return; }
	float *w = fs_w + st->fs_base;
	const long long *ih = fs_ihid + st->fs_base;
	const float *po = pdf_orient + m.prior_off; const unsigned char *pz = pdf_orient_zero + m.prior_off;
	// per-class block of translation priors (2D references with their own prior centre, :2399-2400), else one block
	const int Kp = M.prior_classes(), no = m.nd * m.np;
	const float *pt = pdf_offset + (size_t) p * Kp * T; const unsigned char *tz = pdf_offset_zero + (size_t) p * Kp * T;
	const float min_diff2 = __int_as_float(st->fmin_bits);                                  // :1881
	const long long ov = (long long) NOR * NOT;
	float mx = RB_LOWEST;
	for (long long i = threadIdx.x; i < n; i += blockDim.x)
	{
		const long long ihidden = ih[i] / ov;
		const long long io = ihidden / T;
		const int it = (int) (ihidden % T) + (Kp > 1 ? (int) (io / no) * T : 0);
		const float d = w[i];
		float l;
		if (d < min_diff2 || pz[io] || tz[it]) l = RB_LOWEST;                               // helper.cu:66-74
		else l = po[io] + pt[it] + min_diff2 - d;
		w[i] = l;
		mx = fmaxf(mx, l);
	}
	const float wmax = block_max(mx, fred);
	const float add = 50.f - wmax;                                                          // :2440
	float bv = RB_LOWEST; long long bi = 0x7fffffffffffffffLL;
	__syncthreads();
	for (long long i = threadIdx.x; i < n; i += blockDim.x)
	{
		const float a = w[i] + add;
		const float e = (a < -88.f) ? 0.f : expf(a);
		w[i] = e;
		if (e > bv) { bv = e; bi = i; }
	}
	__syncthreads();
	float maxw; long long maxi;
	block_argmax(bv, bi, am, maxw, maxi);
	SelResult sel = block_significance(w, n, M.adaptive_fraction, 0, sm);
	if (threadIdx.x == 0)
	{
		st->fmin_diff2 = min_diff2;
		st->min_diff2_final = (double) min_diff2 + (double) add;                            // :2444
		st->fmax_weight = maxw; st->fmax_sample = maxi;
		st->fsum_weight = sel.sum_f; st->fsig_weight = sel.sig_w;
		if (sel.sum_f == 0.f && st->status == 0) st->status = RB_ERR_SUMWEIGHT_ZERO;        // :2505
	}
}

int rbk_weights_fine_pool(rb_ctx *ctx, PoolSlot &s)
{
	if (ctx->d_model.do_cc)
	{
		k_weights_cc_fine<<<s.P, WT_THREADS, 0, ctx->stream>>>(s.state.as<RbPartState>(), s.fs_w.as<float>(),
			ctx->d_samp.n_over_rot, ctx->d_samp.n_over_trans, s.counters.as<int>());
		RB_LAUNCH_CHECK(ctx);
		return RB_OK;
	}
	k_weights_fine<<<s.P, WT_THREADS, 0, ctx->stream>>>(s.meta.as<RbPartMeta>(), s.state.as<RbPartState>(),
		s.fs_w.as<float>(), s.fs_ihid.as<long long>(), s.pdf_orient.as<float>(), s.pdf_orient_zero.as<unsigned char>(),
		s.pdf_offset.as<float>(), s.pdf_offset_zero.as<unsigned char>(), ctx->d_model,
		ctx->d_samp.n_trans, ctx->d_samp.n_over_rot, ctx->d_samp.n_over_trans, s.counters.as<int>());
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// ---- stage entry: dense conversion of one array ----------------------------------------------
__global__ void __launch_bounds__(WT_THREADS)
k_convert_stage(float *w, long long n, int T, const float *po, const unsigned char *pz,
                const float *pt, const unsigned char *tz, double adaptive_fraction, int maxsig,
                unsigned char *sig, rb_weights_out *out)
{
	__shared__ SelSmem sm;
	__shared__ ArgMaxSmem am;
	__shared__ float fred[32];
	float mn = FLT_MAX;
	for (long long i = threadIdx.x; i < n; i += blockDim.x) { float d = w[i]; if (d > RB_LOWEST) mn = fminf(mn, d); }
	mn = -block_max(-mn, fred);
	__syncthreads();
	DenseOut o = dense_convert(w, n, T, po, pz, pt, tz, mn, adaptive_fraction, maxsig, sm, am, fred);
	__syncthreads();
	const float sigw = (n == 1) ? 0.f : o.sel.sig_w;
	if (sig) for (long long i = threadIdx.x; i < n; i += blockDim.x) sig[i] = w[i] >= sigw;
	if (threadIdx.x == 0)
	{
		out->min_diff2 = mn; out->max_weight = o.max_weight; out->max_index = o.max_index;
		out->sum_weight = o.sel.sum_f; out->significant_weight = o.sel.sig_w;
		out->nr_significant = (int) (o.sel.n_nonzero - o.sel.thr_idx); out->n_nonzero = (int) o.sel.n_nonzero;
	}
}

int rbk_convert_weights_stage(rb_ctx *ctx, float *d_w, long long n_orient, int n_trans,
                              const float *d_pdf_o, const unsigned char *d_pdf_oz,
                              const float *d_pdf_t, const unsigned char *d_pdf_tz,
                              double adaptive_fraction, int maxsig, int /*filter_zero*/,
                              unsigned char *d_sig, rb_weights_out *d_out)
{
	k_convert_stage<<<1, WT_THREADS, 0, ctx->stream>>>(d_w, n_orient * n_trans, n_trans, d_pdf_o, d_pdf_oz, d_pdf_t, d_pdf_tz,
		adaptive_fraction, maxsig, d_sig, d_out);
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}
