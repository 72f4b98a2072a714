/*
 * TEST INFRASTRUCTURE ONLY — see estep_driver.h.  Nothing under relion_b200/ may use this.
 *
 * Restatement of the host orchestration around the E-step kernels.  Citations are to
 * /root/reference/src/acc/acc_ml_optimiser_impl.h unless another file is named.
 */
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <limits>
#include <algorithm>
#include <omp.h>

#include "estep_driver.h"

namespace {

const float LOWEST = std::numeric_limits<float>::lowest();

inline int iround(double x) { return (int) (x > 0 ? floor(x + 0.5) : -floor(-x + 0.5)); } // ROUND (src/macros.h)

// Mresol_fine / Mresol_coarse (src/ml_optimiser.cpp:5784-5811)
void make_mresol(int n, std::vector<int> &M)
{
	int xs = n / 2 + 1;
	M.assign((size_t) n * xs, -1);
	for (int iy = 0; iy < n; iy++)
	{
		int ip = iy < xs ? iy : iy - n;
		for (int jp = 0; jp < xs; jp++)
		{
			int ires = iround(sqrt((double) (ip * ip + jp * jp)));
			if (ires < xs && !(jp == 0 && ip < 0)) M[(size_t) iy * xs + jp] = ires;
		}
	}
}

// windowFourierTransform, shrinking branch (src/fftw.h:850-856)
template <typename T>
void window_ft(const T *in, int nin, T *out, int nout, int comps)
{
	int xin = nin / 2 + 1, xout = nout / 2 + 1;
	for (int i = 0; i < nout; i++)
	{
		int ip = i < xout ? i : i - nout;
		int iin = ip < 0 ? ip + nin : ip;
		memcpy(out + (size_t) i * xout * comps, in + (size_t) iin * xin * comps, sizeof(T) * xout * comps);
	}
}

struct Shared {
	const ok_kernel_table *K;
	const rb_model *m;
	const rb_sampling *s;
	const ok_projector *refs;
	ok_backprojector *bps;
	int nc, nf, Npc, Npf, nshell;
	std::vector<int> Mres_c, Mres_f;
	std::vector<float> ctx, cty;         // coarse trans (radians/pixel)
	std::vector<float> ftx, fty;         // fine trans
	std::vector<float> coarse_eulers;    // [n_dir*n_psi*9]
	int NOR, NOT;
};

struct Sig { float sum_weight, significant_weight; int64_t thresholdIdx, n_filtered; };

// filterGreaterZeroOnHost + sortOnHost + scanOnHost + findThresholdIdxInCumulativeSum
// (src/acc/utilities.h:301-398, src/acc/acc_helper_functions.h:191-240; callers :2240-2308, :2497-2523)
Sig significance(const float *w, int64_t n, double adaptive_fraction, int maxsig, bool filter_zero, bool exact)
{
	std::vector<float> sorted;
	sorted.reserve(n);
	if (filter_zero) { for (int64_t i = 0; i < n; i++) if (w[i] > 0.f) sorted.push_back(w[i]); }
	else sorted.assign(w, w + n);
	std::sort(sorted.begin(), sorted.end());
	Sig r; r.n_filtered = (int64_t) sorted.size(); r.thresholdIdx = 0; r.sum_weight = 0; r.significant_weight = 0;
	if (sorted.empty()) return r;
	int64_t sz = r.n_filtered;
	int64_t idx = 0;
	if (!exact)
	{
		std::vector<float> cum(sz);
		float sum = 0.f;
		for (int64_t i = 0; i < sz; i++) { sum += sorted[i]; cum[i] = sum; }
		r.sum_weight = cum[sz - 1];
		float thr = (float) ((1 - adaptive_fraction) * (double) r.sum_weight);
		for (int64_t i = 0; i < sz - 1; i++) if (cum[i] <= thr && thr < cum[i + 1]) idx = i + 1;
	}
	else
	{
		std::vector<double> cum(sz);
		double sum = 0.;
		for (int64_t i = 0; i < sz; i++) { sum += (double) sorted[i]; cum[i] = sum; }
		r.sum_weight = (float) cum[sz - 1];
		float thr = (float) ((1 - adaptive_fraction) * (double) r.sum_weight);
		for (int64_t i = 0; i < sz - 1; i++) if (cum[i] <= (double) thr && (double) thr < cum[i + 1]) idx = i + 1;
	}
	r.thresholdIdx = idx;
	if (filter_zero && maxsig > 0 && sz - idx > maxsig) r.thresholdIdx = sz - maxsig;   // :2301-2306
	r.significant_weight = sorted[r.thresholdIdx];
	return r;
}

int run_particle(const Shared &S, const rb_particles *pool, int p, rb_pool_out *out, unsigned flags,
                 bool exact, ok_debug *dbg,
                 std::vector<double> &acc_pdf_direction, std::vector<double> &acc_pdf_class)
{
	const rb_model *m = S.m;
	const rb_sampling *s = S.s;
	const ok_kernel_table *K = S.K;
	const int Kc = m->nr_classes;
	const int T = s->n_trans;
	const int NOR = S.NOR, NOT = S.NOT;
	const bool noprior = (pool->dir_idx == NULL);
	const bool dump = dbg && dbg->particle == p;

	// ---- per-particle orientation lists (sp.nr_dir / sp.nr_psi, :3822-3823) ----
	int nd, np;
	std::vector<int> dirs, psis;
	std::vector<double> dprior, pprior;
	if (noprior)
	{
		nd = s->n_dir; np = s->n_psi;
		dirs.resize(nd); psis.resize(np);
		for (int i = 0; i < nd; i++) dirs[i] = i;
		for (int i = 0; i < np; i++) psis[i] = i;
	}
	else
	{
		nd = pool->dir_off[p + 1] - pool->dir_off[p];
		np = pool->psi_off[p + 1] - pool->psi_off[p];
		dirs.assign(pool->dir_idx + pool->dir_off[p], pool->dir_idx + pool->dir_off[p + 1]);
		psis.assign(pool->psi_idx + pool->psi_off[p], pool->psi_idx + pool->psi_off[p + 1]);
		dprior.assign(pool->dir_prior + pool->dir_off[p], pool->dir_prior + pool->dir_off[p + 1]);
		pprior.assign(pool->psi_prior + pool->psi_off[p], pool->psi_prior + pool->psi_off[p + 1]);
	}
	const int64_t nOrient = (int64_t) nd * np;
	const int64_t nCoarse = (int64_t) Kc * nOrient * T;

	auto pdf_of = [&](int k, int idl, int ipl) -> double {
		if (m->do_skip_rotate) return m->pdf_class[k];                          // acc_projector_plan_impl.h:193-195, acc_ml_optimiser_impl.h:1966
		if (noprior) return m->pdf_direction[(size_t) k * s->n_dir + idl];      // acc_projector_plan_impl.h:196-204
		return dprior[idl] * pprior[ipl];
	};

	const int group = pool->group_id[p], og = pool->optics_group[p];
	const double *sigma2 = m->sigma2_noise + (size_t) og * S.nshell;
	const float scale_correction = m->do_scale_correction ? (float) m->scale_correction[group] : 1.f;   // :1244
	const double highres_Xi2 = pool->highres_Xi2[p];
	const float *Fimg_full = pool->Fimg + (size_t) p * S.Npf * 2;
	const float *Fnomask_full = pool->Fimg_nomask + (size_t) p * S.Npf * 2;
	// rb_particles.pre_shift: the particle's own translation (--skip_align), applied to both transforms once with the phase ramp of a
	// sampled shift (-2 pi shift / n_full, :1239-1241)
	std::vector<float> shifted_img, shifted_nomask;
	if (pool->pre_shift && (pool->pre_shift[2 * p] != 0. || pool->pre_shift[2 * p + 1] != 0.))
	{
		const double dx = pool->pre_shift[2 * p], dy = pool->pre_shift[2 * p + 1];
		const int nf = S.nf, xs = nf / 2 + 1;
		shifted_img.assign(Fimg_full, Fimg_full + (size_t) S.Npf * 2);
		shifted_nomask.assign(Fnomask_full, Fnomask_full + (size_t) S.Npf * 2);
		for (int iy = 0; iy < nf; iy++)
			for (int x = 0; x < xs; x++)
			{
				const int y = iy < xs ? iy : iy - nf;
				const double a = -2. * M_PI * ((double) x * dx + (double) y * dy) / (double) m->ori_size;
				const float c = (float) cos(a), s = (float) sin(a);
				const size_t i = (size_t) iy * xs + x;
				float re = shifted_img[2 * i], im = shifted_img[2 * i + 1];
				shifted_img[2 * i] = c * re - s * im; shifted_img[2 * i + 1] = c * im + s * re;
				re = shifted_nomask[2 * i]; im = shifted_nomask[2 * i + 1];
				shifted_nomask[2 * i] = c * re - s * im; shifted_nomask[2 * i + 1] = c * im + s * re;
			}
		Fimg_full = shifted_img.data(); Fnomask_full = shifted_nomask.data();
	}
	const float *Fctf_full = pool->Fctf ? pool->Fctf + (size_t) p * S.Npf : NULL;

	// ---- image-side arrays for one window size (precalculateShiftedImagesCtfsAndInvSigma2s,
	//      src/ml_optimiser.cpp:6826-6879; pixel correction :1251-1268; buildCorrImage
	//      acc_helper_functions_impl.h:164-196) ----
	struct Win { std::vector<float> re, im, corr, ctf, minvs2; };
	const bool do_cc = m->do_cc != 0;     // (iter == 1 && do_firstiter_cc) || do_always_cc  (:1164)
	auto prep = [&](int n, const std::vector<int> &Mres, Win &w) {
		int Np = n * (n / 2 + 1);
		std::vector<float> F(2 * (size_t) Np), C(Np, 1.f);
		if (n == S.nf) { memcpy(F.data(), Fimg_full, sizeof(float) * 2 * Np); if (Fctf_full) memcpy(C.data(), Fctf_full, sizeof(float) * Np); }
		else { window_ft(Fimg_full, S.nf, F.data(), n, 2); if (Fctf_full) window_ft(Fctf_full, S.nf, C.data(), n, 1); }
		w.re.resize(Np); w.im.resize(Np); w.corr.resize(Np); w.ctf = C; w.minvs2.assign(Np, 0.f);
		// exp_local_sqrtXi2: power of the windowed (masked) transform, all its pixels (src/ml_optimiser.cpp:6846-6856)
		double sqrtXi2 = 0.;
		if (do_cc)
		{
			double sumxi2 = 0.;
			for (int i = 0; i < Np; i++) sumxi2 += (double) F[2 * i] * (double) F[2 * i] + (double) F[2 * i + 1] * (double) F[2 * i + 1];
			sqrtXi2 = sqrt(sumxi2);
		}
		for (int i = 0; i < Np; i++)
		{
			int ires = Mres[i];
			double mi = 0.;
			if (ires > 0 && ires < S.nshell) mi = 1. / (m->sigma2_fudge * sigma2[ires]);
			w.minvs2[i] = (float) mi;
			float pixel_correction = (float) (1.0 / scale_correction);
			if (m->do_ctf_correction && fabs((double) C[i]) > 1e-8 && m->refs_are_ctf_corrected)
				pixel_correction = (float) ((double) pixel_correction / (double) C[i]);
			w.re[i] = (float) ((double) F[2 * i] * pixel_correction);
			w.im[i] = (float) ((double) F[2 * i + 1] * pixel_correction);
			float c = (float) mi;
			if (do_cc) c = (float) (1. / (sqrtXi2 * sqrtXi2));                                 // buildCorrImage :172-174
			if (m->do_ctf_correction && m->refs_are_ctf_corrected) c = (float) (c * ((double) C[i] * (double) C[i]));
			if (m->do_scale_correction) { float ms = (float) m->scale_correction[group]; c *= ms * ms; }
			w.corr[i] = c;
		}
	};

	// =========================== pass 0: coarse (getAllSquaredDifferencesCoarse :1015-1413) ===========================
	Win wc; prep(S.nc, S.Mres_c, wc);
	std::vector<float> Mweight(nCoarse, LOWEST);                                            // :3849
	float min_diff2 = std::numeric_limits<float>::max();
	{
		std::vector<float> eul; std::vector<int64_t> ioc;
		for (int k = 0; k < Kc; k++)
		{
			if (!(m->pdf_class[k] > 0.)) continue;                                          // :1069
			eul.clear(); ioc.clear();
			for (int idl = 0; idl < nd; idl++)
				for (int ipl = 0; ipl < np; ipl++)
					if (pdf_of(k, idl, ipl) > 0.)
					{
						const float *e = &S.coarse_eulers[((size_t) dirs[idl] * s->n_psi + psis[ipl]) * 9];
						eul.insert(eul.end(), e, e + 9);
						ioc.push_back((int64_t) k * nOrient + (int64_t) idl * np + ipl);
					}
			size_t O = ioc.size();
			if (!O) continue;
			std::vector<float> allW(O * T, 0.f);
			const float xi = (float) (highres_Xi2 / 2.);                                    // :1290-1296
			if (!do_cc) for (auto &v : allW) v += xi;                                       // :1287-1297: no Xi2 term with CC
			if (do_cc)
				K->diff2_cc_coarse(&S.refs[k], S.nc / 2 + 1, S.nc, eul.data(), O, S.ctx.data(), S.cty.data(), T,
				                   wc.re.data(), wc.im.data(), wc.corr.data(), allW.data());
			else
			K->diff2_coarse(&S.refs[k], S.nc / 2 + 1, S.nc, eul.data(), O, S.ctx.data(), S.cty.data(), T,
			                wc.re.data(), wc.im.data(), wc.corr.data(), allW.data());
			for (size_t o = 0; o < O; o++)                                                   // mapAllWeightsToMweights (helper.cu:782-796)
				for (int t = 0; t < T; t++)
				{
					Mweight[ioc[o] * T + t] = allW[o * T + t];
					min_diff2 = std::min(min_diff2, allW[o * T + t]);                       // :1407
				}
		}
	}
	if (dump && dbg->coarse_diff2) memcpy(dbg->coarse_diff2, Mweight.data(), sizeof(float) * nCoarse);

	// ---- priors (convertAllSquaredDifferencesToWeights :1915-2175) ----
	std::vector<float> pdf_orientation((size_t) Kc * nOrient), pdf_offset((size_t) Kc * T);
	std::vector<unsigned char> pdf_orientation_zeros((size_t) Kc * nOrient), pdf_offset_zeros((size_t) Kc * T);
	for (int k = 0; k < Kc; k++)
		for (int idl = 0; idl < nd; idl++)
			for (int ipl = 0; ipl < np; ipl++)
			{
				double pdf = pdf_of(k, idl, ipl);
				size_t i = (size_t) k * nOrient + (size_t) idl * np + ipl;
				pdf_orientation_zeros[i] = (pdf == 0);                                       // initOrientations (utilities_impl.h:656-668)
				pdf_orientation[i] = (pdf == 0) ? 0.f : (float) log(pdf);
			}
	double s2off = (m->offset_range > 0.) ? (m->offset_range * m->offset_range) / 9. : m->sigma2_offset;   // :1916-1926
	// rb_particles.pre_shift is part of every sampled offset (with --skip_align it IS the sampled translation, :2138-2139)
	const double oldx = pool->old_offset[2 * p] + (pool->pre_shift ? pool->pre_shift[2 * p] : 0.), oldy = pool->old_offset[2 * p + 1] + (pool->pre_shift ? pool->pre_shift[2 * p + 1] : 0.);
	const double prx = pool->prior_offset[2 * p], pry = pool->prior_offset[2 * p + 1];
	for (int k = 0; k < Kc; k++)
	{
		// :2094-2111: 2D references carry their own prior centre (mymodel.prior_offset_class), else op.prior
		const double prx = m->prior_offset_class ? m->prior_offset_class[2 * k] : pool->prior_offset[2 * p];
		const double pry = m->prior_offset_class ? m->prior_offset_class[2 * k + 1] : pool->prior_offset[2 * p + 1];
		for (int t = 0; t < T; t++)
		{
			// NB: sampling.translations_x are in Angstrom in RELION >= 3.1 and converted per pixel size; the
			// tables handed to us are already in pixels (getTranslationsInPixel), so offsets are in pixels and
			// tdiff2 is converted with pixel_size^2 exactly as :2138-2153
			double offx = oldx + s->trans_x[t], offy = oldy + s->trans_y[t];
			double tdiff2 = (offx - prx) * (offx - prx) / (-2. * s2off) + (offy - pry) * (offy - pry) / (-2. * s2off);
			tdiff2 *= m->pixel_size * m->pixel_size;
			double pdf; bool z;
			if (s2off < 0.0001) { z = tdiff2 > 0.; pdf = z ? 0. : 1.; }
			else { z = false; pdf = tdiff2; }
			pdf_offset_zeros[(size_t) k * T + t] = z;
			pdf_offset[(size_t) k * T + t] = (float) pdf;
		}
	}

	// ---- weights, pass 0 (:2180-2350) ----
	// NB the reference passes ONE pdf_offset block (class 0's) to the coarse kernel for all classes
	// (kernel indexes itrans only, :2187-2196); for 3D references all classes share the same prior, so identical;
	// for 2D references with their own prior centres the coarse weights of every class use the first class' centre.
	std::vector<unsigned char> significant(nCoarse, 0);
	rb_particle_out &po = out->particles[p];
	memset(&po, 0, sizeof(po));
	po.min_diff2_coarse = min_diff2;
	if (do_cc)
	{
		// :2012-2071: the smallest diff2 gets weight one, everything else zero; significant_weight = 0.999, NR_SIGN = 1.
		// (The reference takes the arg-min over the whole Mweight array, lowest()-initialised entries included; the
		// supported case is the one it is used in, a global search where every entry has been computed.)
		int64_t amin = -1;
		for (int64_t i = 0; i < nCoarse; i++)
			if (Mweight[i] > LOWEST && (amin < 0 || Mweight[i] < Mweight[amin])) amin = i;
		if (amin < 0) return RB_ERR_NO_SIGNIFICANT;
		for (int64_t i = 0; i < nCoarse; i++) Mweight[i] = 0.f;
		Mweight[amin] = 1.f;
		significant[amin] = 1;
		po.nr_significant_coarse = 1;
		po.sum_weight_coarse = 1.f;                                                        // :1986
		po.significant_weight_coarse = 0.999f;
	}
	else {
	K->weights_exponent_coarse(pdf_orientation.data(), pdf_orientation_zeros.data(), pdf_offset.data(),
	                           pdf_offset_zeros.data(), Mweight.data(), min_diff2, Kc * nOrient, T, nCoarse);
	float wmax = LOWEST;
	for (int64_t i = 0; i < nCoarse; i++) wmax = std::max(wmax, Mweight[i]);
	K->exponentiate(Mweight.data(), 50.f - wmax, nCoarse);                                   // :2207
	if (nCoarse > 1)
	{
		Sig sg = significance(Mweight.data(), nCoarse, m->adaptive_fraction, m->maximum_significants, true, exact);
		if (sg.n_filtered == 0) return RB_ERR_NO_SIGNIFICANT;                               // ERRFILTEREDZERO :2242
		int64_t nsig = sg.n_filtered - sg.thresholdIdx;
		if (nsig == 0) return RB_ERR_NO_SIGNIFICANT;                                        // ERRNOSIGNIFS :2282
		po.nr_significant_coarse = (int) nsig;
		po.sum_weight_coarse = sg.sum_weight;
		po.significant_weight_coarse = sg.significant_weight;
		for (int64_t i = 0; i < nCoarse; i++) significant[i] = Mweight[i] >= sg.significant_weight;   // arrayOverThreshold
	}
	else { significant[0] = 1; po.nr_significant_coarse = 1; }                              // :2347-2350
	}
	if (dump && dbg->coarse_weights) memcpy(dbg->coarse_weights, Mweight.data(), sizeof(float) * nCoarse);
	if (dump && dbg->coarse_significant) memcpy(dbg->coarse_significant, significant.data(), nCoarse);

	// =========================== pass 1: fine ===========================
	// generateProjectionSetupFine (acc_helper_functions_impl.h:265-314) + makeJobsForDiff2Fine (:26-102)
	const int Tf = T * NOT;
	struct ClassFine {
		std::vector<double> rot, tilt, psi;
		std::vector<int64_t> iorientclass; std::vector<int> iover_rot;
		std::vector<unsigned long> rot_id, rot_idx, trans_idx, job_idx, job_num;
		std::vector<int64_t> ihidden_over;
		size_t firstPos, weightNum;
		std::vector<float> eulers;
	};
	std::vector<ClassFine> cf(Kc);
	size_t newDataSize = 0;
	const int chunk = 4; // D2F_CHUNK_DATA3D for 3D refs too (:1651-1657)
	for (int k = 0; k < Kc; k++)
	{
		ClassFine &c = cf[k];
		c.firstPos = newDataSize; c.weightNum = 0;
		if (!(m->pdf_class[k] > 0.)) continue;
		for (int idl = 0; idl < nd; idl++)
			for (int ipl = 0; ipl < np; ipl++)
			{
				int64_t ioc = (int64_t) k * nOrient + (int64_t) idl * np + ipl;
				bool any = false;
				for (int t = 0; t < T; t++) if (significant[ioc * T + t]) { any = true; break; }
				if (!any) continue;
				size_t g = ((size_t) dirs[idl] * s->n_psi + psis[ipl]) * NOR;
				for (int io = 0; io < NOR; io++)
				{
					if (s->over_rot) { c.rot.push_back(s->over_rot[g + io]); c.tilt.push_back(s->over_tilt[g + io]); c.psi.push_back(s->over_psi[g + io]); }
					else { c.rot.push_back(s->rot[dirs[idl]]); c.tilt.push_back(s->tilt[dirs[idl]]); c.psi.push_back(s->psi[psis[ipl]]); }
					c.iorientclass.push_back(ioc); c.iover_rot.push_back(io);
				}
			}
		size_t On = c.rot.size();
		if (!On) continue;
		// job lists
		unsigned long w = 0;
		c.job_idx.push_back(0); c.job_num.push_back(0);
		size_t kk = 0;
		for (size_t i = 0; i < On; i++)
		{
			c.job_num[kk] = 0;
			long tk = 0;
			for (int j = 0; j < Tf; j++)
			{
				int itrans = j / NOT, iover_trans = j % NOT;
				int64_t ihidden = c.iorientclass[i] * T + itrans;
				if (significant[ihidden])
				{
					c.rot_id.push_back(c.iorientclass[i] % nOrient);
					c.rot_idx.push_back(i);
					c.trans_idx.push_back(j);
					c.ihidden_over.push_back((ihidden * NOR + c.iover_rot[i]) * NOT + iover_trans);
					if (tk >= chunk)
					{
						tk = 0; kk++;
						if (c.job_idx.size() <= kk) { c.job_idx.push_back(0); c.job_num.push_back(0); }
						c.job_idx[kk] = w; c.job_num[kk] = 0;
					}
					tk++; c.job_num[kk]++; w++;
				}
				else if (tk != 0)
				{
					tk = 0; kk++;
					if (c.job_idx.size() <= kk) { c.job_idx.push_back(0); c.job_num.push_back(0); }
					c.job_idx[kk] = w; c.job_num[kk] = 0;
				}
			}
			if (tk > 0)
			{
				kk++;
				if (c.job_idx.size() <= kk) { c.job_idx.push_back(0); c.job_num.push_back(0); }
				c.job_idx[kk] = w; c.job_num[kk] = 0;
			}
		}
		if (c.job_num[kk] != 0) kk += 1;
		c.job_idx.resize(kk); c.job_num.resize(kk);
		c.weightNum = w;
		newDataSize += w;
		// generateEulerMatrices(inverse=true) in double, cast to XFLOAT (acc_helper_functions_impl.h:198-262)
		c.eulers.resize(9 * On);
		for (size_t i = 0; i < On; i++)
		{
			double a = c.rot[i] * M_PI / 180., b = c.tilt[i] * M_PI / 180., g = c.psi[i] * M_PI / 180.;
			double ca = cos(a), sa = sin(a), cb = cos(b), sb = sin(b), cg = cos(g), sg = sin(g);
			double cc = cb * ca, cs = cb * sa, sc = sb * ca, ss = sb * sa;
			double A[9] = { cg * cc - sg * sa, cg * cs + sg * ca, -cg * sb,
			               -sg * cc - cg * sa, -sg * cs + cg * ca, sg * sb,
			                sc, ss, cb };
			float *e = &c.eulers[9 * i];
			if (pool->mat_left || pool->mat_right)
			{   // A = L * A; A = A * R; A = A.inv() (:248-255)
				double B[9];
				if (pool->mat_left)
				{
					const double *L = pool->mat_left;
					for (int r = 0; r < 3; r++)
						for (int q = 0; q < 3; q++) B[r * 3 + q] = L[r * 3] * A[q] + L[r * 3 + 1] * A[3 + q] + L[r * 3 + 2] * A[6 + q];
					for (int q = 0; q < 9; q++) A[q] = B[q];
				}
				if (pool->mat_right)
				{
					const double *R = pool->mat_right;
					for (int r = 0; r < 3; r++)
						for (int q = 0; q < 3; q++) B[r * 3 + q] = A[r * 3] * R[q] + A[r * 3 + 1] * R[3 + q] + A[r * 3 + 2] * R[6 + q];
					for (int q = 0; q < 9; q++) A[q] = B[q];
				}
				const double det = A[0] * (A[4] * A[8] - A[7] * A[5]) - A[1] * (A[3] * A[8] - A[6] * A[5]) + A[2] * (A[3] * A[7] - A[6] * A[4]);
				e[0] = (float) ((A[4] * A[8] - A[7] * A[5]) / det); e[1] = (float) ((A[7] * A[2] - A[1] * A[8]) / det);
				e[2] = (float) ((A[1] * A[5] - A[4] * A[2]) / det); e[3] = (float) ((A[5] * A[6] - A[8] * A[3]) / det);
				e[4] = (float) ((A[8] * A[0] - A[2] * A[6]) / det); e[5] = (float) ((A[2] * A[3] - A[5] * A[0]) / det);
				e[6] = (float) ((A[3] * A[7] - A[6] * A[4]) / det); e[7] = (float) ((A[6] * A[1] - A[0] * A[7]) / det);
				e[8] = (float) ((A[0] * A[4] - A[3] * A[1]) / det);
				continue;
			}
			// inverse of a rotation = transpose
			e[0] = (float) A[0]; e[1] = (float) A[3]; e[2] = (float) A[6];
			e[3] = (float) A[1]; e[4] = (float) A[4]; e[5] = (float) A[7];
			e[6] = (float) A[2]; e[7] = (float) A[5]; e[8] = (float) A[8];
		}
	}
	po.n_fine_samples = (int) newDataSize;
	{ size_t no = 0; for (auto &c : cf) no += c.rot.size(); po.n_fine_orient = (int) no; }

	// getAllSquaredDifferencesFine (:1419-1889)
	Win wf; prep(S.nf, S.Mres_f, wf);
	std::vector<float> fw(newDataSize, 0.f);                                                 // :1438
	for (int k = 0; k < Kc; k++)
	{
		ClassFine &c = cf[k];
		if (!c.weightNum) continue;
		if (do_cc)
			K->diff2_cc_fine(&S.refs[k], S.nf / 2 + 1, S.nf, c.eulers.data(), S.ftx.data(), S.fty.data(),
			                 wf.re.data(), wf.im.data(), wf.corr.data(),
			                 c.rot.size(), Tf, c.job_idx.size(),
			                 c.rot_idx.data(), c.trans_idx.data(), c.job_idx.data(), c.job_num.data(),
			                 fw.data() + c.firstPos);
		else
		K->diff2_fine(&S.refs[k], S.nf / 2 + 1, S.nf, c.eulers.data(), S.ftx.data(), S.fty.data(),
		              wf.re.data(), wf.im.data(), wf.corr.data(), (float) (highres_Xi2 / 2.),
		              c.rot.size(), Tf, c.job_idx.size(),
		              c.rot_idx.data(), c.trans_idx.data(), c.job_idx.data(), c.job_num.data(),
		              fw.data() + c.firstPos);
	}
	if (newDataSize == 0) return RB_ERR_NO_SIGNIFICANT;
	float min_diff2_f = *std::min_element(fw.begin(), fw.end());                             // :1881
	if (dump)
	{
		dbg->fine_count = (int64_t) newDataSize;
		if ((int64_t) newDataSize <= dbg->fine_capacity)
		{
			if (dbg->fine_diff2) memcpy(dbg->fine_diff2, fw.data(), sizeof(float) * newDataSize);
			if (dbg->fine_ihidden_over)
				for (int k = 0; k < Kc; k++)
					for (size_t i = 0; i < cf[k].weightNum; i++) dbg->fine_ihidden_over[cf[k].firstPos + i] = cf[k].ihidden_over[i];
		}
	}

	// weights, pass 1 (:2354-2535)
	float wmaxf = LOWEST;
	double min_diff2_final;
	Sig sf;
	if (do_cc)
	{
		// :2012-2071 with exp_ipass == 1: weight one for the smallest diff2; sum_weight stays at its initial value of one
		// (:1986), significant_weight = 0.999; op.min_diff2 keeps the minimum of the CC values (:1881)
		size_t amin = 0;
		for (size_t i = 1; i < newDataSize; i++) if (fw[i] < fw[amin]) amin = i;
		for (size_t i = 0; i < newDataSize; i++) fw[i] = 0.f;
		fw[amin] = 1.f;
		min_diff2_final = (double) min_diff2_f;
		sf.sum_weight = 1.f; sf.significant_weight = 0.999f; sf.thresholdIdx = 0; sf.n_filtered = 1;
	}
	else {
	for (int k = 0; k < Kc; k++)
	{
		ClassFine &c = cf[k];
		if (!(m->pdf_class[k] > 0.) || !c.weightNum) continue;
		K->exponentiate_weights_fine(&pdf_orientation[(size_t) k * nOrient], &pdf_orientation_zeros[(size_t) k * nOrient],
		                             &pdf_offset[(size_t) k * T], &pdf_offset_zeros[(size_t) k * T],
		                             fw.data() + c.firstPos, min_diff2_f, NOR, NOT,
		                             c.rot_id.data(), c.trans_idx.data(), c.job_idx.data(), c.job_num.data(),
		                             (long) c.job_idx.size());
		for (size_t i = 0; i < c.weightNum; i++) wmaxf = std::max(wmaxf, fw[c.firstPos + i]);
	}
	for (int k = 0; k < Kc; k++)
		if ((m->pdf_class[k] > 0.) && cf[k].weightNum)
			K->exponentiate(fw.data() + cf[k].firstPos, 50.f - wmaxf, cf[k].weightNum);       // :2440
	// op.min_diff2 is RFLOAT; the float kernel argument was (XFLOAT)op.min_diff2 (:2411, :2444)
	min_diff2_final = (double) min_diff2_f + (double) (50.f - wmaxf);
	sf = significance(fw.data(), (int64_t) newDataSize, m->adaptive_fraction, 0, false, exact);
	if (sf.sum_weight == 0.f) return RB_ERR_SUMWEIGHT_ZERO;                                  // :2505
	}
	int64_t amax = 0;
	for (size_t i = 1; i < newDataSize; i++) if (fw[i] > fw[amax]) amax = (int64_t) i;
	if (dump && dbg->fine_weights && (int64_t) newDataSize <= dbg->fine_capacity)
		memcpy(dbg->fine_weights, fw.data(), sizeof(float) * newDataSize);
	int64_t fineIdx = 0;
	for (int k = 0; k < Kc; k++)
		if (amax >= (int64_t) cf[k].firstPos && amax < (int64_t) (cf[k].firstPos + cf[k].weightNum))
			fineIdx = cf[k].ihidden_over[amax - cf[k].firstPos];
	po.best_ihidden_over = fineIdx;
	{   // fineIndexToFineIndices (src/acc/acc_ml_optimiser.h:78-95)
		int64_t t_idx = fineIdx, ov = (int64_t) NOR * NOT;
		po.best_class = (int) (t_idx / (nOrient * T * ov)); t_idx -= (int64_t) po.best_class * nOrient * T * ov;
		po.best_idir = (int) (t_idx / ((int64_t) np * T * ov)); t_idx -= (int64_t) po.best_idir * np * T * ov;
		po.best_ipsi = (int) (t_idx / ((int64_t) T * ov)); t_idx -= (int64_t) po.best_ipsi * T * ov;
		po.best_itrans = (int) (t_idx / ov); t_idx -= (int64_t) po.best_itrans * ov;
		po.best_iover_rot = (int) (t_idx / NOT); t_idx -= (int64_t) po.best_iover_rot * NOT;
		po.best_iover_trans = (int) t_idx;
	}
	po.min_diff2 = (float) min_diff2_final;
	po.max_weight = fw[amax];
	po.sum_weight = sf.sum_weight;
	po.significant_weight = sf.significant_weight;
	po.pmax = po.max_weight / po.sum_weight;                                                 // :2924
	if (po.pmax > 1.f) return RB_ERR_PMAX;
	po.dLL_nolog = log((double) po.sum_weight) - min_diff2_final;                            // :3574

	// =========================== storeWeightedSums (:2553-3667) ===========================
	// collect2jobs (:2631-2855)
	{
		std::vector<float> oox(Tf), ooy(Tf), oo2(Tf);
		for (int it = 0; it < Tf; it++)
		{
			double xs = oldx + s->over_trans_x[it], ys = oldy + s->over_trans_y[it];         // :2703-2704
			oox[it] = (float) xs; ooy[it] = (float) ys;
		}
		for (int k = 0; k < Kc; k++)
		{
			ClassFine &c = cf[k];
			if ((m->pdf_class[k] == 0.) || c.rot.empty()) continue;
			const double cprx = m->prior_offset_class ? m->prior_offset_class[2 * k] : prx;   // :2673-2686
			const double cpry = m->prior_offset_class ? m->prior_offset_class[2 * k + 1] : pry;
			for (int it = 0; it < Tf; it++)
			{
				double dx = cprx - (oldx + s->over_trans_x[it]), dy = cpry - (oldy + s->over_trans_y[it]);
				oo2[it] = (float) (dx * dx + dy * dy);                                       // :2736
			}
			// makeJobsForCollect (acc_helper_functions_impl.h:104-139): one job per run of equal rot_idx
			std::vector<unsigned long> jo, je;
			if (c.weightNum)
			{
				jo.push_back(0); je.push_back(1);
				unsigned long crot = c.rot_idx[0];
				for (size_t n = 1; n < c.weightNum; n++)
				{
					if (c.rot_idx[n] == crot) je.back()++;
					else { jo.push_back(n); je.push_back(1); crot = c.rot_idx[n]; }
				}
			}
			int nb = (int) jo.size();
			std::vector<float> pw(nb), ppx(nb), ppy(nb), ps2(nb);
			K->collect2jobs(nb, oox.data(), ooy.data(), oo2.data(), fw.data() + c.firstPos,
			                po.significant_weight, po.sum_weight, T, NOT, NOR, (unsigned long) NOR * NOT,
			                pw.data(), ppx.data(), ppy.data(), ps2.data(),
			                c.rot_idx.data(), c.trans_idx.data(), jo.data(), je.data());
			for (int n = 0; n < nb; n++)                                                     // :2830-2852
			{
				long iorient = (long) c.rot_id[jo[n]];
				long idir = iorient / np;
				long mydir = noprior ? idir : dirs[idir];
				acc_pdf_direction[(size_t) k * s->n_dir + mydir] += pw[n];
				po.sumw += pw[n];
				acc_pdf_class[k] += pw[n];
				po.wsum_sigma2_offset += m->pixel_size * m->pixel_size * ps2[n];
				if (m->prior_offset_class)                                                   // :2847-2851 (ref_dim == 2)
				{
					acc_pdf_class[Kc + 2 * k] += m->pixel_size * ppx[n];
					acc_pdf_class[Kc + 2 * k + 1] += m->pixel_size * ppy[n];
				}
			}
		}
	}

	if (flags & 1u) return 0; // do_skip_maximization

	// wavg + backprojection (:2940-3495)
	{
		float part_scale = 1.f;
		if (m->do_scale_correction)
		{
			part_scale = (float) m->scale_correction[group];
			if (part_scale > 10000.f) return RB_ERR_ARG;                                     // ERRHIGHSCALE :3064
			if (part_scale < 0.001f) part_scale = 0.001f;                                    // :3069-3078
		}
		std::vector<float> ctfs(S.Npf), minvs2(S.Npf), fr(S.Npf), fi(S.Npf), nr(S.Npf), ni(S.Npf);
		for (int i = 0; i < S.Npf; i++)
		{
			ctfs[i] = m->do_ctf_correction ? (float) ((double) Fctf_full[i] * part_scale) : part_scale;   // :3087-3096
			minvs2[i] = m->do_map ? wf.minvs2[i] : 1.f;                                      // :3110-3115
			fr[i] = Fimg_full[2 * i]; fi[i] = Fimg_full[2 * i + 1];
			nr[i] = Fnomask_full[2 * i]; ni[i] = Fnomask_full[2 * i + 1];
		}
		if (m->do_map) minvs2[0] = (float) (1. / (m->sigma2_fudge * sigma2[0]));             // :2586
		std::vector<float> parts(S.Npf, 0.f), AA((size_t) Kc * S.Npf, 0.f), XA((size_t) Kc * S.Npf, 0.f);
		for (int k = 0; k < Kc; k++)
		{
			ClassFine &c = cf[k];
			if ((m->pdf_class[k] == 0.) || c.rot.empty()) continue;
			size_t On = c.rot.size();
			std::vector<float> sw(On * Tf, LOWEST);                                          // :3247-3252
			for (size_t i = 0; i < c.weightNum; i++) sw[c.rot_idx[i] * Tf + c.trans_idx[i]] = fw[c.firstPos + i];
			for (size_t io = 0; io < On; io++)                                               // orientations the kernels below do work for
			{
				bool any = false;
				for (int t = 0; t < Tf; t++) any |= sw[io * Tf + t] >= po.significant_weight;  // wavg.cuh:106, BP.cuh:278
				po.n_bp_orient += any;
			}
			K->wavg(&S.refs[k], S.nf / 2 + 1, S.nf, c.eulers.data(), On, fr.data(), fi.data(),
			        S.ftx.data(), S.fty.data(), sw.data(), ctfs.data(),
			        parts.data(), &AA[(size_t) k * S.Npf], &XA[(size_t) k * S.Npf],
			        Tf, po.sum_weight, po.significant_weight, part_scale);
			// pseudo half-sets of gradient refinement: iproj_offset = (part_id % 2) * nr_classes (acc_ml_optimiser_impl.h:3395-3400)
			ok_backprojector bpk = S.bps[k + (pool->bp_offset ? pool->bp_offset[p] : 0)];
			if (m->do_grad)                                                                  // :3418 -> backproject3D_SGD / backproject2D_SGD
				(bpk.mdlZ == 1 ? K->backproject2d_sgd : K->backproject_sgd)(&bpk, &S.refs[k], S.nf / 2 + 1, S.nf, nr.data(), ni.data(), S.ftx.data(), S.fty.data(),
				                   sw.data(), minvs2.data(), ctfs.data(), Tf, po.significant_weight, po.sum_weight, c.eulers.data(), On);
			else
			// 2D accumulators (2D classification) go through backproject2D, acc_helper_functions_impl.h:505-577
			(bpk.mdlZ == 1 ? K->backproject2d : K->backproject)(&bpk, S.nf / 2 + 1, S.nf, nr.data(), ni.data(), S.ftx.data(), S.fty.data(),
			               sw.data(), minvs2.data(), ctfs.data(), Tf, po.significant_weight, po.sum_weight,
			               c.eulers.data(), On);
			for (int j = 0; j < S.Npf; j++)                                                  // :3467-3482
			{
				int ires = S.Mres_f[j];
				if (ires > -1 && m->do_scale_correction && m->data_vs_prior_class[(size_t) k * S.nshell + ires] > 3.)
				{
					po.wsum_AA += AA[(size_t) k * S.Npf + j];
					po.wsum_XA += XA[(size_t) k * S.Npf + j];
				}
			}
		}
		float *shell = out->wsum_sigma2_noise ? out->wsum_sigma2_noise + (size_t) p * S.nshell : NULL;
		std::vector<double> sh(S.nshell, 0.);
		for (int j = 0; j < S.Npf; j++)                                                      // :3484-3493
		{
			int ires = S.Mres_f[j];
			if (ires > -1) { sh[ires] += (double) parts[j]; po.wsum_norm_correction += (double) parts[j]; }
		}
		if (shell) for (int i = 0; i < S.nshell; i++) shell[i] = (float) sh[i];
		if (dump)
		{
			if (dbg->wdiff2s_parts) memcpy(dbg->wdiff2s_parts, parts.data(), sizeof(float) * S.Npf);
			if (dbg->wdiff2s_AA) memcpy(dbg->wdiff2s_AA, AA.data(), sizeof(float) * Kc * S.Npf);
			if (dbg->wdiff2s_XA) memcpy(dbg->wdiff2s_XA, XA.data(), sizeof(float) * Kc * S.Npf);
		}
	}
	return 0;
}

} // namespace

extern "C" {

int64_t oracle_significance(const float *weights, int64_t n, double adaptive_fraction, int maximum_significants,
                            int filter_zero, int exact, float *sum_weight, float *significant_weight, int64_t *n_filtered)
{
	Sig r = significance(weights, n, adaptive_fraction, maximum_significants, filter_zero != 0, exact != 0);
	if (sum_weight) *sum_weight = r.sum_weight;
	if (significant_weight) *significant_weight = r.significant_weight;
	if (n_filtered) *n_filtered = r.n_filtered;
	return r.thresholdIdx;
}

// BackProjector::backproject2Dto3D (/root/reference/src/backprojector.cpp:55-357), TRILINEAR branch, no Ewald sphere and no
// magnification matrix, in double like the reference (RFLOAT = double), one image after the other on one thread as
// relion_reconstruct runs it (Reconstructor::backprojectOneParticle, src/reconstructor.cpp:328-744).
// f2d: [count][s][s/2+1] complex fp32 (already CTF-multiplied), mweight: [count][s][s/2+1] fp32, ainv: [count][9] fp32
// INVERTED matrices; data_re/data_im/weight: double [Z][Y][X] centred volumes (STARTINGY = -(Y-1)/2).
int oracle_backproject_posed(double *data_re, double *data_im, double *weight, int xdim, int ydim, int zdim,
                             const float *f2d, const float *mweight, const float *ainv, int s, int count,
                             int r_max, double padding_factor)
{
	const int sh = s / 2 + 1;
	const long rr = (long) floor(r_max * padding_factor + 0.5);
	const double max_r2 = (double) (rr * rr);
	const int starty = -((ydim - 1) / 2), startz = -((zdim - 1) / 2);
	for (int img = 0; img < count; img++)
	{
		const float *e = ainv + (size_t) img * 9;
		const double a00 = e[0] * padding_factor, a01 = e[1] * padding_factor, a10 = e[3] * padding_factor, a11 = e[4] * padding_factor,
		             a20 = e[6] * padding_factor, a21 = e[7] * padding_factor;
		const double AtA_xx = a00 * a00 + a10 * a10 + a20 * a20, AtA_xy = a00 * a01 + a10 * a11 + a20 * a21,
		             AtA_yy = a01 * a01 + a11 * a11 + a21 * a21;
		const float *F = f2d + (size_t) img * s * sh * 2, *W = mweight + (size_t) img * s * sh;
		for (int i = 0; i < s; i++)
		{
			int y, first_allowed_x;
			if (i < sh) { y = i; first_allowed_x = 0; } else { y = i - s; first_allowed_x = 1; }
			const double discr = AtA_xy * AtA_xy * y * y - AtA_xx * (AtA_yy * y * y - max_r2);        // :118-128
			if (discr < 0.0) continue;
			const double d = sqrt(discr) / AtA_xx, q = -AtA_xy * y / AtA_xx;
			int first_x = (int) ceil(q - d), last_x = (int) floor(q + d);
			if (first_x < first_allowed_x) first_x = first_allowed_x;
			if (last_x > sh - 1) last_x = sh - 1;
			for (int x = first_x; x <= last_x; x++)
			{
				double vr = F[2 * ((size_t) i * sh + x)], vi = F[2 * ((size_t) i * sh + x) + 1];
				const double w = W[(size_t) i * sh + x];
				if (w <= 0.) continue;
				double xp = a00 * x + a01 * y, yp = a10 * x + a11 * y, zp = a20 * x + a21 * y;
				if (xp * xp + yp * yp + zp * zp > max_r2) continue;
				if (xp < 0) { xp = -xp; yp = -yp; zp = -zp; vi = -vi; }
				int x0 = (int) floor(xp); const double fx = xp - x0;
				int y0 = (int) floor(yp); const double fy = yp - y0; y0 -= starty;
				int z0 = (int) floor(zp); const double fz = zp - z0; z0 -= startz;
				if (x0 < 0 || x0 + 1 >= xdim || y0 < 0 || y0 + 1 >= ydim || z0 < 0 || z0 + 1 >= zdim) continue;   // :213-218
				const double mfx = 1. - fx, mfy = 1. - fy, mfz = 1. - fz;
				const double dd[8] = {mfz * mfy * mfx, mfz * mfy * fx, mfz * fy * mfx, mfz * fy * fx,
				                      fz * mfy * mfx, fz * mfy * fx, fz * fy * mfx, fz * fy * fx};
				for (int c = 0; c < 8; c++)
				{
					const size_t idx = ((size_t) (z0 + (c >> 2)) * ydim + (y0 + ((c >> 1) & 1))) * xdim + x0 + (c & 1);
					data_re[idx] += dd[c] * vr; data_im[idx] += dd[c] * vi; weight[idx] += dd[c] * w;
				}
			}
		}
	}
	return 0;
}

int oracle_estep_pool(const ok_kernel_table *K, const rb_model *m, const rb_sampling *s,
                      const ok_projector *refs, ok_backprojector *bps,
                      const rb_particles *pool, rb_pool_out *out,
                      unsigned flags, int num_threads, int exact_threshold, ok_debug *dbg)
{
	Shared S;
	S.K = K; S.m = m; S.s = s; S.refs = refs; S.bps = bps;
	S.nc = m->coarse_size; S.nf = m->current_size;
	S.Npc = S.nc * (S.nc / 2 + 1); S.Npf = S.nf * (S.nf / 2 + 1);
	S.nshell = m->ori_size / 2 + 1;
	S.NOR = s->n_over_rot; S.NOT = s->n_over_trans;
	make_mresol(S.nc, S.Mres_c); make_mresol(S.nf, S.Mres_f);
	const int T = s->n_trans, Tf = T * S.NOT;
	S.ctx.resize(T); S.cty.resize(T); S.ftx.resize(Tf); S.fty.resize(Tf);
	for (int t = 0; t < T; t++)                                                              // :1239-1240
	{
		S.ctx[t] = (float) (-2 * M_PI * s->trans_x[t] / (double) m->ori_size);
		S.cty[t] = (float) (-2 * M_PI * s->trans_y[t] / (double) m->ori_size);
	}
	for (int t = 0; t < Tf; t++)                                                             // :1549-1550
	{
		double x = s->over_trans_x ? s->over_trans_x[t] : s->trans_x[t];
		double y = s->over_trans_y ? s->over_trans_y[t] : s->trans_y[t];
		S.ftx[t] = (float) (-2 * M_PI * x / (double) m->ori_size);
		S.fty[t] = (float) (-2 * M_PI * y / (double) m->ori_size);
	}
	{   // coarse matrices: XFLOAT angles -> cpu_kernel_make_eulers_3D (acc_projector_plan_impl.h:246-262)
		size_t n = (size_t) s->n_dir * s->n_psi;
		std::vector<float> a(n), b(n), g(n);
		for (int d = 0; d < s->n_dir; d++)
			for (int q = 0; q < s->n_psi; q++)
			{
				a[(size_t) d * s->n_psi + q] = (float) s->rot[d];
				b[(size_t) d * s->n_psi + q] = (float) s->tilt[d];
				g[(size_t) d * s->n_psi + q] = (float) s->psi[q];
			}
		S.coarse_eulers.resize(n * 9);
		float Lf[9], Rf[9];                                                          // MBL / MBR as XFLOAT (acc_projector_plan_impl.h:232-245)
		for (int i = 0; i < 9; i++) { Lf[i] = pool->mat_left ? (float) pool->mat_left[i] : 0.f; Rf[i] = pool->mat_right ? (float) pool->mat_right[i] : 0.f; }
		K->make_eulers_3d(a.data(), b.data(), g.data(), S.coarse_eulers.data(), n, pool->mat_left ? Lf : NULL, pool->mat_right ? Rf : NULL);
	}
	const int P = pool->n_particles;
	const int Kc = m->nr_classes;
	int status = 0;
	int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
	std::vector<std::vector<double>> tdir(nt, std::vector<double>((size_t) Kc * s->n_dir, 0.)), tcls(nt, std::vector<double>((size_t) 3 * Kc, 0.));   // [Kc] pdf_class + [Kc][2] prior-offset sums
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
	for (int p = 0; p < P; p++)
	{
		int tid = omp_get_thread_num();
		int st = run_particle(S, pool, p, out, flags, exact_threshold != 0, dbg, tdir[tid], tcls[tid]);
		if (st != 0)
		{
#pragma omp critical
			status = st;
		}
	}
	for (int t = 0; t < nt; t++)
	{
		if (out->wsum_pdf_direction) for (size_t i = 0; i < tdir[t].size(); i++) out->wsum_pdf_direction[i] += tdir[t][i];
		if (out->wsum_pdf_class) for (int k = 0; k < Kc; k++) out->wsum_pdf_class[k] += tcls[t][k];
		if (out->wsum_prior_offset_class && m->prior_offset_class)
			for (int k = 0; k < 2 * Kc; k++) out->wsum_prior_offset_class[k] += tcls[t][Kc + k];
	}
	return status;
}

} // extern "C"
