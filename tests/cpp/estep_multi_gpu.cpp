/*
 * E-step + per-iteration reduction + host bookkeeping with NO Python in the loop: drives include/relion_b200_adapter.hpp the way
 * MlOptimiser / MlOptimiserMpi drive the reference's accelerator objects
 *   create   /root/reference/src/ml_optimiser.cpp:3577-3632      MlDeviceBundle / MlOptimiserCuda per device, setupFixedSizedObjects
 *   pools    :4126-4320 expectationSomeParticles                  exp_metadata / exp_imagedata of the pool, doThreadExpectationSomeParticles
 *   reduce   src/ml_optimiser_mpi.cpp:2028-2185                   combineAllWeightedSums (NCCL instead of MPI)
 *   drain    src/ml_optimiser.cpp:3805-3869                       getMdlData += wsum_model.BPref
 * on the mock MlOptimiser of tests/cpp/mock_relion (same member names as RELION's).
 *
 *   estep_multi_gpu <workload.bin> <result.bin> <nranks> [pool_size]
 *
 * Run A: every particle on one GPU.  Run B: the particles split over `nranks` ranks, one MlOptimiser + device bundle each
 * (device r when the box has that many GPUs, else all on device 0 and the reduction summed on the host), reduced with
 * combineAllWeightedSums.  B must reproduce A: identical poses, weighted sums to 1e-5, accumulators to 1e-5 of their maximum.
 * result.bin receives run A's metadata table and packed weighted sums for the Python test to compare with the Python path.
 */
#include "mock_relion/src/ml_optimiser.h"
#include "relion_b200_adapter.hpp"

#include <cuda_runtime_api.h>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>

using namespace relion_b200;

struct Arr { int dtype; std::vector<long long> dims; std::vector<char> raw;
	size_t count() const { size_t n = 1; for (long long d : dims) n *= (size_t) d; return n; }
	const double *f64() const { return (const double *) raw.data(); }
	const int *i32() const { return (const int *) raw.data(); } };

static std::map<std::string, Arr> load_dump(const char *path)
{
	std::map<std::string, Arr> m;
	std::ifstream f(path, std::ios::binary);
	if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
	while (true)
	{
		int nl = 0;
		if (!f.read((char *) &nl, 4)) break;
		std::string name(nl, ' ');
		f.read(&name[0], nl);
		Arr a; int nd = 0;
		f.read((char *) &a.dtype, 4); f.read((char *) &nd, 4);
		a.dims.resize(nd);
		f.read((char *) a.dims.data(), 8 * nd);
		const size_t item = a.dtype == 0 ? 8 : 4;          // 0: float64, 1: float32, 2: int32
		a.raw.resize(a.count() * item);
		f.read(a.raw.data(), (std::streamsize) a.raw.size());
		m[name] = a;
	}
	return m;
}

static double scalar(const std::map<std::string, Arr> &d, const char *k) { return d.at(k).f64()[0]; }

// the optimiser state before an E-step, from the dump
static void build_optimiser(MlOptimiser &o, const std::map<std::string, Arr> &d)
{
	MlModel &m = o.mymodel;
	// "ori_size" is the box of the images (one optics group); "model_ori_size", when present, the model's own box: an optics
	// group whose geometry differs from the model's (src/ml_optimiser.cpp:5735-5777), same pixel size
	const int K = (int) scalar(d, "nr_classes"), img_size = (int) scalar(d, "ori_size");
	const int ori = d.count("model_ori_size") ? (int) scalar(d, "model_ori_size") : img_size, nshell = ori / 2 + 1;
	m.nr_classes = K; m.ori_size = ori; m.pixel_size = scalar(d, "pixel_size"); m.nr_bodies = 1; m.ref_dim = 3; m.data_dim = 2;
	m.nr_groups = (int) scalar(d, "nr_groups"); m.nr_optics_groups = 1; m.sigma2_offset = scalar(d, "sigma2_offset");
	m.avg_norm_correction = scalar(d, "avg_norm_correction"); m.padding_factor = 2.;
	const bool local = scalar(d, "local_search") != 0.;
	m.orientational_prior_mode = local ? PRIOR_ROTTILT_PSI : NOPRIOR;
	m.sigma2_rot = m.sigma2_tilt = m.sigma2_psi = scalar(d, "sigma2_ang");
	const Arr &ref = d.at("refs");                        // [K][Z][Y][X][2] doubles
	const long Z = ref.dims[1], Y = ref.dims[2], X = ref.dims[3];
	m.PPref.resize(K); o.wsum_model.BPref.resize(K);
	for (int k = 0; k < K; k++)
	{
		Projector &p = m.PPref[k];
		p.data.resize(Z, Y, X); p.data.xinit = 0; p.data.yinit = -((Y - 1) / 2); p.data.zinit = -((Z - 1) / 2);
		memcpy(p.data.data, ref.f64() + (size_t) k * Z * Y * X * 2, (size_t) Z * Y * X * 16);
		p.r_max = (int) scalar(d, "r_max"); p.padding_factor = 2.; p.ori_size = ori; p.pad_size = (int) Y;
		BackProjector &b = o.wsum_model.BPref[k];
		b.data.resize(Z, Y, X); b.data.xinit = 0; b.data.yinit = p.data.yinit; b.data.zinit = p.data.zinit;
		b.weight.resize(Z, Y, X); b.r_max = p.r_max; b.padding_factor = 2.; b.ori_size = ori; b.pad_size = (int) Y;
	}
	m.sigma2_noise.resize(1); m.sigma2_noise[0].resize(nshell);
	memcpy(m.sigma2_noise[0].data, d.at("sigma2_noise").f64(), nshell * 8);
	m.scale_correction.assign(d.at("scale_correction").f64(), d.at("scale_correction").f64() + m.nr_groups);
	m.pdf_class.assign(d.at("pdf_class").f64(), d.at("pdf_class").f64() + K);
	const int n_dir = (int) d.at("rot").count();
	m.nr_directions = n_dir;
	m.pdf_direction.resize(K); m.data_vs_prior_class.resize(K); m.prior_offset_class.resize(K);
	for (int k = 0; k < K; k++)
	{
		m.pdf_direction[k].resize(n_dir);
		for (int i = 0; i < n_dir; i++) m.pdf_direction[k].data[i] = 1. / n_dir;
		m.data_vs_prior_class[k].resize(nshell);
		memcpy(m.data_vs_prior_class[k].data, d.at("data_vs_prior_class").f64() + (size_t) k * nshell, nshell * 8);
		m.prior_offset_class[k].resize(2);
	}
	// weighted sums start at zero (MlWsumModel::initZeros)
	MlWsumModel &w = o.wsum_model;
	(MlModel &) w = m;
	w.PPref.clear();
	w.LL = w.ave_Pmax = w.sigma2_offset = w.avg_norm_correction = w.sigma2_rot = w.sigma2_tilt = w.sigma2_psi = 0.;
	w.sigma2_noise[0].initZeros(); w.sumw_ctf2.resize(1); w.sumw_ctf2[0].resize(nshell); w.sumw_stMulti.resize(1); w.sumw_stMulti[0].resize(nshell);
	w.sumw_group.assign(1, 0.); w.wsum_signal_product.assign(m.nr_groups, 0.); w.wsum_reference_power.assign(m.nr_groups, 0.);
	w.pdf_class.assign(K, 0.);
	for (int k = 0; k < K; k++) w.pdf_direction[k].initZeros();
	// experiment
	const int N = (int) d.at("group_id").count();
	o.mydata.group_of_particle.assign(d.at("group_id").i32(), d.at("group_id").i32() + N);
	o.mydata.optics_group_of_particle.assign(N, 0);
	o.mydata.nr_groups = m.nr_groups;
	o.mydata.obsModel.kV.assign(1, 300.); o.mydata.obsModel.Cs.assign(1, 2.7); o.mydata.obsModel.Q0.assign(1, 0.1);
	o.mydata.obsModel.pixel_size.assign(1, m.pixel_size); o.mydata.obsModel.box_size.assign(1, img_size); o.mydata.obsModel.ctf_premultiplied.assign(1, false);
	// sampling tables
	HealpixSampling &s = o.sampling;
	s.healpix_order = (int) scalar(d, "healpix_order"); s.is_3D = true;
	auto vec = [&](const char *k) { return std::vector<RFLOAT>(d.at(k).f64(), d.at(k).f64() + d.at(k).count()); };
	s.rot_angles = vec("rot"); s.tilt_angles = vec("tilt"); s.psi_angles = vec("psi");
	s.n_over_rot = (int) scalar(d, "n_over_rot"); s.n_over_trans = (int) scalar(d, "n_over_trans");
	s.over_rot = vec("over_rot"); s.over_tilt = vec("over_tilt"); s.over_psi = vec("over_psi");
	s.translations_x = vec("trans_x"); s.translations_y = vec("trans_y");           // Angstrom
	s.over_trans_x = vec("over_trans_x"); s.over_trans_y = vec("over_trans_y");
	if (local)
	{
		const int *doff = d.at("dir_off").i32(), *poff = d.at("psi_off").i32(), *didx = d.at("dir_idx").i32(), *pidx = d.at("psi_idx").i32();
		const double *dpr = d.at("dir_prior").f64(), *ppr = d.at("psi_prior").f64(), *md = d.at("metadata").f64();
		const int ncol = (int) d.at("metadata").dims[1];
		for (int p = 0; p < N; p++)
		{
			HealpixSampling::PriorLists l;
			l.dir.assign(didx + doff[p], didx + doff[p + 1]); l.dir_prior.assign(dpr + doff[p], dpr + doff[p + 1]);
			l.psi.assign(pidx + poff[p], pidx + poff[p + 1]); l.psi_prior.assign(ppr + poff[p], ppr + poff[p + 1]);
			s.prior_lists[std::make_tuple(md[(size_t) p * ncol + METADATA_ROT], md[(size_t) p * ncol + METADATA_TILT], md[(size_t) p * ncol + METADATA_PSI])] = l;
		}
	}
	// optimiser flags / sizes
	o.image_full_size.assign(1, img_size); o.image_current_size.assign(1, (int) scalar(d, "current_size")); o.image_coarse_size.assign(1, (int) scalar(d, "coarse_size"));
	// "skip_align": --skip_align (only classify): no oversampling, no priors (src/ml_optimiser.cpp:2382-2389, 2415-2420)
	o.do_skip_align = d.count("skip_align") && scalar(d, "skip_align") != 0.;
	o.iter = 5; o.adaptive_oversampling = o.do_skip_align ? 0 : 1; o.adaptive_fraction = scalar(d, "adaptive_fraction"); o.maximum_significants = -1;
	o.particle_diameter = scalar(d, "particle_diameter"); o.width_mask_edge = (int) scalar(d, "width_mask_edge"); o.sigma2_fudge = 1.;
	o.do_auto_refine = true; o.autosampling_hporder_local_searches = local ? 0 : 99;
	const int cur = o.image_current_size[0], xs = cur / 2 + 1;
	o.Mresol_fine.resize(1); o.Mresol_fine[0].resize(cur, xs);                  // src/ml_optimiser.cpp:5784-5811
	for (int iy = 0; iy < cur; iy++)
		for (int x = 0; x < xs; x++)
		{
			const int y = iy < xs ? iy : iy - cur;
			const int ires = ROUND(sqrt((double) (x * x + y * y)));
			DIRECT_A2D_ELEM(o.Mresol_fine[0], iy, x) = (ires < xs && !(x == 0 && y < 0)) ? ires : -1;
		}
}

// one E-step over particles [p0, p1) in pools of `pool` particles (src/ml_optimiser.cpp:3513-3869 around the device calls)
static void expectation(MlOptimiser &o, const std::map<std::string, Arr> &d, int device, long p0, long p1, int pool, std::vector<double> &metadata_all)
{
	const int ori = o.image_full_size[0], ncol = (int) d.at("metadata").dims[1];   // the images' box
	const float *img = (const float *) d.at("images").raw.data();              // float32 [N][ori][ori]
	MlDeviceBundle *b = new MlDeviceBundle(&o);
	b->setDevice(device);
	b->setupFixedSizedObjects();
	o.accDataBundles.push_back((void *) b);
	std::vector<MlOptimiserCuda *> gpuOptimisers;
	for (int t = 0; t < o.nr_threads; t++) gpuOptimisers.push_back(new MlOptimiserCuda(&o, b, "multi_gpu"));
	b->setupTunableSizedObjects(b->checkFixedSizedObjects(1));
	for (long first = p0; first < p1; first += pool)
	{
		const long last = std::min(p1, first + pool) - 1;
		const int P = (int) (last - first + 1);
		o.exp_my_first_part_id = first; o.exp_my_last_part_id = last;
		// getMetaAndImageDataSubset (:10285-10552): rows and images of the pool
		o.exp_metadata.resize(P, ncol);
		memcpy(o.exp_metadata.data, metadata_all.data() + (size_t) first * ncol, (size_t) P * ncol * 8);
		if (o.do_skip_align)
		{
			// expectationSomeParticles (src/ml_optimiser.cpp:4180-4225): the sampling object is refilled with the pool's own
			// orientations (addOneOrientation) and fractional offsets in Angstrom (addOneTranslation)
			HealpixSampling &s = o.sampling;
			s.rot_angles.clear(); s.tilt_angles.clear(); s.psi_angles.clear(); s.translations_x.clear(); s.translations_y.clear();
			for (int p = 0; p < P; p++)
			{
				s.rot_angles.push_back(DIRECT_A2D_ELEM(o.exp_metadata, p, METADATA_ROT)); s.tilt_angles.push_back(DIRECT_A2D_ELEM(o.exp_metadata, p, METADATA_TILT));
				s.psi_angles.push_back(DIRECT_A2D_ELEM(o.exp_metadata, p, METADATA_PSI));
				const RFLOAT ox = DIRECT_A2D_ELEM(o.exp_metadata, p, METADATA_XOFF), oy = DIRECT_A2D_ELEM(o.exp_metadata, p, METADATA_YOFF);
				s.translations_x.push_back((ox - ROUND(ox)) * o.mymodel.pixel_size); s.translations_y.push_back((oy - ROUND(oy)) * o.mymodel.pixel_size);
			}
		}
		o.exp_imagedata.resize(P, ori, ori);
		for (size_t i = 0; i < (size_t) P * ori * ori; i++) o.exp_imagedata.data[i] = (RFLOAT) img[(size_t) first * ori * ori + i];
		// the OpenMP fan-out (:4280-4282), here one after the other: every thread calls in, thread 0 carries the pool
		for (int t = o.nr_threads - 1; t >= 0; t--) { gpuOptimisers[t]->resetData(); gpuOptimisers[t]->doThreadExpectationSomeParticles(t); }
		// setMetaDataSubset: rows back into the full table
		memcpy(metadata_all.data() + (size_t) first * ncol, o.exp_metadata.data, (size_t) P * ncol * 8);
	}
	b->syncAllBackprojects();
	for (size_t t = 0; t < gpuOptimisers.size(); t++) delete gpuOptimisers[t];
}

static void finish(MlOptimiser &o)
{
	MlDeviceBundle *b = (MlDeviceBundle *) o.accDataBundles[0];
	b->pullBackprojectors();                                                    // :3805-3838
	for (size_t k = 0; k < b->projectors.size(); k++) { b->projectors[k].clear(); b->backprojectors[k].clear(); }
	delete b;
	o.accDataBundles.clear();
}

static double rel_max(const double *a, const double *b, size_t n)
{
	double m = 0., e = 0.;
	for (size_t i = 0; i < n; i++) { m = std::max(m, fabs(a[i])); e = std::max(e, fabs(a[i] - b[i])); }
	return m > 0. ? e / m : e;
}

int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: %s workload.bin result.bin nranks [pool]\n", argv[0]); return 2; }
	try
	{
		const std::map<std::string, Arr> d = load_dump(argv[1]);
		const int nranks = atoi(argv[3]);
		const int pool = argc > 4 ? atoi(argv[4]) : 16;
		const int N = (int) d.at("group_id").count(), ncol = (int) d.at("metadata").dims[1];
		int ndev = 0;
		cudaGetDeviceCount(&ndev);
		const bool real_multi = ndev >= nranks && nranks > 1;

		// ---- run A: one rank ----
		MlOptimiser A;
		build_optimiser(A, d);
		std::vector<double> metaA(d.at("metadata").f64(), d.at("metadata").f64() + (size_t) N * ncol);
		expectation(A, d, 0, 0, N, pool, metaA);
		finish(A);
		std::vector<double> packA;
		WsumPack::pack(A.wsum_model, packA);

		// ---- run B: nranks ranks ----
		std::vector<MlOptimiser> R(nranks);
		std::vector<std::vector<double> > metaR(nranks, std::vector<double>(d.at("metadata").f64(), d.at("metadata").f64() + (size_t) N * ncol));
		std::vector<std::string> errors(nranks);
		for (int r = 0; r < nranks; r++) build_optimiser(R[r], d);
		auto shard = [&](int r, long &a, long &b) { const long base = N / nranks, rem = N % nranks; a = r * base + std::min<long>(r, rem); b = a + base + (r < rem ? 1 : 0); };
		{
			std::vector<std::thread> th;
			for (int r = 0; r < nranks; r++)
				th.emplace_back([&, r]() {
					try { long a, b; shard(r, a, b); expectation(R[r], d, real_multi ? r : 0, a, b, pool, metaR[r]); }
					catch (const std::exception &e) { errors[r] = e.what(); }
				});
			for (auto &t : th) t.join();
		}
		for (int r = 0; r < nranks; r++) if (!errors[r].empty()) { fprintf(stderr, "rank %d: %s\n", r, errors[r].c_str()); return 1; }
		if (real_multi)
		{
			// combineAllWeightedSums over NCCL: one communicator per bundle, the collective calls come from one thread per rank
			std::vector<rb_ctx *> ctxs(nranks);
			for (int r = 0; r < nranks; r++) ctxs[r] = ((MlDeviceBundle *) R[r].accDataBundles[0])->ctx;
			std::vector<rb_comm *> comms(nranks);
			if (rb_comm_create_all(ctxs.data(), nranks, comms.data()) != RB_OK) { fprintf(stderr, "rb_comm_create_all: %s\n", rb_last_error()); return 1; }
			std::vector<std::thread> th;
			for (int r = 0; r < nranks; r++)
				th.emplace_back([&, r]() {
					try { combineAllWeightedSums(&R[r], (MlDeviceBundle *) R[r].accDataBundles[0], comms[r]); }
					catch (const std::exception &e) { errors[r] = e.what(); }
				});
			for (auto &t : th) t.join();
			for (int r = 0; r < nranks; r++) { rb_comm_destroy(comms[r]); if (!errors[r].empty()) { fprintf(stderr, "rank %d: %s\n", r, errors[r].c_str()); return 1; } }
			finish(R[0]);                                                          // rank 0 holds the totals
			for (int r = 1; r < nranks; r++) { delete (MlDeviceBundle *) R[r].accDataBundles[0]; R[r].accDataBundles.clear(); }
		}
		else
		{
			// fewer GPUs than ranks: the ranks shared device 0; sum on the host (what the MPI ring did), same pack order
			std::vector<double> tot, part;
			for (int r = 0; r < nranks; r++)
			{
				finish(R[r]);
				WsumPack::pack(R[r].wsum_model, part);
				if (tot.empty()) tot = part; else for (size_t i = 0; i < tot.size(); i++) tot[i] += part[i];
				if (r > 0)
					for (size_t k = 0; k < R[0].wsum_model.BPref.size(); k++)
						for (long i = 0; i < R[0].wsum_model.BPref[k].data.getSize(); i++)
						{
							R[0].wsum_model.BPref[k].data.data[i].real += R[r].wsum_model.BPref[k].data.data[i].real;
							R[0].wsum_model.BPref[k].data.data[i].imag += R[r].wsum_model.BPref[k].data.data[i].imag;
							R[0].wsum_model.BPref[k].weight.data[i] += R[r].wsum_model.BPref[k].weight.data[i];
						}
			}
			WsumPack::unpack(R[0].wsum_model, tot);
		}
		std::vector<double> packB;
		WsumPack::pack(R[0].wsum_model, packB);

		// ---- B against A ----
		int bad = 0;
		for (int r = 0; r < nranks; r++)
		{
			long a, b; shard(r, a, b);
			for (long p = a; p < b; p++)
				for (int c : {METADATA_ROT, METADATA_TILT, METADATA_PSI, METADATA_XOFF, METADATA_YOFF, METADATA_CLASS, METADATA_NR_SIGN})
					if (metaR[r][(size_t) p * ncol + c] != metaA[(size_t) p * ncol + c]) bad++;
			for (long p = a; p < b; p++)
				for (int c : {METADATA_DLL, METADATA_PMAX, METADATA_NORM})
					if (fabs(metaR[r][(size_t) p * ncol + c] - metaA[(size_t) p * ncol + c]) > 1e-5 * std::max(1., fabs(metaA[(size_t) p * ncol + c]))) bad++;
		}
		const double e_pack = rel_max(packA.data(), packB.data(), packA.size());
		double e_bp = 0.;
		for (size_t k = 0; k < A.wsum_model.BPref.size(); k++)
		{
			e_bp = std::max(e_bp, rel_max((const double *) A.wsum_model.BPref[k].data.data, (const double *) R[0].wsum_model.BPref[k].data.data, 2 * (size_t) A.wsum_model.BPref[k].data.getSize()));
			e_bp = std::max(e_bp, rel_max(A.wsum_model.BPref[k].weight.data, R[0].wsum_model.BPref[k].weight.data, (size_t) A.wsum_model.BPref[k].weight.getSize()));
		}
		printf("particles %d, ranks %d (%s), pool %d: metadata mismatches %d, weighted sums rel %.3g (LL %.9g vs %.9g), accumulators rel %.3g\n",
		       N, nranks, real_multi ? "one GPU each, NCCL" : "shared GPU, host sum", pool, bad, e_pack, packA[0], packB[0], e_bp);
		std::ofstream out(argv[2], std::ios::binary);
		const long long nm = (long long) metaA.size(), np = (long long) packA.size();
		out.write((const char *) &nm, 8); out.write((const char *) metaA.data(), nm * 8);
		out.write((const char *) &np, 8); out.write((const char *) packA.data(), np * 8);
		const bool ok = bad == 0 && e_pack <= 1e-5 && e_bp <= 1e-5 && packA[0] != 0.;
		printf(ok ? "PASS\n" : "FAIL\n");
		return ok ? 0 : 1;
	}
	catch (const std::exception &e)
	{
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
}
