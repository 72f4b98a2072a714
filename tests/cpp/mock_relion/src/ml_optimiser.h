/*
 * TEST INFRASTRUCTURE - a minimal stand-in for RELION's src/ml_optimiser.h (and what it pulls in), written from scratch:
 * only the members include/relion_b200_adapter.hpp reads or writes, with the reference's names, types and macro
 * spellings (/root/reference/src/ml_optimiser.h, ml_model.h, exp_model.h, healpix_sampling.h, multidim_array.h,
 * matrix1d.h, ctf.h, projector.h, backprojector.h), so that the adapter compiles unchanged against either tree.
 * tests/test_adapter_cpp.py::test_adapter_compiles_against_the_reference_headers compiles the same adapter against the
 * real headers.  HealpixSampling here is table driven (the tests fill the tables from relion_b200/sampling.py).
 */
#ifndef MOCK_ML_OPTIMISER_H_
#define MOCK_ML_OPTIMISER_H_

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

typedef double RFLOAT;
#ifndef PI
#define PI 3.14159265358979323846
#endif
#define ROUND(x) (((x) > 0) ? (int) ((x) + 0.5) : (int) ((x) - 0.5))
#define XMIPP_MAX(x, y) (((x) >= (y)) ? (x) : (y))
#define XMIPP_MIN(x, y) (((x) >= (y)) ? (y) : (x))

struct Complex { RFLOAT real, imag; Complex(RFLOAT r = 0, RFLOAT i = 0) : real(r), imag(i) {} };

template <typename T>
class MultidimArray
{
public:
	T *data; long int xdim, ydim, zdim, ndim; long int xinit, yinit, zinit;
	std::vector<T> store;
	MultidimArray() : data(NULL), xdim(0), ydim(0), zdim(0), ndim(0), xinit(0), yinit(0), zinit(0) {}
	MultidimArray(const MultidimArray &o) { *this = o; }
	MultidimArray &operator=(const MultidimArray &o)
	{
		store = o.store; data = store.empty() ? NULL : store.data();
		xdim = o.xdim; ydim = o.ydim; zdim = o.zdim; ndim = o.ndim; xinit = o.xinit; yinit = o.yinit; zinit = o.zinit;
		return *this;
	}
	void resize(long int z, long int y, long int x) { store.assign((size_t) z * y * x, T()); data = store.data(); xdim = x; ydim = y; zdim = z; ndim = 1; }
	void resize(long int y, long int x) { resize(1, y, x); }
	void resize(long int x) { resize(1, 1, x); }
	void initZeros(long int z, long int y, long int x) { resize(z, y, x); }
	void initZeros(long int y, long int x) { resize(1, y, x); }
	void initZeros(long int x) { resize(1, 1, x); }
	void initZeros() { std::fill(store.begin(), store.end(), T()); }
	void setXmippOrigin() { xinit = -(xdim / 2); yinit = -(ydim / 2); zinit = -(zdim / 2); }
	void clear() { store.clear(); data = NULL; xdim = ydim = zdim = ndim = 0; }
	long int getSize() const { return (long int) store.size(); }
};
#define XSIZE(v) ((v).xdim)
#define YSIZE(v) ((v).ydim)
#define ZSIZE(v) ((v).zdim)
#define NZYXSIZE(v) ((v).ndim * (v).zdim * (v).ydim * (v).xdim)
#define MULTIDIM_SIZE(v) NZYXSIZE(v)
#define MULTIDIM_ARRAY(v) ((v).data)
#define STARTINGX(v) ((v).xinit)
#define STARTINGY(v) ((v).yinit)
#define STARTINGZ(v) ((v).zinit)
#define DIRECT_MULTIDIM_ELEM(v, n) ((v).data[(n)])
#define DIRECT_A1D_ELEM(v, i) ((v).data[(i)])
#define DIRECT_A2D_ELEM(v, i, j) ((v).data[(i) * (v).xdim + (j)])
#define DIRECT_A3D_ELEM(v, k, i, j) ((v).data[((k) * (v).ydim + (i)) * (v).xdim + (j)])
#define FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(v) for (long int n = 0; n < NZYXSIZE(v); ++n)

template <typename T>
class Matrix1D
{
public:
	T *vdata; int vdim; std::vector<T> store;
	Matrix1D(int n = 0) : vdata(NULL), vdim(0) { resize(n); }
	Matrix1D(const Matrix1D &o) { *this = o; }
	Matrix1D &operator=(const Matrix1D &o) { store = o.store; vdata = store.empty() ? NULL : store.data(); vdim = o.vdim; return *this; }
	void resize(int n) { store.assign(n, T()); vdata = store.empty() ? NULL : store.data(); vdim = n; }
	void initZeros(int n) { resize(n); }
};
#define XX(v) (v).vdata[0]
#define YY(v) (v).vdata[1]
#define ZZ(v) (v).vdata[2]

// exp_metadata columns (src/ml_optimiser.h:51-80)
#define METADATA_ROT 0
#define METADATA_TILT 1
#define METADATA_PSI 2
#define METADATA_XOFF 3
#define METADATA_YOFF 4
#define METADATA_ZOFF 5
#define METADATA_CLASS 6
#define METADATA_DLL 7
#define METADATA_PMAX 8
#define METADATA_NR_SIGN 9
#define METADATA_NORM 10
#define METADATA_CTF_DEFOCUS_U 11
#define METADATA_CTF_DEFOCUS_V 12
#define METADATA_CTF_DEFOCUS_ANGLE 13
#define METADATA_CTF_BFACTOR 14
#define METADATA_CTF_KFACTOR 15
#define METADATA_CTF_PHASE_SHIFT 16
#define METADATA_ROT_PRIOR 17
#define METADATA_TILT_PRIOR 18
#define METADATA_PSI_PRIOR 19
#define METADATA_XOFF_PRIOR 20
#define METADATA_YOFF_PRIOR 21
#define METADATA_ZOFF_PRIOR 22
#define METADATA_PSI_PRIOR_FLIP_RATIO 23
#define METADATA_ROT_PRIOR_FLIP_RATIO 24
#define METADATA_LINE_LENGTH_BEFORE_BODIES 25
#define METADATA_LINE_LENGTH METADATA_LINE_LENGTH_BEFORE_BODIES
// MlModel::orientational_prior_mode (src/healpix_sampling.h)
#define NOPRIOR 0
#define PRIOR_ROTTILT_PSI 1

class Projector
{
public:
	MultidimArray<Complex> data;
	int ori_size, r_max, pad_size, ref_dim, data_dim; RFLOAT padding_factor;
	Projector() : ori_size(0), r_max(0), pad_size(0), ref_dim(3), data_dim(2), padding_factor(2.) {}
};
class BackProjector : public Projector
{
public:
	MultidimArray<RFLOAT> weight;
};

// src/jaz/image/buffered_image.h: operator()(x, y)
template <class T> class BufferedImage
{
public:
	int w, h; std::vector<T> data;
	BufferedImage() : w(0), h(0) {}
	BufferedImage(int w_, int h_) : w(w_), h(h_), data((size_t) w_ * h_) {}
	const T &operator()(int x, int y) const { return data[(size_t) y * w + x]; }
	T &operator()(int x, int y) { return data[(size_t) y * w + x]; }
};

// src/matrix2d.h: the few members the adapter touches
template <class T> class Matrix2D
{
public:
	int mdimx, mdimy; std::vector<T> mdata;
	Matrix2D() : mdimx(0), mdimy(0) {}
	Matrix2D(int rows, int cols) : mdimx(cols), mdimy(rows), mdata((size_t) rows * cols, T(0)) {}
	void initIdentity(int dim) { mdimx = mdimy = dim; mdata.assign((size_t) dim * dim, T(0)); for (int i = 0; i < dim; i++) mdata[(size_t) i * dim + i] = T(1); }
	T &operator()(int i, int j) { return mdata[(size_t) i * mdimx + j]; }
	const T &operator()(int i, int j) const { return mdata[(size_t) i * mdimx + j]; }
	bool isIdentity() const
	{
		for (int i = 0; i < mdimy; i++) for (int j = 0; j < mdimx; j++)
			if (std::fabs((*this)(i, j) - (i == j ? T(1) : T(0))) > 1e-6) return false;          // XMIPP_EQUAL_ACCURACY
		return true;
	}
	Matrix2D &operator*=(T f) { for (size_t i = 0; i < mdata.size(); i++) mdata[i] *= f; return *this; }
};

class ObservationModel
{
public:
	std::vector<RFLOAT> kV, Cs, Q0, pixel_size; std::vector<int> box_size; std::vector<bool> ctf_premultiplied;
	// anisotropic magnification and scale differences (src/jaz/single_particle/obs_model.cpp:1306-1340)
	bool hasMagMatrices; std::vector<Matrix2D<RFLOAT> > magMatrices;
	Matrix2D<RFLOAT> applyAnisoMag(Matrix2D<RFLOAT> A3D, int og)
	{
		if (!hasMagMatrices) return A3D;
		const Matrix2D<RFLOAT> &M = magMatrices[og];                                                // inverse of the 2x2 block, times A3D
		const RFLOAT det = M(0, 0) * M(1, 1) - M(0, 1) * M(1, 0);
		Matrix2D<RFLOAT> inv; inv.initIdentity(3);
		inv(0, 0) = M(1, 1) / det; inv(0, 1) = -M(0, 1) / det; inv(1, 0) = -M(1, 0) / det; inv(1, 1) = M(0, 0) / det;
		Matrix2D<RFLOAT> out(3, 3);
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) out(i, j) += inv(i, k) * A3D(k, j);
		return out;
	}
	Matrix2D<RFLOAT> applyScaleDifference(Matrix2D<RFLOAT> A3D, int og, int s3D, double angpix3D)
	{
		A3D *= (box_size[og] * pixel_size[og]) / (s3D * angpix3D);
		return A3D;
	}
	// beam tilt / odd Zernike phase correction and detector MTF (src/jaz/single_particle/obs_model.h:45-129, obs_model.cpp:528-626)
	bool hasOddZernike, hasMultipleMtfs;
	std::vector<BufferedImage<Complex> > phaseCorr; std::vector<BufferedImage<RFLOAT> > mtfImage; BufferedImage<RFLOAT> avgMtfImage;
	void demodulatePhase(int og, MultidimArray<Complex> &img, bool do_modulate_instead = false)
	{
		if (!hasOddZernike || (int) phaseCorr.size() <= og) return;
		for (int y = 0; y < img.ydim; y++) for (int x = 0; x < img.xdim; x++)
		{
			Complex &v = DIRECT_A2D_ELEM(img, y, x); const Complex c = phaseCorr[og](x, y);
			const RFLOAT ci = do_modulate_instead ? c.imag : -c.imag;
			v = Complex(v.real * c.real - v.imag * ci, v.real * ci + v.imag * c.real);
		}
	}
	void divideByMtf(int og, MultidimArray<Complex> &img, bool do_multiply_instead = false, bool do_correct_average_mtf = true)
	{
		if (do_correct_average_mtf && !hasMultipleMtfs) return;
		if ((int) mtfImage.size() <= og) return;
		for (int y = 0; y < img.ydim; y++) for (int x = 0; x < img.xdim; x++)
		{
			Complex &v = DIRECT_A2D_ELEM(img, y, x);
			const RFLOAT f = (do_correct_average_mtf ? avgMtfImage(x, y) : 1.) / mtfImage[og](x, y);
			v = Complex(v.real * f, v.imag * f);
		}
	}
	ObservationModel() : hasMagMatrices(false), hasOddZernike(false), hasMultipleMtfs(false) {}
	bool getCtfPremultiplied(int og) const { return ctf_premultiplied[og]; }
	RFLOAT getPixelSize(int og) const { return pixel_size[og]; }
	int getBoxSize(int og) const { return box_size[og]; }
	int numberOfOpticsGroups() const { return (int) kV.size(); }
};

// src/ctf.h: public parameters + setValuesByGroup (kV, Cs, Q0 come from the optics group)
class CTF
{
public:
	RFLOAT kV, DeltafU, DeltafV, azimuthal_angle, Cs, Bfac, scale, phase_shift, Q0;
	void setValuesByGroup(ObservationModel *obs, int opticsGroup, RFLOAT defU, RFLOAT defV, RFLOAT defAng, RFLOAT _Bfac = 0., RFLOAT _scale = 1., RFLOAT _phase_shift = 0., RFLOAT dose = -1.)
	{
		kV = obs->kV[opticsGroup]; Cs = obs->Cs[opticsGroup]; Q0 = obs->Q0[opticsGroup];
		DeltafU = defU; DeltafV = defV; azimuthal_angle = defAng; Bfac = _Bfac; scale = _scale; phase_shift = _phase_shift;
	}
};

class Experiment
{
public:
	ObservationModel obsModel;
	std::vector<int> group_of_particle, optics_group_of_particle; int nr_groups;
	Experiment() : nr_groups(1) {}
	long int getGroupId(long int part_id) { return group_of_particle[part_id]; }
	int getOpticsGroup(long int part_id) { return optics_group_of_particle[part_id]; }
	int getOpticsImageSize(int og) { return obsModel.getBoxSize(og); }
	RFLOAT getOpticsPixelSize(int og) { return obsModel.getPixelSize(og); }
	RFLOAT getImagePixelSize(long int part_id) { return obsModel.getPixelSize(getOpticsGroup(part_id)); }
	int numberOfImagesInParticle(long int) { return 1; }
	int numberOfGroups() { return nr_groups; }
	int numberOfOpticsGroups() { return obsModel.numberOfOpticsGroups(); }
};

// table-driven stand-in for HealpixSampling (src/healpix_sampling.h): angles in degrees, translations in Angstrom
class HealpixSampling
{
public:
	int healpix_order; bool is_3D;
	std::vector<RFLOAT> rot_angles, tilt_angles, psi_angles, translations_x, translations_y, translations_z;
	int n_over_rot, n_over_trans;                                     // oversampling order 1: 8 / 4 (3D), 2 / 4 (2D)
	std::vector<RFLOAT> over_rot, over_tilt, over_psi;                // [(idir * npsi + ipsi) * n_over_rot + io]
	std::vector<RFLOAT> over_trans_x, over_trans_y;                   // [itrans * n_over_trans + io] in Angstrom
	// local searches: lists per (prior_rot, prior_tilt, prior_psi), filled by the test
	struct PriorLists { std::vector<int> dir, psi; std::vector<RFLOAT> dir_prior, psi_prior; };
	std::map<std::tuple<RFLOAT, RFLOAT, RFLOAT>, PriorLists> prior_lists;
	HealpixSampling() : healpix_order(0), is_3D(true), n_over_rot(1), n_over_trans(1) {}

	long int NrDirections(int oversampling_order = 0, const std::vector<int> *pointer_dir_nonzeroprior = NULL)
	{
		const long int n = (pointer_dir_nonzeroprior && !pointer_dir_nonzeroprior->empty()) ? (long int) pointer_dir_nonzeroprior->size() : (long int) rot_angles.size();
		return oversampling_order == 0 ? n : n * (is_3D ? 4 : 1);
	}
	long int NrPsiSamplings(int oversampling_order = 0, const std::vector<int> *pointer_psi_nonzeroprior = NULL)
	{
		const long int n = (pointer_psi_nonzeroprior && !pointer_psi_nonzeroprior->empty()) ? (long int) pointer_psi_nonzeroprior->size() : (long int) psi_angles.size();
		return oversampling_order == 0 ? n : n * 2;
	}
	long int NrTranslationalSamplings(int oversampling_order = 0) { return (long int) translations_x.size() * (oversampling_order == 0 ? 1 : n_over_trans); }
	int oversamplingFactorOrientations(int oversampling_order) { return oversampling_order == 0 ? 1 : n_over_rot; }
	int oversamplingFactorTranslations(int oversampling_order) { return oversampling_order == 0 ? 1 : n_over_trans; }
	void getTranslationsInPixel(long int itrans, int oversampling_order, RFLOAT my_pixel_size, std::vector<RFLOAT> &x, std::vector<RFLOAT> &y,
	                            std::vector<RFLOAT> &z, bool do_helical_refine = false)
	{
		x.clear(); y.clear(); z.clear();
		if (oversampling_order == 0) { x.push_back(translations_x[itrans] / my_pixel_size); y.push_back(translations_y[itrans] / my_pixel_size); return; }
		for (int io = 0; io < n_over_trans; io++)
		{
			x.push_back(over_trans_x[itrans * n_over_trans + io] / my_pixel_size);
			y.push_back(over_trans_y[itrans * n_over_trans + io] / my_pixel_size);
		}
	}
	void getOrientations(long int idir, long int ipsi, int oversampling_order, std::vector<RFLOAT> &my_rot, std::vector<RFLOAT> &my_tilt,
	                     std::vector<RFLOAT> &my_psi, std::vector<int> &pointer_dir_nonzeroprior, std::vector<RFLOAT> &directions_prior,
	                     std::vector<int> &pointer_psi_nonzeroprior, std::vector<RFLOAT> &psi_prior)
	{
		const long int gd = pointer_dir_nonzeroprior.empty() ? idir : pointer_dir_nonzeroprior[idir];
		const long int gp = pointer_psi_nonzeroprior.empty() ? ipsi : pointer_psi_nonzeroprior[ipsi];
		my_rot.clear(); my_tilt.clear(); my_psi.clear();
		if (oversampling_order == 0) { my_rot.push_back(rot_angles[gd]); my_tilt.push_back(tilt_angles[gd]); my_psi.push_back(psi_angles[gp]); return; }
		const size_t g = ((size_t) gd * psi_angles.size() + gp) * n_over_rot;
		for (int io = 0; io < n_over_rot; io++) { my_rot.push_back(over_rot[g + io]); my_tilt.push_back(over_tilt[g + io]); my_psi.push_back(over_psi[g + io]); }
	}
	void selectOrientationsWithNonZeroPriorProbability(RFLOAT prior_rot, RFLOAT prior_tilt, RFLOAT prior_psi, RFLOAT sigma_rot, RFLOAT sigma_tilt,
	                                                   RFLOAT sigma_psi, std::vector<int> &pointer_dir_nonzeroprior, std::vector<RFLOAT> &directions_prior,
	                                                   std::vector<int> &pointer_psi_nonzeroprior, std::vector<RFLOAT> &psi_prior,
	                                                   bool do_bimodal_search_psi = false, RFLOAT sigma_cutoff = 3., RFLOAT sigma_tilt_from_ninety = -1.,
	                                                   RFLOAT sigma_psi_from_zero = -1.)
	{
		const PriorLists &l = prior_lists.at(std::make_tuple(prior_rot, prior_tilt, prior_psi));
		pointer_dir_nonzeroprior = l.dir; directions_prior = l.dir_prior; pointer_psi_nonzeroprior = l.psi; psi_prior = l.psi_prior;
	}
};

class MlModel
{
public:
	int ref_dim, data_dim, ori_size; RFLOAT pixel_size; int current_size;
	int nr_classes, nr_bodies, nr_groups, nr_optics_groups; long long int nr_directions;
	RFLOAT padding_factor, LL, ave_Pmax, avg_norm_correction, sigma2_offset, sigma2_rot, sigma2_tilt, sigma2_psi;
	int orientational_prior_mode;
	std::vector<Projector> PPref;
	std::vector<MultidimArray<RFLOAT> > sigma2_noise, data_vs_prior_class, pdf_direction;
	std::vector<RFLOAT> scale_correction, pdf_class;
	std::vector<Matrix1D<RFLOAT> > prior_offset_class;
	MlModel() : ref_dim(3), data_dim(2), ori_size(0), pixel_size(1.), current_size(0), nr_classes(1), nr_bodies(1), nr_groups(1), nr_optics_groups(1),
	            nr_directions(0), padding_factor(2.), LL(0), ave_Pmax(0), avg_norm_correction(1.), sigma2_offset(0), sigma2_rot(0), sigma2_tilt(0),
	            sigma2_psi(0), orientational_prior_mode(NOPRIOR) {}
};

class MlWsumModel : public MlModel
{
public:
	std::vector<BackProjector> BPref;
	std::vector<RFLOAT> sumw_group, wsum_signal_product, wsum_reference_power;
	std::vector<MultidimArray<RFLOAT> > sumw_ctf2, sumw_stMulti;
};

class MlOptimiser
{
public:
	MlModel mymodel; MlWsumModel wsum_model; Experiment mydata; HealpixSampling sampling;
	MultidimArray<RFLOAT> exp_metadata, exp_imagedata;
	long int exp_my_first_part_id, exp_my_last_part_id;
	std::vector<int> image_coarse_size, image_current_size, image_full_size;
	std::vector<MultidimArray<int> > Mresol_fine, Mresol_coarse;
	int iter, adaptive_oversampling, maximum_significants, nr_threads, nr_pool, random_seed;
	RFLOAT adaptive_fraction, particle_diameter, sigma2_fudge, offset_range_x, offset_range_y, offset_range_z;
	int width_mask_edge, autosampling_hporder_local_searches;
	bool do_ctf_correction, refs_are_ctf_corrected, do_scale_correction, do_norm_correction, do_map, do_zero_mask, do_firstiter_cc, do_always_cc,
	     do_skip_maximization, do_skip_align, do_skip_rotate, do_auto_refine, do_helical_refine, do_gpu, ctf_phase_flipped, only_flip_phases, intact_ctf_first_peak,
	     do_grad, grad_pseudo_halfsets;                  // src/ml_optimiser.h:357-363
	std::vector<void *> accDataBundles, gpuOptimisers;      // src/ml_optimiser.h:109
	MlOptimiser() : exp_my_first_part_id(0), exp_my_last_part_id(-1), iter(2), adaptive_oversampling(1), maximum_significants(-1), nr_threads(1), nr_pool(1), random_seed(0),
	                adaptive_fraction(0.999), particle_diameter(-1.), sigma2_fudge(1.), offset_range_x(-1.), offset_range_y(-1.), offset_range_z(-1.),
	                width_mask_edge(5), autosampling_hporder_local_searches(4),
	                do_ctf_correction(true), refs_are_ctf_corrected(true), do_scale_correction(true), do_norm_correction(true), do_map(true), do_zero_mask(true),
	                do_firstiter_cc(false), do_always_cc(false), do_skip_maximization(false), do_skip_align(false), do_skip_rotate(false), do_auto_refine(true),
	                do_helical_refine(false), do_gpu(true), ctf_phase_flipped(false), only_flip_phases(false), intact_ctf_first_peak(false),
	                do_grad(false), grad_pseudo_halfsets(false) {}
};

#endif
