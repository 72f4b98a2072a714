// TEST INFRASTRUCTURE ONLY (the checker).  C entry points around the REFERENCE's own host classes, compiled from the sources
// where they lie under /root/reference (oracle/Makefile, target `refrecon` -> oracle/_ref/librefrecon.so; nothing is copied):
//   Projector::computeFourierTransformMap            src/projector.cpp      (row f3: reference preparation)
//   BackProjector::reconstruct / symmetrise /
//     updateSSNRarrays / set2DFourierTransform        src/backprojector.cpp  (row f2, and the posed back-projection of a12)
//   softMaskOutsideMap                                src/mask.cpp           (rows f1 / f2)
//   FourierTransformer, CenterFFT, windowFourierTransform, shiftImageInFourierTransform   src/fftw.{h,cpp}   (row f1)
//   getSpectrum-style power spectrum of a particle    restated below from src/ml_optimiser.cpp (the function itself lives in
//                                                     the 10 000-line optimiser translation unit, which cannot be built alone)
// FFTW is replaced by oracle/fftw_shim.cpp (a double-precision DFT with FFTW's conventions); TIFF entry points are stubs that
// abort.  The numpy restatements oracle/reconstruct.py, oracle/prepare.py and relion_b200/synth.py are pinned against these
// functions by tests/test_reference_host.py, and tests/golden/host_*.npz holds their outputs for machines without
// /root/reference (tools/make_host_golden.py).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "src/projector.h"
#include "src/backprojector.h"
#include "src/mask.h"
#include "src/fftw.h"
#include "src/symmetries.h"

// ---- symbols the reference's utility sources expect from its main programs / libtiff ---------------------------------
std::string pipeline_control_outputname = "";
const char *g_RELION_VERSION = "oracle";
extern "C" {
struct tiff;
int TIFFGetField(tiff *, unsigned, ...) { abort(); }
int TIFFGetFieldDefaulted(tiff *, unsigned, ...) { abort(); }
int TIFFSetDirectory(tiff *, unsigned short) { abort(); }
void *_TIFFmalloc(long) { abort(); }
void _TIFFfree(void *) { abort(); }
long TIFFStripSize(tiff *) { abort(); }
long TIFFReadEncodedStrip(tiff *, unsigned, void *, long) { abort(); }
tiff *TIFFOpen(const char *, const char *) { abort(); }
unsigned TIFFNumberOfStrips(tiff *) { abort(); }
void TIFFClose(tiff *) { abort(); }
}

namespace {
std::string g_err;
template <typename F> int guarded(F f)
{
	try { f(); return 0; }
	catch (RelionError &e) { g_err = e.msg; return 1; }
	catch (std::exception &e) { g_err = e.what(); return 1; }
}

void load_bp(BackProjector &bp, const double *re, const double *im, const double *w)
{
	FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(bp.data)
	{
		DIRECT_MULTIDIM_ELEM(bp.data, n).real = re[n];
		DIRECT_MULTIDIM_ELEM(bp.data, n).imag = im[n];
		DIRECT_MULTIDIM_ELEM(bp.weight, n) = w[n];
	}
}
void store_bp(const BackProjector &bp, double *re, double *im, double *w)
{
	FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(bp.data)
	{
		re[n] = DIRECT_MULTIDIM_ELEM(bp.data, n).real;
		im[n] = DIRECT_MULTIDIM_ELEM(bp.data, n).imag;
		w[n] = DIRECT_MULTIDIM_ELEM(bp.weight, n);
	}
}
}  // namespace

extern "C" {

const char *refrec_last_error() { return g_err.c_str(); }

// Projector::computeFourierTransformMap.  vol: [ori]^ref_dim real-space map (C order, unshifted: voxel (0,..) first).
// dims_out: {Z, Y, X, startZ, startY, r_max}; data_out: interleaved complex [Z][Y][X] (capacity in complex elements).
int refrec_ft_map(const double *vol, int ori_size, int ref_dim, int current_size, double padding_factor, int data_dim,
                  int do_gridding, int *dims_out, double *data_out, long long capacity, double *power_spectrum)
{
	return guarded([&] {
		Projector p(ori_size, TRILINEAR, (float) padding_factor, 10, data_dim);
		MultidimArray<RFLOAT> v, ps;
		if (ref_dim == 3) v.resize(ori_size, ori_size, ori_size); else v.resize(ori_size, ori_size);
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(v) DIRECT_MULTIDIM_ELEM(v, n) = vol[n];
		v.setXmippOrigin();
		p.computeFourierTransformMap(v, ps, current_size, 1, do_gridding != 0);
		dims_out[0] = (int) ZSIZE(p.data); dims_out[1] = (int) YSIZE(p.data); dims_out[2] = (int) XSIZE(p.data);
		dims_out[3] = (int) STARTINGZ(p.data); dims_out[4] = (int) STARTINGY(p.data); dims_out[5] = p.r_max;
		if ((long long) MULTIDIM_SIZE(p.data) > capacity) REPORT_ERROR("refrec_ft_map: output capacity too small");
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(p.data)
		{
			data_out[2 * n] = DIRECT_MULTIDIM_ELEM(p.data, n).real;
			data_out[2 * n + 1] = DIRECT_MULTIDIM_ELEM(p.data, n).imag;
		}
		if (power_spectrum)
			for (int i = 0; i < ori_size / 2 + 1; i++) power_spectrum[i] = i < (int) XSIZE(ps) ? DIRECT_A1D_ELEM(ps, i) : 0.;
	});
}

// padded accumulator size for (ori_size, current_size, padding factor): {Z, Y, X, r_max}
int refrec_bp_dims(int ori_size, int ref_dim, int current_size, double padding_factor, int *dims_out)
{
	return guarded([&] {
		BackProjector bp(ori_size, ref_dim, "C1", TRILINEAR, (float) padding_factor, 10, 0, 1.9, 15, 2, true);
		bp.initZeros(current_size);
		dims_out[0] = (int) ZSIZE(bp.data); dims_out[1] = (int) YSIZE(bp.data); dims_out[2] = (int) XSIZE(bp.data); dims_out[3] = bp.r_max;
	});
}

// BackProjector::reconstruct on given accumulators (centred arrays [Z][Y][X] as the class holds them).  vol_out: [ori]^ref_dim.
int refrec_reconstruct(const double *re, const double *im, const double *w, int ori_size, int ref_dim, int current_size,
                       double padding_factor, int skip_gridding, int max_iter_preweight, int do_map, const double *tau2, int n_tau2,
                       double tau2_fudge, double normalise, int minres_map, double *vol_out)
{
	return guarded([&] {
		BackProjector bp(ori_size, ref_dim, "C1", TRILINEAR, (float) padding_factor, 10, 0, 1.9, 15, 2, skip_gridding != 0);
		bp.initZeros(current_size);
		load_bp(bp, re, im, w);
		MultidimArray<RFLOAT> t2, vol;
		t2.resize(n_tau2);
		for (int i = 0; i < n_tau2; i++) DIRECT_A1D_ELEM(t2, i) = tau2 ? tau2[i] : 0.;
		bp.reconstruct(vol, max_iter_preweight, do_map != 0, t2, tau2_fudge, normalise, minres_map, false);
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(vol) vol_out[n] = DIRECT_MULTIDIM_ELEM(vol, n);
	});
}

// BackProjector::symmetrise (point groups; no helical symmetry) in place
int refrec_symmetrise(double *re, double *im, double *w, int ori_size, int ref_dim, int current_size, double padding_factor, const char *sym)
{
	return guarded([&] {
		BackProjector bp(ori_size, ref_dim, sym, TRILINEAR, (float) padding_factor, 10, 0, 1.9, 15, 2, true);
		bp.initZeros(current_size);
		load_bp(bp, re, im, w);
		bp.symmetrise(1, 0., 0., 1);
		store_bp(bp, re, im, w);
	});
}

// BackProjector::symmetrise with helical symmetry (twist in degrees, rise in pixels) in place
int refrec_symmetrise_helical(double *re, double *im, double *w, int ori_size, int ref_dim, int current_size, double padding_factor,
                              const char *sym, int nr_helical_asu, double helical_twist, double helical_rise)
{
	return guarded([&] {
		BackProjector bp(ori_size, ref_dim, sym, TRILINEAR, (float) padding_factor, 10, 0, 1.9, 15, 2, true);
		bp.initZeros(current_size);
		load_bp(bp, re, im, w);
		bp.symmetrise(nr_helical_asu, helical_twist, helical_rise, 1);
		store_bp(bp, re, im, w);
	});
}

// the rotation matrices SymList hands to symmetrise (R of get_matrices, 3x3 row-major each); returns their number or -1
int refrec_sym_matrices(const char *sym, double *R_out, int capacity)
{
	int nsym = -1;
	if (guarded([&] {
		SymList SL;
		SL.read_sym_file(sym);
		Matrix2D<RFLOAT> L(4, 4), R(4, 4);
		nsym = SL.SymsNo();
		if (nsym > capacity) REPORT_ERROR("refrec_sym_matrices: capacity too small");
		for (int i = 0; i < nsym; i++)
		{
			SL.get_matrices(i, L, R);
			for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R_out[9 * i + 3 * r + c] = R(r, c);
		}
	})) return -1;
	return nsym;
}

// BackProjector::updateSSNRarrays; spectra are [ori_size/2 + 1]
int refrec_update_ssnr(const double *w, int ori_size, int ref_dim, int current_size, double padding_factor, double tau2_fudge,
                       double *tau2_io, double *sigma2_out, double *dvp_out, double *cov_out, const double *fsc, const double *avgctf2,
                       int update_tau2_with_fsc, int is_whole_instead_of_half)
{
	return guarded([&] {
		BackProjector bp(ori_size, ref_dim, "C1", TRILINEAR, (float) padding_factor, 10, 0, 1.9, 15, 2, true);
		bp.initZeros(current_size);
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(bp.weight) DIRECT_MULTIDIM_ELEM(bp.weight, n) = w[n];
		const int ns = ori_size / 2 + 1;
		MultidimArray<RFLOAT> t2(ns), s2, dvp, cov, f(ns), a(ns);
		for (int i = 0; i < ns; i++)
		{
			DIRECT_A1D_ELEM(t2, i) = tau2_io[i];
			DIRECT_A1D_ELEM(f, i) = fsc ? fsc[i] : 0.;
			DIRECT_A1D_ELEM(a, i) = avgctf2 ? avgctf2[i] : 1.;
		}
		bp.updateSSNRarrays(tau2_fudge, t2, s2, dvp, cov, f, a, update_tau2_with_fsc != 0, is_whole_instead_of_half != 0, avgctf2 != NULL);
		for (int i = 0; i < ns; i++)
		{
			tau2_io[i] = DIRECT_A1D_ELEM(t2, i); sigma2_out[i] = DIRECT_A1D_ELEM(s2, i);
			dvp_out[i] = DIRECT_A1D_ELEM(dvp, i); cov_out[i] = DIRECT_A1D_ELEM(cov, i);
		}
	});
}

// BackProjector::set2DFourierTransform for n_img images into a 3D accumulator (relion_reconstruct's inner call):
// imgs: [n_img][n][n/2+1] interleaved complex (FFTW layout), A: [n_img][3][3] (the matrix handed to the call: the Euler
// matrix A3D of the particle; backproject2Dto3D inverts it itself, src/backprojector.cpp:78-83), weights: [n_img][n][n/2+1] or NULL.  re/im/w out: centred [Z][Y][X].
int refrec_backproject(const double *imgs, const double *A, const double *weights, int n_img, int n, int ori_size, int current_size,
                       double padding_factor, double *re, double *im, double *w)
{
	return guarded([&] {
		BackProjector bp(ori_size, 3, "C1", TRILINEAR, (float) padding_factor, 10, 0, 1.9, 15, 2, true);
		bp.initZeros(current_size);
		const int xs = n / 2 + 1;
		MultidimArray<Complex> F(n, xs);
		MultidimArray<RFLOAT> W(n, xs);
		Matrix2D<RFLOAT> M(3, 3);
		for (int i = 0; i < n_img; i++)
		{
			FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(F)
			{
				DIRECT_MULTIDIM_ELEM(F, n).real = imgs[2 * ((size_t) i * MULTIDIM_SIZE(F) + n)];
				DIRECT_MULTIDIM_ELEM(F, n).imag = imgs[2 * ((size_t) i * MULTIDIM_SIZE(F) + n) + 1];
				DIRECT_MULTIDIM_ELEM(W, n) = weights ? weights[(size_t) i * MULTIDIM_SIZE(F) + n] : 1.;
			}
			for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) M(r, c) = A[9 * (size_t) i + 3 * r + c];
			bp.set2DFourierTransform(F, M, &W);
		}
		store_bp(bp, re, im, w);
	});
}

// softMaskOutsideMap(vol, radius, cosine_width) in place; dim 2 or 3, n^dim values
int refrec_soft_mask(double *vol, int n, int dim, double radius, double cosine_width)
{
	return guarded([&] {
		MultidimArray<RFLOAT> v;
		if (dim == 3) v.resize(n, n, n); else v.resize(n, n);
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(v) DIRECT_MULTIDIM_ELEM(v, n) = vol[n];
		v.setXmippOrigin();
		softMaskOutsideMap(v, radius, cosine_width);
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(v) vol[n] = DIRECT_MULTIDIM_ELEM(v, n);
	});
}

// The transform of one particle image as getFourierTransformsAndCtfs makes it (src/ml_optimiser.cpp): CenterFFT(img, true),
// FourierTransformer::FourierTransform (normalised by the number of pixels), windowFourierTransform to current_size; with
// shift != 0 also shiftImageInFourierTransform by (sx, sy).  out: [current_size][current_size/2+1] interleaved complex.
int refrec_image_ft(const double *img, int n, int current_size, double sx, double sy, double *out)
{
	return guarded([&] {
		MultidimArray<RFLOAT> I(n, n);
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(I) DIRECT_MULTIDIM_ELEM(I, n) = img[n];
		I.setXmippOrigin();
		CenterFFT(I, true);
		FourierTransformer tr;
		MultidimArray<Complex> Faux, Fimg, Fsh;
		tr.FourierTransform(I, Faux);
		windowFourierTransform(Faux, Fimg, current_size);
		if (sx != 0. || sy != 0.)
		{
			shiftImageInFourierTransform(Fimg, Fsh, (RFLOAT) n, sx, sy);
			Fimg = Fsh;
		}
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(Fimg)
		{
			out[2 * n] = DIRECT_MULTIDIM_ELEM(Fimg, n).real;
			out[2 * n + 1] = DIRECT_MULTIDIM_ELEM(Fimg, n).imag;
		}
	});
}

// getFSC(map1, map2): [n/2 + 1]
int refrec_fsc(const double *m1, const double *m2, int n, double *fsc_out)
{
	return guarded([&] {
		MultidimArray<RFLOAT> a(n, n, n), b(n, n, n), f;
		FOR_ALL_DIRECT_ELEMENTS_IN_MULTIDIMARRAY(a) { DIRECT_MULTIDIM_ELEM(a, n) = m1[n]; DIRECT_MULTIDIM_ELEM(b, n) = m2[n]; }
		getFSC(a, b, f);
		for (int i = 0; i < n / 2 + 1; i++) fsc_out[i] = i < (int) XSIZE(f) ? DIRECT_A1D_ELEM(f, i) : 0.;
	});
}

}  // extern "C"
