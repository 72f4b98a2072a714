// stub of fftw3.h (declarations only): lets RELION's headers parse without FFTW installed
#pragma once
#include <stddef.h>
extern "C" {
typedef double fftw_complex[2];
typedef float fftwf_complex[2];
typedef struct fftw_plan_s *fftw_plan;
typedef struct fftwf_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_MEASURE (0U)
fftw_plan fftw_plan_dft_r2c(int, const int *, double *, fftw_complex *, unsigned);
fftw_plan fftw_plan_dft_c2r(int, const int *, fftw_complex *, double *, unsigned);
fftw_plan fftw_plan_dft(int, const int *, fftw_complex *, fftw_complex *, int, unsigned);
fftwf_plan fftwf_plan_dft_r2c(int, const int *, float *, fftwf_complex *, unsigned);
fftwf_plan fftwf_plan_dft_c2r(int, const int *, fftwf_complex *, float *, unsigned);
fftwf_plan fftwf_plan_dft(int, const int *, fftwf_complex *, fftwf_complex *, int, unsigned);
void fftw_execute(const fftw_plan); void fftwf_execute(const fftwf_plan);
void fftw_execute_dft_r2c(const fftw_plan, double *, fftw_complex *); void fftw_execute_dft_c2r(const fftw_plan, fftw_complex *, double *);
void fftwf_execute_dft_r2c(const fftwf_plan, float *, fftwf_complex *); void fftwf_execute_dft_c2r(const fftwf_plan, fftwf_complex *, float *);
void fftw_destroy_plan(fftw_plan); void fftwf_destroy_plan(fftwf_plan);
void fftw_cleanup(void); void fftwf_cleanup(void);
int fftw_init_threads(void); int fftwf_init_threads(void);
void fftw_plan_with_nthreads(int); void fftwf_plan_with_nthreads(int);
void fftw_cleanup_threads(void); void fftwf_cleanup_threads(void);
void *fftw_malloc(size_t); void fftw_free(void *);
}
