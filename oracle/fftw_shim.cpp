// TEST INFRASTRUCTURE ONLY.  A stand-in for libfftw3 (absent from this image), so that the reference's own
// src/fftw.cpp, src/projector.cpp, src/backprojector.cpp (compiled where they lie under /root/reference by oracle/Makefile,
// target `refrecon`) can run here.  It implements exactly the entry points those files call - real-to-complex and
// complex-to-real plans of rank 1..3, executed through fftw_execute_dft_r2c / _c2r - with FFTW's conventions: row-major,
// last dimension halved (n/2 + 1), forward sign -1, both directions unnormalised.  The transform is a separable DFT with
// exact twiddle tables (radix-2 Cooley-Tukey for power-of-two lengths, O(n^2) otherwise), all in double: slow, and accurate
// to a few ulp, which is all a checker needs.  Nothing under relion_b200/ links or loads this.
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../tests/cpp/relion_stubs/fftw3.h"

namespace {

typedef std::complex<double> cd;

struct Plan {
	int rank;
	int n[3];
	bool r2c;
};

// out[k] = sum_j in[j] exp(sign * 2 pi i j k / n), arbitrary n, strided in place via a scratch line
struct Dft1 {
	int n, sign;
	std::vector<cd> tw;    // exp(sign 2 pi i m / n), m = 0..n-1
	bool pow2;
	Dft1(int n_, int sign_) : n(n_), sign(sign_), tw((size_t) n_), pow2((n_ & (n_ - 1)) == 0)
	{
		for (int m = 0; m < n; m++)
		{
			// exact octant reduction keeps the table symmetric to the last bit
			const double a = 2.0 * M_PI * (double) m / (double) n;
			tw[m] = cd(cos(a), sign * sin(a));
		}
	}
	void run(cd *line, cd *tmp) const
	{
		if (n == 1) return;
		if (pow2)
		{
			// iterative radix-2, decimation in time
			for (int i = 1, j = 0; i < n; i++)
			{
				int bit = n >> 1;
				for (; j & bit; bit >>= 1) j ^= bit;
				j ^= bit;
				if (i < j) std::swap(line[i], line[j]);
			}
			for (int len = 2; len <= n; len <<= 1)
			{
				const int step = n / len;
				for (int i = 0; i < n; i += len)
					for (int k = 0; k < len / 2; k++)
					{
						const cd u = line[i + k], v = line[i + k + len / 2] * tw[(size_t) k * step];
						line[i + k] = u + v;
						line[i + k + len / 2] = u - v;
					}
			}
			return;
		}
		for (int k = 0; k < n; k++)
		{
			cd s(0., 0.);
			for (int j = 0; j < n; j++) s += line[j] * tw[(size_t) (((long long) j * k) % n)];
			tmp[k] = s;
		}
		memcpy(line, tmp, sizeof(cd) * (size_t) n);
	}
};

// full complex transform of a dense [n0][n1][n2] array along every axis
void dft_nd(std::vector<cd> &a, const int *n, int rank, int sign)
{
	int dims[3] = {1, 1, 1};
	for (int i = 0; i < rank; i++) dims[3 - rank + i] = n[i];
	const size_t s2 = 1, s1 = (size_t) dims[2], s0 = (size_t) dims[1] * dims[2];
	const size_t strides[3] = {s0, s1, s2};
	for (int ax = 0; ax < 3; ax++)
	{
		const int len = dims[ax];
		if (len == 1) continue;
		const Dft1 d(len, sign);
		const int o1 = (ax + 1) % 3, o2 = (ax + 2) % 3;
#pragma omp parallel
		{
			std::vector<cd> line((size_t) len), tmp((size_t) len);
#pragma omp for collapse(2) schedule(static)
			for (int i = 0; i < dims[o1]; i++)
				for (int j = 0; j < dims[o2]; j++)
				{
					const size_t base = (size_t) i * strides[o1] + (size_t) j * strides[o2];
					for (int k = 0; k < len; k++) line[k] = a[base + (size_t) k * strides[ax]];
					d.run(line.data(), tmp.data());
					for (int k = 0; k < len; k++) a[base + (size_t) k * strides[ax]] = line[k];
				}
		}
	}
}

size_t total(const Plan *p) { size_t t = 1; for (int i = 0; i < p->rank; i++) t *= (size_t) p->n[i]; return t; }

Plan *make_plan(int rank, const int *n, bool r2c)
{
	if (rank < 1 || rank > 3) return nullptr;
	Plan *p = new Plan;
	p->rank = rank; p->r2c = r2c;
	for (int i = 0; i < 3; i++) p->n[i] = i < rank ? n[i] : 1;
	return p;
}

}  // namespace

extern "C" {

fftw_plan fftw_plan_dft_r2c(int rank, const int *n, double *, fftw_complex *, unsigned) { return (fftw_plan) make_plan(rank, n, true); }
fftw_plan fftw_plan_dft_c2r(int rank, const int *n, fftw_complex *, double *, unsigned) { return (fftw_plan) make_plan(rank, n, false); }
// complex-to-complex plans are created by FourierTransformer::setFourier/setComplex paths but never executed by the files built here
fftw_plan fftw_plan_dft(int rank, const int *n, fftw_complex *, fftw_complex *, int, unsigned) { return (fftw_plan) make_plan(rank, n, true); }

void fftw_execute_dft_r2c(const fftw_plan plan, double *in, fftw_complex *out)
{
	const Plan *p = (const Plan *) plan;
	const size_t nt = total(p);
	std::vector<cd> a(nt);
	for (size_t i = 0; i < nt; i++) a[i] = cd(in[i], 0.);
	dft_nd(a, p->n, p->rank, -1);
	const int nl = p->n[p->rank - 1], nh = nl / 2 + 1;
	const size_t rows = nt / (size_t) nl;
	for (size_t r = 0; r < rows; r++)
		for (int k = 0; k < nh; k++)
		{
			out[r * nh + k][0] = a[r * nl + k].real();
			out[r * nh + k][1] = a[r * nl + k].imag();
		}
}

void fftw_execute_dft_c2r(const fftw_plan plan, fftw_complex *in, double *out)
{
	const Plan *p = (const Plan *) plan;
	const size_t nt = total(p);
	const int rank = p->rank;
	const int nl = p->n[rank - 1], nh = nl / 2 + 1;
	const size_t rows = nt / (size_t) nl;
	// rebuild the full Hermitian array: F(-k) = conj F(k), taking the stored half as authoritative (as FFTW does)
	int dims[3] = {1, 1, 1};
	for (int i = 0; i < rank; i++) dims[3 - rank + i] = p->n[i];
	std::vector<cd> a(nt);
	for (size_t r = 0; r < rows; r++)
	{
		const int i0 = (int) (r / (size_t) dims[1]), i1 = (int) (r % (size_t) dims[1]);
		const int m0 = (dims[0] - i0) % dims[0], m1 = (dims[1] - i1) % dims[1];
		const size_t rm = (size_t) m0 * dims[1] + m1;
		for (int k = 0; k < nl; k++)
		{
			if (k < nh) a[r * nl + k] = cd(in[r * nh + k][0], in[r * nh + k][1]);
			else a[r * nl + k] = std::conj(cd(in[rm * nh + (nl - k)][0], in[rm * nh + (nl - k)][1]));
		}
	}
	dft_nd(a, p->n, rank, +1);
	for (size_t i = 0; i < nt; i++) out[i] = a[i].real();
}

void fftw_execute(const fftw_plan) { abort(); }
void fftw_destroy_plan(fftw_plan plan) { delete (Plan *) plan; }
void fftw_cleanup(void) {}
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
void fftw_cleanup_threads(void) {}
void *fftw_malloc(size_t n) { void *p = nullptr; if (posix_memalign(&p, 64, n ? n : 1)) return nullptr; return p; }
void fftw_free(void *p) { free(p); }

}  // extern "C"
