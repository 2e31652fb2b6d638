/* TEST INFRASTRUCTURE (oracle build only): FFTW3-API stand-in, see fftw3.h.
 * Three passes of a dense N x N DFT (twiddles from the full-precision pi, as
 * FFTW itself would use).  O(N^4) per volume: a correct but slow substitute. */
#include "fftw3.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct oracle_fftw_plan {
	int n, howmany, idist, odist, sign;
	fftw_complex *in, *out;
	double *wre, *wim; /* [n*n]: w^(j*k) */
	double *t0re, *t0im, *t1re, *t1im;
};

void *fftw_malloc(size_t n)
{
	void *p = NULL;
	if (posix_memalign(&p, 64, n ? n : 64) != 0) {
		return NULL;
	}
	return p;
}

void fftw_free(void *p)
{
	free(p);
}

fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed, int istride, int idist,
                             fftw_complex *out, const int *onembed, int ostride, int odist,
                             int sign, unsigned flags)
{
	(void)inembed; (void)onembed; (void)flags;
	if (rank != 3 || n[0] != n[1] || n[1] != n[2] || istride != 1 || ostride != 1) {
		fprintf(stderr, "fftw shim: unsupported plan\n");
		exit(EXIT_FAILURE);
	}
	struct oracle_fftw_plan *p = calloc(1, sizeof(*p));
	p->n = n[0]; p->howmany = howmany; p->idist = idist; p->odist = odist; p->sign = sign;
	p->in = in; p->out = out;
	int N = p->n;
	size_t vol = (size_t)N * N * N;
	p->wre = malloc(sizeof(double) * N * N);
	p->wim = malloc(sizeof(double) * N * N);
	const double pi = acos(-1.0);
	for (int j = 0; j < N; j++) {
		for (int k = 0; k < N; k++) {
			double a = 2.0 * pi * (double)((j * k) % N) / (double)N;
			p->wre[j * N + k] = cos(a);
			p->wim[j * N + k] = (sign < 0 ? -1.0 : 1.0) * sin(a);
		}
	}
	p->t0re = malloc(sizeof(double) * vol); p->t0im = malloc(sizeof(double) * vol);
	p->t1re = malloc(sizeof(double) * vol); p->t1im = malloc(sizeof(double) * vol);
	return p;
}

/* out[j][r] = sum_k w[j][k] in[k][r] for r in [0,len): transform over the leading
 * axis of a [N][len] block */
static void dft_leading(const struct oracle_fftw_plan *p, int len,
                        const double *ire, const double *iim, double *ore, double *oim)
{
	int N = p->n;
	for (int j = 0; j < N; j++) {
		double *orow = ore + (size_t)j * len, *oimrow = oim + (size_t)j * len;
		memset(orow, 0, sizeof(double) * len);
		memset(oimrow, 0, sizeof(double) * len);
		for (int k = 0; k < N; k++) {
			const double wr = p->wre[j * N + k], wi = p->wim[j * N + k];
			const double *a = ire + (size_t)k * len, *b = iim + (size_t)k * len;
			for (int r = 0; r < len; r++) {
				orow[r] += wr * a[r] - wi * b[r];
				oimrow[r] += wr * b[r] + wi * a[r];
			}
		}
	}
}

void fftw_execute(const fftw_plan p)
{
	int N = p->n;
	size_t N2 = (size_t)N * N, vol = N2 * N;
	for (int h = 0; h < p->howmany; h++) {
		const fftw_complex *in = p->in + (size_t)h * p->idist;
		fftw_complex *out = p->out + (size_t)h * p->odist;
		for (size_t i = 0; i < vol; i++) {
			p->t0re[i] = creal(in[i]);
			p->t0im[i] = cimag(in[i]);
		}
		/* axis 0 */
		dft_leading(p, (int)N2, p->t0re, p->t0im, p->t1re, p->t1im);
		/* axis 1: for each i0, block [N][N] */
		for (int i0 = 0; i0 < N; i0++) {
			dft_leading(p, N, p->t1re + i0 * N2, p->t1im + i0 * N2, p->t0re + i0 * N2, p->t0im + i0 * N2);
		}
		/* axis 2: per row of length N */
		for (size_t row = 0; row < N2; row++) {
			const double *a = p->t0re + row * N, *b = p->t0im + row * N;
			for (int j = 0; j < N; j++) {
				double sr = 0.0, si = 0.0;
				const double *wr = p->wre + j * N, *wi = p->wim + j * N;
				for (int k = 0; k < N; k++) {
					sr += wr[k] * a[k] - wi[k] * b[k];
					si += wr[k] * b[k] + wi[k] * a[k];
				}
				out[row * N + j] = sr + si * I;
			}
		}
	}
}

void fftw_destroy_plan(fftw_plan p)
{
	if (p != NULL) {
		free(p->wre); free(p->wim);
		free(p->t0re); free(p->t0im); free(p->t1re); free(p->t1im);
		free(p);
	}
}
