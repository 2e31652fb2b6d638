/* TEST INFRASTRUCTURE (oracle build only): FFTW3-API stand-in, see fftw3.h.
 *
 * Row-column 3-D DFT for odd N.  Each 1-D pass uses the even/odd split of the DFT matrix,
 *     out[j]   = E_j - i s O_j,   out[N-j] = E_j + i s O_j,   j = 1..(N-1)/2,   s = -sign,
 *     E_j = in[0] + sum_k cos(2 pi jk/N) (in[k] + in[N-k]),   O_j = sum_k sin(2 pi jk/N) (in[k] - in[N-k]),
 * i.e. (N-1)^2/2 real-by-complex multiply-adds per line — a quarter of the dense product and in the
 * range of what FFTW spends on these prime lengths (31, 41, 61) — so that the timed CPU baseline is not
 * handicapped by the stand-in.  Twiddles use the full-precision pi, as FFTW does.  Hot loops get an
 * AVX2 clone selected by the dynamic linker (the GPU box's host CPU is not known at build time). */
#include "fftw3.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#define SHIM_CLONES __attribute__((target_clones("arch=x86-64-v3", "default")))
#else
#define SHIM_CLONES
#endif

struct oracle_fftw_plan {
	int n, h, howmany, idist, odist, sign;
	fftw_complex *in, *out;
	double *c, *s; /* [h+1][h+1]: cos / sin of 2 pi j k / n */
	double *t0re, *t0im, *t1re, *t1im;
	double *ure, *uim, *vre, *vim; /* [h+1][n*n] scratch for the symmetric / antisymmetric parts */
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
	if (rank != 3 || n[0] != n[1] || n[1] != n[2] || istride != 1 || ostride != 1 || (n[0] % 2) == 0) {
		fprintf(stderr, "fftw shim: unsupported plan\n");
		exit(EXIT_FAILURE);
	}
	struct oracle_fftw_plan *p = calloc(1, sizeof(*p));
	p->n = n[0]; p->h = (n[0] - 1) / 2; p->howmany = howmany; p->idist = idist; p->odist = odist; p->sign = sign;
	p->in = in; p->out = out;
	const int N = p->n, h = p->h;
	const size_t vol = (size_t)N * N * N, plane = (size_t)N * N;
	p->c = malloc(sizeof(double) * (h + 1) * (h + 1));
	p->s = malloc(sizeof(double) * (h + 1) * (h + 1));
	const double pi = acos(-1.0);
	for (int j = 0; j <= h; j++) {
		for (int k = 0; k <= h; k++) {
			double a = 2.0 * pi * (double)((j * k) % N) / (double)N;
			p->c[j * (h + 1) + k] = cos(a);
			p->s[j * (h + 1) + k] = sin(a);
		}
	}
	p->t0re = malloc(sizeof(double) * vol); p->t0im = malloc(sizeof(double) * vol);
	p->t1re = malloc(sizeof(double) * vol); p->t1im = malloc(sizeof(double) * vol);
	p->ure = malloc(sizeof(double) * (h + 1) * plane); p->uim = malloc(sizeof(double) * (h + 1) * plane);
	p->vre = malloc(sizeof(double) * (h + 1) * plane); p->vim = malloc(sizeof(double) * (h + 1) * plane);
	return p;
}

/* 1-D DFT over the leading axis of a [N][len] block (split re/im), len contiguous */
SHIM_CLONES
static void dft_leading(const struct oracle_fftw_plan *p, int len,
                        const double *restrict ire, const double *restrict iim,
                        double *restrict ore, double *restrict oim)
{
	const int N = p->n, h = p->h;
	const double sg = (p->sign < 0) ? 1.0 : -1.0; /* forward: out[j] = E - iO */
	double *ure = p->ure, *uim = p->uim, *vre = p->vre, *vim = p->vim;
	for (int k = 1; k <= h; k++) {
		const double *a = ire + (size_t)k * len, *b = iim + (size_t)k * len;
		const double *c = ire + (size_t)(N - k) * len, *d = iim + (size_t)(N - k) * len;
		double *ur = ure + (size_t)k * len, *ui = uim + (size_t)k * len, *vr = vre + (size_t)k * len, *vi = vim + (size_t)k * len;
		for (int r = 0; r < len; r++) {
			ur[r] = a[r] + c[r]; ui[r] = b[r] + d[r];
			vr[r] = a[r] - c[r]; vi[r] = b[r] - d[r];
		}
	}
	/* j = 0 */
	for (int r = 0; r < len; r++) {
		double sr = ire[r], si = iim[r];
		for (int k = 1; k <= h; k++) {
			sr += ure[(size_t)k * len + r];
			si += uim[(size_t)k * len + r];
		}
		ore[r] = sr; oim[r] = si;
	}
	for (int j = 1; j <= h; j++) {
		double *o1r = ore + (size_t)j * len, *o1i = oim + (size_t)j * len;
		double *o2r = ore + (size_t)(N - j) * len, *o2i = oim + (size_t)(N - j) * len;
		/* accumulate E in o1, O in o2 */
		for (int r = 0; r < len; r++) {
			o1r[r] = ire[r]; o1i[r] = iim[r];
			o2r[r] = 0.0; o2i[r] = 0.0;
		}
		for (int k = 1; k <= h; k++) {
			const double cc = p->c[j * (h + 1) + k], ss = p->s[j * (h + 1) + k];
			const double *ur = ure + (size_t)k * len, *ui = uim + (size_t)k * len;
			const double *vr = vre + (size_t)k * len, *vi = vim + (size_t)k * len;
			for (int r = 0; r < len; r++) {
				o1r[r] += cc * ur[r]; o1i[r] += cc * ui[r];
				o2r[r] += ss * vr[r]; o2i[r] += ss * vi[r];
			}
		}
		/* out[j] = E - i sg O = (Er + sg Oi, Ei - sg Or); out[N-j] = E + i sg O */
		for (int r = 0; r < len; r++) {
			const double er = o1r[r], ei = o1i[r], orr = o2r[r], oi = o2i[r];
			o1r[r] = er + sg * oi; o1i[r] = ei - sg * orr;
			o2r[r] = er - sg * oi; o2i[r] = ei + sg * orr;
		}
	}
}

/* transpose the last two axes of every [N][N] plane so that the innermost axis becomes a leading one */
SHIM_CLONES
static void transpose_planes(int N, const double *restrict in, double *restrict out)
{
	for (int i0 = 0; i0 < N; i0++) {
		const double *a = in + (size_t)i0 * N * N;
		double *b = out + (size_t)i0 * N * N;
		for (int i1 = 0; i1 < N; i1++) {
			for (int i2 = 0; i2 < N; i2++) {
				b[(size_t)i2 * N + i1] = a[(size_t)i1 * N + i2];
			}
		}
	}
}

void fftw_execute(const fftw_plan p)
{
	const int N = p->n;
	const size_t N2 = (size_t)N * N, vol = N2 * N;
	for (int hh = 0; hh < p->howmany; hh++) {
		const fftw_complex *in = p->in + (size_t)hh * p->idist;
		fftw_complex *out = p->out + (size_t)hh * p->odist;
		for (size_t i = 0; i < vol; i++) {
			p->t0re[i] = creal(in[i]);
			p->t0im[i] = cimag(in[i]);
		}
		/* axis 0 */
		dft_leading(p, (int)N2, p->t0re, p->t0im, p->t1re, p->t1im);
		/* axis 1: within each i0 plane the leading axis is i1 */
		for (int i0 = 0; i0 < N; i0++) {
			dft_leading(p, N, p->t1re + i0 * N2, p->t1im + i0 * N2, p->t0re + i0 * N2, p->t0im + i0 * N2);
		}
		/* axis 2: transpose planes, transform the (now leading) axis, transpose back */
		transpose_planes(N, p->t0re, p->t1re);
		transpose_planes(N, p->t0im, p->t1im);
		for (int i0 = 0; i0 < N; i0++) {
			dft_leading(p, N, p->t1re + i0 * N2, p->t1im + i0 * N2, p->t0re + i0 * N2, p->t0im + i0 * N2);
		}
		for (int i0 = 0; i0 < N; i0++) {
			for (int i1 = 0; i1 < N; i1++) {
				for (int i2 = 0; i2 < N; i2++) {
					const size_t src = (size_t)i0 * N2 + (size_t)i2 * N + i1; /* [i0][i2][i1] after the transpose */
					out[(size_t)i0 * N2 + (size_t)i1 * N + i2] = p->t0re[src] + p->t0im[src] * I;
				}
			}
		}
	}
}

void fftw_destroy_plan(fftw_plan p)
{
	if (p != NULL) {
		free(p->c); free(p->s);
		free(p->t0re); free(p->t0im); free(p->t1re); free(p->t1im);
		free(p->ure); free(p->uim); free(p->vre); free(p->vim);
		free(p);
	}
}
