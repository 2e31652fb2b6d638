/* TEST INFRASTRUCTURE — plain-C restatement of the reference's FFT-SAXS scoring path (the oracle "port").
 *
 * Independent of oracle/_ref (it needs neither /root/reference nor GMP/FFTW) and of the product's
 * factorised formulation: this file keeps the REFERENCE's own structure — per z the translation matrix,
 * per (z, beta2) the ligand pre-sum, per cell the l-contraction into sum2[m][m1][m2] and a direct 3-D DFT
 * evaluated at the listed grid points only (the reference's simple_ft branch; its FFTW branch computes the
 * same transform).  Every function cites the reference lines it follows.  Pinned by tests/test_cpu_port.py
 * against the reference's goldens and against the compiled reference.
 *
 * Build: make -C oracle port   ->  oracle/liboracle_port.so   (gcc -std=c11 -O2 -ffp-contract=off, libquadmath)
 */
#include <math.h>
#include <quadmath.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PORT_PI 3.14159265358 /* src/define.h:13-15: what M_PI is under -std=c11 */

#define SXS_HD static inline
#include "port_fit.h"

static inline int lm_index(int l, int m) { return l * (l + 1) + m; }            /* src/common.h:49-51 */
static inline int index_LM(int N, int l, int m) { return N * (m + N - 1) + l; } /* src/borrowed.h:17-20 */
static inline int index_d(int L, int l, int m, int m1) { return (2 * L + 1) * ((2 * L + 1) * l + (L + m)) + (L + m1); }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------ src/sfbessel.c:40-60 */
double port_sbessel(int l, double x)
{
	if (x > 0.0) {
		double tlp1 = 2.0 * (double)l + 1.0;
		double dfact = 1.0;
		for (int i = 1; i <= (int)tlp1; i += 2) {
			dfact *= (double)i; /* doublefact(), src/sfbessel.c:22-37 */
		}
		double summand = 1.0 / dfact;
		double summ = summand;
		for (int k = 1; fabs(summand / summ) > 0.00001; k++) {
			summand *= (-1.0) * x * x / (2.0 * (double)k * (2.0 * (double)k + tlp1));
			summ += summand;
		}
		return summ * pow(x, l);
	}
	return l == 0 ? 1.0 : 0.0;
}

/* ------------------------------------------------------------------ src/saxs_utils.c:50-63 */
void port_mkarray(double begin, double end, int qnum, double *q)
{
	double step = (end - begin) / (qnum - 1);
	q[0] = begin;
	for (int i = 1; i < qnum; i++) {
		q[i] = q[i - 1] + step;
	}
}

/* ------------------------------------------------------------------ src/borrowed.c:19-65 */
static void spherical_norm(int N, double *Y)
{
	for (int l = 0; l < N; ++l) {
		double fact = 1.;
		double tmp0 = (2 * l + 1.) / (4. * PORT_PI);
		for (int m = 0; m <= l; ++m) {
			if (m > 0) fact *= (l - m + 1) * (l + m);
			double tmp = sqrt(tmp0 / fact);
			double sgn = (m % 2 == 0) ? (1) : (-1);
			Y[index_LM(N, l, m)] = tmp;
			Y[index_LM(N, l, -m)] = sgn * tmp;
		}
	}
}

static void assoc_legendre(int N, double x, double *P)
{
	double y = sqrt(1. - x * x);
	P[index_LM(N, 0, 0)] = 1;
	if (N == 1) return;
	P[index_LM(N, 1, 0)] = x;
	for (int l = 2; l < N; ++l) {
		P[index_LM(N, l, 0)] = (2. * l - 1.) / l * x * P[index_LM(N, l - 1, 0)] - (l - 1.) / l * P[index_LM(N, l - 2, 0)];
	}
	for (int m = 1; m < N - 1; ++m) {
		P[index_LM(N, m, m)] = -1. * (2. * m - 1.) * y * P[index_LM(N, m - 1, m - 1)];
		P[index_LM(N, m + 1, m)] = (2. * m + 1.) * x * P[index_LM(N, m, m)];
		for (int l = m + 2; l < N; ++l) {
			P[index_LM(N, l, m)] = (2. * l - 1.) / (l - m) * x * P[index_LM(N, l - 1, m)] - (l + m - 1.) / (l - m) * P[index_LM(N, l - 2, m)];
		}
	}
	P[index_LM(N, N - 1, N - 1)] = -1. * (2. * (N - 1.) - 1.) * y * P[index_LM(N, N - 2, N - 2)];
}

/* ------------------------------------------------------------------ src/pdb2spf.c:9-152
 * Per-atom inputs are coordinates and the three q-independent form factors (vacuum, dummy, h2o*SASA).
 * coef[((c*qnum + q)*(L+1)^2 + lm)*2 + {re,im}] */
void port_expand(int natoms, const double *xyz, const double *fv, const double *fd, const double *fw,
                 const double *qvals, int qnum, int L, double *coef)
{
	const int nb = L + 1, lm_n = nb * nb, leg = nb * (2 * L + 1);
	double *ynorm = calloc(nb * (2 * nb - 1), sizeof(double));
	double *P = calloc(leg, sizeof(double));
	double *hre = calloc(lm_n, sizeof(double)), *him = calloc(lm_n, sizeof(double));
	spherical_norm(nb, ynorm);
	memset(coef, 0, sizeof(double) * 3 * qnum * lm_n * 2);
	for (int a = 0; a < natoms; a++) {
		const double x = xyz[3 * a], y = xyz[3 * a + 1], z = xyz[3 * a + 2];
		const double r = sqrt(x * x + y * y + z * z);          /* cart2sph, :9-22 */
		const double theta = acos(z / r);
		double fi;
		if (y > 0.0) fi = acos(x / sqrt(x * x + y * y));
		else fi = -acos(x / sqrt(x * x + y * y)) + 2.0 * PORT_PI;
		memset(P, 0, sizeof(double) * leg);
		assoc_legendre(nb, cos(theta), P);
		for (int l = 0; l <= L; l++) {
			for (int m = -l; m <= l; m++) {
				double h = ynorm[index_LM(nb, l, m)] * P[index_LM(nb, l, abs(m))];
				hre[lm_index(l, m)] = h * cos(fi * m);
				him[lm_index(l, m)] = h * sin(-fi * m);
			}
		}
		for (int q = 0; q < qnum; q++) {
			double *V = coef + ((size_t)(0 * qnum + q) * lm_n) * 2, *D = coef + ((size_t)(1 * qnum + q) * lm_n) * 2,
			       *W = coef + ((size_t)(2 * qnum + q) * lm_n) * 2;
			for (int l = 0; l <= L; l++) {
				const double bes = port_sbessel(l, qvals[q] * r);
				for (int m = -l; m <= l; m++) {
					const int i = lm_index(l, m);
					const double vr = bes * hre[i], vi = bes * him[i];
					V[2 * i] += fv[a] * vr; V[2 * i + 1] += fv[a] * vi;
					D[2 * i] += fd[a] * vr; D[2 * i + 1] += fd[a] * vi;
					W[2 * i] += fw[a] * vr; W[2 * i + 1] += fw[a] * vi;
				}
			}
		}
	}
	const double four_pi = 4.0 * PORT_PI;                       /* times 4 pi i^l, :55-57,118-147 */
	const double cre[4] = {four_pi, 0.0, -four_pi, 0.0}, cim[4] = {0.0, four_pi, 0.0, -four_pi};
	for (int c = 0; c < 3; c++) {
		for (int q = 0; q < qnum; q++) {
			double *A = coef + ((size_t)(c * qnum + q) * lm_n) * 2;
			for (int l = 0; l <= L; l++) {
				for (int m = -l; m <= l; m++) {
					const int i = lm_index(l, m);
					const double re = A[2 * i], im = A[2 * i + 1];
					A[2 * i] = cre[l % 4] * re - cim[l % 4] * im;
					A[2 * i + 1] = cre[l % 4] * im + cim[l % 4] * re;
				}
			}
		}
	}
	free(ynorm); free(P); free(hre); free(him);
}

/* ------------------------------------------------------------------ src/borrowed.c:243-313 */
void port_wigner_d(int L, double beta, double *D)
{
	memset(D, 0, sizeof(double) * (L + 1) * (2 * L + 1) * (2 * L + 1));
#define DW(l, m, m1) D[index_d(L, (l), (m), (m1))]
	DW(0, 0, 0) = 1.0;
	DW(1, -1, -1) = 1.0 * (1.0 + cos(beta)) / 2.0;
	DW(1, -1, 0) = 1.0 * sin(beta) / sqrt(2.0);
	DW(1, -1, 1) = 1.0 * (1.0 - cos(beta)) / 2.0;
	DW(1, 0, -1) = -1.0 * sin(beta) / sqrt(2.0);
	DW(1, 0, 0) = 1.0 * cos(beta);
	DW(1, 0, 1) = 1.0 * sin(beta) / sqrt(2.0);
	DW(1, 1, -1) = 1.0 * (1.0 - cos(beta)) / 2.0;
	DW(1, 1, 0) = -1.0 * sin(beta) / sqrt(2.0);
	DW(1, 1, 1) = 1.0 * (1.0 + cos(beta)) / 2.0;
	for (int l = 2; l <= L; ++l) {
		for (int m = -l; m <= l; ++m) {
			double fact, d1 = 1, d2 = 1, d3 = 1, d4 = 1;
			for (int k = 1; k <= 2 * l; ++k) {
				fact = sqrt(k / (((k <= l + m) ? k : 1.0) * ((k <= l - m) ? k : 1.0)));
				d1 *= fact * ((k <= l + m) ? cos(beta / 2.0) : 1.0) * ((k <= l - m) ? -sin(beta / 2.0) : 1.0);
				d2 *= fact * ((k <= l - m) ? cos(beta / 2.0) : 1.0) * ((k <= l + m) ? sin(beta / 2.0) : 1.0);
				d3 *= fact * ((k <= l + m) ? cos(beta / 2.0) : 1.0) * ((k <= l - m) ? sin(beta / 2.0) : 1.0);
				d4 *= fact * ((k <= l - m) ? cos(beta / 2.0) : 1.0) * ((k <= l + m) ? -sin(beta / 2.0) : 1.0);
			}
			DW(l, l, m) = d1; DW(l, -l, m) = d2; DW(l, m, l) = d3; DW(l, m, -l) = d4;
			for (int m1 = -l; m1 <= l; ++m1) {
				int j = l - 1;
				if ((m1 > -l) && (m1 < l) && (m > -l) && (m < l)) {
					double val = 1.0, val1 = -1.0;
					val *= (j + 1) * (2.0 * j + 1) / sqrt(((j + 1) * (j + 1) - m * m) * ((j + 1) * (j + 1) - m1 * m1));
					val *= (cos(beta) - (double)(m * m1) / (double)(j * (j + 1))) * DW(j, m, m1);
					val1 *= sqrt((j * j - m * m) * (j * j - m1 * m1));
					val1 *= (double)(j + 1.0) * DW(j - 1, m, m1) / (double)j;
					val1 /= sqrt(((j + 1.0) * (j + 1.0) - (double)(m * m)) * ((j + 1.0) * (j + 1.0) - (double)(m1 * m1)));
					val += val1;
					DW(l, m, m1) = val;
				}
			}
		}
	}
#undef DW
}

/* ------------------------------------------------------------------ src/borrowed.c:86-222 (Racah sum; the reference
 * evaluates it in GMP floats, here IEEE binary128: error < 1e-26 up to L = 40) */
double port_wigner_3j(int j1, int j2, int j3, int m1, int m2, int m3)
{
	static __float128 f[401];
	static int ready = 0;
	if (!ready) {
		f[0] = 1;
		for (int i = 1; i <= 400; i++) f[i] = f[i - 1] * i;
		ready = 1;
	}
	int t1 = j2 - m1 - j3, t2 = j1 + m2 - j3, t3 = j1 + j2 - j3, t4 = j1 - m1, t5 = j2 + m2;
	int tmin = imax(0, imax(t1, t2)), tmax = imin(t3, imin(t4, t5));
	__float128 w = 0;
	for (int t = tmin; t <= tmax; ++t) {
		__float128 den = f[t4 - t] * f[t5 - t] * f[t3 - t] * f[t - t2] * f[t - t1] * f[t];
		w += ((t % 2 == 0) ? (__float128)1 : (__float128)-1) / den;
	}
	__float128 b = f[j3 + m3] * f[j3 - m3] * f[j2 - m2] * f[j2 + m2] * f[j1 - m1] * f[j1 + m1] / f[j1 + j2 + j3 + 1] *
	               f[-j1 + j2 + j3] * f[j1 - j2 + j3] * f[j1 + j2 - j3];
	w *= sqrtq(b);
	if ((j1 - j2 - m3) % 2 != 0) w = -w;
	return (double)w;
}

/* ------------------------------------------------------------------ src/fftsaxs.c:3-6,138-174 */
void port_dsymb(int L, double *d_symb)
{
	const int nb = L + 1, N = 2 * L + 1;
	memset(d_symb, 0, sizeof(double) * nb * nb * nb * N);
	for (int l = 0; l < nb; l++) {
		double k1 = 2 * l + 1;
		for (int l1 = 0; l1 < nb; l1++) {
			double k2 = sqrt((2 * l1 + 1) * k1);
			for (int p = abs(l - l1); p <= l + l1; p++) {
				double k3 = (2 * p + 1) * k2 * port_wigner_3j(l, p, l1, 0, 0, 0);
				for (int m = -imin(l, l1); m <= imin(l, l1); m++) {
					d_symb[((size_t)lm_index(l, m) * nb + l1) * N + p] = k3 * port_wigner_3j(l, p, l1, -m, 0, m);
				}
			}
		}
	}
}

/* ------------------------------------------------------------------ src/min_saxs.c:353-389,116-124 */
void port_opt_params(const double *eq, const double *ei, const double *ee, int en, const double *qvals, int qnum,
                     double rm, double *a, double *scal)
{
	memset(a, 0, sizeof(double) * 6 * qnum);
	int j = 0;
	for (int i = 0; i < qnum; i++) {
		double lower = qvals[0], upper = qvals[i];
		if (i > 0) lower = qvals[i - 1];
		while (j < en && eq[j] >= lower && eq[j] <= upper) { /* the reference's scan is unguarded (:369) */
			a[i * 6 + 0] += ei[j] * ei[j] / (ee[j] * ee[j]);
			a[i * 6 + 1] += ei[j] / (ee[j] * ee[j]);
			a[i * 6 + 2] += eq[j] * ei[j] / (ee[j] * ee[j]);
			a[i * 6 + 3] += 1.0 / (ee[j] * ee[j]);
			a[i * 6 + 4] += eq[j] / (ee[j] * ee[j]);
			a[i * 6 + 5] += eq[j] * eq[j] / (ee[j] * ee[j]);
			j++;
		}
	}
	double norm = (double)j;
	for (int i = 0; i < 6 * qnum; i++) a[i] /= norm;
	scal[0] = rm;
	scal[1] = pow(4.0 * PORT_PI / 3.0, 1.5) * rm * rm / (16.0 * PORT_PI);
	scal[2] = ei[0];
}

/* ------------------------------------------------------------------ src/min_saxs.c:153-259 on explicit cross terms
 * x[npts][6][qnum] -> out[npts][4] = chi, c1, c2, evaluations */
void port_fit(const double *x, int npts, const double *a, const double *scal, const double *qvals, int qnum, double *out)
{
	double *row = malloc(sizeof(double) * 6 * qnum);
	for (int p = 0; p < npts; p++) {
		for (int c = 0; c < 6; c++)
			for (int q = 0; q < qnum; q++) row[q * 6 + c] = x[((size_t)p * 6 + c) * qnum + q];
		double s, c1, c2;
		int nfg;
		port_fit_point(row, 1, 6, a, qvals, qnum, scal[1], scal[2], &s, &c1, &c2, &nfg);
		out[4 * p] = s; out[4 * p + 1] = c1; out[4 * p + 2] = c2; out[4 * p + 3] = nfg;
	}
	free(row);
}

/* ------------------------------------------------------------------ src/fftsaxs.c:608-986, skip = 1 semantics.
 * Also returns, when cross != NULL, the six cross terms of every row before the peak rescale:
 * cross[(row*6 + k)*qnum + q] (what fill_const/fill_var leave in the profile). */
void port_scores(double *scores, double *c1, double *c2, const long long *index, int nout, const double *coefA,
                 const double *coefB, const double *a, const double *scal, const double *qvals, int qnum,
                 const double *zvals, int znum, int L, double *cross)
{
	const int nb = L + 1, N = 2 * L + 1, lm_n = nb * nb;
	const long long n3 = (long long)N * N * N, size5d = n3 * nb * nb;
	const double beta_step = (PORT_PI - 0.0) / (nb - 1);          /* :632-636 */
#define CA(c, q, i, ri) coefA[(((size_t)(c) * qnum + (q)) * lm_n + (i)) * 2 + (ri)]
#define CB(c, q, i, ri) coefB[(((size_t)(c) * qnum + (q)) * lm_n + (i)) * 2 + (ri)]

	/* comp_const_int x6, :27-50,638-643; fill_const doubles VD, VW, DW (:76-81) */
	static const int P1[6] = {0, 0, 0, 1, 1, 2}, P2[6] = {0, 1, 2, 1, 2, 2};
	double *cst = calloc((size_t)6 * qnum, sizeof(double));
	for (int k = 0; k < 6; k++) {
		for (int q = 0; q < qnum; q++) {
			double in = 0.0;
			for (int i = 0; i < lm_n; i++) {
				in += CA(P1[k], q, i, 0) * CA(P2[k], q, i, 0) + CA(P1[k], q, i, 1) * CA(P2[k], q, i, 1);
				in += CB(P1[k], q, i, 0) * CB(P2[k], q, i, 0) + CB(P1[k], q, i, 1) * CB(P2[k], q, i, 1);
			}
			cst[k * qnum + q] = (P1[k] != P2[k]) ? in * 2.0 : in;
		}
	}

	double *d_symb = malloc(sizeof(double) * nb * nb * nb * N);
	port_dsymb(L, d_symb);
	double *dw = malloc(sizeof(double) * (size_t)nb * nb * N * N);
	for (int i = 0; i < nb; i++) port_wigner_d(L, 0.0 + i * beta_step, dw + (size_t)i * nb * N * N);

	double *bessel = malloc(sizeof(double) * qnum * N);
	double *t_re = calloc((size_t)qnum * nb * nb * nb, sizeof(double)), *t_im = calloc((size_t)qnum * nb * nb * nb, sizeof(double));
	const size_t s1n = (size_t)qnum * N * N * nb;
	double *s1 = calloc(s1n * 6, sizeof(double));       /* sum1 for the current (z, beta2): [c][re|im][q][m][m2][l] */
	double *sum2 = malloc(sizeof(double) * 6 * n3 * 2);  /* [comp][m][m1][m2][re|im], negative indices wrapped */
	double *cross_buf = cross ? cross : calloc((size_t)nout * 6 * qnum, sizeof(double));
	double *e_re = malloc(sizeof(double) * N), *e_im = malloc(sizeof(double) * N);
	for (int i = 0; i < N; i++) {                         /* simple_ft phases, :536-548 */
		e_re[i] = cos(i * (2 * PORT_PI / N));
		e_im[i] = -sin(i * (2 * PORT_PI / N));
	}
	int *cell_rows = malloc(sizeof(int) * nout);

	for (int zi = 0; zi < znum; zi++) {
		int any = 0;
		for (int i = 0; i < nout; i++) any |= (index[i] >= 0 && index[i] / size5d == zi);
		if (!any) continue;
		/* besselj_n :110-121, fill_t_matrix :182-243 */
		for (int q = 0; q < qnum; q++)
			for (int p = 0; p < N; p++) bessel[q * N + p] = port_sbessel(p, zvals[zi] * qvals[q]);
		static const double ipr[4] = {1.0, 0.0, -1.0, 0.0}, ipi[4] = {0.0, 1.0, 0.0, -1.0};
		for (int q = 0; q < qnum; q++) {
			for (int m = 0; m <= L; m++) {
				const double sg = (m % 2) ? -1.0 : 1.0;
				for (int l = m; l <= L; l++) {
					for (int l1 = l; l1 <= L; l1++) {
						double re = 0.0, im = 0.0;
						const double *dr = d_symb + ((size_t)lm_index(l, m) * nb + l1) * N;
						for (int p = abs(l - l1); p <= l + l1; p++) {
							double val = sg * dr[p] * bessel[q * N + p];
							re += ipr[p % 4] * val;
							im += ipi[p % 4] * val;
						}
						size_t b = ((size_t)q * nb + m) * nb;
						t_re[(b + l) * nb + l1] = re; t_im[(b + l) * nb + l1] = im;
						t_re[(b + l1) * nb + l] = re; t_im[(b + l1) * nb + l] = im;
					}
				}
			}
		}
		for (int b2 = 0; b2 < nb; b2++) {
			any = 0;
			for (int i = 0; i < nout; i++)
				any |= (index[i] >= 0 && index[i] / size5d == zi && (index[i] / n3) % nb == b2);
			if (!any) continue;
			/* compute_sum1 x3, :251-333 */
			const double *d2 = dw + (size_t)b2 * nb * N * N;
			for (int c = 0; c < 3; c++) {
				double *sre = s1 + (size_t)(2 * c) * s1n, *sim = s1 + (size_t)(2 * c + 1) * s1n;
				for (int q = 0; q < qnum; q++) {
					for (int m = -L; m <= L; m++) {
						for (int m2 = -L; m2 <= L; m2++) {
							size_t so = (((size_t)q * N + m + L) * N + m2 + L) * nb;
							for (int l = abs(m); l <= L; l++) {
								double vr = 0.0, vi = 0.0;
								size_t to = (((size_t)q * nb + abs(m)) * nb + l) * nb;
								for (int l1 = imax(abs(m2), abs(m)); l1 <= L; l1++) {
									double dv = d2[index_d(L, l1, m, m2)];
									double br = CB(c, q, lm_index(l1, m2), 0), bi = CB(c, q, lm_index(l1, m2), 1);
									vr += dv * (br * t_re[to + l1] - bi * t_im[to + l1]);
									vi -= dv * (br * t_im[to + l1] + bi * t_re[to + l1]);
								}
								sre[so + l] = vr;
								sim[so + l] = vi;
							}
						}
					}
				}
			}
			for (int b1 = 0; b1 < nb; b1++) {
				int nrows = 0;
				for (int i = 0; i < nout; i++) {
					if (index[i] >= 0 && index[i] / size5d == zi && (index[i] / n3) % nb == b2 &&
					    (index[i] / (n3 * nb)) % nb == b1)
						cell_rows[nrows++] = i;
				}
				if (nrows == 0) continue;
				const double *d1 = dw + (size_t)b1 * nb * N * N;
				for (int q = 0; q < qnum; q++) {
					/* compute_sum2, :416-524 */
					for (int m = -L; m <= L; m++) {
						for (int m1 = -L; m1 <= L; m1++) {
							for (int m2 = -L; m2 <= L; m2++) {
								size_t so = (((size_t)q * N + m + L) * N + m2 + L) * nb;
								size_t o = (((size_t)(m < 0 ? N + m : m) * N + (m1 < 0 ? N + m1 : m1)) * N + (m2 < 0 ? N + m2 : m2));
								double vr[6] = {0}, vi[6] = {0};
								for (int l = imax(abs(m1), abs(m)); l <= L; l++) {
									double dv = d1[index_d(L, l, m, m1)];
									int ai = lm_index(l, m1);
									double Vr = CA(0, q, ai, 0), Vi = CA(0, q, ai, 1), Dr = CA(1, q, ai, 0), Di = CA(1, q, ai, 1),
									       Wr = CA(2, q, ai, 0), Wi = CA(2, q, ai, 1);
									double svr = s1[0 * s1n + so + l], svi = s1[1 * s1n + so + l], sdr = s1[2 * s1n + so + l],
									       sdi = s1[3 * s1n + so + l], swr = s1[4 * s1n + so + l], swi = s1[5 * s1n + so + l];
									vr[0] += dv * (Vr * svr - Vi * svi);
									vi[0] += dv * (Vr * svi + Vi * svr);
									vr[1] += dv * (Vr * sdr + Dr * svr - Vi * sdi - Di * svi);
									vi[1] += dv * (Vr * sdi + Dr * svi + Vi * sdr + Di * svr);
									vr[2] += dv * (Vr * swr + Wr * svr - Vi * swi - Wi * svi);
									vi[2] += dv * (Vr * swi + Wr * svi + Vi * swr + Wi * svr);
									vr[3] += dv * (Dr * sdr - Di * sdi);
									vi[3] += dv * (Dr * sdi + Di * sdr);
									vr[4] += dv * (Dr * swr + Wr * sdr - Di * swi - Wi * sdi);
									vi[4] += dv * (Dr * swi + Wr * sdi + Di * swr + Wi * sdr);
									vr[5] += dv * (Wr * swr - Wi * swi);
									vi[5] += dv * (Wr * swi + Wi * swr);
								}
								for (int k = 0; k < 6; k++) {
									sum2[((size_t)k * n3 + o) * 2] = vr[k];
									sum2[((size_t)k * n3 + o) * 2 + 1] = vi[k];
								}
							}
						}
					}
					/* simple_ft at the listed points (:529-606) + fill_const/fill_var (:52-108) */
					for (int r = 0; r < nrows; r++) {
						long long pt = index[cell_rows[r]] % n3;
						int g2 = (int)(pt % N), g1 = (int)((pt / N) % N), a2 = (int)(pt / ((long long)N * N));
						double val[6] = {0};
						long long eid = 0;
						size_t mm = 0;
						for (int m = 0; m < N; m++) {
							for (int m1 = 0; m1 < N; m1++) {
								for (int m2 = 0; m2 < N; m2++) {
									double er = e_re[eid % N], ei_ = e_im[eid % N];
									for (int k = 0; k < 6; k++)
										val[k] += er * sum2[((size_t)k * n3 + mm) * 2] - ei_ * sum2[((size_t)k * n3 + mm) * 2 + 1];
									eid += g2;
									mm++;
								}
								eid += g1;
							}
							eid += a2;
						}
						for (int k = 0; k < 6; k++) {
							double v = cst[k * qnum + q];
							v += 2.0 * val[k];
							cross_buf[((size_t)cell_rows[r] * 6 + k) * qnum + q] = v;
						}
					}
				}
			}
		}
	}
	/* sxs_fit_params (:153-194, :909): the reference fits once per distinct grid point of a cell and copies the
	 * result to every row that names the point; fitting each row from its own (identical) cross terms gives the
	 * same numbers.  Rows whose z digit is outside the table keep their incoming values. */
	for (int i = 0; i < nout; i++) {
		if (index[i] < 0 || index[i] / size5d >= znum) continue;
		double o[4];
		port_fit(cross_buf + (size_t)i * 6 * qnum, 1, a, scal, qvals, qnum, o);
		scores[i] = o[0]; c1[i] = o[1]; c2[i] = o[2];
	}
	if (!cross) free(cross_buf);
	free(cst); free(d_symb); free(dw); free(bessel); free(t_re); free(t_im); free(s1); free(sum2);
	free(e_re); free(e_im); free(cell_rows);
#undef CA
#undef CB
}
