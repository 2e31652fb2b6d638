/* Special-function tables of the scoring path, all host-side one-off work:
 * spherical-harmonic norms, associated Legendre values, Wigner small-d matrices on the beta grid,
 * Wigner-3j products ("d-symbols") of the z-translation operator.
 *
 * Arithmetic follows the reference's expressions (src/borrowed.c, src/fftsaxs.c:138-174) so the
 * tables agree with it to the last bit wherever it uses doubles; the 3j symbols, which the reference
 * evaluates in GMP floats, are evaluated here in IEEE binary128 (error < 1e-26 up to L = 40) and
 * rounded to double once, like the reference's mpf_get_d.  Built with -ffp-contract=off.
 */
#include <quadmath.h>

#include "borrowed.h"
#include "sfbessel.h"
#include "sxs_tables.h"

/* ----------------------------------------------------------- Y_lm norms */

double *generate_spherical_norm(int N)
{
	double *norm = (double *)calloc((size_t)N * (2 * N - 1), sizeof(double));
	CHECK_PTR(norm);
	for (int l = 0; l < N; ++l) {
		double ratio = 1.;                       /* (l+m)!/(l-m)! built incrementally */
		double lead = (2 * l + 1.) / (4. * M_PI);
		for (int m = 0; m <= l; ++m) {
			if (m > 0) {
				ratio *= (l - m + 1) * (l + m);
			}
			double v = sqrt(lead / ratio);
			double sgn = (m % 2 == 0) ? (1) : (-1);
			norm[index_LM(N, l, m)] = v;
			norm[index_LM(N, l, -m)] = sgn * v;
		}
	}
	return norm;
}

/* ------------------------------------------------- associated Legendre */

void fill_assoc_Legendre_array(int N, double x, double *P)
{
	double y = sqrt(1. - x * x);

	P[index_LM(N, 0, 0)] = 1;
	if (N == 1) {
		return;
	}
	P[index_LM(N, 1, 0)] = x;
	for (int l = 2; l < N; ++l) {
		P[index_LM(N, l, 0)] = (2. * l - 1.) / l * x * P[index_LM(N, (l - 1), 0)] - (l - 1.) / l * P[index_LM(N, (l - 2), 0)];
	}
	for (int m = 1; m < N - 1; ++m) {
		P[index_LM(N, m, m)] = -1. * (2. * m - 1.) * y * P[index_LM(N, (m - 1), (m - 1))];
		P[index_LM(N, (m + 1), m)] = (2. * m + 1.) * x * P[index_LM(N, m, m)];
		for (int l = m + 2; l < N; ++l) {
			P[index_LM(N, l, m)] = (2. * l - 1.) / (l - m) * x * P[index_LM(N, (l - 1), m)] -
			                       (l + m - 1.) / (l - m) * P[index_LM(N, (l - 2), m)];
		}
	}
	P[index_LM(N, (N - 1), (N - 1))] = -1. * (2. * (N - 1.) - 1.) * y * P[index_LM(N, (N - 2), (N - 2))];
}

/* ------------------------------------------------------ Wigner small-d */

struct d_array *allocate_d_array(const int L)
{
	struct d_array *d = (struct d_array *)calloc(1, sizeof(struct d_array));
	CHECK_PTR(d);
	d->L = L;
	d->data = (double *)calloc((size_t)(L + 1) * (2 * L + 1) * (2 * L + 1), sizeof(double));
	CHECK_PTR(d->data);
	return d;
}

void deallocate_d_array(struct d_array *d)
{
	if (d != NULL) {
		free(d->data);
		free(d);
	}
}

/* Kostelec & Rockmore (2003): closed forms for l <= 1, the four border lines of every l by the
 * product formula (eq. 26), the interior by the three-term recurrence in l (eq. 28).
 * Operation order as src/borrowed.c:243-313. */
struct d_array *generate_d_array(const int L, const double beta)
{
	struct d_array *d = allocate_d_array(L);
	double *D = d->data;
#define DW(l, m, m1) D[index_d_array(L, (l), (m), (m1))]

	DW(0, 0, 0) = 1.0;
	if (L >= 1) {
		DW(1, -1, -1) = 1.0 * (1.0 + cos(beta)) / 2.0;
		DW(1, -1, 0) = 1.0 * sin(beta) / sqrt(2.0);
		DW(1, -1, 1) = 1.0 * (1.0 - cos(beta)) / 2.0;
		DW(1, 0, -1) = -1.0 * sin(beta) / sqrt(2.0);
		DW(1, 0, 0) = 1.0 * cos(beta);
		DW(1, 0, 1) = 1.0 * sin(beta) / sqrt(2.0);
		DW(1, 1, -1) = 1.0 * (1.0 - cos(beta)) / 2.0;
		DW(1, 1, 0) = -1.0 * sin(beta) / sqrt(2.0);
		DW(1, 1, 1) = 1.0 * (1.0 + cos(beta)) / 2.0;
	}

	for (int l = 2; l <= L; ++l) {
		for (int m = -l; m <= l; ++m) {
			double e1 = 1, e2 = 1, e3 = 1, e4 = 1;
			for (int k = 1; k <= 2 * l; ++k) {
				const int in_p = (k <= l + m), in_n = (k <= l - m);
				double w = sqrt(k / ((in_p ? k : 1.0) * (in_n ? k : 1.0)));
				e1 *= w * (in_p ? cos(beta / 2.0) : 1.0) * (in_n ? -sin(beta / 2.0) : 1.0);
				e2 *= w * (in_n ? cos(beta / 2.0) : 1.0) * (in_p ? sin(beta / 2.0) : 1.0);
				e3 *= w * (in_p ? cos(beta / 2.0) : 1.0) * (in_n ? sin(beta / 2.0) : 1.0);
				e4 *= w * (in_n ? cos(beta / 2.0) : 1.0) * (in_p ? -sin(beta / 2.0) : 1.0);
			}
			DW(l, l, m) = e1;
			DW(l, -l, m) = e2;
			DW(l, m, l) = e3;
			DW(l, m, -l) = e4;

			for (int m1 = -l; m1 <= l; ++m1) {
				int j = l - 1;
				if ((m1 > -l) && (m1 < l) && (m > -l) && (m < l)) {
					double up = 1.0;
					up *= (j + 1) * (2.0 * j + 1) / sqrt(((j + 1) * (j + 1) - m * m) * ((j + 1) * (j + 1) - m1 * m1));
					up *= (cos(beta) - (double)(m * m1) / (double)(j * (j + 1))) * DW(j, m, m1);

					double dn = -1.0;
					dn *= sqrt((j * j - m * m) * (j * j - m1 * m1));
					dn *= (double)(j + 1.0) * DW(j - 1, m, m1) / (double)j;
					dn /= sqrt(((j + 1.0) * (j + 1.0) - (double)(m * m)) * ((j + 1.0) * (j + 1.0) - (double)(m1 * m1)));

					up += dn;
					DW(l, m, m1) = up;
				}
			}
		}
	}
#undef DW
	return d;
}

/* ------------------------------------------------------------ Wigner 3j */

#define SXS_FACT_MAX 400
static __float128 g_fact[SXS_FACT_MAX + 1];
static int g_fact_ready = 0;

static void fact_init(void)
{
	if (!g_fact_ready) {
		g_fact[0] = 1;
		for (int i = 1; i <= SXS_FACT_MAX; i++) {
			g_fact[i] = g_fact[i - 1] * i;
		}
		g_fact_ready = 1;
	}
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* Racah's single-sum formula (the reference's wigner_3j_symbol_arb, src/borrowed.c:86-222,
 * evaluates the same expression in GMP floats).  Integer arguments only. */
double sxs_wigner_3j(int j1, int j2, int j3, int m1, int m2, int m3)
{
	fact_init();
	if (j1 + j2 + j3 + 1 > SXS_FACT_MAX) {
		ERROR_MSG("3j arguments exceed the factorial table");
	}
	const int t1 = j2 - m1 - j3;
	const int t2 = j1 + m2 - j3;
	const int t3 = j1 + j2 - j3;
	const int t4 = j1 - m1;
	const int t5 = j2 + m2;
	const int tmin = imax(0, imax(t1, t2));
	const int tmax = imin(t3, imin(t4, t5));

	__float128 sum = 0;
	for (int t = tmin; t <= tmax; ++t) {
		__float128 den = g_fact[t] * g_fact[t - t1] * g_fact[t - t2] * g_fact[t3 - t] * g_fact[t4 - t] * g_fact[t5 - t];
		sum += ((t % 2 == 0) ? (__float128)1 : (__float128)-1) / den;
	}
	__float128 tri = g_fact[j1 + m1] * g_fact[j1 - m1] * g_fact[j2 + m2] * g_fact[j2 - m2] * g_fact[j3 + m3] *
	                 g_fact[j3 - m3] / g_fact[j1 + j2 + j3 + 1] * g_fact[-j1 + j2 + j3] * g_fact[j1 - j2 + j3] *
	                 g_fact[j1 + j2 - j3];
	__float128 res = sum * sqrtq(tri);
	if ((j1 - j2 - m3) % 2 != 0) {
		res = -res;
	}
	return (double)res;
}

/* ------------------------------------------------------- cached L tables */

#define SXS_LMAX_CACHE 128
static struct sxs_l_tables *g_tables[SXS_LMAX_CACHE + 1];

static void build_dsymb(double *dsymb, int L)
{
	const int nb = L + 1, N = 2 * L + 1;
	for (int l = 0; l < nb; l++) {
		double k1 = 2 * l + 1;
		for (int l1 = 0; l1 < nb; l1++) {
			double k2 = sqrt((2 * l1 + 1) * k1);
			for (int p = abs(l - l1); p <= l + l1; p++) {
				double k3 = (2 * p + 1) * k2 * sxs_wigner_3j(l, p, l1, 0, 0, 0);
				int mm = imin(l, l1);
				for (int m = -mm; m <= mm; m++) {
					double k4 = sxs_wigner_3j(l, p, l1, -m, 0, m);
					dsymb[((size_t)lm_index(l, m) * nb + l1) * N + p] = k3 * k4;
				}
			}
		}
	}
}

const struct sxs_l_tables *sxs_l_tables_get(int L)
{
	if (L < 1 || L > SXS_LMAX_CACHE) {
		ERROR_MSG("unsupported expansion order L");
	}
	if (g_tables[L] != NULL) {
		return g_tables[L];
	}
	const int nb = L + 1, N = 2 * L + 1;
	struct sxs_l_tables *t = (struct sxs_l_tables *)calloc(1, sizeof(*t));
	CHECK_PTR(t);
	t->L = L;
	t->dsymb = (double *)calloc((size_t)nb * nb * nb * N, sizeof(double));
	t->dwig = (double *)calloc((size_t)nb * nb * N * N, sizeof(double));
	t->twiddle = (double *)calloc((size_t)N * 2, sizeof(double));
	t->inv_dfact = (double *)calloc((size_t)N, sizeof(double));
	CHECK_PTR(t->dsymb); CHECK_PTR(t->dwig); CHECK_PTR(t->twiddle); CHECK_PTR(t->inv_dfact);

	build_dsymb(t->dsymb, L);

	/* beta grid: 0 .. pi in L steps (src/fftsaxs.c:632-636,706-708), truncated pi */
	const double beta_bgn = 0.0, beta_end = M_PI;
	const double beta_step = (beta_end - beta_bgn) / (nb - 1);
	for (int i = 0; i < nb; i++) {
		struct d_array *d = generate_d_array(L, beta_bgn + i * beta_step);
		memcpy(t->dwig + (size_t)i * nb * N * N, d->data, sizeof(double) * nb * N * N);
		deallocate_d_array(d);
	}

	/* angular DFT phases as the reference's direct transform builds them (src/fftsaxs.c:536-548) */
	const double step = 2 * M_PI / N;
	for (int i = 0; i < N; i++) {
		t->twiddle[2 * i] = cos(i * step);
		t->twiddle[2 * i + 1] = -sin(i * step);
	}

	t->ynorm = generate_spherical_norm(L + 1);
	for (int p = 0; p < N; p++) {
		t->inv_dfact[p] = 1.0 / sxs_odd_double_factorial(2 * p + 1);
	}
	g_tables[L] = t;
	return t;
}

void sxs_fill_bessel_table(double *bessel, const double *zvals, int znum, const double *qvals, int qnum, int L)
{
	const int N = 2 * L + 1;
	for (int zi = 0; zi < znum; zi++) {
		for (int q = 0; q < qnum; q++) {
			for (int p = 0; p < N; p++) {
				bessel[((size_t)zi * qnum + q) * N + p] = sxs_sbessel(p, zvals[zi] * qvals[q]);
			}
		}
	}
}
