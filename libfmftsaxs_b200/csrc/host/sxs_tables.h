/* Internal: host-side precomputed tables handed to the CUDA layer. */
#ifndef SXS_TABLES_H
#define SXS_TABLES_H
#ifdef __cplusplus
extern "C" {
#endif

/* Tables that depend only on L; built once per L and cached for the life of the process. */
struct sxs_l_tables {
	int L;
	double *dsymb;   /* [(L+1)^2][L+1][2L+1]: d_lm(l1,p), src/fftsaxs.c:138-174 */
	double *dwig;    /* [L+1 betas][L+1][2L+1][2L+1]: d^l_{m m1}(k*pi/L), src/fftsaxs.c:706-708 */
	double *twiddle; /* [2L+1][2]: cos(k*step), -sin(k*step), step = 2*pi/(2L+1), src/fftsaxs.c:536-548 */
	double *ynorm;   /* [(L+1)(2L+1)]: generate_spherical_norm(L+1) */
	double *inv_dfact; /* [2L+1]: 1/(2p+1)!! */
};

const struct sxs_l_tables *sxs_l_tables_get(int L);

/* j_p(q z), p = 0..2L: bessel[(zi*qnum + q)*(2L+1) + p], src/fftsaxs.c:110-121. */
void sxs_fill_bessel_table(double *bessel, const double *zvals, int znum, const double *qvals, int qnum, int L);

double sxs_odd_double_factorial(int n);

#ifdef __cplusplus
}
#endif
#endif
