/* Special-function tables (spherical-harmonic norms, associated Legendre, Wigner small-d, 3j);
 * interface of src/borrowed.h minus its GMP types: the 3j symbol is returned as a double. */
#ifndef FMFTSAXS_BORROWED_H
#define FMFTSAXS_BORROWED_H
#include "common.h"
#ifdef __cplusplus
extern "C" {
#endif

/* (l, m) -> slot in the N x (2N-1) tables used by the two functions below (src/borrowed.h:17-20). */
static inline int index_LM(int N, int l, int m) { return N * (m + N - 1) + l; }

/* sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) with the Condon-Shortley sign folded into negative m; src/borrowed.c:19-38. */
double *generate_spherical_norm(int N);
/* P_l^m(x), m >= 0, l < N by the standard upward recurrences; src/borrowed.c:40-65. */
void fill_assoc_Legendre_array(int N, double x, double *P);

struct d_array {
	int L;
	double *data; /* (L+1)(2L+1)^2 doubles */
};

static inline int index_d_array(int L, int l, int m, int m1)
{
	return (2 * L + 1) * ((2 * L + 1) * l + (L + m)) + (L + m1);
}

struct d_array *allocate_d_array(const int L);
void deallocate_d_array(struct d_array *d);
/* Wigner small-d d^l_{m m1}(beta), l <= L; src/borrowed.c:243-313. */
struct d_array *generate_d_array(const int L, const double beta);

/* Wigner 3j by the Racah sum, evaluated in binary128 (replaces the GMP routine src/borrowed.c:86-222). */
double sxs_wigner_3j(int j1, int j2, int j3, int m1, int m2, int m3);

#ifdef __cplusplus
}
#endif
#endif
