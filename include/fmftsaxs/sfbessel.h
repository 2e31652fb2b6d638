/* Spherical Bessel function exactly as the reference evaluates it (src/sfbessel.h:19). */
#ifndef FMFTSAXS_SFBESSEL_H
#define FMFTSAXS_SFBESSEL_H
#include "common.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Truncated ascending series (relative cut 1e-5), src/sfbessel.c:40-60; x <= 0 gives delta_{l0}. */
double sxs_sbessel(int order, double x);
#ifdef __cplusplus
}
#endif
#endif
