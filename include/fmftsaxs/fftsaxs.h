/* FFT-SAXS dimer scoring entry point; interface of src/fftsaxs.h:95-105.
 *
 * Grid index of a pose: ((((z*nb + b1)*nb + b2)*N + a2)*N + g1)*N + g2, nb = L+1, N = 2L+1,
 * beta_k = k*pi/L, alpha/gamma_k = k*2pi/N (gamma_A negated), z = index into `zvals`. */
#ifndef FMFTSAXS_FFTSAXS_H
#define FMFTSAXS_FFTSAXS_H
#include "common.h"
#include "sfbessel.h"
#include "pdb2spf.h"
#include "min_saxs.h"
#include "profile.h"
#include "index.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Scores every pose of `index_list` (length nout) and writes chi, c1, c2 into the caller's arrays.
 * Entries whose z digit is outside [0, znum) keep their incoming values.  `skip` is accepted for
 * compatibility: listed poses get the same values with skip = 0 or 1 (src/fftsaxs.c:716-733,867-872).
 * Runs on the CUDA devices selected by SXS_CUDA_DEVICES (default: all visible), z-steps sharded. */
void sxs_compute_saxs_scores(double *scores_list, double *c1_list, double *c2_list, int *index_list, int nout,
                             struct sxs_spf_full *A, struct sxs_spf_full *B, struct sxs_opt_params *params,
                             double *qvals, int qnum, double *zvals, int znum, int L, int skip);

/* Same with 64-bit flat indices (int overflows beyond 9 z-steps at L = 30). */
void sxs_compute_saxs_scores64(double *scores_list, double *c1_list, double *c2_list, const long long *index_list,
                               long long nout, struct sxs_spf_full *A, struct sxs_spf_full *B,
                               struct sxs_opt_params *params, double *qvals, int qnum, double *zvals, int znum,
                               int L, int skip);

#ifdef __cplusplus
}
#endif
#endif
