/* Scattering curve container with the six partial cross terms; interface of src/profile.h. */
#ifndef FMFTSAXS_PROFILE_H
#define FMFTSAXS_PROFILE_H
#include "common.h"
#include "pdb2spf.h"
#ifdef __cplusplus
extern "C" {
#endif

struct sxs_profile {
	int qnum;
	double rerr;

	double *in;
	double *err;
	double *qvals; /* borrowed, never freed by the profile */

	double *VV, *VD, *VW, *DD, *DW, *WW;

	double score;
	double scale;
	double c1;
	double c2;

	double b1, g1, a2, b2, g2;
};

struct sxs_profile *sxs_profile_create(double *qvals, int qnum, int cross_terms_flag);
void sxs_profile_init(struct sxs_profile *profile, double *qvals, int qnum, int cross_terms_flag);
void sxs_profile_alloc_cross_terms(struct sxs_profile *profile, int qnum);
void sxs_profile_destroy(struct sxs_profile *profile);
void sxs_profile_free(struct sxs_profile *profile);
/* rows "%.4f %.4f %.4f" = q, I, err (src/profile.c:66-85) */
void sxs_profile_write(char *path, struct sxs_profile *profile);
struct sxs_profile *sxs_profile_read(char *path);
struct sxs_profile *sxs_profile_fread(FILE *f);
/* I(q) = sum_lm |A^v - G A^d + c2 A^w|^2 with G taken at qvals[0] for every q (src/profile.c:186-233). */
void sxs_profile_from_spf(struct sxs_profile *profile, struct sxs_spf_full *s, double c1, double c2);

#ifdef __cplusplus
}
#endif
#endif
