/* (c1, c2) fit of a cross-term profile against an experimental curve; interface of src/min_saxs.h.
 * The minimisation itself runs on the GPU (kernel K4); there is no host optimiser in this library. */
#ifndef FMFTSAXS_MIN_SAXS_H
#define FMFTSAXS_MIN_SAXS_H
#include "common.h"
#include "profile.h"
#include "pdb2spf.h"
#ifdef __cplusplus
extern "C" {
#endif

/* The reference embeds 371 KB of L-BFGS-B scratch here (src/min_saxs.h:32-46); every GPU thread
 * carries its own optimiser state instead, so only the physical parameters remain. */
struct sxs_opt_params {
	double rm;
	double mult; /* (4 pi / 3)^(3/2) rm^2 / (16 pi) */
	double peak; /* first experimental intensity */
	double *a;   /* 6 moments per model q-bin */
	int last_nfg; /* objective evaluations of the most recent single fit */
};

struct sxs_opt_params *sxs_opt_params_create(struct sxs_profile *exp, double *qvals, int qnum, double rm);
void sxs_opt_params_init(struct sxs_opt_params *params, struct sxs_profile *exp, double *qvals, int qnum, double rm);
void sxs_opt_params_destroy(struct sxs_opt_params *params);
void sxs_opt_params_free(struct sxs_opt_params *params);

void sxs_fit_params(struct sxs_profile **profiles, struct sxs_opt_params *params, int *mask, int n);
void sxs_lbfgs_fitting(struct sxs_profile *profile, struct sxs_opt_params *params);
double *scoring_helper(struct sxs_profile *exp, int qnum, double *qvals);
void sxs_compile_intensity(struct sxs_profile *profile, double rm, double c1, double c2);
double sxs_best_scale(struct sxs_profile *profile, struct sxs_opt_params *params, double c1, double c2);
void sxs_spf2cross_terms(struct sxs_profile *profile, struct sxs_spf_full *s);
void sxs_spf2fitted_profile(struct sxs_profile *profile, struct sxs_spf_full *s, struct sxs_opt_params *params);

#ifdef __cplusplus
}
#endif
#endif
