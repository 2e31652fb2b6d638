/* Host side of the (c1, c2) fit: experiment compression and parameter container.  The optimiser
 * (reference: sxs_lbfgs_fitting + vendored L-BFGS-B) runs only on the GPU, one fit per thread. */
#include "min_saxs.h"
#include "sxs_host.h"

/* Bins the experimental points into the model's q intervals (q_{i-1}, q_i] (first bin [q_0, q_0]) and
 * keeps six moments per bin, all divided by the number of points consumed (src/min_saxs.c:353-389).
 * The scan stops at the end of the experimental arrays; the reference relies on whatever follows
 * them in memory to stop it. */
double *scoring_helper(struct sxs_profile *exp, int qnum, double *qvals)
{
	double *a = (double *)calloc((size_t)6 * qnum, sizeof(double));
	CHECK_PTR(a);
	int j = 0;
	for (int i = 0; i < qnum; i++) {
		double lower = qvals[0];
		double upper = qvals[i];
		if (i > 0) {
			lower = qvals[i - 1];
		}
		while (j < exp->qnum && exp->qvals[j] >= lower && exp->qvals[j] <= upper) {
			const double w = exp->err[j] * exp->err[j];
			a[i * 6 + 0] += exp->in[j] * exp->in[j] / w;
			a[i * 6 + 1] += exp->in[j] / w;
			a[i * 6 + 2] += exp->qvals[j] * exp->in[j] / w;
			a[i * 6 + 3] += 1.0 / w;
			a[i * 6 + 4] += exp->qvals[j] / w;
			a[i * 6 + 5] += exp->qvals[j] * exp->qvals[j] / w;
			j++;
		}
	}
	const double norm = (double)j;
	for (int i = 0; i < 6 * qnum; i++) {
		a[i] /= norm;
	}
	return a;
}

struct sxs_opt_params *sxs_opt_params_create(struct sxs_profile *exp, double *qvals, int qnum, double rm)
{
	struct sxs_opt_params *p = (struct sxs_opt_params *)calloc(1, sizeof(*p));
	sxs_opt_params_init(p, exp, qvals, qnum, rm);
	return p;
}

void sxs_opt_params_init(struct sxs_opt_params *p, struct sxs_profile *exp, double *qvals, int qnum, double rm)
{
	if (p != NULL) {
		p->a = scoring_helper(exp, qnum, qvals);
		p->rm = rm;
		p->mult = pow(4.0 * M_PI / 3.0, 1.5) * rm * rm / (16.0 * M_PI);
		p->peak = exp->in[0];
		p->last_nfg = 0;
	}
}

void sxs_opt_params_destroy(struct sxs_opt_params *p)
{
	if (p != NULL) {
		sxs_myfree(p->a);
		p->a = NULL;
	}
}

void sxs_opt_params_free(struct sxs_opt_params *p)
{
	sxs_opt_params_destroy(p);
	sxs_myfree(p);
}

static void pack_cross(const struct sxs_profile *p, double *dst)
{
	const int n = p->qnum;
	memcpy(dst + 0 * n, p->VV, sizeof(double) * n);
	memcpy(dst + 1 * n, p->VD, sizeof(double) * n);
	memcpy(dst + 2 * n, p->VW, sizeof(double) * n);
	memcpy(dst + 3 * n, p->DD, sizeof(double) * n);
	memcpy(dst + 4 * n, p->DW, sizeof(double) * n);
	memcpy(dst + 5 * n, p->WW, sizeof(double) * n);
}

/* Fit `count` profiles in one kernel launch.  rescale: the peak normalisation of sxs_fit_params. */
static void fit_batch(struct sxs_profile **list, int count, struct sxs_opt_params *params, int rescale)
{
	if (count == 0) {
		return;
	}
	const int qnum = list[0]->qnum;
	double *cross = (double *)calloc((size_t)count * 6 * qnum, sizeof(double));
	double *out = (double *)calloc((size_t)count * 4, sizeof(double));
	CHECK_PTR(cross);
	CHECK_PTR(out);
	for (int i = 0; i < count; i++) {
		pack_cross(list[i], cross + (size_t)i * 6 * qnum);
	}
	SXS_CUDA_CHECK(sxs_cuda_fit_profiles(sxs_host_default_device(), cross, count, params->a, list[0]->qvals, qnum,
	                                     params->mult, params->peak, rescale, out));
	for (int i = 0; i < count; i++) {
		struct sxs_profile *p = list[i];
		if (rescale) { /* the reference scales the stored cross terms in place (src/min_saxs.c:181-188) */
			double s = p->VV[0] + p->DD[0] + p->WW[0] + p->VW[0] - p->VD[0] - p->DW[0];
			s = params->peak / s;
			for (int q = 0; q < qnum; q++) {
				p->VV[q] *= s; p->DD[q] *= s; p->WW[q] *= s;
				p->VW[q] *= s; p->VD[q] *= s; p->DW[q] *= s;
			}
		}
		p->score = out[4 * i + 0];
		p->c1 = out[4 * i + 1];
		p->c2 = out[4 * i + 2];
		params->last_nfg = (int)out[4 * i + 3];
		p->scale = sxs_best_scale(p, params, p->c1, p->c2);
		sxs_compile_intensity(p, params->rm, p->c1, p->c2);
	}
	free(cross);
	free(out);
}

void sxs_fit_params(struct sxs_profile **profiles, struct sxs_opt_params *params, int *mask, int n)
{
	struct sxs_profile **list = (struct sxs_profile **)calloc((size_t)(n > 0 ? n : 1), sizeof(*list));
	CHECK_PTR(list);
	int count = 0;
	for (int i = 0; i < n; i++) {
		if (mask[i] == 1) {
			list[count++] = profiles[i];
		}
	}
	fit_batch(list, count, params, 1);
	free(list);
}

void sxs_lbfgs_fitting(struct sxs_profile *profile, struct sxs_opt_params *params)
{
	fit_batch(&profile, 1, params, 0);
}

double sxs_best_scale(struct sxs_profile *profile, struct sxs_opt_params *params, double c1, double c2)
{
	const int qnum = profile->qnum;
	double *cross = (double *)calloc((size_t)6 * qnum, sizeof(double));
	CHECK_PTR(cross);
	pack_cross(profile, cross);
	double out[4];
	SXS_CUDA_CHECK(sxs_cuda_fit_eval(sxs_host_default_device(), cross, params->a, profile->qvals, qnum, params->mult,
	                                 c1, c2, out));
	free(cross);
	return out[0];
}

/* I(q) from the stored cross terms at (c1, c2) (src/min_saxs.c:321-350); output formatting only. */
void sxs_compile_intensity(struct sxs_profile *profile, double rm, double c1, double c2)
{
	const double mult = pow(4.0 * M_PI / 3.0, 3.0 / 2.0) * rm * rm / (16.0 * M_PI);
	const double corr = -mult * (c1 * c1 - 1.0);
	const double *q = profile->qvals;
	for (int i = 0; i < profile->qnum; i++) {
		const double G = c1 * c1 * c1 * exp(corr * q[i] * q[i]);
		profile->in[i] = profile->VV[i] - G * profile->VD[i] + c2 * profile->VW[i] + G * G * profile->DD[i] -
		                 G * c2 * profile->DW[i] + c2 * c2 * profile->WW[i];
		profile->err[i] = profile->rerr * profile->in[i];
	}
}

void sxs_spf2cross_terms(struct sxs_profile *profile, struct sxs_spf_full *s)
{
	const int qnum = profile->qnum, L = s->L;
	double *coef = (double *)calloc((size_t)3 * qnum * (L + 1) * (L + 1) * 2, sizeof(double));
	double *out = (double *)calloc((size_t)6 * qnum, sizeof(double));
	CHECK_PTR(coef);
	CHECK_PTR(out);
	sxs_spf_full_pack(s, coef);
	SXS_CUDA_CHECK(sxs_cuda_self_terms(sxs_host_default_device(), coef, qnum, L, out));
	memcpy(profile->VV, out + 0 * qnum, sizeof(double) * qnum);
	memcpy(profile->VD, out + 1 * qnum, sizeof(double) * qnum);
	memcpy(profile->VW, out + 2 * qnum, sizeof(double) * qnum);
	memcpy(profile->DD, out + 3 * qnum, sizeof(double) * qnum);
	memcpy(profile->DW, out + 4 * qnum, sizeof(double) * qnum);
	memcpy(profile->WW, out + 5 * qnum, sizeof(double) * qnum);
	free(coef);
	free(out);
}

void sxs_spf2fitted_profile(struct sxs_profile *profile, struct sxs_spf_full *s, struct sxs_opt_params *params)
{
	sxs_spf2cross_terms(profile, s);
	sxs_lbfgs_fitting(profile, params);
}
