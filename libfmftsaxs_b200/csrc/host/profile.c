/* Scattering-curve container, text I/O, and I(q) of a single molecule. */
#include "profile.h"
#include "sxs_host.h"

struct sxs_profile *sxs_profile_create(double *qvals, int qnum, int cross_terms_flag)
{
	struct sxs_profile *p = (struct sxs_profile *)calloc(1, sizeof(*p));
	sxs_profile_init(p, qvals, qnum, cross_terms_flag);
	return p;
}

void sxs_profile_init(struct sxs_profile *p, double *qvals, int qnum, int cross_terms_flag)
{
	if (p == NULL) {
		return;
	}
	p->qnum = qnum;
	p->rerr = REL_ERR;
	p->qvals = qvals;
	p->in = (double *)calloc(qnum, sizeof(double));
	p->err = (double *)calloc(qnum, sizeof(double));
	p->score = INFINITY;
	p->c1 = C1_DEFAULT;
	p->c2 = C2_DEFAULT;
	p->scale = 1.0;
	p->VV = p->VD = p->VW = p->DD = p->DW = p->WW = NULL;
	if (cross_terms_flag == 1) {
		sxs_profile_alloc_cross_terms(p, qnum);
	}
}

void sxs_profile_alloc_cross_terms(struct sxs_profile *p, int qnum)
{
	p->VV = (double *)calloc(qnum, sizeof(double));
	p->VD = (double *)calloc(qnum, sizeof(double));
	p->VW = (double *)calloc(qnum, sizeof(double));
	p->DD = (double *)calloc(qnum, sizeof(double));
	p->DW = (double *)calloc(qnum, sizeof(double));
	p->WW = (double *)calloc(qnum, sizeof(double));
}

void sxs_profile_destroy(struct sxs_profile *p)
{
	p->qvals = NULL;
	sxs_myfree(p->in);
	sxs_myfree(p->err);
	sxs_myfree(p->VV); sxs_myfree(p->VD); sxs_myfree(p->VW);
	sxs_myfree(p->DD); sxs_myfree(p->DW); sxs_myfree(p->WW);
	p->in = p->err = p->VV = p->VD = p->VW = p->DD = p->DW = p->WW = NULL;
}

void sxs_profile_free(struct sxs_profile *p)
{
	if (p != NULL) {
		sxs_profile_destroy(p);
		free(p);
	}
}

void sxs_profile_write(char *path, struct sxs_profile *p)
{
	FILE *f = sxs_myfopen(path, "w");
	for (int i = 0; i < p->qnum; i++) {
		double err = p->err[i] > 0.0 ? p->err[i] : p->in[i] * p->rerr;
		fprintf(f, "%.4f %.4f %.4f\n", p->qvals[i], p->in[i], err);
	}
	fclose(f);
}

/* "q I err" triplets; the q array is allocated here and owned by the caller (src/profile.c:98-124). */
struct sxs_profile *sxs_profile_fread(FILE *f)
{
	if (f == NULL) {
		return NULL;
	}
	int n = 0, got;
	double a, b, c;
	while ((got = fscanf(f, "%lf %lf %lf", &a, &b, &c)) != EOF) {
		if (got != 3) {
			ERROR_MSG("Wrong input file format.");
		}
		n++;
	}
	rewind(f);
	/* one spare slot: scoring_helper scans one element past the end in the reference (src/min_saxs.c:369) */
	double *qvals = (double *)calloc((size_t)n + 1, sizeof(double));
	struct sxs_profile *p = sxs_profile_create(qvals, n, 0);
	for (int i = 0; i < n; i++) {
		if (fscanf(f, "%lf %lf %lf\n", &qvals[i], &p->in[i], &p->err[i]) != 3) {
			ERROR_MSG("Wrong input file format.");
		}
	}
	return p;
}

struct sxs_profile *sxs_profile_read(char *path)
{
	FILE *f = fopen(path, "r");
	struct sxs_profile *p = sxs_profile_fread(f);
	if (f != NULL) {
		fclose(f);
	}
	return p;
}

void sxs_profile_from_spf(struct sxs_profile *profile, struct sxs_spf_full *s, double c1, double c2)
{
	CHECK_PTR(profile);
	CHECK_PTR(s);
	assert(profile->qnum == s->qnum);
	const int qnum = profile->qnum, L = s->L;
	const double rm = s->rm;
	const double mult = pow(4.0 * M_PI / 3.0, 3.0 / 2.0) * rm * rm / (16.0 * M_PI);

	double *coef = (double *)calloc((size_t)3 * qnum * (L + 1) * (L + 1) * 2, sizeof(double));
	CHECK_PTR(coef);
	sxs_spf_full_pack(s, coef);
	SXS_CUDA_CHECK(sxs_cuda_profile_from_spf(sxs_host_default_device(), coef, qnum, L, mult, profile->qvals, c1, c2,
	                                         profile->in));
	free(coef);
	for (int q = 0; q < qnum; q++) {
		profile->err[q] = profile->in[q] * profile->rerr;
	}
	profile->c1 = c1;
	profile->c2 = c2;
}
