/* Amplitude expansion A_lm(q) = 4 pi i^l sum_j f_j j_l(q r_j) Y*_lm(Omega_j), host side.
 *
 * The host prepares exactly the per-atom scalars the reference derives with libm before its atom loop
 * (cart2sph, src/pdb2spf.c:9-22; form-factor lookup :66,86-87; water factor :40-43,71-74) and hands them
 * to the CUDA kernel K1 (sxs_cuda_expand), which does the Legendre/Y_lm/Bessel work and the atom sums.
 */
#include "pdb2spf.h"
#include "sxs_host.h"
#include "sxs_tables.h"

void sxs_spf_full_pack(const struct sxs_spf_full *s, double *coef)
{
	const int lm_n = (s->L + 1) * (s->L + 1);
	struct sxs_spf_sing **comp[3] = {s->V, s->D, s->W};
	for (int c = 0; c < 3; c++) {
		for (int q = 0; q < s->qnum; q++) {
			double *dst = coef + ((size_t)c * s->qnum + q) * lm_n * 2;
			const double *re = comp[c][q]->re, *im = comp[c][q]->im;
			for (int i = 0; i < lm_n; i++) {
				dst[2 * i] = re[i];
				dst[2 * i + 1] = im[i];
			}
		}
	}
}

void sxs_spf_full_unpack(struct sxs_spf_full *s, const double *coef)
{
	const int lm_n = (s->L + 1) * (s->L + 1);
	struct sxs_spf_sing **comp[3] = {s->V, s->D, s->W};
	for (int c = 0; c < 3; c++) {
		for (int q = 0; q < s->qnum; q++) {
			const double *src = coef + ((size_t)c * s->qnum + q) * lm_n * 2;
			double *re = comp[c][q]->re, *im = comp[c][q]->im;
			for (int i = 0; i < lm_n; i++) {
				re[i] = src[2 * i];
				im[i] = src[2 * i + 1];
			}
		}
	}
}

void atom_grp2spf_inplace(struct sxs_spf_full *spf_coefs, struct mol_atom_group *ag,
                          struct saxs_form_factor_table *ff_table, double *qvals, int qnum, int L, double *saxs_sa)
{
	CHECK_PTR(spf_coefs);
	CHECK_PTR(ag);
	sxs_spf_full_reset(spf_coefs);
	spf_coefs->rm = mol_atom_group_average_radius(ag);

	const int natoms = (int)ag->natoms;
	double h2o_ff = 0.0;
	if (saxs_sa != NULL) {
		h2o_ff = ff_table->factors[s_OH2].zero_ff;
	}

	double *buf = (double *)calloc((size_t)natoms * 6, sizeof(double));
	CHECK_PTR(buf);
	double *r = buf, *ct = buf + natoms, *fi = buf + 2 * (size_t)natoms;
	double *fv = buf + 3 * (size_t)natoms, *fd = buf + 4 * (size_t)natoms, *fw = buf + 5 * (size_t)natoms;

	for (int j = 0; j < natoms; j++) {
		const double x = ag->coords[j].X, y = ag->coords[j].Y, z = ag->coords[j].Z;
		/* cart2sph, then cos(theta) as the reference feeds it to the Legendre recurrences (:68-69) */
		const double rad = sqrt(x * x + y * y + z * z);
		const double theta = acos(z / rad);
		double phi;
		if (y > 0.0) {
			phi = acos(x / sqrt(x * x + y * y));
		} else {
			phi = -acos(x / sqrt(x * x + y * y)) + 2.0 * M_PI;
		}
		r[j] = rad;
		ct[j] = cos(theta);
		fi[j] = phi;

		const struct saxs_form_factor *ff = get_ff(ff_table, ag, (size_t)j);
		if (ff == NULL) {
			ERROR_MSG("atom without a form factor (the reference dereferences NULL here)");
		}
		fv[j] = ff->vacuum_ff;
		fd[j] = ff->dummy_ff;
		fw[j] = (saxs_sa != NULL) ? h2o_ff * saxs_sa[j] : 0.0;
	}

	const struct sxs_l_tables *t = sxs_l_tables_get(L);
	const size_t ncoef = (size_t)3 * qnum * (L + 1) * (L + 1) * 2;
	double *coef = (double *)calloc(ncoef, sizeof(double));
	CHECK_PTR(coef);
	SXS_CUDA_CHECK(sxs_cuda_expand(sxs_host_default_device(), natoms, r, ct, fi, fv, fd, fw, qvals, qnum, L,
	                               t->ynorm, t->inv_dfact, 4.0 * M_PI, coef));
	sxs_spf_full_unpack(spf_coefs, coef);
	free(coef);
	free(buf);
}

struct sxs_spf_full *atom_grp2spf(struct mol_atom_group *ag, struct saxs_form_factor_table *ff_table,
                                  double *qvals, int qnum, int L, int water)
{
	struct sxs_spf_full *s = sxs_spf_full_create(L, qnum);
	double *sa = NULL;
	if (water == 1) {
		sa = (double *)calloc(ag->natoms, sizeof(double));
		CHECK_PTR(sa);
		sxs_faccs(sa, ag, 1.4);
	}
	atom_grp2spf_inplace(s, ag, ff_table, qvals, qnum, L, sa);
	free(sa);
	return s;
}

/* ------------------------------------------------------------ containers */

struct sxs_spf_sing *sxs_spf_sing_create(int L)
{
	struct sxs_spf_sing *s = (struct sxs_spf_sing *)calloc(1, sizeof(*s));
	sxs_spf_sing_init(s, L);
	return s;
}

void sxs_spf_sing_init(struct sxs_spf_sing *s, int L)
{
	if (s != NULL) {
		s->L = L;
		s->re = (double *)calloc((size_t)(L + 1) * (L + 1), sizeof(double));
		s->im = (double *)calloc((size_t)(L + 1) * (L + 1), sizeof(double));
	}
}

void sxs_spf_sing_destroy(struct sxs_spf_sing *s)
{
	if (s != NULL) {
		sxs_myfree(s->re);
		sxs_myfree(s->im);
		s->re = s->im = NULL;
	}
}

void sxs_spf_sing_free(struct sxs_spf_sing *s)
{
	sxs_spf_sing_destroy(s);
	sxs_myfree(s);
}

struct sxs_spf_full *sxs_spf_full_create(int L, int qnum)
{
	struct sxs_spf_full *s = (struct sxs_spf_full *)calloc(1, sizeof(*s));
	sxs_spf_full_init(s, L, qnum);
	return s;
}

void sxs_spf_full_init(struct sxs_spf_full *s, int L, int qnum)
{
	if (s == NULL) {
		return;
	}
	s->L = L;
	s->qnum = qnum;
	s->V = (struct sxs_spf_sing **)calloc(qnum, sizeof(struct sxs_spf_sing *));
	s->D = (struct sxs_spf_sing **)calloc(qnum, sizeof(struct sxs_spf_sing *));
	s->W = (struct sxs_spf_sing **)calloc(qnum, sizeof(struct sxs_spf_sing *));
	for (int i = 0; i < qnum; i++) {
		s->V[i] = sxs_spf_sing_create(L);
		s->D[i] = sxs_spf_sing_create(L);
		s->W[i] = sxs_spf_sing_create(L);
	}
}

void sxs_spf_full_destroy(struct sxs_spf_full *s)
{
	if (s == NULL) {
		return;
	}
	for (int i = 0; i < s->qnum; i++) {
		sxs_spf_sing_free(s->V[i]);
		sxs_spf_sing_free(s->D[i]);
		sxs_spf_sing_free(s->W[i]);
	}
	sxs_myfree(s->V);
	sxs_myfree(s->D);
	sxs_myfree(s->W);
	s->V = s->D = s->W = NULL;
	s->qnum = 0;
}

void sxs_spf_full_free(struct sxs_spf_full *s)
{
	sxs_spf_full_destroy(s);
	sxs_myfree(s);
}

void sxs_spf_full_reset(struct sxs_spf_full *s)
{
	const size_t bytes = sizeof(double) * (s->L + 1) * (s->L + 1);
	for (int i = 0; i < s->qnum; i++) {
		memset(s->V[i]->re, 0, bytes); memset(s->V[i]->im, 0, bytes);
		memset(s->D[i]->re, 0, bytes); memset(s->D[i]->im, 0, bytes);
		memset(s->W[i]->re, 0, bytes); memset(s->W[i]->im, 0, bytes);
	}
}

void sxs_spf_full_write(char *path, struct sxs_spf_full *s)
{
	FILE *f = sxs_myfopen(path, "w");
	fprintf(f, "% i % i % .4f\n", s->L, s->qnum, s->rm);
	const int lm_n = (s->L + 1) * (s->L + 1);
	for (int q = 0; q < s->qnum; q++) {
		for (int i = 0; i < lm_n; i++) { /* (l, m) ascending == flat order */
			fprintf(f, "% .4e % .4e % .4e % .4e % .4e % .4e\n", s->V[q]->re[i], s->V[q]->im[i], s->D[q]->re[i],
			        s->D[q]->im[i], s->W[q]->re[i], s->W[q]->im[i]);
		}
	}
	fclose(f);
}

struct sxs_spf_full *sxs_spf_full_fread(FILE *f)
{
	if (f == NULL) {
		return NULL;
	}
	int L, qnum;
	double rm;
	if (fscanf(f, "%i %i %lf\n", &L, &qnum, &rm) != 3) {
		ERROR_MSG("Wrong input file format.");
	}
	struct sxs_spf_full *s = sxs_spf_full_create(L, qnum);
	s->rm = rm;
	const int lm_n = (L + 1) * (L + 1);
	for (int q = 0; q < qnum; q++) {
		for (int i = 0; i < lm_n; i++) {
			if (fscanf(f, "%lf %lf %lf %lf %lf %lf\n", &s->V[q]->re[i], &s->V[q]->im[i], &s->D[q]->re[i],
			           &s->D[q]->im[i], &s->W[q]->re[i], &s->W[q]->im[i]) != 6) {
				ERROR_MSG("Wrong input file format.");
			}
		}
	}
	return s;
}

struct sxs_spf_full *sxs_spf_full_read(char *path)
{
	FILE *f = fopen(path, "r");
	struct sxs_spf_full *s = sxs_spf_full_fread(f);
	if (f != NULL) {
		fclose(f);
	}
	return s;
}
