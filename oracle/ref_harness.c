/* TEST INFRASTRUCTURE — not product code.
 *
 * Flat-array entry points around the UNMODIFIED reference sources (compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/libsxsref.so).  They exist so
 * tests/ and bench.py's cpu_baseline leg can drive the reference through ctypes
 * without re-declaring its pointer-of-pointer structs.  Nothing under
 * libfmftsaxs_b200/ may link or load this.
 *
 * Coefficient layout used by every function here (and by the product C-ABI):
 *   coef[((c*qnum + q)*(L+1)^2 + lm)*2 + {0:re,1:im}],  c = 0:V 1:D 2:W, lm = l(l+1)+m
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fftsaxs.h"
#include "index.h"
#include "min_saxs.h"
#include "pdb2spf.h"
#include "profile.h"
#include "saxs_utils.h"
#include "sfbessel.h"
#include "borrowed.h"

#include "mol2/atom_group.h"
#include "mol2/pdb.h"
#include "mol2/prms.h"

static void spf_to_flat(const struct sxs_spf_full *s, double *coef)
{
	int lm_n = (s->L + 1) * (s->L + 1);
	struct sxs_spf_sing **comp[3] = {s->V, s->D, s->W};
	for (int c = 0; c < 3; c++) {
		for (int q = 0; q < s->qnum; q++) {
			for (int i = 0; i < lm_n; i++) {
				size_t o = (((size_t)c * s->qnum + q) * lm_n + i) * 2;
				coef[o] = comp[c][q]->re[i];
				coef[o + 1] = comp[c][q]->im[i];
			}
		}
	}
}

static struct sxs_spf_full *flat_to_spf(const double *coef, int qnum, int L, double rm)
{
	struct sxs_spf_full *s = sxs_spf_full_create(L, qnum);
	s->rm = rm;
	int lm_n = (L + 1) * (L + 1);
	struct sxs_spf_sing **comp[3] = {s->V, s->D, s->W};
	for (int c = 0; c < 3; c++) {
		for (int q = 0; q < qnum; q++) {
			for (int i = 0; i < lm_n; i++) {
				size_t o = (((size_t)c * qnum + q) * lm_n + i) * 2;
				comp[c][q]->re[i] = coef[o];
				comp[c][q]->im[i] = coef[o + 1];
			}
		}
	}
	return s;
}

static struct mol_atom_group *make_group(int natoms, const double *xyz, const char *const *res,
                                         const char *const *atm, const double *radius)
{
	struct mol_atom_group *ag = mol_atom_group_create((size_t)natoms);
	for (int i = 0; i < natoms; i++) {
		ag->coords[i].X = xyz[3 * i];
		ag->coords[i].Y = xyz[3 * i + 1];
		ag->coords[i].Z = xyz[3 * i + 2];
		ag->vdw_radius[i] = radius[i];
		ag->residue_name[i] = strdup(res[i]);
		ag->atom_name[i] = strdup(atm[i]);
	}
	return ag;
}

/* sxs_sbessel, src/sfbessel.c:40-60 */
double ref_sbessel(int l, double x)
{
	return sxs_sbessel(l, x);
}

/* sxs_mkarray, src/saxs_utils.c:50-63 */
void ref_mkarray(double begin, double end, int qnum, double *out)
{
	double *q = sxs_mkarray(begin, end, qnum);
	memcpy(out, q, sizeof(double) * qnum);
	free(q);
}

/* atom_grp2spf_inplace (src/pdb2spf.c:24-152) on explicit atoms.
 * water_mode: 0 = no hydration term (saxs_sa NULL), 1 = use sa[] as given,
 *             2 = sxs_faccs(…, 1.4) (src/saxs_utils.c:30-48), result copied back into sa[] if non-NULL. */
int ref_expand(const char *map_path, int natoms, const double *xyz, const char *const *res,
               const char *const *atm, const double *radius, double *sa, int water_mode,
               const double *qvals, int qnum, int L, double *coef, double *rm)
{
	struct saxs_form_factor_table *ff = default_ff_table(map_path);
	struct mol_atom_group *ag = make_group(natoms, xyz, res, atm, radius);
	struct sxs_spf_full *s = sxs_spf_full_create(L, qnum);
	double *sa_use = NULL;
	if (water_mode == 1) {
		sa_use = sa;
	} else if (water_mode == 2) {
		sa_use = calloc((size_t)natoms, sizeof(double));
		sxs_faccs(sa_use, ag, 1.4);
		if (sa != NULL) {
			memcpy(sa, sa_use, sizeof(double) * natoms);
		}
	}
	atom_grp2spf_inplace(s, ag, ff, (double *)qvals, qnum, L, sa_use);
	spf_to_flat(s, coef);
	*rm = s->rm;
	if (water_mode == 2) {
		free(sa_use);
	}
	sxs_spf_full_free(s);
	mol_atom_group_free(ag);
	return 0;
}

/* Per-atom form factors the way src/pdb2spf.c:66,86-87 resolves them. ff[3*i+{0,1}] = vacuum, dummy; ff[3*i+2] = h2o zero_ff. */
int ref_form_factors(const char *map_path, int natoms, const char *const *res, const char *const *atm, double *ff)
{
	struct saxs_form_factor_table *t = default_ff_table(map_path);
	double *xyz = calloc(3 * (size_t)natoms, sizeof(double));
	double *rad = calloc((size_t)natoms, sizeof(double));
	struct mol_atom_group *ag = make_group(natoms, xyz, res, atm, rad);
	int bad = 0;
	for (int i = 0; i < natoms; i++) {
		const struct saxs_form_factor *f = get_ff(t, ag, (size_t)i);
		if (f == NULL) {
			bad++;
			ff[3 * i] = ff[3 * i + 1] = 0.0;
		} else {
			ff[3 * i] = f->vacuum_ff;
			ff[3 * i + 1] = f->dummy_ff;
		}
		ff[3 * i + 2] = t->factors[s_OH2].zero_ff;
	}
	mol_atom_group_free(ag);
	free(xyz);
	free(rad);
	return bad;
}

/* Load a PDB the way tools/correlate.c:83-105 does.  centre: 0 none, 1 centre of extrema (receptor), 2 centroid (ligand).
 * Returns natoms; fills up to cap atoms.  names are written as 8-byte records. */
int ref_load_pdb(const char *pdb_path, const char *prm_path, int centre, int cap, double *xyz, double *radius,
                 char *res8, char *atm8, double *shift)
{
	struct mol_prms *prms = mol_prms_read((char *)prm_path);
	struct mol_atom_group *ag = mol_read_pdb((char *)pdb_path);
	if (prms == NULL || ag == NULL) {
		return -1;
	}
	mol_atom_group_add_prms(ag, prms);
	struct mol_vector3 c = {0, 0, 0};
	if (centre == 1) {
		center_of_extrema(&c, ag);
	} else if (centre == 2) {
		centroid(&c, ag);
	}
	MOL_VEC_MULT_SCALAR(c, c, -1.0);
	if (centre != 0) {
		mol_atom_group_translate(ag, &c);
	}
	if (shift != NULL) {
		shift[0] = c.X; shift[1] = c.Y; shift[2] = c.Z;
	}
	int n = (int)ag->natoms;
	for (int i = 0; i < n && i < cap; i++) {
		xyz[3 * i] = ag->coords[i].X;
		xyz[3 * i + 1] = ag->coords[i].Y;
		xyz[3 * i + 2] = ag->coords[i].Z;
		radius[i] = ag->vdw_radius[i];
		memset(res8 + 8 * i, 0, 8);
		memset(atm8 + 8 * i, 0, 8);
		strncpy(res8 + 8 * i, ag->residue_name[i], 7);
		strncpy(atm8 + 8 * i, ag->atom_name[i], 7);
	}
	mol_atom_group_free(ag);
	mol_prms_free(prms);
	return n;
}

/* scoring_helper + sxs_opt_params_init, src/min_saxs.c:108-124,353-389. a_out[6*qnum], scal[3] = {rm, mult, peak}. */
void ref_opt_params(const double *exp_q, const double *exp_in, const double *exp_err, int exp_n,
                    const double *qvals, int qnum, double rm, double *a_out, double *scal)
{
	double *eq = malloc(sizeof(double) * (exp_n + 1));
	memcpy(eq, exp_q, sizeof(double) * exp_n);
	eq[exp_n] = -1.0; /* sentinel: the reference's scan (src/min_saxs.c:369) is unguarded */
	struct sxs_profile *exp = sxs_profile_create(eq, exp_n, 0);
	memcpy(exp->in, exp_in, sizeof(double) * exp_n);
	memcpy(exp->err, exp_err, sizeof(double) * exp_n);
	struct sxs_opt_params *p = sxs_opt_params_create(exp, (double *)qvals, qnum, rm);
	memcpy(a_out, p->a, sizeof(double) * 6 * qnum);
	scal[0] = p->rm;
	scal[1] = p->mult;
	scal[2] = p->peak;
	sxs_opt_params_free(p);
	sxs_profile_free(exp);
	free(eq);
}

/* sxs_profile_read, src/profile.c:87-124: returns number of rows, fills up to cap. */
int ref_profile_read(const char *path, int cap, double *q, double *in, double *err)
{
	struct sxs_profile *p = sxs_profile_read((char *)path);
	if (p == NULL) {
		return -1;
	}
	int n = p->qnum;
	for (int i = 0; i < n && i < cap; i++) {
		q[i] = p->qvals[i];
		in[i] = p->in[i];
		err[i] = p->err[i];
	}
	free(p->qvals);
	sxs_profile_free(p);
	return n;
}

static struct sxs_opt_params *params_from_flat(const double *a, int qnum, const double *scal)
{
	struct sxs_opt_params *p = calloc(1, sizeof(struct sxs_opt_params));
	p->a = malloc(sizeof(double) * 6 * qnum);
	memcpy(p->a, a, sizeof(double) * 6 * qnum);
	p->rm = scal[0];
	p->mult = scal[1];
	p->peak = scal[2];
	return p;
}

/* sxs_compute_saxs_scores, src/fftsaxs.c:608-986. */
void ref_scores(double *scores, double *c1, double *c2, const int *index_list, int nout,
                const double *coefA, const double *coefB, const double *a, const double *scal,
                const double *qvals, int qnum, const double *zvals, int znum, int L, int skip)
{
	struct sxs_spf_full *A = flat_to_spf(coefA, qnum, L, 0.0);
	struct sxs_spf_full *B = flat_to_spf(coefB, qnum, L, 0.0);
	struct sxs_opt_params *p = params_from_flat(a, qnum, scal);
	sxs_compute_saxs_scores(scores, c1, c2, (int *)index_list, nout, A, B, p, (double *)qvals, qnum,
	                        (double *)zvals, znum, L, skip);
	sxs_opt_params_free(p);
	sxs_spf_full_free(A);
	sxs_spf_full_free(B);
}

/* sxs_fit_params on explicit cross terms (src/min_saxs.c:153-259): x[npts][6][qnum] in the order
 * VV,VD,VW,DD,DW,WW (unscaled, as fill_const/fill_var leave them).  rescale != 0 applies the peak
 * rescale of sxs_fit_params first.  out[npts][4] = {score, c1, c2, nfg (isave[33])}. */
void ref_fit(const double *x, int npts, const double *a, const double *scal, const double *qvals, int qnum,
             int rescale, double *out)
{
	struct sxs_opt_params *p = params_from_flat(a, qnum, scal);
	struct sxs_profile *prof = sxs_profile_create((double *)qvals, qnum, 1);
	int mask = 1;
	for (int i = 0; i < npts; i++) {
		const double *xi = x + (size_t)i * 6 * qnum;
		memcpy(prof->VV, xi + 0 * qnum, sizeof(double) * qnum);
		memcpy(prof->VD, xi + 1 * qnum, sizeof(double) * qnum);
		memcpy(prof->VW, xi + 2 * qnum, sizeof(double) * qnum);
		memcpy(prof->DD, xi + 3 * qnum, sizeof(double) * qnum);
		memcpy(prof->DW, xi + 4 * qnum, sizeof(double) * qnum);
		memcpy(prof->WW, xi + 5 * qnum, sizeof(double) * qnum);
		if (rescale) {
			sxs_fit_params(&prof, p, &mask, 1);
		} else {
			sxs_lbfgs_fitting(prof, p);
		}
		out[4 * i + 0] = prof->score;
		out[4 * i + 1] = prof->c1;
		out[4 * i + 2] = prof->c2;
		out[4 * i + 3] = (double)p->isave[33];
	}
	sxs_profile_free(prof);
	sxs_opt_params_free(p);
}

/* sxs_profile_from_spf, src/profile.c:186-233. */
void ref_profile_from_spf(const double *coef, int qnum, int L, double rm, const double *qvals, double c1, double c2,
                          double *in, double *err)
{
	struct sxs_spf_full *s = flat_to_spf(coef, qnum, L, rm);
	struct sxs_profile *p = sxs_profile_create((double *)qvals, qnum, 0);
	sxs_profile_from_spf(p, s, c1, c2);
	memcpy(in, p->in, sizeof(double) * qnum);
	memcpy(err, p->err, sizeof(double) * qnum);
	sxs_profile_free(p);
	sxs_spf_full_free(s);
}

/* sxs_spf2fitted_profile, src/min_saxs.c:507-511 (tests/saxs_test.c:365-407). out3 = {score,c1,c2}. */
void ref_fitted_profile(const double *coef, int qnum, int L, const double *a, const double *scal, const double *qvals,
                        double *in, double *err, double *out3)
{
	struct sxs_spf_full *s = flat_to_spf(coef, qnum, L, scal[0]);
	struct sxs_opt_params *p = params_from_flat(a, qnum, scal);
	struct sxs_profile *prof = sxs_profile_create((double *)qvals, qnum, 1);
	sxs_spf2fitted_profile(prof, s, p);
	memcpy(in, prof->in, sizeof(double) * qnum);
	memcpy(err, prof->err, sizeof(double) * qnum);
	out3[0] = prof->score;
	out3[1] = prof->c1;
	out3[2] = prof->c2;
	sxs_profile_free(prof);
	sxs_opt_params_free(p);
	sxs_spf_full_free(s);
}

/* sxs_ft2euler, src/index.c:38-75. tv[3], rm[9] row-major, ref_lig[3] -> out[6] = z,b1,g1,a2,b2,g2 */
void ref_ft2euler(const double *tv, const double *rm, const double *ref_lig, double *out)
{
	struct mol_vector3 t = {tv[0], tv[1], tv[2]}, r = {ref_lig[0], ref_lig[1], ref_lig[2]};
	struct mol_matrix3 m = {rm[0], rm[1], rm[2], rm[3], rm[4], rm[5], rm[6], rm[7], rm[8]};
	struct sxs_euler e;
	sxs_ft2euler(&e, &t, &m, &r);
	out[0] = e.z; out[1] = e.b1; out[2] = e.g1; out[3] = e.a2; out[4] = e.b2; out[5] = e.g2;
}

void ref_ft_file2euler_file(const char *eu, const char *ft, const char *rm, const double *ref_lig)
{
	struct mol_vector3 r = {ref_lig[0], ref_lig[1], ref_lig[2]};
	sxs_ft_file2euler_file(eu, ft, rm, &r);
}

/* generate_d_array, src/borrowed.c:243-313: out[(L+1)*(2L+1)^2] */
void ref_wigner_d(int L, double beta, double *out)
{
	struct d_array *d = generate_d_array(L, beta);
	memcpy(out, d->data, sizeof(double) * (L + 1) * (2 * L + 1) * (2 * L + 1));
	deallocate_d_array(d);
}

/* wigner_3j_symbol_arb, src/borrowed.c:86-222 */
double ref_wigner_3j(int j1, int j2, int j3, int m1, int m2, int m3)
{
	int mx = j1 + j2 + j3 + 1;
	mpf_t *fact = generate_desc_fact_array_arb(mx);
	mpf_t w;
	mpf_init(w);
	wigner_3j_symbol_arb(w, j1, j2, j3, m1, m2, m3, (const mpf_t *)fact);
	double v = mpf_get_d(w);
	mpf_clear(w);
	for (int i = 0; i <= mx; i++) {
		mpf_clear(fact[i]);
	}
	free(fact);
	return v;
}
