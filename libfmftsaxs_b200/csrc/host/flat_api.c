/* Flat-array adapters over the reference-shaped API (see sxs_flat.h). */
#define _POSIX_C_SOURCE 200809L
#include "sxs_flat.h"

#include "fftsaxs.h"
#include "index.h"
#include "min_saxs.h"
#include "pdb2spf.h"
#include "profile.h"
#include "sxs_tables.h"
#include "mol2/pdb.h"
#include "mol2/prms.h"

static struct sxs_spf_full *spf_from_flat(const double *coef, int qnum, int L, double rm)
{
	struct sxs_spf_full *s = sxs_spf_full_create(L, qnum);
	s->rm = rm;
	sxs_spf_full_unpack(s, coef);
	return s;
}

static struct mol_atom_group *group_from_flat(int natoms, const double *xyz, const char *const *res,
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

int sxs_flat_load_pdb(const char *pdb_path, const char *prm_path, int centre, int cap, double *xyz, double *radius,
                      char *res8, char *atm8, double *shift3)
{
	struct mol_prms *prms = mol_prms_read(prm_path);
	struct mol_atom_group *ag = mol_read_pdb(pdb_path);
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
	if (shift3 != NULL) {
		shift3[0] = c.X; shift3[1] = c.Y; shift3[2] = c.Z;
	}
	const int n = (int)ag->natoms;
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

int sxs_flat_expand(const char *map_path, int natoms, const double *xyz, const char *const *res,
                    const char *const *atm, const double *radius, double *sa, int water_mode, const double *qvals,
                    int qnum, int L, double *coef, double *rm)
{
	struct saxs_form_factor_table *ff = default_ff_table(map_path);
	struct mol_atom_group *ag = group_from_flat(natoms, xyz, res, atm, radius);
	struct sxs_spf_full *s = sxs_spf_full_create(L, qnum);
	double *sa_use = NULL;
	if (water_mode == 1) {
		sa_use = sa;
	} else if (water_mode == 2) {
		sa_use = (double *)calloc((size_t)natoms, sizeof(double));
		sxs_faccs(sa_use, ag, 1.4);
		if (sa != NULL) {
			memcpy(sa, sa_use, sizeof(double) * natoms);
		}
	}
	atom_grp2spf_inplace(s, ag, ff, (double *)qvals, qnum, L, sa_use);
	sxs_spf_full_pack(s, coef);
	*rm = s->rm;
	if (water_mode == 2) {
		free(sa_use);
	}
	sxs_spf_full_free(s);
	mol_atom_group_free(ag);
	return 0;
}

int sxs_flat_profile_read(const char *path, int cap, double *q, double *in, double *err)
{
	struct sxs_profile *p = sxs_profile_read((char *)path);
	if (p == NULL) {
		return -1;
	}
	const int n = p->qnum;
	for (int i = 0; i < n && i < cap; i++) {
		q[i] = p->qvals[i];
		in[i] = p->in[i];
		err[i] = p->err[i];
	}
	free(p->qvals);
	sxs_profile_free(p);
	return n;
}

void sxs_flat_opt_params(const double *exp_q, const double *exp_in, const double *exp_err, int exp_n,
                         const double *qvals, int qnum, double rm, double *a_out, double *scal3)
{
	double *eq = (double *)malloc(sizeof(double) * exp_n);
	memcpy(eq, exp_q, sizeof(double) * exp_n);
	struct sxs_profile *exp = sxs_profile_create(eq, exp_n, 0);
	memcpy(exp->in, exp_in, sizeof(double) * exp_n);
	memcpy(exp->err, exp_err, sizeof(double) * exp_n);
	struct sxs_opt_params *p = sxs_opt_params_create(exp, (double *)qvals, qnum, rm);
	memcpy(a_out, p->a, sizeof(double) * 6 * qnum);
	scal3[0] = p->rm;
	scal3[1] = p->mult;
	scal3[2] = p->peak;
	sxs_opt_params_free(p);
	sxs_profile_free(exp);
	free(eq);
}

static struct sxs_opt_params *params_from_flat(const double *a, int qnum, const double *scal3)
{
	struct sxs_opt_params *p = (struct sxs_opt_params *)calloc(1, sizeof(*p));
	p->a = (double *)malloc(sizeof(double) * 6 * qnum);
	memcpy(p->a, a, sizeof(double) * 6 * qnum);
	p->rm = scal3[0];
	p->mult = scal3[1];
	p->peak = scal3[2];
	return p;
}

void sxs_flat_scores(double *scores, double *c1, double *c2, const int *index_list, int nout, const double *coefA,
                     const double *coefB, const double *a, const double *scal3, const double *qvals, int qnum,
                     const double *zvals, int znum, int L, int skip)
{
	struct sxs_spf_full *A = spf_from_flat(coefA, qnum, L, 0.0);
	struct sxs_spf_full *B = spf_from_flat(coefB, qnum, L, 0.0);
	struct sxs_opt_params *p = params_from_flat(a, qnum, scal3);
	sxs_compute_saxs_scores(scores, c1, c2, (int *)index_list, nout, A, B, p, (double *)qvals, qnum, (double *)zvals,
	                        znum, L, skip);
	sxs_opt_params_free(p);
	sxs_spf_full_free(A);
	sxs_spf_full_free(B);
}

void sxs_flat_scores64(double *scores, double *c1, double *c2, const long long *index_list, long long nout,
                       const double *coefA, const double *coefB, const double *a, const double *scal3,
                       const double *qvals, int qnum, const double *zvals, int znum, int L, int skip)
{
	struct sxs_spf_full *A = spf_from_flat(coefA, qnum, L, 0.0);
	struct sxs_spf_full *B = spf_from_flat(coefB, qnum, L, 0.0);
	struct sxs_opt_params *p = params_from_flat(a, qnum, scal3);
	sxs_compute_saxs_scores64(scores, c1, c2, index_list, nout, A, B, p, (double *)qvals, qnum, (double *)zvals, znum,
	                          L, skip);
	sxs_opt_params_free(p);
	sxs_spf_full_free(A);
	sxs_spf_full_free(B);
}

void sxs_flat_profile_from_spf(const double *coef, int qnum, int L, double rm, const double *qvals, double c1,
                               double c2, double *in, double *err)
{
	struct sxs_spf_full *s = spf_from_flat(coef, qnum, L, rm);
	struct sxs_profile *p = sxs_profile_create((double *)qvals, qnum, 0);
	sxs_profile_from_spf(p, s, c1, c2);
	memcpy(in, p->in, sizeof(double) * qnum);
	memcpy(err, p->err, sizeof(double) * qnum);
	sxs_profile_free(p);
	sxs_spf_full_free(s);
}

void sxs_flat_fitted_profile(const double *coef, int qnum, int L, const double *a, const double *scal3,
                             const double *qvals, double *in, double *err, double *out3)
{
	struct sxs_spf_full *s = spf_from_flat(coef, qnum, L, scal3[0]);
	struct sxs_opt_params *p = params_from_flat(a, qnum, scal3);
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

void sxs_flat_ft2euler(const double *tv, const double *rm, const double *ref_lig, double *out6)
{
	struct mol_vector3 t = {tv[0], tv[1], tv[2]}, r = {ref_lig[0], ref_lig[1], ref_lig[2]};
	struct mol_matrix3 m = {rm[0], rm[1], rm[2], rm[3], rm[4], rm[5], rm[6], rm[7], rm[8]};
	struct sxs_euler e;
	sxs_ft2euler(&e, &t, &m, &r);
	out6[0] = e.z; out6[1] = e.b1; out6[2] = e.g1; out6[3] = e.a2; out6[4] = e.b2; out6[5] = e.g2;
}

void sxs_flat_euler_to_index(const double *euler, const int *z_index, int n, int L, int *index)
{
	for (int i = 0; i < n; i++) {
		struct sxs_euler e = {euler[6 * i], euler[6 * i + 1], euler[6 * i + 2], euler[6 * i + 3], euler[6 * i + 4],
		                      euler[6 * i + 5]};
		index[i] = sxs_euler_to_index(&e, z_index[i], L);
	}
}

long long sxs_flat_ft_rows_to_indices(long long *index, int *ft_id, int *order, const int *rot_id, const double *trans,
                                      long long n, const double *rot_mats, int nrot, const double *ref_lig,
                                      const double *zvals, int znum, int L, int nthreads)
{
	struct mol_matrix3_list rots;
	rots.size = (size_t)nrot;
	rots.members = (struct mol_matrix3 *)malloc(sizeof(struct mol_matrix3) * (size_t)(nrot > 0 ? nrot : 1));
	CHECK_PTR(rots.members);
	for (int i = 0; i < nrot; i++) {
		const double *m = rot_mats + 9 * (size_t)i;
		struct mol_matrix3 r = {m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]};
		rots.members[i] = r;
	}
	struct mol_vector3 rl = {ref_lig[0], ref_lig[1], ref_lig[2]};
	const long long kept = sxs_ft_rows_to_indices64(index, ft_id, order, rot_id, trans, n, &rots, &rl, zvals, znum, L,
	                                                nthreads);
	free(rots.members);
	return kept;
}

void sxs_flat_ft_file2euler_file(const char *eu_path, const char *ft_path, const char *rm_path, const double *ref_lig)
{
	struct mol_vector3 rl = {ref_lig[0], ref_lig[1], ref_lig[2]};
	sxs_ft_file2euler_file(eu_path, ft_path, rm_path, &rl);
}

void sxs_flat_wigner_d(int L, double beta, double *out)
{
	struct d_array *d = generate_d_array(L, beta);
	memcpy(out, d->data, sizeof(double) * (L + 1) * (2 * L + 1) * (2 * L + 1));
	deallocate_d_array(d);
}

void sxs_flat_tables(int L, double *dsymb, double *dwig, double *twiddle)
{
	const struct sxs_l_tables *t = sxs_l_tables_get(L);
	const size_t nb = L + 1, N = 2 * L + 1;
	if (dsymb) memcpy(dsymb, t->dsymb, sizeof(double) * nb * nb * nb * N);
	if (dwig) memcpy(dwig, t->dwig, sizeof(double) * nb * nb * N * N);
	if (twiddle) memcpy(twiddle, t->twiddle, sizeof(double) * 2 * N);
}

void sxs_flat_bessel_table(const double *zvals, int znum, const double *qvals, int qnum, int L, double *bessel)
{
	sxs_fill_bessel_table(bessel, zvals, znum, qvals, qnum, L);
}
