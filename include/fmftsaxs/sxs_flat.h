/* Flat-array convenience entry points over the reference-shaped C API, for hosts that cannot build
 * the pointer-of-pointer containers (Python ctypes, Fortran, …).  Each one constructs the reference's
 * structs and calls the public function named in its comment — nothing here has its own numerics.
 * Coefficient layout: coef[((c*qnum + q)*(L+1)^2 + lm)*2 + {re,im}], c = V, D, W. */
#ifndef FMFTSAXS_SXS_FLAT_H
#define FMFTSAXS_SXS_FLAT_H
#ifdef __cplusplus
extern "C" {
#endif

/* mol_read_pdb + mol_atom_group_add_prms (+ centring: 0 none, 1 centre of extrema, 2 centroid);
 * names are returned as 8-byte zero-padded records.  Returns the atom count (or -1). */
int sxs_flat_load_pdb(const char *pdb_path, const char *prm_path, int centre, int cap, double *xyz, double *radius,
                      char *res8, char *atm8, double *shift3);

/* atom_grp2spf_inplace / atom_grp2spf.  water_mode 0: no hydration term; 1: use sa[]; 2: sxs_faccs(…,1.4),
 * fractions copied back to sa[] when non-NULL. */
int sxs_flat_expand(const char *map_path, int natoms, const double *xyz, const char *const *res,
                    const char *const *atm, const double *radius, double *sa, int water_mode, const double *qvals,
                    int qnum, int L, double *coef, double *rm);

/* sxs_profile_read */
int sxs_flat_profile_read(const char *path, int cap, double *q, double *in, double *err);
/* sxs_opt_params_create: a_out[6*qnum], scal3 = {rm, mult, peak} */
void sxs_flat_opt_params(const double *exp_q, const double *exp_in, const double *exp_err, int exp_n,
                         const double *qvals, int qnum, double rm, double *a_out, double *scal3);
/* sxs_compute_saxs_scores / sxs_compute_saxs_scores64 */
void sxs_flat_scores(double *scores, double *c1, double *c2, const int *index_list, int nout, const double *coefA,
                     const double *coefB, const double *a, const double *scal3, const double *qvals, int qnum,
                     const double *zvals, int znum, int L, int skip);
void sxs_flat_scores64(double *scores, double *c1, double *c2, const long long *index_list, long long nout,
                       const double *coefA, const double *coefB, const double *a, const double *scal3,
                       const double *qvals, int qnum, const double *zvals, int znum, int L, int skip);
/* sxs_profile_from_spf */
void sxs_flat_profile_from_spf(const double *coef, int qnum, int L, double rm, const double *qvals, double c1,
                               double c2, double *in, double *err);
/* sxs_spf2fitted_profile: out3 = {chi, c1, c2} */
void sxs_flat_fitted_profile(const double *coef, int qnum, int L, const double *a, const double *scal3,
                             const double *qvals, double *in, double *err, double *out3);
/* sxs_ft2euler: tv[3], rm[9] row-major, ref_lig[3] -> out6 = z, b1, g1, a2, b2, g2 */
void sxs_flat_ft2euler(const double *tv, const double *rm, const double *ref_lig, double *out6);
/* sxs_ft_rows_to_indices64 with the rotation set as rot_mats[nrot][9] (row-major) and ref_lig[3]; returns kept rows */
long long sxs_flat_ft_rows_to_indices(long long *index, int *ft_id, int *order, const int *rot_id, const double *trans,
                                      long long n, const double *rot_mats, int nrot, const double *ref_lig,
                                      const double *zvals, int znum, int L, int nthreads);
/* sxs_ft_file2euler_file with ref_lig[3] */
void sxs_flat_ft_file2euler_file(const char *eu_path, const char *ft_path, const char *rm_path, const double *ref_lig);
/* sxs_euler_to_index on Euler rows euler[n][6] with explicit z indices */
void sxs_flat_euler_to_index(const double *euler, const int *z_index, int n, int L, int *index);
/* host tables, for inspection: generate_d_array, sxs_wigner_3j, sxs_sbessel */
void sxs_flat_wigner_d(int L, double beta, double *out);
/* scoring-plan tables of one L (cached): copies of dsymb [(L+1)^3 (2L+1)], dwig [(L+1)^2 (2L+1)^2], twiddle [2(2L+1)] */
void sxs_flat_tables(int L, double *dsymb, double *dwig, double *twiddle);
/* j_p(q z) table as the scoring call builds it */
void sxs_flat_bessel_table(const double *zvals, int znum, const double *qvals, int qnum, int L, double *bessel);

/* Test access to the host partition of a pose list over z shards (csrc/host/partition.c; what
 * sxs_compute_saxs_scores does with several devices).  Exactly one of idx32 / idx64 is non-NULL; cell5 = grid points per
 * z step.  Fills z_lo/z_hi/rows per shard and, shard after shard, the positions and indices of the shards' rows
 * (pos_concat, sub_concat: nout entries each).  Returns the number of shards, negated when the list was left whole. */
int sxs_flat_partition_rows(const int *idx32, const long long *idx64, long long nout, long long cell5, int znum,
                            int nshard_max, int *z_lo, int *z_hi, long long *rows, long long *pos_concat,
                            long long *sub_concat);

#ifdef __cplusplus
}
#endif
#endif
