/* Grid indexing and (translation, rotation) -> Euler conversion; interface of src/index.h. */
#ifndef FMFTSAXS_INDEX_H
#define FMFTSAXS_INDEX_H
#include "common.h"
#include "saxs_utils.h"
#include "mol2/vector.h"
#include "mol2/matrix.h"
#include "mol2/lists.h"
#ifdef __cplusplus
extern "C" {
#endif

struct sxs_euler {
	double z, b1, g1, a2, b2, g2;
};

struct sxs_index {
	int z, b1, g1, a2, b2, g2;
};

size_t sxs_assemble_index(struct sxs_index *id, int nbeta, int L);
void sxs_disassemble_index(struct sxs_index *id, size_t index, int nbeta, int L);
void sxs_ft2euler(struct sxs_euler *euler, struct mol_vector3 *tv, struct mol_matrix3 *rm, struct mol_vector3 *ref_lig);
void sxs_ft_file2euler_file(const char *eu_path, const char *ft_path, const char *rm_path, struct mol_vector3 *ref_lig);

/* Snap one Euler row to the flat grid index exactly as tools/correlate.c:219-242 does
 * (a2, g2 reflected with the truncated pi, round-half-away, carry instead of modulo). */
int sxs_euler_to_index(const struct sxs_euler *euler, int z_index, int L);

/* In-memory hand-off replacing the Euler text file between sxs_ft_file2euler_file (src/index.c:77-121) and
 * the read-back / snapping loop of tools/correlate.c:202-251, with identical results: every row goes through
 * sxs_ft2euler, the six numbers are quantised exactly as the "% .3f" text round trip quantises them, the
 * row is kept when its z matches zvals[j] within 0.001, and its flat index is sxs_euler_to_index(.., j, L).
 * rot_id[i], trans[3*i..] = rotation index and translation of ft row i.  For each kept row k (input order):
 * index[k], ft_id[k] = rot_id, order[k] = i (the serial number counts every row).  Arrays need room for n
 * entries.  Rows are converted by `nthreads` host threads (<= 0: one per online core, at most 64).
 * Returns the number of kept rows. */
long long sxs_ft_rows_to_indices(int *index, int *ft_id, int *order, const int *rot_id, const double *trans,
                                 long long n, const struct mol_matrix3_list *rots, struct mol_vector3 *ref_lig,
                                 const double *zvals, int znum, int L, int nthreads);
/* same with 64-bit flat indices (needed from L = 20 on, see sxs_compute_saxs_scores64) */
long long sxs_ft_rows_to_indices64(long long *index, int *ft_id, int *order, const int *rot_id, const double *trans,
                                   long long n, const struct mol_matrix3_list *rots, struct mol_vector3 *ref_lig,
                                   const double *zvals, int znum, int L, int nthreads);
long long sxs_euler_to_index64(const struct sxs_euler *euler, int z_index, int L);

/* The whole file route of the tool in one threaded pass: reads the ft file once, writes the Euler side file
 * (byte-identical to sxs_ft_file2euler_file's; eu_path may be NULL) and returns, for every row whose quantised z lies
 * on zvals, what tools/correlate.c:202-251 would have read back from that file: *index (flat grid index, 64-bit),
 * *ft_id (rotation index), *order (row number counting every row); the three arrays are malloc'ed here.  Returns the
 * number of kept rows, or -1 without touching anything when the input is not plain enough for the fast parser (a row
 * with other than ten tokens, a token fscanf would split differently, inf/nan/hex numbers): the caller then takes the
 * slow route, which reports malformed files as the reference does. */
long long sxs_ft_file_to_indices(const char *eu_path, const char *ft_path, const char *rm_path, struct mol_vector3 *ref_lig,
                                 const double *zvals, int znum, int L, int nthreads, long long **index, int **ft_id,
                                 int **order);
/* Output rows of the tool, "%-6d\t%d\t%.3lf\t%.3lf\t%.3lf\n" (tools/correlate.c:369-375), formatted by nthreads threads. */
void sxs_write_score_rows(const char *path, long long n, const int *order, const int *ft_id, const double *score,
                          const double *c1, const double *c2, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
