/* Small helpers of the path; interface of src/saxs_utils.h. */
#ifndef FMFTSAXS_SAXS_UTILS_H
#define FMFTSAXS_SAXS_UTILS_H
#include "common.h"
#include "mol2/atom_group.h"
#include "mol2/sasa.h"
#include "mol2/transform.h"
#ifdef __cplusplus
extern "C" {
#endif
double mol_atom_group_max_dist(const struct mol_atom_group *ag);
double mol_atom_group_average_radius(const struct mol_atom_group *ag);
/* q[i] = q[i-1] + step, accumulated (src/saxs_utils.c:50-63). */
double *sxs_mkarray(double begin, double end, int qnum);
void sxs_fill_active_rotation_matrix(struct mol_matrix3 *rm, double alpha, double beta, double gamma);
void sxs_mult_rot_mats(struct mol_matrix3 *c, struct mol_matrix3 *a, struct mol_matrix3 *b);
/* exposed fraction of each atom's solvent-expanded sphere (src/saxs_utils.c:30-48). */
void sxs_faccs(double *fractional_sa, const struct mol_atom_group *ag, double r_solv);
#ifdef __cplusplus
}
#endif
#endif
