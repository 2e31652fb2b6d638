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

#ifdef __cplusplus
}
#endif
#endif
