/* mini-libmol2: the subset of the (un-vendored, unpinned) libmol2 API that the
 * FFT-SAXS scoring path touches.  libmol2 is an external dependency of the
 * reference (CMakeLists.txt:104 `find_package(mol2 REQUIRED)`), its sources are
 * not in /root/reference, so this is a restatement of its documented
 * behaviour, pinned by tests/data/ref_spf (header rm=1.5823, V/D columns).
 *
 * Call sites served: tools/correlate.c:74-105,130-134, tools/single_saxs.c:40-70,
 * src/saxs_utils.c:30-48 (accs), src/index.c:95-96 (matrix list),
 * src/form_factor_table.c:170,185 (is_whitespace_line, free_if_not_null).
 */
#ifndef FMFTSAXS_MOL2_MINI_H
#define FMFTSAXS_MOL2_MINI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

struct mol_vector3 {
	double X, Y, Z;
};

struct mol_matrix3 {
	double m11, m12, m13;
	double m21, m22, m23;
	double m31, m32, m33;
};

struct mol_matrix3_list {
	size_t size;
	struct mol_matrix3 *members;
};

#define MOL_VEC_ADD(DST, U, V) do { (DST).X = (U).X + (V).X; (DST).Y = (U).Y + (V).Y; (DST).Z = (U).Z + (V).Z; } while (0)
#define MOL_VEC_SUB(DST, U, V) do { (DST).X = (U).X - (V).X; (DST).Y = (U).Y - (V).Y; (DST).Z = (U).Z - (V).Z; } while (0)
#define MOL_VEC_MULT_SCALAR(DST, U, A) do { (DST).X = (U).X * (A); (DST).Y = (U).Y * (A); (DST).Z = (U).Z * (A); } while (0)
#define MOL_VEC_SQ_NORM(U) ((U).X * (U).X + (U).Y * (U).Y + (U).Z * (U).Z)
#define MOL_VEC_EUCLIDEAN_DIST_SQ(U, V) \
	(((U).X - (V).X) * ((U).X - (V).X) + ((U).Y - (V).Y) * ((U).Y - (V).Y) + ((U).Z - (V).Z) * ((U).Z - (V).Z))

struct mol_atom_group {
	size_t natoms;
	struct mol_vector3 *coords;
	double *vdw_radius;
	double *charge;
	char **atom_name;    /* 4-character PDB field, blanks kept */
	char **residue_name; /* up to 4 characters, blanks kept */
};

struct mol_prm_atom {
	int id;
	char typemaj[8]; /* residue */
	char typemin[8]; /* atom */
	int subid;
	double r;
	double q;
};

struct mol_prms {
	size_t natoms;
	struct mol_prm_atom *atoms;
};

/* PDB / parameter I/O */
struct mol_atom_group *mol_read_pdb(const char *path);
struct mol_prms *mol_prms_read(const char *path);
void mol_prms_free(struct mol_prms *prms);
/* Assigns vdw_radius/charge by (residue, atom) lookup; an unmatched atom is an error (exit). */
void mol_atom_group_add_prms(struct mol_atom_group *ag, const struct mol_prms *prms);

/* Atom-group helpers */
struct mol_atom_group *mol_atom_group_create(size_t natoms);
void mol_atom_group_free(struct mol_atom_group *ag);
struct mol_atom_group *mol_atom_group_join(const struct mol_atom_group *a, const struct mol_atom_group *b);
void mol_atom_group_translate(struct mol_atom_group *ag, const struct mol_vector3 *t);
/* dst->coords[i] = R * src->coords[i] + t (libmol2 transform.h; used by tools/score_ft_naive.c:161-164) */
void mol_atom_group_move_in_copy(const struct mol_atom_group *src, struct mol_atom_group *dst,
                                 const struct mol_matrix3 *rotation, const struct mol_vector3 *translation);
void centroid(struct mol_vector3 *c, const struct mol_atom_group *ag);
void center_of_extrema(struct mol_vector3 *c, const struct mol_atom_group *ag);

/* Solvent accessible surface.  `as[i]` receives, for cont_acc != 0, the
 * "contact" area  4*pi*r_i^2 * (exposed fraction of the sphere of radius
 * r_i + r_solv), which is what src/saxs_utils.c:38-47 divides by 4*pi*r_i^2. */
void accs(double *as, const struct mol_atom_group *ag, double r_solv, short cont_acc);

/* Rotation-matrix list: one row-major 3x3 per line, optional leading integer. */
struct mol_matrix3_list *mol_matrix3_list_from_file(const char *path);
void mol_matrix3_list_free(struct mol_matrix3_list *list);

bool is_whitespace_line(const char *line);
#define free_if_not_null(p) do { if ((p) != NULL) free(p); } while (0)

#ifdef __cplusplus
}
#endif

#endif
