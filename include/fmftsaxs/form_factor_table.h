/* Zero-angle form factors per atom type and the (residue, atom) -> type map;
 * interface of src/form_factor_table.h restricted to what the scoring path calls. */
#ifndef FMFTSAXS_FORM_FACTOR_TABLE_H
#define FMFTSAXS_FORM_FACTOR_TABLE_H
#include "common.h"
#include "mol2/atom_group.h"
#ifdef __cplusplus
extern "C" {
#endif

enum saxs_ff_type {
	H, He,
	Li, Be, B, C, N, O, F, Ne,
	Na, Mg, Al, Si, P, S, Cl, Ar,
	K, Ca, Cr, Mn, Fe, Co, Ni, Cu, Zn, Se, Br,
	Io, Ir, Pt, Au, Hg, SINGLE_ATOM_SIZE = 34,
	CH = 34, CH2 = 35, CH3 = 36, NH = 37, NH2 = 38, NH3 = 39, OH = 40, s_OH2 = 41, SH = 42, PO4 = 43,
	HEAVY_ATOM_SIZE = 43, s_UNK = 99
};

struct saxs_form_factor {
	double zero_ff;
	double vacuum_ff;
	double dummy_ff;
	double *values;
};

struct saxs_ff_map_entry {
	char residue_name[8];
	char atom_name[8];
	enum saxs_ff_type type;
};

struct saxs_form_factor_table {
	struct saxs_form_factor factors[HEAVY_ATOM_SIZE];
	struct saxs_ff_map_entry *map; /* open-addressing hash, `map_cap` slots */
	size_t map_cap;
	size_t map_len;
};

/* Process-wide singleton built from the mapping file on first call (src/form_factor_table.c:138-188). */
struct saxs_form_factor_table *default_ff_table(const char *type_mapping_file);
const char *ff_type_to_string(enum saxs_ff_type type);
/* NULL when the pair is unmapped or maps to an unknown type (src/form_factor_table.c:294-339). */
const struct saxs_form_factor *get_ff(const struct saxs_form_factor_table *table,
                                      const struct mol_atom_group *ag, size_t atom_index);
#ifdef __cplusplus
}
#endif
#endif
