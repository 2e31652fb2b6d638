/* Zero-angle form factors and the (residue, atom) -> scatterer type map.
 *
 * On the scoring path form factors are q-independent constants (src/pdb2spf.c:40-43,86-87):
 * vacuum_ff and dummy_ff per atom, and the water factor zero_ff[OH2] scaled by the SASA fraction.
 * Values are the reference's tables (src/form_factor_table.c:8-49), listed here per type.
 */
#define _POSIX_C_SOURCE 200809L
#include "form_factor_table.h"

struct ff_row {
	const char *label; /* name used in the mapping file, NULL if the reference cannot parse one */
	double zero, vacuum, dummy;
};

/* order = enum saxs_ff_type */
static const struct ff_row k_rows[HEAVY_ATOM_SIZE] = {
	/* H   */ {"H", -0.720147, 0.999953, 1.7201},
	/* He  */ {"HE", -0.720228, 0.999872, 1.7201},
	/* Li  */ {NULL, 1.591, 2.99, 1.399},
	/* Be  */ {NULL, 2.591, 3.99, 1.399},
	/* B   */ {NULL, 3.591, 4.99, 1.399},
	/* C   */ {"C", 0.50824, 5.9992, 5.49096},
	/* N   */ {"N", 6.16294, 6.9946, 0.83166},
	/* O   */ {"O", 4.94998, 7.9994, 3.04942},
	/* F   */ {NULL, 7.591, 8.99, 1.399},
	/* Ne  */ {"NE", 6.993, 9.999, 3.006},
	/* Na  */ {"SOD+", 7.9864, 10.9924, 3.006},
	/* Mg  */ {"MG2+", 8.9805, 11.9865, 3.006},
	/* Al  */ {NULL, 9.984, 12.99, 3.006},
	/* Si  */ {NULL, 10.984, 13.99, 3.006},
	/* P   */ {"P", 13.0855, 14.9993, 1.91382},
	/* S   */ {"S", 9.36656, 15.9998, 6.63324},
	/* Cl  */ {NULL, 13.984, 16.99, 3.006},
	/* Ar  */ {NULL, 16.591, 17.99, 1.399},
	/* K   */ {"K", 15.984, 18.99, 3.006},
	/* Ca  */ {"CAL2+", 14.9965, 18.0025, 3.006},
	/* Cr  */ {NULL, 20.984, 23.99, 3.006},
	/* Mn  */ {NULL, 21.984, 24.99, 3.006},
	/* Fe  */ {"FE2+", 20.9946, 24.0006, 3.006},
	/* Co  */ {NULL, 23.984, 26.99, 3.006},
	/* Ni  */ {NULL, 24.984, 27.99, 3.006},
	/* Cu  */ {NULL, 25.984, 28.99, 3.006},
	/* Zn  */ {"ZN2+", 24.9936, 27.9996, 3.006},
	/* Se  */ {"SE", 30.9825, 33.99, 3.006},
	/* Br  */ {NULL, 31.984, 34.99, 3.006},
	/* I   */ {NULL, 49.16, 52.99, 3.83},
	/* Ir  */ {NULL, 70.35676, 76.99, 6.63324},
	/* Pt  */ {NULL, 71.35676, 77.99, 6.63324},
	/* Au  */ {"AU", 72.324, 78.9572, 6.63324},
	/* Hg  */ {NULL, 73.35676, 79.99, 6.63324},
	/* CH  */ {"CH", -0.211907, 6.99915, 7.21106},
	/* CH2 */ {"CH2", -0.932054, 7.99911, 8.93116},
	/* CH3 */ {"CH3", -1.6522, 8.99906, 10.6513},
	/* NH  */ {"NH", 5.44279, 7.99455, 2.55176},
	/* NH2 */ {"NH2", 4.72265, 8.99451, 4.27186},
	/* NH3 */ {"NH3", 4.0025, 9.99446, 5.99196},
	/* OH  */ {"OH", 4.22983, 8.99935, 4.76952},
	/* OH2 */ {NULL, 3.50968, 9.9993, 6.48962},
	/* SH  */ {"SH", 8.64641, 16.9998, 8.35334},
};

/* The reference's parser (src/form_factor_table.c:76-107) knows exactly the labelled rows above;
 * anything else (e.g. "PO4", "OH2") becomes s_UNK and get_ff() then yields NULL. */
static enum saxs_ff_type type_from_label(const char *name)
{
	for (int i = 0; i < HEAVY_ATOM_SIZE; i++) {
		if (k_rows[i].label != NULL && !strcmp(name, k_rows[i].label)) {
			return (enum saxs_ff_type)i;
		}
	}
	return s_UNK;
}

const char *ff_type_to_string(enum saxs_ff_type type)
{
	if ((int)type >= 0 && (int)type < HEAVY_ATOM_SIZE) {
		return k_rows[type].label;
	}
	return NULL;
}

static size_t key_hash(const char *res, const char *atm)
{
	size_t h = 1469598103934665603ULL;
	for (const char *p = res; *p; p++) { h = (h ^ (unsigned char)*p) * 1099511628211ULL; }
	h = (h ^ 0x1f) * 1099511628211ULL;
	for (const char *p = atm; *p; p++) { h = (h ^ (unsigned char)*p) * 1099511628211ULL; }
	return h;
}

static void map_put(struct saxs_form_factor_table *t, const char *res, const char *atm, enum saxs_ff_type type)
{
	size_t i = key_hash(res, atm) & (t->map_cap - 1);
	for (;;) {
		struct saxs_ff_map_entry *e = &t->map[i];
		if (e->residue_name[0] == '\0' && e->atom_name[0] == '\0') {
			strncpy(e->residue_name, res, 7);
			strncpy(e->atom_name, atm, 7);
			e->type = type;
			t->map_len++;
			return;
		}
		if (!strncmp(e->residue_name, res, 7) && !strncmp(e->atom_name, atm, 7)) {
			e->type = type; /* later rows override, like kh_put + assignment */
			return;
		}
		i = (i + 1) & (t->map_cap - 1);
	}
}

static const struct saxs_ff_map_entry *map_get(const struct saxs_form_factor_table *t, const char *res, const char *atm)
{
	size_t i = key_hash(res, atm) & (t->map_cap - 1);
	for (;;) {
		const struct saxs_ff_map_entry *e = &t->map[i];
		if (e->residue_name[0] == '\0' && e->atom_name[0] == '\0') {
			return NULL;
		}
		if (!strncmp(e->residue_name, res, 7) && !strncmp(e->atom_name, atm, 7)) {
			return e;
		}
		i = (i + 1) & (t->map_cap - 1);
	}
}

struct saxs_form_factor_table *default_ff_table(const char *type_mapping_file)
{
	static struct saxs_form_factor_table *table = NULL;
	if (table != NULL) {
		return table;
	}
	FILE *fp = fopen(type_mapping_file, "r");
	if (fp == NULL) {
		fprintf(stderr, "Could not open file %s\n", type_mapping_file);
		exit(1);
	}
	table = (struct saxs_form_factor_table *)calloc(1, sizeof(*table));
	CHECK_PTR(table);
	for (int i = 0; i < HEAVY_ATOM_SIZE; i++) {
		table->factors[i].zero_ff = k_rows[i].zero;
		table->factors[i].vacuum_ff = k_rows[i].vacuum;
		table->factors[i].dummy_ff = k_rows[i].dummy;
	}
	table->map_cap = 4096;
	table->map = (struct saxs_ff_map_entry *)calloc(table->map_cap, sizeof(struct saxs_ff_map_entry));
	CHECK_PTR(table->map);

	char *line = NULL;
	size_t len = 0;
	while (getline(&line, &len, fp) != -1) {
		if (line[0] == '#' || is_whitespace_line(line)) {
			continue;
		}
		char res[8] = {0}, atm[8] = {0}, type_name[16] = {0};
		if (sscanf(line, "%7s %7s %15s", res, atm, type_name) != 3) {
			continue;
		}
		if (table->map_len * 2 >= table->map_cap) {
			ERROR_MSG("form-factor mapping file too large");
		}
		map_put(table, res, atm, type_from_label(type_name));
	}
	free(line);
	fclose(fp);
	return table;
}

static void strip_to(char *dst, const char *src)
{
	size_t n = 0;
	while (*src && *src == ' ') {
		src++;
	}
	while (*src && *src != ' ' && n < 7) {
		dst[n++] = *src++;
	}
	dst[n] = '\0';
}

const struct saxs_form_factor *get_ff(const struct saxs_form_factor_table *table,
                                      const struct mol_atom_group *ag, size_t atom_index)
{
	char res[8], atm[8];
	strip_to(res, ag->residue_name[atom_index]);
	strip_to(atm, ag->atom_name[atom_index]);
	const struct saxs_ff_map_entry *e = map_get(table, res, atm);
	if (e == NULL) {
		fprintf(stderr, "mapping for (%s, %s) not found\n", res, atm);
		return NULL;
	}
	if (e->type == s_UNK) {
		printf("(%s, %s) maps to UNK\n", res, atm);
		return NULL;
	}
	return &table->factors[e->type];
}
