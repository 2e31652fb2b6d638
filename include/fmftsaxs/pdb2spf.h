/* Spherical-harmonic amplitude expansion of an atom group; interface of src/pdb2spf.h.
 * The per-atom sums run on the GPU (kernel K1, sxs_cuda_expand); these containers are the
 * host-side view the reference API exposes. */
#ifndef FMFTSAXS_PDB2SPF_H
#define FMFTSAXS_PDB2SPF_H
#include "common.h"
#include "borrowed.h"
#include "saxs_utils.h"
#include "form_factor_table.h"
#include "sfbessel.h"
#ifdef __cplusplus
extern "C" {
#endif

/* (L+1)^2 complex coefficients A_lm at one q (src/pdb2spf.h:30-35). */
struct sxs_spf_sing {
	int L;
	double *re;
	double *im;
};

/* Vacuum / excluded-volume ("dummy") / hydration-shell amplitudes over the whole q grid (src/pdb2spf.h:41-53). */
struct sxs_spf_full {
	int L;
	int qnum;
	double rm; /* mean van der Waals radius, zero-radius hydrogens included */
	struct sxs_spf_sing **V;
	struct sxs_spf_sing **D;
	struct sxs_spf_sing **W;
};

/* water == 1 adds the hydration term with SASA fractions from sxs_faccs(…, 1.4) (src/pdb2spf.c:154-175). */
struct sxs_spf_full *atom_grp2spf(struct mol_atom_group *ag, struct saxs_form_factor_table *ff_table,
                                  double *qvals, int qnum, int L, int water);
/* Explicit SASA fractions (NULL: no hydration term); resets `spf_coefs` first (src/pdb2spf.c:24-152). */
void atom_grp2spf_inplace(struct sxs_spf_full *spf_coefs, struct mol_atom_group *ag,
                          struct saxs_form_factor_table *ff_table, double *qvals, int qnum, int L,
                          double *saxs_sa);

struct sxs_spf_sing *sxs_spf_sing_create(int L);
void sxs_spf_sing_init(struct sxs_spf_sing *s, int L);
void sxs_spf_sing_free(struct sxs_spf_sing *s);
void sxs_spf_sing_destroy(struct sxs_spf_sing *s);
struct sxs_spf_full *sxs_spf_full_create(int L, int qnum);
void sxs_spf_full_init(struct sxs_spf_full *s, int L, int qnum);
void sxs_spf_full_free(struct sxs_spf_full *s);
void sxs_spf_full_destroy(struct sxs_spf_full *s);
void sxs_spf_full_reset(struct sxs_spf_full *s);
/* Text dump "L qnum rm" + one "% .4e" x6 row per (q, l, m) (src/pdb2spf.c:267-350). */
void sxs_spf_full_write(char *path, struct sxs_spf_full *s);
struct sxs_spf_full *sxs_spf_full_fread(FILE *f);
struct sxs_spf_full *sxs_spf_full_read(char *path);

/* Flat packing used by the CUDA C-ABI: coef[((c*qnum + q)*(L+1)^2 + lm)*2 + {re,im}], c = V,D,W. */
void sxs_spf_full_pack(const struct sxs_spf_full *s, double *coef);
void sxs_spf_full_unpack(struct sxs_spf_full *s, const double *coef);

#ifdef __cplusplus
}
#endif
#endif
