/* single_saxs — SAXS profile of a receptor + ligand pair as one rigid body at given (c1, c2).
 *
 *   single_saxs MAPPING_PRM ATOMPRM REC LIG C1 C2 L PROFILE_OUT
 *
 * Contract of the reference tool (tools/single_saxs.c:23-76): the two PDBs are joined, centred on the
 * joint centroid, expanded with the hydration term, and I(q) is written as "%.4f %.4f %.4f" rows on the
 * 50-point grid 0 … 0.5 Å^-1.  The expansion (K1) and the profile sum run on the device.
 */
#include "common.h"

#include "mol2/atom_group.h"
#include "mol2/pdb.h"
#include "mol2/prms.h"

#include "form_factor_table.h"
#include "min_saxs.h"
#include "pdb2spf.h"
#include "profile.h"
#include "saxs_utils.h"

enum { ARG_MAP = 1, ARG_PRM, ARG_REC, ARG_LIG, ARG_C1, ARG_C2, ARG_L, ARG_OUT, ARG_COUNT };

/* both structures with radii attached, as one atom group whose centroid is the origin */
static struct mol_atom_group *read_complex(const char *rec_path, const char *lig_path, const char *prm_path)
{
	struct mol_prms *prms = mol_prms_read(prm_path);
	if (prms == NULL) {
		ERROR_MSG("cannot read atom parameter file");
	}
	struct mol_atom_group *part[2] = {mol_read_pdb(rec_path), mol_read_pdb(lig_path)};
	for (int k = 0; k < 2; k++) {
		if (part[k] == NULL) {
			ERROR_MSG("cannot read PDB file");
		}
		mol_atom_group_add_prms(part[k], prms);
	}
	struct mol_atom_group *both = mol_atom_group_join(part[0], part[1]);
	mol_atom_group_free(part[0]);
	mol_atom_group_free(part[1]);
	mol_prms_free(prms);

	struct mol_vector3 shift;
	centroid(&shift, both);
	MOL_VEC_MULT_SCALAR(shift, shift, -1.0);
	mol_atom_group_translate(both, &shift);
	return both;
}

int main(int argc, char **argv)
{
	if (argc != ARG_COUNT) {
		fprintf(stderr, "Usage: single_saxs MAPPING_PRM ATOMPRM REC_PATH LIG_PATH C1 C2 L PROFILE_PATH\n");
		return EXIT_FAILURE;
	}
	const int qnum = 50;
	const int L = atoi(argv[ARG_L]);
	double *qvals = sxs_mkarray(0.0, 0.5, qnum);
	struct saxs_form_factor_table *ff = default_ff_table(argv[ARG_MAP]);
	struct mol_atom_group *both = read_complex(argv[ARG_REC], argv[ARG_LIG], argv[ARG_PRM]);

	SXS_PRINTF("Computing coefficients ...\n");
	struct sxs_spf_full *spf = atom_grp2spf(both, ff, qvals, qnum, L, 1);
	struct sxs_profile *profile = sxs_profile_create(qvals, qnum, 1);
	sxs_profile_from_spf(profile, spf, atof(argv[ARG_C1]), atof(argv[ARG_C2]));
	sxs_profile_write(argv[ARG_OUT], profile);
	SXS_PRINTF("Profile is written into %s\n", argv[ARG_OUT]);

	sxs_profile_free(profile);
	sxs_spf_full_free(spf);
	mol_atom_group_free(both);
	free(qvals);
	return EXIT_SUCCESS;
}
