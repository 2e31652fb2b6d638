/* single_saxs — SAXS profile of a receptor + ligand pair as one rigid body at given (c1, c2).
 *
 *   single_saxs MAPPING_PRM ATOMPRM REC LIG C1 C2 L PROFILE_OUT
 *
 * Contract of the reference tool (tools/single_saxs.c:23-76): the two PDBs are joined, centred on the
 * joint centroid, expanded with the hydration term, and I(q) is written as "%.4f %.4f %.4f" rows on the
 * 50-point grid 0 … 0.5 Å^-1.
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

int main(int argc, char **argv)
{
	if (argc != 9) {
		fprintf(stderr, "Usage: single_saxs MAPPING_PRM ATOMPRM REC_PATH LIG_PATH C1 C2 L PROFILE_PATH\n");
		return EXIT_FAILURE;
	}
	const double c1 = atof(argv[5]), c2 = atof(argv[6]);
	const int L = atoi(argv[7]);
	const int qnum = 50;
	double *qvals = sxs_mkarray(0.0, 0.5, qnum);

	struct mol_prms *prms = mol_prms_read(argv[2]);
	if (prms == NULL) {
		ERROR_MSG("cannot read atom parameter file");
	}
	struct saxs_form_factor_table *ff = default_ff_table(argv[1]);
	struct mol_atom_group *rec = mol_read_pdb(argv[3]);
	struct mol_atom_group *lig = mol_read_pdb(argv[4]);
	if (rec == NULL || lig == NULL) {
		ERROR_MSG("cannot read PDB file");
	}
	mol_atom_group_add_prms(rec, prms);
	mol_atom_group_add_prms(lig, prms);
	struct mol_atom_group *both = mol_atom_group_join(rec, lig);

	struct mol_vector3 com;
	centroid(&com, both);
	MOL_VEC_MULT_SCALAR(com, com, -1.0);
	mol_atom_group_translate(both, &com);

	SXS_PRINTF("Computing coefficients ...\n");
	struct sxs_profile *profile = sxs_profile_create(qvals, qnum, 1);
	struct sxs_spf_full *spf = atom_grp2spf(both, ff, qvals, qnum, L, 1);
	sxs_profile_from_spf(profile, spf, c1, c2);
	sxs_profile_write(argv[8], profile);
	SXS_PRINTF("Profile is written into %s\n", argv[8]);

	mol_atom_group_free(both);
	mol_atom_group_free(rec);
	mol_atom_group_free(lig);
	mol_prms_free(prms);
	sxs_profile_free(profile);
	sxs_spf_full_free(spf);
	free(qvals);
	return EXIT_SUCCESS;
}
