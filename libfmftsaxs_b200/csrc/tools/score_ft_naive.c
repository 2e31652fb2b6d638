/* score_ft_naive — real-space check of the FFT scoring path: every pose is built atom by atom, expanded and fitted.
 *
 *   score_ft_naive MAPPING_PRM ATOMPRM FTFILE RMFILE REC LIG SAXS_PROFILE L OUTPUT
 *
 * Contract of the reference tool (tools/score_ft_naive.c:68-232): the ft file is converted to Euler rows (written
 * to ./euler_list); for each row the receptor is turned by R(0, b1, g1), the ligand by R(a2, b2, g2) and moved z along
 * the axis, both are joined and centred, the complex is expanded (the hydration weights are those of the joined
 * input structures, computed once) and (c1, c2) are fitted; OUTPUT rows are "id<TAB>chi<TAB>c1<TAB>c2" (3 decimals).
 * It shares no kernel of the translate/rotate/angular-transform chain with `correlate`: expansion (K1), self terms and
 * the fit (K4) run on the device once per pose, so the two tools agree only as far as the truncation at L allows.
 */
#include "common.h"

#include "fftsaxs.h"
#include "index.h"
#include "saxs_utils.h"

#include "mol2/atom_group.h"
#include "mol2/pdb.h"
#include "mol2/prms.h"
#include "mol2/transform.h"

static struct mol_atom_group *load(const char *path, struct mol_prms *prms)
{
	struct mol_atom_group *ag = mol_read_pdb(path);
	if (ag == NULL) {
		ERROR_MSG("cannot read PDB file");
	}
	mol_atom_group_add_prms(ag, prms);
	return ag;
}

int main(int argc, char **argv)
{
	if (argc != 10) {
		fprintf(stderr, "Usage: score_ft_naive MAPPING_PRM ATOMPRM FT_PATH RM_PATH REC_PATH LIG_PATH REF_PROFILE L OUTPUT\n");
		return EXIT_FAILURE;
	}
	const char *map_path = argv[1], *prm_path = argv[2], *ft_path = argv[3], *rm_path = argv[4];
	const char *rec_path = argv[5], *lig_path = argv[6], *exp_path = argv[7], *out_path = argv[9];
	const int L = atoi(argv[8]);
	const int qnum = 50;
	char eul_path[] = "euler_list";
	double *qvals = sxs_mkarray(0.0, 0.5, qnum);

	struct mol_prms *prms = mol_prms_read(prm_path);
	if (prms == NULL) {
		ERROR_MSG("cannot read atom parameter file");
	}
	struct saxs_form_factor_table *ff = default_ff_table(map_path);

	/* receptor on its centre of extrema, ligand on its centroid; one working copy of each */
	struct mol_atom_group *rec = load(rec_path, prms), *rec_moved = load(rec_path, prms);
	struct mol_vector3 coe;
	center_of_extrema(&coe, rec);
	MOL_VEC_MULT_SCALAR(coe, coe, -1.0);
	mol_atom_group_translate(rec, &coe);
	mol_atom_group_translate(rec_moved, &coe);

	struct mol_atom_group *lig = load(lig_path, prms), *lig_moved = load(lig_path, prms);
	struct mol_vector3 com;
	centroid(&com, lig);
	MOL_VEC_MULT_SCALAR(com, com, -1.0);
	mol_atom_group_translate(lig, &com);
	mol_atom_group_translate(lig_moved, &com);

	struct mol_vector3 ref_lig;
	MOL_VEC_SUB(ref_lig, com, coe);
	MOL_VEC_MULT_SCALAR(ref_lig, ref_lig, -1.0);

	struct sxs_profile *exp_profile = sxs_profile_read((char *)exp_path);
	const double rec_rad = mol_atom_group_average_radius(rec), lig_rad = mol_atom_group_average_radius(lig);
	const double join_rad = (rec_rad * rec->natoms + lig_rad * lig->natoms) / (rec->natoms + lig->natoms);
	struct sxs_opt_params *params = sxs_opt_params_create(exp_profile, qvals, qnum, join_rad);
	struct sxs_profile *profile = sxs_profile_create(qvals, qnum, 1);

	sxs_ft_file2euler_file(eul_path, (char *)ft_path, (char *)rm_path, &ref_lig);
	FILE *euler = sxs_myfopen(eul_path, "r");

	struct sxs_spf_full *coefs = sxs_spf_full_create(L, qnum);
	/* hydration weights of the joined input structures, kept for every pose */
	struct mol_atom_group *join = mol_atom_group_join(rec_moved, lig_moved);
	double *saxs_sa = (double *)malloc(join->natoms * sizeof(double));
	sxs_faccs(saxs_sa, join, 1.4);
	mol_atom_group_free(join);

	FILE *out = sxs_myfopen((char *)out_path, "w");
	printf("Output will be written to %s\n", out_path);
	const clock_t t0 = clock();
	int eu_id, iline = 0;
	double z, b1, g1, a2, b2, g2;
	struct mol_matrix3 rec_rm, lig_rm;
	struct mol_vector3 rec_tv = {0.0, 0.0, 0.0}, lig_tv = {0.0, 0.0, 0.0};
	while (fscanf(euler, "%d %lf %lf %lf %lf %lf %lf", &eu_id, &z, &b1, &g1, &a2, &b2, &g2) == 7) {
		printf("Processing line %i\n", iline++);
		lig_tv.Z = z;
		sxs_fill_active_rotation_matrix(&rec_rm, 0.0, b1, g1);
		sxs_fill_active_rotation_matrix(&lig_rm, a2, b2, g2);
		mol_atom_group_move_in_copy(lig, lig_moved, &lig_rm, &lig_tv);
		mol_atom_group_move_in_copy(rec, rec_moved, &rec_rm, &rec_tv);
		join = mol_atom_group_join(rec_moved, lig_moved);
		struct mol_vector3 jc;
		centroid(&jc, join);
		MOL_VEC_MULT_SCALAR(jc, jc, -1.0);
		mol_atom_group_translate(join, &jc);
		atom_grp2spf_inplace(coefs, join, ff, qvals, qnum, L, saxs_sa);   /* K1 on the device */
		sxs_spf2fitted_profile(profile, coefs, params);                   /* self terms + K4 on the device */
		mol_atom_group_free(join);
		fprintf(out, "%d\t%.3f\t%.3f\t%.3f\n", eu_id, profile->score, profile->c1, profile->c2);
	}
	printf("Total time elapsed: %.4f\n", (double)(clock() - t0) / CLOCKS_PER_SEC);
	fclose(out);
	fclose(euler);

	free(saxs_sa);
	mol_prms_free(prms);
	sxs_spf_full_free(coefs);
	sxs_profile_free(exp_profile);
	sxs_profile_free(profile);
	sxs_opt_params_free(params);
	mol_atom_group_free(rec);
	mol_atom_group_free(rec_moved);
	mol_atom_group_free(lig);
	mol_atom_group_free(lig_moved);
	free(ff);
	free(qvals);
	return EXIT_SUCCESS;
}
