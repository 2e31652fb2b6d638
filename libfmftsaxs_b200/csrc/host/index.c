/* Flat grid index <-> digits, and PIPER-style (translation, rotation) -> (z, five Euler angles). */
#include "index.h"

size_t sxs_assemble_index(struct sxs_index *id, int nbeta, int L)
{
	const int n = 2 * L + 1;
	return ((((id->z * nbeta + id->b1) * nbeta + id->b2) * n + id->a2) * n + id->g1) * n + id->g2;
}

/* Digits are peeled with 32-bit arithmetic after the first division, as src/index.c:9-29 does. */
void sxs_disassemble_index(struct sxs_index *id, size_t index, int nbeta, int L)
{
	const int n = 2 * L + 1;
	int rest;
	id->g2 = index % n;
	rest = index / n;
	id->g1 = rest % n;
	rest /= n;
	id->a2 = rest % n;
	rest /= n;
	id->b2 = rest % nbeta;
	rest /= nbeta;
	id->b1 = rest % nbeta;
	id->z = rest / nbeta;
}

static double clamp_unit(double a)
{
	if (a < -1.0) { return -1.0; }
	if (a > 1.0) { return 1.0; }
	return a;
}

/* src/index.c:38-75.  The receptor is turned by (0, b1, g1) so that the ligand centre lies on +z at
 * the rounded distance z; the ligand's own rotation is composed with it and read back as z-y-z angles. */
void sxs_ft2euler(struct sxs_euler *euler, struct mol_vector3 *tv, struct mol_matrix3 *rm, struct mol_vector3 *ref_lig)
{
	struct mol_vector3 vec;
	MOL_VEC_ADD(vec, *tv, *ref_lig);

	const double z = round(sqrt(MOL_VEC_SQ_NORM(vec)));
	const double b1 = acos(clamp_unit(vec.Z / z));
	double g1 = acos(clamp_unit(-vec.X / (z * sin(b1))));
	if (vec.Y / (z * sin(b1)) < 0.0) {
		g1 = 2 * M_PI - g1;
	}

	struct mol_matrix3 rec_rm, lig_rm;
	sxs_fill_active_rotation_matrix(&rec_rm, 0.0, b1, g1);
	sxs_mult_rot_mats(&lig_rm, &rec_rm, rm);

	const double b2 = acos(clamp_unit(lig_rm.m33));
	double a2 = acos(clamp_unit(lig_rm.m13 / sin(b2)));
	if (lig_rm.m23 / sin(b2) < 0.0) {
		a2 = 2 * M_PI - a2;
	}
	double g2 = acos(clamp_unit(-lig_rm.m31 / sin(b2)));
	if (lig_rm.m32 / sin(b2) < 0.0) {
		g2 = 2 * M_PI - g2;
	}

	euler->z = z;
	euler->b1 = b1;
	euler->g1 = g1;
	euler->a2 = a2;
	euler->b2 = b2;
	euler->g2 = g2;
}

/* ft rows: rotation index, translation xyz, six ignored numbers.  Output rows carry three decimals;
 * that quantisation decides the grid cell later, so the format string is part of the contract. */
void sxs_ft_file2euler_file(const char *eu_path, const char *ft_path, const char *rm_path, struct mol_vector3 *ref_lig)
{
	FILE *ft = fopen(ft_path, "r");
	FILE *eu = fopen(eu_path, "w");
	if (ft == NULL || eu == NULL) {
		ERROR_MSG("cannot open ft or Euler file");
	}
	struct mol_matrix3_list *rots = mol_matrix3_list_from_file(rm_path);
	if (rots == NULL) {
		ERROR_MSG("cannot read rotation file");
	}
	int id;
	double v[9];
	int got;
	while ((got = fscanf(ft, "%d %lf %lf %lf %lf %lf %lf %lf %lf %lf\n", &id, &v[0], &v[1], &v[2], &v[3], &v[4],
	                     &v[5], &v[6], &v[7], &v[8])) != EOF) {
		if (got != 10) {
			ERROR_MSG("Wrong input file format.");
		}
		if (id < 0 || (size_t)id >= rots->size) {
			ERROR_MSG("rotation index outside the rotation file");
		}
		struct mol_vector3 t = {v[0], v[1], v[2]};
		struct sxs_euler e;
		sxs_ft2euler(&e, &t, &rots->members[id], ref_lig);
		fprintf(eu, "%d\t% .3f\t% .3f\t% .3f\t% .3f\t% .3f\t% .3f\n", id, e.z, e.b1, e.g1, e.a2, e.b2, e.g2);
	}
	mol_matrix3_list_free(rots);
	fclose(eu);
	fclose(ft);
}

int sxs_euler_to_index(const struct sxs_euler *euler, int z_index, int L)
{
	const int nbeta = L + 1, n = 2 * L + 1;
	const double b_step = M_PI / L;
	const double a_step = 2.0 * M_PI / n;
	const double a2 = 2 * M_PI - euler->a2;
	const double g2 = 2 * M_PI - euler->g2;

	int id = z_index * nbeta;
	id = (id + (int)(round(euler->b1 / b_step))) * nbeta;
	id = (id + (int)(round(euler->b2 / b_step))) * n;
	id = (id + (int)(round(a2 / a_step))) * n;
	id = (id + (int)(round(euler->g1 / a_step))) * n;
	id = id + (int)(round(g2 / a_step));
	return id;
}

long long sxs_euler_to_index64(const struct sxs_euler *euler, int z_index, int L)
{
	const long long nbeta = L + 1, n = 2 * L + 1;
	const double b_step = M_PI / L;
	const double a_step = 2.0 * M_PI / n;
	const double a2 = 2 * M_PI - euler->a2;
	const double g2 = 2 * M_PI - euler->g2;

	long long id = z_index * nbeta;
	id = (id + (int)(round(euler->b1 / b_step))) * nbeta;
	id = (id + (int)(round(euler->b2 / b_step))) * n;
	id = (id + (int)(round(a2 / a_step))) * n;
	id = (id + (int)(round(euler->g1 / a_step))) * n;
	id = id + (int)(round(g2 / a_step));
	return id;
}
