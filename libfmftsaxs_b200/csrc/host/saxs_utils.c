/* Helper routines of the path (q grid, mean radius, SASA fraction, Euler rotation matrices). */
#include "saxs_utils.h"

double mol_atom_group_max_dist(const struct mol_atom_group *ag)
{
	double best = 0.0;
	for (size_t i = 0; i < ag->natoms; i++) {
		for (size_t j = i + 1; j < ag->natoms; j++) {
			best = fmax(best, MOL_VEC_EUCLIDEAN_DIST_SQ(ag->coords[i], ag->coords[j]));
		}
	}
	return sqrt(best);
}

/* src/saxs_utils.c:19-28: plain mean over all atoms, zero-radius hydrogens included. */
double mol_atom_group_average_radius(const struct mol_atom_group *ag)
{
	double sum = 0.0;
	for (size_t i = 0; i < ag->natoms; i++) {
		sum += ag->vdw_radius[i];
	}
	return sum / ag->natoms;
}

/* src/saxs_utils.c:30-48: contact area / (4 pi r^2) with the truncated pi; 0 for r == 0 or NaN. */
void sxs_faccs(double *fractional_sa, const struct mol_atom_group *ag, double r_solv)
{
	const double fourpi = 4 * M_PI;
	accs(fractional_sa, ag, r_solv, 1);
	for (size_t i = 0; i < ag->natoms; i++) {
		double ri = ag->vdw_radius[i];
		if (ri == 0 || isnan(fractional_sa[i])) {
			fractional_sa[i] = 0.0;
		} else {
			fractional_sa[i] = fractional_sa[i] / (fourpi * ri * ri);
		}
	}
}

/* src/saxs_utils.c:50-63: the grid is built by repeated addition, so q[qnum-1] != end exactly. */
double *sxs_mkarray(double begin, double end, int qnum)
{
	if (begin > end || begin < 0 || end < 0) {
		fprintf(stderr, "WRONG Q VALUES WHEN CREATING ARRAY");
		exit(EXIT_FAILURE);
	}
	double step = (end - begin) / (qnum - 1);
	double *q = (double *)calloc(qnum, sizeof(double));
	CHECK_PTR(q);
	q[0] = begin;
	for (int i = 1; i < qnum; i++) {
		q[i] = q[i - 1] + step;
	}
	return q;
}

/* z-y-z active rotation (src/saxs_utils.c:65-79). */
void sxs_fill_active_rotation_matrix(struct mol_matrix3 *rm, double alpha, double beta, double gamma)
{
	const double ca = cos(alpha), sa = sin(alpha);
	const double cb = cos(beta), sb = sin(beta);
	const double cg = cos(gamma), sg = sin(gamma);
	rm->m11 = cg * cb * ca - sg * sa;
	rm->m21 = cg * cb * sa + sg * ca;
	rm->m31 = -cg * sb;
	rm->m12 = -sg * cb * ca - cg * sa;
	rm->m22 = -sg * cb * sa + cg * ca;
	rm->m32 = sg * sb;
	rm->m13 = sb * ca;
	rm->m23 = sb * sa;
	rm->m33 = cb;
}

/* c = a b (src/saxs_utils.c:81-94).  mol_matrix3 is nine consecutive doubles in row-major order; every entry is the
 * left-to-right sum of its three products, and c may alias a or b. */
void sxs_mult_rot_mats(struct mol_matrix3 *c, struct mol_matrix3 *a, struct mol_matrix3 *b)
{
	const double *x = &a->m11, *y = &b->m11;
	double r[9];
	for (int row = 0; row < 3; row++) {
		for (int col = 0; col < 3; col++) {
			r[3 * row + col] = x[3 * row] * y[col] + x[3 * row + 1] * y[3 + col] + x[3 * row + 2] * y[6 + col];
		}
	}
	double *out = &c->m11;
	for (int k = 0; k < 9; k++) {
		out[k] = r[k];
	}
}
