/* mini-libmol2 implementation (see mol2/mini.h for scope and provenance).
 *
 * libmol2 itself is absent from the reference tree (un-vendored dependency),
 * so everything here restates published behaviour:
 *   - PDB ATOM/HETATM fixed columns (wwPDB format v3.3),
 *   - the `atom id RES ATOM subid radius charge` rows of the atom prm file
 *     (prms/atoms.0.0.6.prm.ms.3cap+0.5ace.Hr0rec:10-12),
 *   - centroid = mean, center_of_extrema = (min+max)/2,
 *   - accs = Lee & Richards (1971) z-slice accessible-surface integration with
 *     the classic integration increment P = 0.01 (100 slices per sphere),
 *     scaled to the van der Waals sphere when cont_acc != 0.
 * Pinned by tests/data/ref_spf: rm header (prm radii incl. zero-radius H) and
 * the W columns (SASA).
 */
#define _POSIX_C_SOURCE 200809L
#include "mol2/mini.h"

#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

static void *xcalloc(size_t n, size_t sz)
{
	void *p = calloc(n ? n : 1, sz);
	if (p == NULL) {
		fprintf(stderr, "[Error] mol2_mini: out of memory\n");
		exit(EXIT_FAILURE);
	}
	return p;
}

bool is_whitespace_line(const char *line)
{
	for (; *line; ++line) {
		if (!isspace((unsigned char)*line)) {
			return false;
		}
	}
	return true;
}

/* ------------------------------------------------------------------ atoms */

struct mol_atom_group *mol_atom_group_create(size_t natoms)
{
	struct mol_atom_group *ag = xcalloc(1, sizeof(*ag));
	ag->natoms = natoms;
	ag->coords = xcalloc(natoms, sizeof(struct mol_vector3));
	ag->vdw_radius = xcalloc(natoms, sizeof(double));
	ag->charge = xcalloc(natoms, sizeof(double));
	ag->atom_name = xcalloc(natoms, sizeof(char *));
	ag->residue_name = xcalloc(natoms, sizeof(char *));
	return ag;
}

void mol_atom_group_free(struct mol_atom_group *ag)
{
	if (ag == NULL) {
		return;
	}
	for (size_t i = 0; i < ag->natoms; i++) {
		free(ag->atom_name[i]);
		free(ag->residue_name[i]);
	}
	free(ag->atom_name);
	free(ag->residue_name);
	free(ag->coords);
	free(ag->vdw_radius);
	free(ag->charge);
	free(ag);
}

static char *dup_field(const char *line, size_t len, size_t from, size_t to)
{
	/* columns are 1-based inclusive */
	char buf[16];
	size_t n = 0;
	for (size_t c = from; c <= to && c <= len && n + 1 < sizeof(buf); c++) {
		buf[n++] = line[c - 1];
	}
	buf[n] = '\0';
	return strdup(buf);
}

static double num_field(const char *line, size_t len, size_t from, size_t to)
{
	char buf[32];
	size_t n = 0;
	for (size_t c = from; c <= to && c <= len && n + 1 < sizeof(buf); c++) {
		buf[n++] = line[c - 1];
	}
	buf[n] = '\0';
	return atof(buf);
}

struct mol_atom_group *mol_read_pdb(const char *path)
{
	FILE *f = fopen(path, "r");
	if (f == NULL) {
		fprintf(stderr, "[Error] mol_read_pdb: cannot open %s\n", path);
		return NULL;
	}

	char *line = NULL;
	size_t cap = 0;
	size_t natoms = 0;
	while (getline(&line, &cap, f) != -1) {
		if (!strncmp(line, "ATOM  ", 6) || !strncmp(line, "HETATM", 6)) {
			natoms++;
		}
	}
	rewind(f);

	struct mol_atom_group *ag = mol_atom_group_create(natoms);
	size_t i = 0;
	while (getline(&line, &cap, f) != -1 && i < natoms) {
		if (strncmp(line, "ATOM  ", 6) && strncmp(line, "HETATM", 6)) {
			continue;
		}
		size_t len = strlen(line);
		while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) {
			line[--len] = '\0';
		}
		ag->atom_name[i] = dup_field(line, len, 13, 16);
		ag->residue_name[i] = dup_field(line, len, 18, 21);
		ag->coords[i].X = num_field(line, len, 31, 38);
		ag->coords[i].Y = num_field(line, len, 39, 46);
		ag->coords[i].Z = num_field(line, len, 47, 54);
		i++;
	}
	free(line);
	fclose(f);
	return ag;
}

struct mol_atom_group *mol_atom_group_join(const struct mol_atom_group *a, const struct mol_atom_group *b)
{
	struct mol_atom_group *ag = mol_atom_group_create(a->natoms + b->natoms);
	for (size_t k = 0; k < ag->natoms; k++) {
		const struct mol_atom_group *src = k < a->natoms ? a : b;
		size_t i = k < a->natoms ? k : k - a->natoms;
		ag->coords[k] = src->coords[i];
		ag->vdw_radius[k] = src->vdw_radius[i];
		ag->charge[k] = src->charge[i];
		ag->atom_name[k] = strdup(src->atom_name[i]);
		ag->residue_name[k] = strdup(src->residue_name[i]);
	}
	return ag;
}

void mol_atom_group_translate(struct mol_atom_group *ag, const struct mol_vector3 *t)
{
	for (size_t i = 0; i < ag->natoms; i++) {
		MOL_VEC_ADD(ag->coords[i], ag->coords[i], *t);
	}
}

void mol_atom_group_move_in_copy(const struct mol_atom_group *src, struct mol_atom_group *dst,
                                 const struct mol_matrix3 *r, const struct mol_vector3 *t)
{
	const size_t n = src->natoms < dst->natoms ? src->natoms : dst->natoms;
	for (size_t i = 0; i < n; i++) {
		const struct mol_vector3 v = src->coords[i];
		dst->coords[i].X = r->m11 * v.X + r->m12 * v.Y + r->m13 * v.Z + t->X;
		dst->coords[i].Y = r->m21 * v.X + r->m22 * v.Y + r->m23 * v.Z + t->Y;
		dst->coords[i].Z = r->m31 * v.X + r->m32 * v.Y + r->m33 * v.Z + t->Z;
	}
}

void centroid(struct mol_vector3 *c, const struct mol_atom_group *ag)
{
	struct mol_vector3 s = {0.0, 0.0, 0.0};
	for (size_t i = 0; i < ag->natoms; i++) {
		MOL_VEC_ADD(s, s, ag->coords[i]);
	}
	MOL_VEC_MULT_SCALAR(*c, s, 1.0 / (double)ag->natoms);
}

void center_of_extrema(struct mol_vector3 *c, const struct mol_atom_group *ag)
{
	struct mol_vector3 lo = ag->coords[0], hi = ag->coords[0];
	for (size_t i = 1; i < ag->natoms; i++) {
		const struct mol_vector3 *p = &ag->coords[i];
		lo.X = fmin(lo.X, p->X); hi.X = fmax(hi.X, p->X);
		lo.Y = fmin(lo.Y, p->Y); hi.Y = fmax(hi.Y, p->Y);
		lo.Z = fmin(lo.Z, p->Z); hi.Z = fmax(hi.Z, p->Z);
	}
	c->X = 0.5 * (lo.X + hi.X);
	c->Y = 0.5 * (lo.Y + hi.Y);
	c->Z = 0.5 * (lo.Z + hi.Z);
}

/* ------------------------------------------------------------------- prms */

struct mol_prms *mol_prms_read(const char *path)
{
	FILE *f = fopen(path, "r");
	if (f == NULL) {
		fprintf(stderr, "[Error] mol_prms_read: cannot open %s\n", path);
		return NULL;
	}
	struct mol_prms *prms = xcalloc(1, sizeof(*prms));
	size_t cap_atoms = 1024;
	prms->atoms = xcalloc(cap_atoms, sizeof(struct mol_prm_atom));

	char *line = NULL;
	size_t cap = 0;
	while (getline(&line, &cap, f) != -1) {
		struct mol_prm_atom a;
		char maj[32], min[32];
		if (sscanf(line, " atom %d %31s %31s %d %lf %lf", &a.id, maj, min, &a.subid, &a.r, &a.q) != 6) {
			continue;
		}
		strncpy(a.typemaj, maj, sizeof(a.typemaj) - 1);
		a.typemaj[sizeof(a.typemaj) - 1] = '\0';
		strncpy(a.typemin, min, sizeof(a.typemin) - 1);
		a.typemin[sizeof(a.typemin) - 1] = '\0';
		if (prms->natoms == cap_atoms) {
			cap_atoms *= 2;
			prms->atoms = realloc(prms->atoms, cap_atoms * sizeof(struct mol_prm_atom));
			if (prms->atoms == NULL) {
				fprintf(stderr, "[Error] mol_prms_read: out of memory\n");
				exit(EXIT_FAILURE);
			}
		}
		prms->atoms[prms->natoms++] = a;
	}
	free(line);
	fclose(f);
	return prms;
}

void mol_prms_free(struct mol_prms *prms)
{
	if (prms != NULL) {
		free(prms->atoms);
		free(prms);
	}
}

static void strip_copy(char *dst, size_t dstlen, const char *src)
{
	size_t n = 0;
	while (*src == ' ') {
		src++;
	}
	while (*src && *src != ' ' && n + 1 < dstlen) {
		dst[n++] = *src++;
	}
	dst[n] = '\0';
}

void mol_atom_group_add_prms(struct mol_atom_group *ag, const struct mol_prms *prms)
{
	for (size_t i = 0; i < ag->natoms; i++) {
		char res[8], atm[8];
		strip_copy(res, sizeof(res), ag->residue_name[i]);
		strip_copy(atm, sizeof(atm), ag->atom_name[i]);
		const struct mol_prm_atom *hit = NULL;
		for (size_t k = 0; k < prms->natoms; k++) {
			if (!strcmp(prms->atoms[k].typemaj, res) && !strcmp(prms->atoms[k].typemin, atm)) {
				hit = &prms->atoms[k];
				break;
			}
		}
		if (hit == NULL) {
			fprintf(stderr, "[Error] mol_atom_group_add_prms: no parameters for (%s, %s), atom %zu\n",
			        res, atm, i);
			exit(EXIT_FAILURE);
		}
		ag->vdw_radius[i] = hit->r;
		ag->charge[i] = hit->q;
	}
}

/* ---------------------------------------------------------------- matrices */

struct mol_matrix3_list *mol_matrix3_list_from_file(const char *path)
{
	FILE *f = fopen(path, "r");
	if (f == NULL) {
		fprintf(stderr, "[Error] mol_matrix3_list_from_file: cannot open %s\n", path);
		return NULL;
	}
	struct mol_matrix3_list *list = xcalloc(1, sizeof(*list));
	size_t cap_m = 1024;
	list->members = xcalloc(cap_m, sizeof(struct mol_matrix3));

	char *line = NULL;
	size_t cap = 0;
	while (getline(&line, &cap, f) != -1) {
		double v[10];
		int n = sscanf(line, "%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf",
		               &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8], &v[9]);
		if (n < 9) {
			continue;
		}
		const double *m = (n == 10) ? v + 1 : v; /* optional leading index */
		if (list->size == cap_m) {
			cap_m *= 2;
			list->members = realloc(list->members, cap_m * sizeof(struct mol_matrix3));
			if (list->members == NULL) {
				fprintf(stderr, "[Error] mol_matrix3_list_from_file: out of memory\n");
				exit(EXIT_FAILURE);
			}
		}
		struct mol_matrix3 *d = &list->members[list->size++];
		d->m11 = m[0]; d->m12 = m[1]; d->m13 = m[2];
		d->m21 = m[3]; d->m22 = m[4]; d->m23 = m[5];
		d->m31 = m[6]; d->m32 = m[7]; d->m33 = m[8];
	}
	free(line);
	fclose(f);
	return list;
}

void mol_matrix3_list_free(struct mol_matrix3_list *list)
{
	if (list != NULL) {
		free(list->members);
		free(list);
	}
}

/* -------------------------------------------------------------------- SASA */

struct arc {
	double ti, tf;
};

static int arc_cmp(const void *a, const void *b)
{
	double x = ((const struct arc *)a)->ti, y = ((const struct arc *)b)->ti;
	return (x > y) - (x < y);
}

/* The per-atom part of accs: atoms ir0 .. ir1-1 of the compact list.  Every atom's area depends on the read-only
 * arrays only, so the atoms are split over host threads and the result does not depend on their number. */
struct accs_job {
	size_t ir0, ir1, n;
	const double *x, *y, *z, *r, *rsq;
	const double *lo;
	const int *dim, *cell_start, *cell_items;
	const size_t *ind;
	double edge, r_solv;
	short cont_acc;
	double *as;
};

static void *accs_range(void *arg)
{
	const struct accs_job *jb = (const struct accs_job *)arg;
	static const double P = 0.01;
	const double pi = acos(-1.0);
	const double pix2 = 2.0 * pi;
	const double *x = jb->x, *y = jb->y, *z = jb->z, *r = jb->r, *rsq = jb->rsq, *lo = jb->lo;
	const int *dim = jb->dim, *cell_start = jb->cell_start, *cell_items = jb->cell_items;
	const size_t *ind = jb->ind;
	const double edge = jb->edge, r_solv = jb->r_solv;
	const short cont_acc = jb->cont_acc;
	double *as = jb->as;

	size_t nb_cap = 256;
	int *nb = xcalloc(nb_cap, sizeof(int));
	double *nb_d = xcalloc(nb_cap, sizeof(double)), *nb_dsq = xcalloc(nb_cap, sizeof(double));
	double *nb_dx = xcalloc(nb_cap, sizeof(double)), *nb_dy = xcalloc(nb_cap, sizeof(double));
	struct arc *arcs = xcalloc(2 * nb_cap, sizeof(struct arc));

	const int nzp = (int)(1.0 / P + 0.5);

	for (size_t ir = jb->ir0; ir < jb->ir1; ir++) {
		const double xr = x[ir], yr = y[ir], zr = z[ir], rr = r[ir], rrsq = rsq[ir];
		const double rrx2 = 2.0 * rr;

		/* gather neighbours whose expanded spheres overlap this one */
		size_t nnb = 0;
		int cx = (int)((xr - lo[0]) / edge), cy = (int)((yr - lo[1]) / edge), cz = (int)((zr - lo[2]) / edge);
		for (int ix = cx - 1; ix <= cx + 1; ix++) {
			if (ix < 0 || ix >= dim[0]) continue;
			for (int iy = cy - 1; iy <= cy + 1; iy++) {
				if (iy < 0 || iy >= dim[1]) continue;
				for (int iz = cz - 1; iz <= cz + 1; iz++) {
					if (iz < 0 || iz >= dim[2]) continue;
					int c = (ix * dim[1] + iy) * dim[2] + iz;
					for (int s = cell_start[c]; s < cell_start[c + 1]; s++) {
						int in = cell_items[s];
						if ((size_t)in == ir) continue;
						double dx = xr - x[in], dy = yr - y[in], dz = zr - z[in];
						double dsq = dx * dx + dy * dy;
						double rsum = rr + r[in];
						if (dsq + dz * dz >= rsum * rsum) continue;
						if (nnb == nb_cap) {
							nb_cap *= 2;
							nb = realloc(nb, nb_cap * sizeof(int));
							nb_d = realloc(nb_d, nb_cap * sizeof(double));
							nb_dsq = realloc(nb_dsq, nb_cap * sizeof(double));
							nb_dx = realloc(nb_dx, nb_cap * sizeof(double));
							nb_dy = realloc(nb_dy, nb_cap * sizeof(double));
							arcs = realloc(arcs, 2 * nb_cap * sizeof(struct arc));
							if (!nb || !nb_d || !nb_dsq || !nb_dx || !nb_dy || !arcs) {
								fprintf(stderr, "[Error] accs: out of memory\n");
								exit(EXIT_FAILURE);
							}
						}
						nb[nnb] = in;
						nb_dx[nnb] = dx;
						nb_dy[nnb] = dy;
						nb_dsq[nnb] = dsq;
						nb_d[nnb] = sqrt(dsq);
						nnb++;
					}
				}
			}
		}

		double area = 0.0;
		if (nnb == 0) {
			area = pix2 * rrx2;
		} else {
			const double zres = rrx2 / nzp;
			double zgrid = zr - rr - zres / 2.0;
			for (int i = 0; i < nzp; i++) {
				zgrid += zres;
				double rsec2r = rrsq - (zgrid - zr) * (zgrid - zr);
				if (rsec2r < 0.0) {
					rsec2r = 0.000001;
				}
				const double rsecr = sqrt(rsec2r);
				size_t karc = 0;
				bool buried = false;
				for (size_t j = 0; j < nnb && !buried; j++) {
					const int in = nb[j];
					const double rsec2n = rsq[in] - (zgrid - z[in]) * (zgrid - z[in]);
					if (rsec2n <= 0.0) {
						continue;
					}
					const double rsecn = sqrt(rsec2n);
					if (nb_d[j] >= rsecr + rsecn) {
						continue; /* circles apart */
					}
					const double b = rsecr - rsecn;
					if (nb_d[j] <= fabs(b)) {
						if (b <= 0.0) {
							buried = true; /* this circle lies inside the neighbour's */
						}
						continue;
					}
					double arg = (nb_dsq[j] + rsec2r - rsec2n) / (2.0 * nb_d[j] * rsecr);
					if (arg > 1.0) arg = 1.0;
					if (arg < -1.0) arg = -1.0;
					const double alpha = acos(arg);
					const double beta = atan2(nb_dy[j], nb_dx[j]) + pi;
					double ti = beta - alpha, tf = beta + alpha;
					if (ti < 0.0) ti += pix2;
					if (tf > pix2) tf -= pix2;
					arcs[karc].ti = ti;
					if (tf < ti) { /* occluded arc crosses 0: split */
						arcs[karc].tf = pix2;
						karc++;
						arcs[karc].ti = 0.0;
					}
					arcs[karc].tf = tf;
					karc++;
				}
				if (buried) {
					continue;
				}
				double arcsum;
				if (karc == 0) {
					arcsum = pix2;
				} else {
					qsort(arcs, karc, sizeof(struct arc), arc_cmp);
					arcsum = arcs[0].ti;
					double t = arcs[0].tf;
					for (size_t k = 1; k < karc; k++) {
						if (t < arcs[k].ti) {
							arcsum += arcs[k].ti - t;
						}
						if (arcs[k].tf > t) {
							t = arcs[k].tf;
						}
					}
					arcsum += pix2 - t;
				}
				area += arcsum * zres;
			}
		}

		if (cont_acc) {
			as[ind[ir]] = area * (rr - r_solv) * (rr - r_solv) / rr;
		} else {
			as[ind[ir]] = area * rr;
		}
	}

	free(arcs); free(nb); free(nb_d); free(nb_dsq); free(nb_dx); free(nb_dy);
	return NULL;
}

/* Lee & Richards slice integration.  Atoms with zero radius neither receive
 * area nor occlude.  Slices are perpendicular to z, NZP = 1/P + 0.5 of them
 * per (expanded) sphere, sampled at slice mid-planes. */
void accs(double *as, const struct mol_atom_group *ag, double r_solv, short cont_acc)
{
	const size_t n_all = ag->natoms;

	for (size_t i = 0; i < n_all; i++) {
		as[i] = 0.0;
	}

	/* compact list of atoms with non-zero radius */
	size_t n = 0;
	size_t *ind = xcalloc(n_all, sizeof(size_t));
	double rmax = 0.0;
	for (size_t i = 0; i < n_all; i++) {
		if (ag->vdw_radius[i] > 0.0) {
			ind[n++] = i;
			rmax = fmax(rmax, ag->vdw_radius[i] + r_solv);
		}
	}
	if (n == 0) {
		free(ind);
		return;
	}
	double *x = xcalloc(n, sizeof(double)), *y = xcalloc(n, sizeof(double)), *z = xcalloc(n, sizeof(double));
	double *r = xcalloc(n, sizeof(double)), *rsq = xcalloc(n, sizeof(double));
	double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
	for (size_t k = 0; k < n; k++) {
		const struct mol_vector3 *p = &ag->coords[ind[k]];
		x[k] = p->X; y[k] = p->Y; z[k] = p->Z;
		r[k] = ag->vdw_radius[ind[k]] + r_solv;
		rsq[k] = r[k] * r[k];
		lo[0] = fmin(lo[0], x[k]); hi[0] = fmax(hi[0], x[k]);
		lo[1] = fmin(lo[1], y[k]); hi[1] = fmax(hi[1], y[k]);
		lo[2] = fmin(lo[2], z[k]); hi[2] = fmax(hi[2], z[k]);
	}

	/* cubic cell list with edge 2*rmax: neighbours are in the 27 surrounding cells */
	const double edge = 2.0 * rmax;
	int dim[3];
	for (int d = 0; d < 3; d++) {
		dim[d] = (int)((hi[d] - lo[d]) / edge) + 1;
	}
	size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
	int *cell_of = xcalloc(n, sizeof(int));
	int *cell_start = xcalloc(ncell + 1, sizeof(int));
	int *cell_items = xcalloc(n, sizeof(int));
	for (size_t k = 0; k < n; k++) {
		int cx = (int)((x[k] - lo[0]) / edge), cy = (int)((y[k] - lo[1]) / edge), cz = (int)((z[k] - lo[2]) / edge);
		cell_of[k] = (cx * dim[1] + cy) * dim[2] + cz;
		cell_start[cell_of[k] + 1]++;
	}
	for (size_t c = 0; c < ncell; c++) {
		cell_start[c + 1] += cell_start[c];
	}
	int *fill = xcalloc(ncell, sizeof(int));
	for (size_t k = 0; k < n; k++) {
		cell_items[cell_start[cell_of[k]] + fill[cell_of[k]]++] = (int)k;
	}
	free(fill);

	/* host threads: SXS_HOST_THREADS, else the online processors, at least 64 atoms per thread */
	int nthreads = getenv("SXS_HOST_THREADS") ? atoi(getenv("SXS_HOST_THREADS")) : (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (nthreads > 64) nthreads = 64;
	if ((size_t)nthreads > n / 64 + 1) nthreads = (int)(n / 64 + 1);
	if (nthreads < 1) nthreads = 1;
	struct accs_job jobs[64];
	pthread_t th[64];
	for (int k = 0; k < nthreads; k++) {
		struct accs_job *jb = &jobs[k];
		jb->ir0 = n * (size_t)k / (size_t)nthreads; jb->ir1 = n * (size_t)(k + 1) / (size_t)nthreads; jb->n = n;
		jb->x = x; jb->y = y; jb->z = z; jb->r = r; jb->rsq = rsq; jb->lo = lo; jb->dim = dim;
		jb->cell_start = cell_start; jb->cell_items = cell_items; jb->ind = ind;
		jb->edge = edge; jb->r_solv = r_solv; jb->cont_acc = cont_acc; jb->as = as;
	}
	int started = 0;
	for (int k = 1; k < nthreads; k++) {
		if (pthread_create(&th[k], NULL, accs_range, &jobs[k]) != 0) {
			break;
		}
		started = k;
	}
	accs_range(&jobs[0]);
	for (int k = started + 1; k < nthreads; k++) {
		accs_range(&jobs[k]); /* a thread that could not be created: its atoms here */
	}
	for (int k = 1; k <= started; k++) {
		pthread_join(th[k], NULL);
	}

	free(cell_of); free(cell_start); free(cell_items);
	free(x); free(y); free(z); free(r); free(rsq); free(ind);
}
