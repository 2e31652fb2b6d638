/* In-memory ft rows -> flat grid indices (include/fmftsaxs/index.h, sxs_ft_rows_to_indices): the Euler text file
 * of the reference's tool chain without the file.  Kept apart from index.c because the thread/sysconf
 * declarations need a POSIX feature macro, and index.c must see <math.h> exactly as the reference's -std=c11
 * build does (no M_PI from libc, src/define.h:13-15).  Nothing here uses pi. */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include <limits.h>

#include "index.h"

/* see ft_fast.c: off-table is a value of its own, rows with nan-derived (negative) indices are kept like the
 * reference-shaped route keeps them */
#define SXS_OFF_TABLE LLONG_MIN

/* what fprintf("% .3f") followed by fscanf("%lf") makes of x (src/index.c:114, tools/correlate.c:214) */
static double through_text(double x)
{
	char buf[64];
	snprintf(buf, sizeof(buf), "% .3f", x);
	return strtod(buf, NULL);
}

struct rows_job {
	long long i0, i1;
	const int *rot_id;
	const double *trans;
	const struct mol_matrix3_list *rots;
	struct mol_vector3 *ref_lig;
	const double *zvals;
	int znum, L;
	long long *flat; /* per input row: flat index, or SXS_OFF_TABLE when the row is dropped */
	int bad;
};

static void *rows_main(void *arg)
{
	struct rows_job *j = (struct rows_job *)arg;
	for (long long i = j->i0; i < j->i1; i++) {
		const int id = j->rot_id[i];
		if (id < 0 || (size_t)id >= j->rots->size) {
			j->bad = 1;
			return NULL;
		}
		struct mol_vector3 t = {j->trans[3 * i], j->trans[3 * i + 1], j->trans[3 * i + 2]};
		struct sxs_euler e;
		sxs_ft2euler(&e, &t, &j->rots->members[id], j->ref_lig);
		e.z = through_text(e.z);
		e.b1 = through_text(e.b1);
		e.g1 = through_text(e.g1);
		e.a2 = through_text(e.a2);
		e.b2 = through_text(e.b2);
		e.g2 = through_text(e.g2);
		long long flat = SXS_OFF_TABLE;
		for (int k = 0; k < j->znum; k++) {
			if (j->zvals[k] > e.z - 0.001 && j->zvals[k] < e.z + 0.001) {
				flat = sxs_euler_to_index64(&e, k, j->L);
				break; /* the 1 A table of the tool matches at most one z */
			}
		}
		j->flat[i] = flat;
	}
	return NULL;
}

static long long rows_to_flat(long long *flat, const int *rot_id, const double *trans, long long n,
                              const struct mol_matrix3_list *rots, struct mol_vector3 *ref_lig, const double *zvals,
                              int znum, int L, int nthreads)
{
	if (nthreads <= 0) {
		nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	}
	if (nthreads > 64) nthreads = 64;
	if (nthreads < 1) nthreads = 1;
	if ((long long)nthreads > n / 4096 + 1) nthreads = (int)(n / 4096 + 1);
	struct rows_job jobs[64];
	pthread_t th[64];
	for (int k = 0; k < nthreads; k++) {
		struct rows_job *j = &jobs[k];
		j->i0 = n * k / nthreads; j->i1 = n * (k + 1) / nthreads;
		j->rot_id = rot_id; j->trans = trans; j->rots = rots; j->ref_lig = ref_lig;
		j->zvals = zvals; j->znum = znum; j->L = L; j->flat = flat; j->bad = 0;
	}
	if (nthreads == 1) {
		rows_main(&jobs[0]);
	} else {
		for (int k = 0; k < nthreads; k++) {
			if (pthread_create(&th[k], NULL, rows_main, &jobs[k]) != 0) {
				ERROR_MSG("pthread_create failed");
			}
		}
		for (int k = 0; k < nthreads; k++) {
			pthread_join(th[k], NULL);
		}
	}
	for (int k = 0; k < nthreads; k++) {
		if (jobs[k].bad) {
			ERROR_MSG("rotation index outside the rotation file");
		}
	}
	return n;
}

long long sxs_ft_rows_to_indices64(long long *index, int *ft_id, int *order, const int *rot_id, const double *trans,
                                   long long n, const struct mol_matrix3_list *rots, struct mol_vector3 *ref_lig,
                                   const double *zvals, int znum, int L, int nthreads)
{
	CHECK_PTR(rots);
	if (n <= 0) {
		return 0;
	}
	long long *flat = (long long *)malloc(sizeof(long long) * (size_t)n);
	CHECK_PTR(flat);
	rows_to_flat(flat, rot_id, trans, n, rots, ref_lig, zvals, znum, L, nthreads);
	long long kept = 0;
	for (long long i = 0; i < n; i++) {
		if (flat[i] != SXS_OFF_TABLE) {
			index[kept] = flat[i];
			ft_id[kept] = rot_id[i];
			order[kept] = (int)i;
			kept++;
		}
	}
	free(flat);
	return kept;
}

long long sxs_ft_rows_to_indices(int *index, int *ft_id, int *order, const int *rot_id, const double *trans,
                                 long long n, const struct mol_matrix3_list *rots, struct mol_vector3 *ref_lig,
                                 const double *zvals, int znum, int L, int nthreads)
{
	CHECK_PTR(rots);
	if (n <= 0) {
		return 0;
	}
	long long *flat = (long long *)malloc(sizeof(long long) * (size_t)n);
	CHECK_PTR(flat);
	rows_to_flat(flat, rot_id, trans, n, rots, ref_lig, zvals, znum, L, nthreads);
	long long kept = 0;
	for (long long i = 0; i < n; i++) {
		if (flat[i] != SXS_OFF_TABLE) {
			index[kept] = (int)flat[i]; /* 32-bit packing like the tool's `int id` (tools/correlate.c:225-240) */
			ft_id[kept] = rot_id[i];
			order[kept] = (int)i;
			kept++;
		}
	}
	free(flat);
	return kept;
}
