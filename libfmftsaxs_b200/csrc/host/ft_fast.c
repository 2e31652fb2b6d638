/* ft file -> Euler text file + flat grid indices in one threaded pass (include/fmftsaxs/index.h,
 * sxs_ft_file_to_indices), and a threaded writer for the tool's output rows.
 *
 * The reference tool chain reads the ft file twice with fscanf, writes the Euler file with fprintf and reads it back
 * with fscanf (src/index.c:77-121, tools/correlate.c:169-251): ~17 number conversions per row on one core, which is
 * what a large run spends its time on once the scoring itself takes a fraction of a second.  Here the file is read
 * once, cut at line ends into one piece per thread, and every thread parses its rows (strtol/strtod: the conversions
 * fscanf performs), converts them (sxs_ft2euler), formats the Euler line exactly as the reference does and takes the
 * quantised angles back from that very text, so that the side file is byte-identical and the indices are the ones the
 * file would give.  Anything unusual in the input (a token fscanf would split differently, a row with other than ten
 * tokens) makes the function step aside: it returns -1 and the caller takes the reference-shaped slow route.
 * Kept apart from index.c for the POSIX feature macro (see index_rows.c). */
#define _POSIX_C_SOURCE 200809L
#include <ctype.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <limits.h>

#include "index.h"

/* "this row's z is not on the table" — a value of its own, not the sign of the index: a degenerate geometry (beta = 0:
 * 0 / 0 angles, "nan" in the Euler file) gives the tool's `(int)round(nan)` digits and with them a negative index for a
 * row that DID match a z; the reference-shaped route keeps such a row (the CUDA layer leaves out-of-range indices
 * untouched), so this route keeps it too */
#define SXS_OFF_TABLE LLONG_MIN

struct piece {
	const char *beg, *end; /* whole lines */
	const struct mol_matrix3_list *rots;
	struct mol_vector3 *ref_lig;
	const double *zvals;
	int znum, L, want_text;
	/* results */
	long long nrows, cap;
	long long *flat; /* per row: flat index, or SXS_OFF_TABLE when the row's z is not on the table */
	int *rot_id;
	char *text; size_t text_len, text_cap;
	int status; /* 0 ok, 1 step aside, 2 rotation index out of range */
};

static int plain_number(const char *s, const char *e)
{
	int digits = 0;
	for (const char *p = s; p < e; p++) {
		const char c = *p;
		if (c >= '0' && c <= '9') {
			digits++;
		} else if (!(c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E')) {
			return 0;
		}
	}
	return digits > 0;
}

static void *piece_main(void *arg)
{
	struct piece *pc = (struct piece *)arg;
	const char *p = pc->beg;
	char line[256];
	while (p < pc->end) {
		/* ten whitespace-separated tokens make a row, wherever the line ends fall (fscanf does not look at them) */
		const char *tok[10], *tend[10];
		int nt = 0;
		while (nt < 10) {
			while (p < pc->end && isspace((unsigned char)*p)) p++;
			if (p >= pc->end) break;
			tok[nt] = p;
			while (p < pc->end && !isspace((unsigned char)*p)) p++;
			tend[nt] = p;
			nt++;
		}
		if (nt == 0) {
			break;
		}
		if (nt != 10) {
			pc->status = 1;
			return NULL;
		}
		for (int k = 0; k < 10; k++) {
			if (!plain_number(tok[k], tend[k]) || tend[k] - tok[k] > 40) {
				pc->status = 1;
				return NULL;
			}
		}
		char *endp;
		char buf[10][48];
		for (int k = 0; k < 10; k++) {
			memcpy(buf[k], tok[k], (size_t)(tend[k] - tok[k]));
			buf[k][tend[k] - tok[k]] = '\0';
		}
		/* the six ignored columns: "%lf" must consume each token whole ("12.3-4.5", "1.5.3", "--" would be split or
		 * refused by fscanf: such a file goes to the reference-shaped route, which reports it as the reference does) */
		for (int k = 4; k < 10; k++) {
			(void)strtod(buf[k], &endp);
			if (endp == buf[k] || *endp != '\0') {
				pc->status = 1;
				return NULL;
			}
		}
		const long id = strtol(buf[0], &endp, 10);
		if (*endp != '\0') { /* "%d" would stop inside the token */
			pc->status = 1;
			return NULL;
		}
		double v[3];
		for (int k = 0; k < 3; k++) {
			v[k] = strtod(buf[k + 1], &endp);
			if (*endp != '\0') {
				pc->status = 1;
				return NULL;
			}
		}
		if (id < 0 || (size_t)id >= pc->rots->size) {
			pc->status = 2;
			return NULL;
		}
		struct mol_vector3 t = {v[0], v[1], v[2]};
		struct sxs_euler e;
		sxs_ft2euler(&e, &t, &pc->rots->members[id], pc->ref_lig);
		/* the Euler row of src/index.c:114, and the numbers a reader of that row gets */
		const int len = snprintf(line, sizeof(line), "%d\t% .3f\t% .3f\t% .3f\t% .3f\t% .3f\t% .3f\n", (int)id, e.z, e.b1, e.g1,
		                         e.a2, e.b2, e.g2);
		if (len <= 0 || len >= (int)sizeof(line)) {
			pc->status = 1;
			return NULL;
		}
		const char *q = strchr(line, '\t') + 1;
		double *dst[6] = {&e.z, &e.b1, &e.g1, &e.a2, &e.b2, &e.g2};
		for (int k = 0; k < 6; k++) {
			*dst[k] = strtod(q, &endp);
			q = endp + 1;
		}
		long long flat = SXS_OFF_TABLE;
		for (int k = 0; k < pc->znum; k++) {
			if (pc->zvals[k] > e.z - 0.001 && pc->zvals[k] < e.z + 0.001) {
				flat = sxs_euler_to_index64(&e, k, pc->L);
				break; /* the 1 A table of the tool matches at most one z */
			}
		}
		if (pc->nrows == pc->cap) {
			pc->cap = pc->cap ? pc->cap * 2 : 4096;
			pc->flat = (long long *)realloc(pc->flat, sizeof(long long) * (size_t)pc->cap);
			pc->rot_id = (int *)realloc(pc->rot_id, sizeof(int) * (size_t)pc->cap);
			CHECK_PTR(pc->flat); CHECK_PTR(pc->rot_id);
		}
		pc->flat[pc->nrows] = flat;
		pc->rot_id[pc->nrows] = (int)id;
		pc->nrows++;
		if (pc->want_text) {
			if (pc->text_len + (size_t)len > pc->text_cap) {
				pc->text_cap = pc->text_cap ? pc->text_cap * 2 : (1u << 20);
				pc->text = (char *)realloc(pc->text, pc->text_cap);
				CHECK_PTR(pc->text);
			}
			memcpy(pc->text + pc->text_len, line, (size_t)len);
			pc->text_len += (size_t)len;
		}
	}
	return NULL;
}

static int thread_count(int nthreads, long long work_units)
{
	if (nthreads <= 0) {
		nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	}
	if (nthreads > 64) nthreads = 64;
	if (nthreads < 1) nthreads = 1;
	if ((long long)nthreads > work_units / 4096 + 1) nthreads = (int)(work_units / 4096 + 1);
	return nthreads;
}

long long sxs_ft_file_to_indices(const char *eu_path, const char *ft_path, const char *rm_path, struct mol_vector3 *ref_lig,
                                 const double *zvals, int znum, int L, int nthreads, long long **index, int **ft_id,
                                 int **order)
{
	FILE *ft = fopen(ft_path, "rb");
	if (ft == NULL) {
		ERROR_MSG("cannot open ft file");
	}
	fseek(ft, 0, SEEK_END);
	const long size = ftell(ft);
	fseek(ft, 0, SEEK_SET);
	char *data = (char *)malloc((size_t)size + 1);
	CHECK_PTR(data);
	if (size > 0 && fread(data, 1, (size_t)size, ft) != (size_t)size) {
		ERROR_MSG("cannot read ft file");
	}
	fclose(ft);
	data[size] = '\0';
	struct mol_matrix3_list *rots = mol_matrix3_list_from_file(rm_path);
	if (rots == NULL) {
		ERROR_MSG("cannot read rotation file");
	}

	nthreads = thread_count(nthreads, size / 64);
	struct piece pcs[64];
	pthread_t th[64];
	memset(pcs, 0, sizeof(pcs));
	const char *cut = data;
	for (int k = 0; k < nthreads; k++) {
		struct piece *pc = &pcs[k];
		pc->beg = cut;
		const char *stop = data + (size_t)size * (size_t)(k + 1) / (size_t)nthreads;
		if (k == nthreads - 1) {
			stop = data + size;
		} else {
			if (stop < cut) stop = cut;
			while (stop < data + size && *stop != '\n') stop++; /* pieces end at line ends */
			if (stop < data + size) stop++;
		}
		pc->end = stop;
		cut = stop;
		pc->rots = rots; pc->ref_lig = ref_lig; pc->zvals = zvals; pc->znum = znum; pc->L = L;
		pc->want_text = eu_path != NULL;
	}
	if (nthreads == 1) {
		piece_main(&pcs[0]);
	} else {
		for (int k = 0; k < nthreads; k++) {
			if (pthread_create(&th[k], NULL, piece_main, &pcs[k]) != 0) {
				ERROR_MSG("pthread_create failed");
			}
		}
		for (int k = 0; k < nthreads; k++) {
			pthread_join(th[k], NULL);
		}
	}
	int status = 0;
	long long total = 0;
	for (int k = 0; k < nthreads; k++) {
		if (pcs[k].status > status) status = pcs[k].status;
		total += pcs[k].nrows;
	}
	long long kept = -1;
	if (status == 2) {
		ERROR_MSG("rotation index outside the rotation file");
	}
	if (status == 0) {
		if (eu_path != NULL) {
			FILE *eu = fopen(eu_path, "wb");
			if (eu == NULL) {
				ERROR_MSG("cannot open Euler file");
			}
			for (int k = 0; k < nthreads; k++) {
				if (pcs[k].text_len > 0 && fwrite(pcs[k].text, 1, pcs[k].text_len, eu) != pcs[k].text_len) {
					ERROR_MSG("cannot write Euler file");
				}
			}
			fclose(eu);
		}
		*index = (long long *)malloc(sizeof(long long) * (size_t)(total ? total : 1));
		*ft_id = (int *)malloc(sizeof(int) * (size_t)(total ? total : 1));
		*order = (int *)malloc(sizeof(int) * (size_t)(total ? total : 1));
		CHECK_PTR(*index); CHECK_PTR(*ft_id); CHECK_PTR(*order);
		kept = 0;
		long long row = 0;
		for (int k = 0; k < nthreads; k++) {
			for (long long i = 0; i < pcs[k].nrows; i++, row++) {
				if (pcs[k].flat[i] != SXS_OFF_TABLE) {
					(*index)[kept] = pcs[k].flat[i];
					(*ft_id)[kept] = pcs[k].rot_id[i];
					(*order)[kept] = (int)row; /* the serial number counts every row */
					kept++;
				}
			}
		}
	}
	for (int k = 0; k < nthreads; k++) {
		free(pcs[k].flat); free(pcs[k].rot_id); free(pcs[k].text);
	}
	mol_matrix3_list_free(rots);
	free(data);
	return kept;
}

/* ------------------------------------------------------------------ output rows */

struct out_job {
	long long i0, i1;
	const int *order, *ft_id;
	const double *score, *c1, *c2;
	char *text; size_t len;
};

static void *out_main(void *arg)
{
	struct out_job *j = (struct out_job *)arg;
	const size_t cap = (size_t)(j->i1 - j->i0) * 96 + 16;
	j->text = (char *)malloc(cap);
	CHECK_PTR(j->text);
	j->len = 0;
	for (long long i = j->i0; i < j->i1; i++) {
		/* tools/correlate.c:369-375 */
		const int n = snprintf(j->text + j->len, cap - j->len, "%-6d\t%d\t%.3lf\t%.3lf\t%.3lf\n", j->order[i], j->ft_id[i],
		                       j->score[i], j->c1[i], j->c2[i]);
		if (n < 0 || (size_t)n >= cap - j->len) {
			ERROR_MSG("output row does not fit its buffer");
		}
		j->len += (size_t)n;
	}
	return NULL;
}

void sxs_write_score_rows(const char *path, long long n, const int *order, const int *ft_id, const double *score,
                          const double *c1, const double *c2, int nthreads)
{
	FILE *out = fopen(path, "wb");
	if (out == NULL) {
		ERROR_MSG("cannot open output file");
	}
	nthreads = thread_count(nthreads, n);
	struct out_job jobs[64];
	pthread_t th[64];
	for (int k = 0; k < nthreads; k++) {
		struct out_job *j = &jobs[k];
		j->i0 = n * k / nthreads; j->i1 = n * (k + 1) / nthreads;
		j->order = order; j->ft_id = ft_id; j->score = score; j->c1 = c1; j->c2 = c2; j->text = NULL; j->len = 0;
	}
	if (nthreads == 1) {
		out_main(&jobs[0]);
	} else {
		for (int k = 0; k < nthreads; k++) {
			if (pthread_create(&th[k], NULL, out_main, &jobs[k]) != 0) {
				ERROR_MSG("pthread_create failed");
			}
		}
		for (int k = 0; k < nthreads; k++) {
			pthread_join(th[k], NULL);
		}
	}
	for (int k = 0; k < nthreads; k++) {
		if (jobs[k].len > 0 && fwrite(jobs[k].text, 1, jobs[k].len, out) != jobs[k].len) {
			ERROR_MSG("cannot write output file");
		}
		free(jobs[k].text);
	}
	fclose(out);
}
