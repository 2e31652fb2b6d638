/* sxs_compute_saxs_scores: host orchestration of the GPU scoring pipeline.
 *
 * Work split (reference: src/fftsaxs.c:608-986, single-threaded, optional MPI over z in the caller):
 *   host, once per L  : 3j d-symbols, Wigner-d on the beta grid, DFT phases (tables.c, cached)
 *   host, per call    : j_p(q z) table with the reference-exact series (tiny: znum*qnum*(2L+1))
 *   GPU               : everything else — rotated coefficient tables, z-translation, cross terms of every
 *                       listed grid point, the (c1,c2) fit, scatter back to list order.
 * z-steps are sharded over the CUDA devices named by SXS_CUDA_DEVICES (default all), one host thread
 * per device, in place of the reference's MPI decomposition (tools/correlate.c:140-147); each device
 * writes only the rows of its own z range, so no reduction is needed.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <unistd.h>

#include "fftsaxs.h"
#include "sxs_host.h"
#include "sxs_tables.h"

int sxs_host_device_list(int *devices, int cap)
{
	const int visible = sxs_cuda_device_count();
	int n = 0;
	const char *env = getenv("SXS_CUDA_DEVICES");
	if (env != NULL && *env) {
		const char *p = env;
		while (*p && n < cap) {
			char *end;
			long v = strtol(p, &end, 10);
			if (end == p) {
				break;
			}
			int seen = 0;
			for (int k = 0; k < n; k++) {
				seen |= devices[k] == (int)v; /* a device named twice is one device: its plan is not shareable */
			}
			if (v >= 0 && v < visible && !seen) {
				devices[n++] = (int)v;
			}
			p = (*end == ',') ? end + 1 : end;
		}
	} else {
		for (int d = 0; d < visible && n < cap; d++) {
			devices[n++] = d;
		}
	}
	if (n == 0) {
		ERROR_MSG("no usable CUDA device: this library has no CPU path");
	}
	return n;
}

int sxs_host_default_device(void)
{
	int dev[64];
	sxs_host_device_list(dev, 64);
	return dev[0];
}

/* One cached plan per device; rebuilt when (L, q grid) changes. */
#define SXS_MAX_DEV 64
static struct {
	sxs_cuda_plan *plan;
	int L, qnum;
	double *qvals;
} g_plans[SXS_MAX_DEV];

static sxs_cuda_plan *plan_for(int device, int L, int qnum, const double *qvals, const struct sxs_l_tables *t)
{
	if (device < 0 || device >= SXS_MAX_DEV) {
		ERROR_MSG("device id out of range");
	}
	if (g_plans[device].plan != NULL && g_plans[device].L == L && g_plans[device].qnum == qnum &&
	    !memcmp(g_plans[device].qvals, qvals, sizeof(double) * qnum)) {
		return g_plans[device].plan;
	}
	if (g_plans[device].plan != NULL) {
		sxs_cuda_plan_destroy(g_plans[device].plan);
		free(g_plans[device].qvals);
		g_plans[device].plan = NULL;
	}
	sxs_cuda_plan *p = sxs_cuda_plan_create(device, L, qnum, qvals, t->dsymb, t->dwig, t->twiddle);
	if (p == NULL) {
		fprintf(stderr, "[Error] sxs_cuda_plan_create: %s\n", sxs_cuda_last_error());
		exit(EXIT_FAILURE);
	}
	g_plans[device].plan = p;
	g_plans[device].L = L;
	g_plans[device].qnum = qnum;
	g_plans[device].qvals = (double *)malloc(sizeof(double) * qnum);
	memcpy(g_plans[device].qvals, qvals, sizeof(double) * qnum);
	return p;
}

struct shard_job {
	sxs_cuda_plan *plan;
	const double *coefA, *coefB, *a, *bessel;
	double mult, peak;
	int znum, z_lo, z_hi;
	const int *idx32;
	const long long *idx64;
	long long nout;
	double *scores, *c1, *c2;
	/* several devices: the rows of this shard's z range, extracted by partition_rows() */
	int compact;
	long long rows;      /* rows of the shard */
	long long *pos;      /* their positions in the caller's list */
	void *sub;           /* their indices (int or long long), in the plan's pinned buffer */
	double *out;         /* [3][rows] results, same buffer */
	int rc;
	char err[512]; /* the CUDA layer's message of this thread (its error buffer is thread-local) */
};

/* ---- rows -> z shards on host threads ---------------------------------------------------------------------------
 * With several devices every device gets the rows of its own z range as a compact list, so that uploads, the key sort
 * and the scatter scale with the shard and not with the whole list (the reference's MPI ranks do the same filtering
 * while reading the Euler file, tools/correlate.c:169-251).  Two passes over the list, both split over host threads:
 * (1) z digit of every row (kept, two bytes per row) and a histogram per thread — the shard boundaries follow from the
 * summed histogram; (2) every thread writes its rows into the shards' lists at offsets known from the histograms, so
 * the lists keep the input order.  The z digit is index / cell5; it is taken through a reciprocal and corrected with
 * two multiplications instead of a 64-bit division per row (the r2x 8-GPU run spent more time in those divisions —
 * three passes, two of them once per device — than on the GPUs). */
#define SXS_Z_NONE 0xFFFFu
struct part_job {
	const int *idx32;
	const long long *idx64;
	long long i0, i1, cell5;
	int znum, nshard;
	unsigned short *zdig;
	long long *cnt;            /* [znum] rows per z digit in [i0, i1) */
	const int *shard_of_z;     /* pass 2 */
	long long *off;            /* [nshard] write offsets of this thread, pass 2 */
	struct shard_job *jobs;
};

static void *part_pass1(void *arg)
{
	struct part_job *j = (struct part_job *)arg;
	const double inv = 1.0 / (double)j->cell5;
	for (long long i = j->i0; i < j->i1; i++) {
		const long long v = j->idx32 != NULL ? (long long)j->idx32[i] : j->idx64[i];
		unsigned short zd = SXS_Z_NONE;
		if (v >= 0) {
			long long z = (long long)((double)v * inv);
			if (z * j->cell5 > v) {
				z--;
			} else if ((z + 1) * j->cell5 <= v) {
				z++;
			}
			if (z < j->znum) {
				zd = (unsigned short)z;
				j->cnt[z]++;
			}
		}
		j->zdig[i] = zd;
	}
	return NULL;
}

static void *part_pass2(void *arg)
{
	struct part_job *j = (struct part_job *)arg;
	for (long long i = j->i0; i < j->i1; i++) {
		const unsigned short zd = j->zdig[i];
		if (zd == SXS_Z_NONE) {
			continue;
		}
		const int s = j->shard_of_z[zd];
		if (s < 0) {
			continue;
		}
		struct shard_job *sj = &j->jobs[s];
		const long long k = j->off[s]++;
		sj->pos[k] = i;
		if (j->idx32 != NULL) {
			((int *)sj->sub)[k] = j->idx32[i];
		} else {
			((long long *)sj->sub)[k] = j->idx64[i];
		}
	}
	return NULL;
}

static int host_threads(long long n)
{
	int t = getenv("SXS_HOST_THREADS") ? atoi(getenv("SXS_HOST_THREADS")) : (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (t > 32) t = 32;
	if ((long long)t > n / 65536 + 1) t = (int)(n / 65536 + 1);
	return t < 1 ? 1 : t;
}

static void run_threads(void *(*fn)(void *), struct part_job *pj, int nt)
{
	pthread_t th[32];
	int started = 0;
	for (int k = 1; k < nt; k++) {
		if (pthread_create(&th[k], NULL, fn, &pj[k]) != 0) {
			break;
		}
		started = k;
	}
	fn(&pj[0]);
	for (int k = started + 1; k < nt; k++) {
		fn(&pj[k]);
	}
	for (int k = 1; k <= started; k++) {
		pthread_join(th[k], NULL);
	}
}

static int shard_score_compact(struct shard_job *j)
{
	const long long n = j->rows;
	if (n == 0) {
		return 0;
	}
	int rc;
	if (j->idx32 != NULL) {
		rc = sxs_cuda_plan_score_i32(j->plan, (const int *)j->sub, n, j->z_lo, j->z_hi, j->out, j->out + n, j->out + 2 * n);
	} else {
		rc = sxs_cuda_plan_score_i64(j->plan, (const long long *)j->sub, n, j->z_lo, j->z_hi, j->out, j->out + n, j->out + 2 * n);
	}
	if (rc == 0) {
		/* every row of the compact list lies in [z_lo, z_hi): all of them were scored */
		for (long long k = 0; k < n; k++) {
			j->scores[j->pos[k]] = j->out[k]; j->c1[j->pos[k]] = j->out[n + k]; j->c2[j->pos[k]] = j->out[2 * n + k];
		}
	}
	return rc;
}

static void *shard_main(void *arg)
{
	struct shard_job *j = (struct shard_job *)arg;
	j->rc = sxs_cuda_plan_set_molecules(j->plan, j->coefA, j->coefB);
	if (j->rc == 0) j->rc = sxs_cuda_plan_set_experiment(j->plan, j->a, j->mult, j->peak);
	if (j->rc == 0) j->rc = sxs_cuda_plan_set_translations(j->plan, j->bessel, j->znum);
	if (j->rc == 0) {
		if (j->compact) {
			j->rc = shard_score_compact(j);
		} else if (j->idx32 != NULL) {
			j->rc = sxs_cuda_plan_score_i32(j->plan, j->idx32, j->nout, j->z_lo, j->z_hi, j->scores, j->c1, j->c2);
		} else {
			j->rc = sxs_cuda_plan_score_i64(j->plan, j->idx64, j->nout, j->z_lo, j->z_hi, j->scores, j->c1, j->c2);
		}
	}
	if (j->rc != 0 && j->err[0] == 0) {
		snprintf(j->err, sizeof(j->err), "%s", sxs_cuda_last_error());
	}
	return NULL;
}

/* The plan cache and the per-device plans are process-wide state: concurrent callers take turns (the reference
 * itself is non-reentrant, SURVEY 8b "Threading"). */
static pthread_mutex_t g_score_lock = PTHREAD_MUTEX_INITIALIZER;

static void score_impl(double *scores, double *c1, double *c2, const int *idx32, const long long *idx64,
                       long long nout, struct sxs_spf_full *A, struct sxs_spf_full *B, struct sxs_opt_params *params,
                       double *qvals, int qnum, double *zvals, int znum, int L)
{
	CHECK_PTR(A);
	CHECK_PTR(B);
	CHECK_PTR(params);
	if (nout <= 0 || znum <= 0) {
		return;
	}
	pthread_mutex_lock(&g_score_lock);
	const struct sxs_l_tables *t = sxs_l_tables_get(L);
	const int N = 2 * L + 1, nb = L + 1;

	const size_t ncoef = (size_t)3 * qnum * nb * nb * 2;
	double *coefA = (double *)malloc(sizeof(double) * ncoef);
	double *coefB = (double *)malloc(sizeof(double) * ncoef);
	CHECK_PTR(coefA); CHECK_PTR(coefB);
	sxs_spf_full_pack(A, coefA);
	sxs_spf_full_pack(B, coefB);
	/* j_p(q z) with the reference-exact series: znum*qnum*(2L+1) values, ~10 ms for the 64 x 50 x 31 table of config 3 —
	 * kept between calls with the same (z table, q grid, L) (callers score list after list on one table) */
	static struct { double *tab, *z, *q; int znum, qnum, L; } bc;
	if (bc.tab == NULL || bc.znum != znum || bc.qnum != qnum || bc.L != L || memcmp(bc.z, zvals, sizeof(double) * znum) ||
	    memcmp(bc.q, qvals, sizeof(double) * qnum)) {
		free(bc.tab); free(bc.z); free(bc.q);
		bc.tab = (double *)malloc(sizeof(double) * (size_t)znum * qnum * N);
		bc.z = (double *)malloc(sizeof(double) * znum);
		bc.q = (double *)malloc(sizeof(double) * qnum);
		CHECK_PTR(bc.tab); CHECK_PTR(bc.z); CHECK_PTR(bc.q);
		memcpy(bc.z, zvals, sizeof(double) * znum);
		memcpy(bc.q, qvals, sizeof(double) * qnum);
		bc.znum = znum; bc.qnum = qnum; bc.L = L;
		sxs_fill_bessel_table(bc.tab, zvals, znum, qvals, qnum, L);
	}
	const double *bessel = bc.tab;

	/* rows per z digit -> contiguous z ranges of roughly equal row count, one per device */
	int dev[SXS_MAX_DEV];
	int ndev = sxs_host_device_list(dev, SXS_MAX_DEV);
	long long *per_z = (long long *)calloc((size_t)znum, sizeof(long long));
	CHECK_PTR(per_z);
	const long long cell5 = (long long)nb * nb * N * N * N;
	long long total = 0;
	unsigned short *zdig = NULL;
	struct part_job pj[32];
	long long *cnt_all = NULL;
	int nt = 1;
	if (ndev == 1 || znum >= (int)SXS_Z_NONE) {
		/* one device takes the whole z table: no need to look at the list on the host */
		ndev = 1;
		per_z[0] = total = nout;
	} else {
		nt = host_threads(nout);
		zdig = (unsigned short *)malloc(sizeof(unsigned short) * (size_t)nout);
		cnt_all = (long long *)calloc((size_t)nt * znum, sizeof(long long));
		CHECK_PTR(zdig); CHECK_PTR(cnt_all);
		for (int k = 0; k < nt; k++) {
			memset(&pj[k], 0, sizeof(pj[k]));
			pj[k].idx32 = idx32; pj[k].idx64 = idx64; pj[k].cell5 = cell5; pj[k].znum = znum; pj[k].zdig = zdig;
			pj[k].i0 = nout * k / nt; pj[k].i1 = nout * (k + 1) / nt;
			pj[k].cnt = cnt_all + (size_t)k * znum;
		}
		run_threads(part_pass1, pj, nt);
		for (int k = 0; k < nt; k++) {
			for (int z = 0; z < znum; z++) {
				per_z[z] += pj[k].cnt[z];
				total += pj[k].cnt[z];
			}
		}
	}
	int nz_used = 0;
	for (int z = 0; z < znum; z++) {
		nz_used += per_z[z] > 0;
	}
	if (ndev > nz_used) {
		ndev = nz_used > 0 ? nz_used : 1;
	}
	const int whole = ndev == 1;
	if (whole) {
		per_z[0] = 0; /* the single shard below spans [0, znum) regardless of the histogram */
	}

	struct shard_job jobs[SXS_MAX_DEV];
	pthread_t threads[SXS_MAX_DEV];
	int *shard_of_z = (int *)malloc(sizeof(int) * (size_t)znum);
	CHECK_PTR(shard_of_z);
	for (int z = 0; z < znum; z++) {
		shard_of_z[z] = -1;
	}
	int z_next = 0;
	long long done = 0;
	int njobs = 0;
	for (int d = 0; d < ndev; d++) {
		long long want = (total * (d + 1)) / ndev;
		int z_lo = z_next;
		const long long done_before = done;
		while (z_next < znum && (done < want || d == ndev - 1)) {
			done += per_z[z_next++];
			if (d < ndev - 1 && done >= want) {
				break;
			}
		}
		if (d == ndev - 1) {
			z_next = znum;
		}
		if (z_next == z_lo) {
			continue;
		}
		struct shard_job *j = &jobs[njobs];
		memset(j, 0, sizeof(*j));
		j->plan = plan_for(dev[d], L, qnum, qvals, t);
		j->coefA = coefA; j->coefB = coefB; j->a = params->a; j->bessel = bessel;
		j->mult = params->mult; j->peak = params->peak;
		j->znum = znum; j->z_lo = z_lo; j->z_hi = z_next;
		j->idx32 = idx32; j->idx64 = idx64; j->nout = nout;
		j->scores = scores; j->c1 = c1; j->c2 = c2;
		j->rows = done - done_before;
		for (int z = z_lo; z < z_next; z++) {
			shard_of_z[z] = njobs;
		}
		njobs++;
	}
	if (!whole && njobs > 0) {
		/* the shards' lists: positions on the heap, indices and results in each plan's pinned buffer */
		const size_t isz = idx32 != NULL ? sizeof(int) : sizeof(long long);
		for (int k = 0; k < njobs; k++) {
			struct shard_job *j = &jobs[k];
			j->compact = 1;
			const size_t n = (size_t)(j->rows > 0 ? j->rows : 1);
			const size_t sub_bytes = (isz * n + 15) & ~(size_t)15;
			j->pos = (long long *)malloc(sizeof(long long) * n);
			unsigned char *buf = (unsigned char *)sxs_cuda_plan_host_buffer(j->plan, sub_bytes + sizeof(double) * 3 * n);
			if (j->pos == NULL || buf == NULL) {
				ERROR_MSG("out of host memory for the z shards of the pose list");
			}
			j->sub = buf;
			j->out = (double *)(buf + sub_bytes);
		}
		long long *off_all = (long long *)calloc((size_t)nt * njobs, sizeof(long long));
		CHECK_PTR(off_all);
		for (int s = 0; s < njobs; s++) {
			long long run = 0;
			for (int k = 0; k < nt; k++) {
				off_all[(size_t)k * njobs + s] = run;
				for (int z = jobs[s].z_lo; z < jobs[s].z_hi; z++) {
					run += pj[k].cnt[z];
				}
			}
		}
		for (int k = 0; k < nt; k++) {
			pj[k].shard_of_z = shard_of_z; pj[k].off = off_all + (size_t)k * njobs; pj[k].jobs = jobs; pj[k].nshard = njobs;
		}
		run_threads(part_pass2, pj, nt);
		free(off_all);
	}
	free(zdig);
	free(cnt_all);
	free(shard_of_z);
	if (njobs == 1) {
		shard_main(&jobs[0]);
	} else {
		for (int k = 0; k < njobs; k++) {
			if (pthread_create(&threads[k], NULL, shard_main, &jobs[k]) != 0) {
				ERROR_MSG("pthread_create failed");
			}
		}
		for (int k = 0; k < njobs; k++) {
			pthread_join(threads[k], NULL);
		}
	}
	for (int k = 0; k < njobs; k++) {
		free(jobs[k].pos);
	}
	for (int k = 0; k < njobs; k++) {
		if (jobs[k].rc != 0) {
			fprintf(stderr, "[Error] sxs_compute_saxs_scores: CUDA layer failed: %s\n", jobs[k].err);
			exit(EXIT_FAILURE);
		}
	}
	free(per_z);
	free(coefA);
	free(coefB);
	pthread_mutex_unlock(&g_score_lock);
}

void sxs_compute_saxs_scores(double *scores_list, double *c1_list, double *c2_list, int *index_list, int nout,
                             struct sxs_spf_full *A, struct sxs_spf_full *B, struct sxs_opt_params *params,
                             double *qvals, int qnum, double *zvals, int znum, int L, int skip)
{
	(void)skip;
	score_impl(scores_list, c1_list, c2_list, index_list, NULL, nout, A, B, params, qvals, qnum, zvals, znum, L);
}

void sxs_compute_saxs_scores64(double *scores_list, double *c1_list, double *c2_list, const long long *index_list,
                               long long nout, struct sxs_spf_full *A, struct sxs_spf_full *B,
                               struct sxs_opt_params *params, double *qvals, int qnum, double *zvals, int znum,
                               int L, int skip)
{
	(void)skip;
	score_impl(scores_list, c1_list, c2_list, NULL, index_list, nout, A, B, params, qvals, qnum, zvals, znum, L);
}
