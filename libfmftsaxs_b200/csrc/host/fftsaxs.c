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

	/* rows per z digit -> contiguous z ranges of roughly equal row count, one per device (partition.c) */
	int dev[SXS_MAX_DEV];
	const int ndev = sxs_host_device_list(dev, SXS_MAX_DEV);
	const long long cell5 = (long long)nb * nb * N * N * N;
	struct sxs_partition part;
	sxs_partition_plan(&part, idx32, idx64, nout, cell5, znum, ndev < SXS_PART_MAX ? ndev : SXS_PART_MAX);

	struct shard_job jobs[SXS_MAX_DEV];
	pthread_t threads[SXS_MAX_DEV];
	const int njobs = part.nshard;
	for (int k = 0; k < njobs; k++) {
		struct shard_job *j = &jobs[k];
		memset(j, 0, sizeof(*j));
		j->plan = plan_for(dev[k], L, qnum, qvals, t);
		j->coefA = coefA; j->coefB = coefB; j->a = params->a; j->bessel = bessel;
		j->mult = params->mult; j->peak = params->peak;
		j->znum = znum; j->z_lo = part.z_lo[k]; j->z_hi = part.z_hi[k];
		j->idx32 = idx32; j->idx64 = idx64; j->nout = nout;
		j->scores = scores; j->c1 = c1; j->c2 = c2;
		j->rows = part.rows[k];
	}
	if (!part.whole) {
		/* the shards' lists: positions on the heap, indices and results in each plan's pinned buffer */
		const size_t isz = idx32 != NULL ? sizeof(int) : sizeof(long long);
		long long *pos[SXS_PART_MAX];
		void *sub[SXS_PART_MAX];
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
			pos[k] = j->pos;
			sub[k] = j->sub;
		}
		sxs_partition_fill(&part, pos, sub);
	}
	sxs_partition_free(&part);
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
