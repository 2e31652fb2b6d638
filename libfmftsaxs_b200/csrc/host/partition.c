/* Pose list -> z shards on host threads (used by sxs_compute_saxs_scores when several devices share one list;
 * csrc/host/fftsaxs.c).  No CUDA in here: tests/test_cpu_host.py drives it through sxs_flat_partition_rows. */
#define _GNU_SOURCE
#include <pthread.h>
#include <unistd.h>

#include "sxs_host.h"

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
	long long *const *pos;     /* [nshard] positions of the shard's rows in the caller's list */
	void *const *sub;          /* [nshard] their indices (int or long long like the list) */
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
		const long long k = j->off[s]++;
		j->pos[s][k] = i;
		if (j->idx32 != NULL) {
			((int *)j->sub[s])[k] = j->idx32[i];
		} else {
			((long long *)j->sub[s])[k] = j->idx64[i];
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


void sxs_partition_plan(struct sxs_partition *P, const int *idx32, const long long *idx64, long long nout,
                        long long cell5, int znum, int nshard_max)
{
	memset(P, 0, sizeof(*P));
	P->idx32 = idx32; P->idx64 = idx64; P->nout = nout; P->cell5 = cell5; P->znum = znum;
	if (nshard_max > SXS_PART_MAX) nshard_max = SXS_PART_MAX;
	if (nshard_max <= 1 || znum >= (int)SXS_Z_NONE || nout <= 0) {
		/* one shard takes the whole z table: no need to look at the list on the host */
		P->whole = 1;
		P->nshard = 1;
		P->z_lo[0] = 0; P->z_hi[0] = znum; P->rows[0] = nout;
		return;
	}
	const int nt = host_threads(nout);
	P->nt = nt;
	P->zdig = (unsigned short *)malloc(sizeof(unsigned short) * (size_t)nout);
	P->cnt_all = (long long *)calloc((size_t)nt * znum, sizeof(long long));
	long long *per_z = (long long *)calloc((size_t)znum, sizeof(long long));
	CHECK_PTR(P->zdig); CHECK_PTR(P->cnt_all); CHECK_PTR(per_z);
	struct part_job pj[32];
	for (int k = 0; k < nt; k++) {
		memset(&pj[k], 0, sizeof(pj[k]));
		pj[k].idx32 = idx32; pj[k].idx64 = idx64; pj[k].cell5 = cell5; pj[k].znum = znum; pj[k].zdig = P->zdig;
		pj[k].i0 = nout * k / nt; pj[k].i1 = nout * (k + 1) / nt;
		pj[k].cnt = P->cnt_all + (size_t)k * znum;
	}
	run_threads(part_pass1, pj, nt);
	long long total = 0;
	int nz_used = 0;
	for (int z = 0; z < znum; z++) {
		for (int k = 0; k < nt; k++) {
			per_z[z] += pj[k].cnt[z];
		}
		total += per_z[z];
		nz_used += per_z[z] > 0;
	}
	int nsh = nshard_max;
	if (nsh > nz_used) {
		nsh = nz_used > 0 ? nz_used : 1;
	}
	/* contiguous z ranges: shard d ends where the running row count reaches total * (d + 1) / nsh */
	int z_next = 0;
	long long done = 0;
	for (int d = 0; d < nsh; d++) {
		const long long want = (total * (d + 1)) / nsh;
		const int z_lo = z_next;
		const long long done_before = done;
		while (z_next < znum && (done < want || d == nsh - 1)) {
			done += per_z[z_next++];
			if (d < nsh - 1 && done >= want) {
				break;
			}
		}
		if (d == nsh - 1) {
			z_next = znum;
		}
		if (z_next == z_lo) {
			continue;
		}
		P->z_lo[P->nshard] = z_lo; P->z_hi[P->nshard] = z_next; P->rows[P->nshard] = done - done_before;
		P->nshard++;
	}
	free(per_z);
	if (P->nshard == 0) { /* no row on the table at all */
		P->nshard = 1;
		P->z_lo[0] = 0; P->z_hi[0] = znum; P->rows[0] = 0;
	}
}

void sxs_partition_fill(struct sxs_partition *P, long long *const *pos, void *const *sub)
{
	if (P->whole) {
		return;
	}
	const int nt = P->nt, ns = P->nshard, znum = P->znum;
	int *shard_of_z = (int *)malloc(sizeof(int) * (size_t)znum);
	long long *off_all = (long long *)calloc((size_t)nt * ns, sizeof(long long));
	CHECK_PTR(shard_of_z); CHECK_PTR(off_all);
	for (int z = 0; z < znum; z++) {
		shard_of_z[z] = -1;
	}
	for (int s = 0; s < ns; s++) {
		for (int z = P->z_lo[s]; z < P->z_hi[s]; z++) {
			shard_of_z[z] = s;
		}
		long long run = 0;
		for (int k = 0; k < nt; k++) {
			off_all[(size_t)k * ns + s] = run;
			for (int z = P->z_lo[s]; z < P->z_hi[s]; z++) {
				run += P->cnt_all[(size_t)k * znum + z];
			}
		}
	}
	struct part_job pj[32];
	for (int k = 0; k < nt; k++) {
		memset(&pj[k], 0, sizeof(pj[k]));
		pj[k].idx32 = P->idx32; pj[k].idx64 = P->idx64; pj[k].cell5 = P->cell5; pj[k].znum = znum; pj[k].zdig = P->zdig;
		pj[k].i0 = P->nout * k / nt; pj[k].i1 = P->nout * (k + 1) / nt;
		pj[k].shard_of_z = shard_of_z; pj[k].off = off_all + (size_t)k * ns; pj[k].pos = pos; pj[k].sub = sub; pj[k].nshard = ns;
	}
	run_threads(part_pass2, pj, nt);
	free(off_all);
	free(shard_of_z);
}

void sxs_partition_free(struct sxs_partition *P)
{
	free(P->zdig);
	free(P->cnt_all);
	P->zdig = NULL;
	P->cnt_all = NULL;
}

/* test access (include/fmftsaxs/sxs_flat.h) */
int sxs_flat_partition_rows(const int *idx32, const long long *idx64, long long nout, long long cell5, int znum,
                            int nshard_max, int *z_lo, int *z_hi, long long *rows, long long *pos_concat,
                            long long *sub_concat)
{
	struct sxs_partition P;
	sxs_partition_plan(&P, idx32, idx64, nout, cell5, znum, nshard_max);
	for (int s = 0; s < P.nshard; s++) {
		z_lo[s] = P.z_lo[s]; z_hi[s] = P.z_hi[s]; rows[s] = P.rows[s];
	}
	if (!P.whole) {
		long long *pos[SXS_PART_MAX];
		void *sub[SXS_PART_MAX];
		void *tmp[SXS_PART_MAX];
		long long at = 0;
		for (int s = 0; s < P.nshard; s++) {
			pos[s] = pos_concat + at;
			tmp[s] = malloc((idx32 != NULL ? sizeof(int) : sizeof(long long)) * (size_t)(P.rows[s] > 0 ? P.rows[s] : 1));
			CHECK_PTR(tmp[s]);
			sub[s] = tmp[s];
			at += P.rows[s];
		}
		sxs_partition_fill(&P, pos, sub);
		at = 0;
		for (int s = 0; s < P.nshard; s++) {
			for (long long k = 0; k < P.rows[s]; k++) {
				sub_concat[at + k] = idx32 != NULL ? (long long)((int *)tmp[s])[k] : ((long long *)tmp[s])[k];
			}
			at += P.rows[s];
			free(tmp[s]);
		}
	}
	const int ns = P.whole ? -P.nshard : P.nshard; /* negative: the list was not partitioned (one shard) */
	sxs_partition_free(&P);
	return ns;
}
