/* Internal helpers shared by the host sources. */
#ifndef SXS_HOST_H
#define SXS_HOST_H
#include "common.h"
#include "sxs_cuda.h"

/* Device used by single-device entry points: first id of SXS_CUDA_DEVICES, else 0. */
int sxs_host_default_device(void);
/* Parses SXS_CUDA_DEVICES ("0,1,2"); default = all visible devices.  Returns the count. */
int sxs_host_device_list(int *devices, int cap);

/* Pose list -> contiguous z ranges with roughly equal row counts, and the compact list of every range (partition.c).
 * plan: one threaded pass (z digit of every row, histogram) and the ranges; fill: a second threaded pass that writes
 * every shard's row positions and indices, in input order, into the caller's buffers (rows[s] entries each). */
#define SXS_PART_MAX 64
struct sxs_partition {
	int nshard;                        /* shards that hold at least one z step */
	int z_lo[SXS_PART_MAX], z_hi[SXS_PART_MAX];
	long long rows[SXS_PART_MAX];
	int whole;                         /* 1: a single shard spans the whole table, the list was not looked at */
	const int *idx32;
	const long long *idx64;
	long long nout, cell5;
	int znum, nt;
	unsigned short *zdig;
	long long *cnt_all;
};
void sxs_partition_plan(struct sxs_partition *P, const int *idx32, const long long *idx64, long long nout,
                        long long cell5, int znum, int nshard_max);
void sxs_partition_fill(struct sxs_partition *P, long long *const *pos, void *const *sub);
void sxs_partition_free(struct sxs_partition *P);

#define SXS_CUDA_CHECK(call) do {                                                      \
	if ((call) != 0) {                                                                 \
		fprintf(stderr, "[Error] %s, function %s, line %i: CUDA layer: %s\n",          \
		        __FILE__, __func__, __LINE__, sxs_cuda_last_error());                  \
		exit(EXIT_FAILURE);                                                            \
	}                                                                                  \
} while (0)

#endif
