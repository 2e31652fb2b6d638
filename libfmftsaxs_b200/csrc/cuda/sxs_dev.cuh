/* Internal declarations shared by the CUDA translation units (sm_100a only). */
#ifndef SXS_DEV_CUH
#define SXS_DEV_CUH

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sxs_cuda.h"

extern "C" void sxs_cuda_set_error(const char *fmt, ...);

#define SXS_CK(call)                                                                                   \
	do {                                                                                               \
		cudaError_t e_ = (call);                                                                       \
		if (e_ != cudaSuccess) {                                                                       \
			sxs_cuda_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			return -1;                                                                                 \
		}                                                                                              \
	} while (0)

#define SXS_CK_LAUNCH()                                                                               \
	do {                                                                                              \
		cudaError_t e_ = cudaGetLastError();                                                          \
		if (e_ != cudaSuccess) {                                                                      \
			sxs_cuda_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
			return -1;                                                                                \
		}                                                                                             \
	} while (0)

/* number of (m, l) pairs with 0 <= m <= l <= L, and the packed index of one */
__host__ __device__ inline int sxs_ml_count(int L) { return (L + 1) * (L + 2) / 2; }
__host__ __device__ inline int sxs_ml_index(int L, int m, int l) { return m * (L + 1) - m * (m - 1) / 2 + (l - m); }

/* Rows of the rotated/translated tables hold N = 2L+1 complex values and are padded to a multiple of 8 (128 bytes),
 * so that the 8 values g = 8k .. 8k+7 — one "band" — are exactly one cache line. */
__host__ __device__ inline int sxs_row_pad(int N) { return (N + 7) & ~7; }

/* Sort key of a pose (z, b1, b2, a2, g1, g2 = the digits of the reference's flat index, src/index.c:9-29):
 * (z, b2, b1, g2 / 8, g1, g2 % 8, a2), most significant first.  Cells (z, b2, b1) are contiguous key ranges of
 * sxs_keys_per_cell(N) keys, points that differ only in a2 are neighbours (key / N equal), and inside a cell the
 * 128-byte band g2 / 8 of the ligand operand comes first. */
struct sxs_pose_digits {
	int z, b1, b2, a2, g1, g2;
};
__host__ __device__ inline unsigned long long sxs_keys_per_cell(int N)
{
	return (unsigned long long)(sxs_row_pad(N) / 8) * N * 8 * N;
}
__host__ __device__ inline unsigned long long sxs_key_pack(int nb, int N, const sxs_pose_digits &d)
{
	const unsigned long long nband = sxs_row_pad(N) / 8;
	unsigned long long key = ((unsigned long long)d.z * nb + d.b2) * nb + d.b1;
	key = ((key * nband + d.g2 / 8) * N + d.g1) * 8 + d.g2 % 8;
	return key * N + d.a2;
}
__host__ __device__ inline sxs_pose_digits sxs_key_unpack(int nb, int N, unsigned long long key)
{
	const int nband = sxs_row_pad(N) / 8;
	sxs_pose_digits d;
	d.a2 = (int)(key % N); key /= N;
	d.g2 = (int)(key % 8); key /= 8;
	d.g1 = (int)(key % N); key /= N;
	d.g2 += 8 * (int)(key % nband); key /= nband;
	d.b1 = (int)(key % nb); key /= nb;
	d.b2 = (int)(key % nb);
	d.z = (int)(key / nb);
	return d;
}

/* Cross terms between K3 and K4: one contiguous row of 6*qnum doubles per distinct grid point, node-major,
 * X[(p*qnum + q)*6 + k] (k = VV, VD, VW, DD, DW, WW), 16-byte aligned.  K4 gives every lane its own row and streams it
 * through a per-lane shared-memory ring with 16-byte asynchronous copies, so a lane can take its next point the moment
 * its fit ends.  (Round 1 kept tiles of 32 points for coalesced warp loads, which ties the 32 fits of a warp together:
 * with 2..60 evaluations per fit only 56 % of the lanes were active in the objective, profiles/r2_k4_notes.md.) */
__host__ __device__ inline size_t sxs_x_index(long long p, int qnum, int q, int k)
{
	return ((size_t)p * qnum + q) * 6 + k;
}

/* ---- launchers implemented in sxs_exact.cu (compiled with -fmad=false) ---- */

/* K4: one fit per point.  x: cross terms in the layout of sxs_x_index; res[p*4] = chi, c1, c2,
 * evaluations.  d_ticket: one device word used as the work queue head (reset by the launcher). */
int sxs_launch_fit(const double *d_x, long long npts, const double *d_a, const double *d_qvals, int qnum, double mult,
                   double peak, int rescale, double *d_res, unsigned long long *d_ticket, cudaStream_t stream);

/* Self terms of the pair (A, B) in comp_const_int order (src/fftsaxs.c:27-50,638-643); out[k*qnum + q],
 * VD/VW/DW already doubled as fill_const applies them (src/fftsaxs.c:76-81). */
int sxs_launch_pair_const(const double *d_coefA, const double *d_coefB, int qnum, int L, double *d_out,
                          cudaStream_t stream);

#endif
