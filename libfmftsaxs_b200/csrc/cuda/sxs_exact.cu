/* Kernels whose arithmetic must follow the reference operation for operation; this translation
 * unit is compiled with -fmad=false so that nvcc never fuses a multiply into an add.
 *
 *   k_fit            K4, the fused per-conformation (c1, c2) fit: peak rescale + L-BFGS-B + sqrt(f)
 *                    (src/min_saxs.c:153-259 and the vendored lbfgsb/src/ of the reference)
 *   k_fit_eval       best scale / objective / gradient at one point (src/min_saxs.c:3-105,261-319)
 *   k_pair_const     I_A + I_B self terms (src/fftsaxs.c:27-50)
 *   k_self_terms     six self cross terms of one molecule (src/min_saxs.c:437-499)
 *   k_profile        I(q) of one molecule (src/profile.c:186-233)
 *
 *   k_ft_rows_to_index   ft rows -> flat grid indices (src/index.c:38-75,114, tools/correlate.c:214-247)
 *
 * K4 evaluates the objective as the reference does — two passes over the 6*qnum cross terms of the point per
 * evaluation, exp() by the reference libm's algorithm (exp_glibc.h) — and is bit-identical to the vendored L-BFGS-B on
 * equal cross terms; 7-13 evaluations per fit on real data, ~20 on the synthetic bench workload.
 */
#include <math.h>

#include "sxs_dev.cuh"

#define SXS_HD __host__ __device__ __forceinline__
#define SXS_ROWMAJOR_VEC 1
/* The one-pass form of round 1 (algebraically equal, 27 % cheaper per evaluation; see fit_eval.h)
 * left 17 of 1131 / 45 of 70 000 real 4G9S rows beyond 1e-6 in c2 against 3 / 10 for the two-pass form
 * (gpurun_out/r2a_pytest_*.txt): the kernel does not offer it any more. */
#include "fit_point.h"

#ifndef SXS_FIT_THREADS
#define SXS_FIT_THREADS 256
#endif
#ifndef SXS_FIT_MINBLOCKS
#define SXS_FIT_MINBLOCKS 1
#endif

/* part B of the optimiser runs when NUM/DEN of the block's running fits wait for it (or none can evaluate) */
#ifndef SXS_FIT_BATCH_NUM
#define SXS_FIT_BATCH_NUM 3
#endif
#ifndef SXS_FIT_BATCH_DEN
#define SXS_FIT_BATCH_DEN 4
#endif

/* 2^(k/128) table of the libm-faithful exp (exp_glibc.h); k_fit copies it to shared memory */
__device__ const uint64_t d_exp_tab[SXS_EXP_TABLE_ENTRIES] = SXS_EXP_TABLE_INIT;

/* Per-block tables of the fit, statically placed so that every device function addresses them as shared memory
 * (through a generic pointer handed across a call the objective's loads became generic LD + R2UR/UMOV address
 * traffic: 20 % of its instructions, r2g ncu). */
#define SXS_FIT_MAXQ 128
__shared__ double fs_a[6 * SXS_FIT_MAXQ];   /* compressed experiment a[q*6 + 0..5] */
__shared__ double fs_q[SXS_FIT_MAXQ];       /* q grid */
__shared__ double fs_rq[SXS_FIT_MAXQ];      /* 1 / (q_i - q_{i-1}) */
__shared__ double fs_dq[SXS_FIT_MAXQ];      /* q_i - q_{i-1}, q_{-1} = -1 */
__shared__ uint64_t fs_etab[SXS_EXP_TABLE_ENTRIES];
__shared__ int fs_fast; /* |corr q^2| < 512 for every c1 of the box and every node: the exp core needs no range test */

__device__ __forceinline__ void fit_tables_fill(const double *a, const double *qvals, int qnum, double mult, int allow_fast)
{
	if (threadIdx.x == 0) {
		/* |c1^2 - 1| <= max(1 - L1^2, U1^2 - 1) inside the box the optimiser never leaves */
		const double span = fmax(1.0 - SXS_C1_LOWER * SXS_C1_LOWER, SXS_C1_UPPER * SXS_C1_UPPER - 1.0);
		double qq = 0.0;
		for (int i = 0; i < qnum; i++) {
			qq = fmax(qq, qvals[i] * qvals[i]);
		}
		const double bound = fabs(mult) * span * qq;
		fs_fast = (allow_fast && bound < 500.0) ? 1 : 0; /* false for NaN as well */
	}
	for (int i = threadIdx.x; i < SXS_EXP_TABLE_ENTRIES; i += blockDim.x) {
		fs_etab[i] = d_exp_tab[i];
	}
	for (int i = threadIdx.x; i < 6 * qnum; i += blockDim.x) {
		fs_a[i] = a[i];
	}
	for (int i = threadIdx.x; i < qnum; i += blockDim.x) {
		fs_q[i] = qvals[i];
		fs_dq[i] = qvals[i] - (i > 0 ? qvals[i - 1] : -1.0);
		fs_rq[i] = 1.0 / (qvals[i] - (i > 0 ? qvals[i - 1] : -1.0));
	}
}

__device__ __forceinline__ void fit_store(const struct lq_state *st, double *__restrict__ res, long long p)
{
	res[p * 4 + 0] = sqrt(st->f);
	res[p * 4 + 1] = st->x[1];
	res[p * 4 + 2] = st->x[2];
	res[p * 4 + 3] = (double)st->nfgv;
}

/* ---- K4 -------------------------------------------------------------------------------------------------------
 * Every lane owns one fit: it runs the reverse-communication optimiser (lbfgsb_lean.h, all state in registers) until
 * the optimiser asks for the objective at a trial (c1, c2); the lanes of a warp that want the objective then evaluate
 * it together.  A lane whose fit ends takes the next point from the ticket counter at once.
 *
 * The objective walks the point's row of 6*qnum cross terms twice (best scale, then f and gradient: the reference's
 * two passes, src/min_saxs.c:261-319 and :3-105).  The row is streamed from L2/HBM through a per-lane ring in shared
 * memory with 16-byte asynchronous copies (cp.async, three per node), SXS_FIT_RING - 1 nodes ahead of the arithmetic.
 * A node is straight-line code: exp through its core without the range test, the pass-1 quotient through the
 * reciprocal table (both bit-exact, fit_eval.h) — 323 executed instructions per node and evaluation (127 + 196) against
 * 362 for the form with the two range branches, no branch and no spill in either loop (r2n / r2s ncu).
 *
 * Measured and not kept (profiles/r2_k4_notes.md, r2o-r2ae; none of these is left in the code): two to four nodes per
 * trip (the fixed-latency "wait" share of the objective's stall samples falls from 45 % to 21 %, but a trip needs ~60
 * more registers inside a kernel that is at 255: its spills cost what the interleaving wins), rows prefetched through
 * registers instead of the ring (32 lanes = 32 cache lines per load: L1 thrashes), G(c1, q_i) kept in shared memory for
 * pass 2 (one exp per node less, but 100 KB less L1 for the kernel's spills), the optimiser's matrices parked in
 * shared memory across the objective, G of node i + 1 formed during node i, the six terms of node i + 1 carried in
 * registers.  Every one of them lands within 57-72 ms per 1.12 M fits against 55 for this form and 63 for the form
 * before it.
 *
 * The iteration boundary (part B: BFGS update, Cauchy point, subspace step) runs when NUM/DEN of the warp's waiting
 * lanes wait for it, so that it executes with most lanes active. */
#ifndef SXS_FIT_RING
#define SXS_FIT_RING 4
#endif

struct fit_eval_args {
	const double *row;
	int qnum;
	double mult, scale;
};

/* dynamic shared memory of k_fit: the lanes' row rings, SXS_FIT_RING slots of 48 bytes per lane; the slots of one ring
 * position are lane-contiguous (48-byte lane stride: conflict-free 16-byte shared loads per quarter warp) */
extern __shared__ __align__(16) double s_fit_dyn[];
#define SXS_FIT_SLOT_BYTES (SXS_FIT_THREADS * 48)

__device__ __forceinline__ void fit_cp16(unsigned dst_smem, const double *src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void fit_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void fit_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void fit_take(unsigned src, double scale, struct sxs_six *x)
{
	double2 p0, p1, p2;
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p0.x), "=d"(p0.y) : "r"(src));
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p1.x), "=d"(p1.y) : "r"(src + 16));
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p2.x), "=d"(p2.y) : "r"(src + 32));
	x->vv = p0.x * scale; x->vd = p0.y * scale;
	x->vw = p1.x * scale; x->dd = p1.y * scale;
	x->dw = p2.x * scale; x->ww = p2.y * scale;
}

/* f, g of the fit at (c1, c2): sxs_fit_eval_fast (fit_eval.h) with the row arriving through the ring.
 *
 * The stream is the row twice (pass 1, pass 2); position s is requested SXS_FIT_RING - 1 positions ahead of its use
 * (with one position of lead the wait for the copy was 26 % of the loop's stall samples, r2r ncu).  The node loops are
 * ROLLED on purpose (unrolled by the ring size the kernel ran 2.5x slower on instruction-cache misses,
 * profiles/r2_k4_notes.md); ring slot and row pointer advance by pointer increments.  Node 0 of each pass is peeled:
 * it also seeds the running values (begin), and a test inside the loop would split the trip's straight-line code.
 *
 * A real call (noinline): the caller holds ~150 registers of optimiser state that are dead weight in here; behind a
 * call boundary the loops get a register allocation of their own. */
#ifdef SXS_FIT_EVAL_INLINE
#define SXS_FIT_EVAL_LINKAGE __device__ __forceinline__
#else
#define SXS_FIT_EVAL_LINKAGE __device__ __noinline__
#endif
SXS_FIT_EVAL_LINKAGE void fit_eval_fast(const struct fit_eval_args *ap, double c1, double c2, double *out3)
{
	struct sxs_fit_ctx ctx;
	ctx.x = ap->row; ctx.stride = 1; ctx.qstride = 6;
	ctx.a = fs_a; ctx.qvals = fs_q; ctx.qnum = ap->qnum; ctx.mult = ap->mult; ctx.scale = ap->scale;
	ctx.rq = fs_rq; ctx.dq = fs_dq; ctx.etab = fs_etab;
	const double *row = ap->row;
	const double scale = ap->scale;
	const int Q = ap->qnum;
	const double corr = -ap->mult * (c1 * c1 - 1.0);
	const double c1_cube = c1 * c1 * c1;
	const unsigned ring = (unsigned)__cvta_generic_to_shared(s_fit_dyn) + threadIdx.x * 48u;
	const unsigned ring_end = ring + SXS_FIT_RING * SXS_FIT_SLOT_BYTES;

	/* positions 0 .. RING-2 */
	{
		const int total = 2 * Q;
#pragma unroll
		for (int t = 0; t < SXS_FIT_RING - 1; t++) {
			if (t < total) {
				const double *src = row + (t % Q) * 6;
				const unsigned dst = ring + (unsigned)t * SXS_FIT_SLOT_BYTES;
				fit_cp16(dst, src);
				fit_cp16(dst + 16, src + 2);
				fit_cp16(dst + 32, src + 4);
			}
			fit_cp_commit();
		}
	}
	struct sxs_six x;
	struct sxs_scale_run sr;
	struct sxs_grad_run gr;
	unsigned cur = ring;                                            /* slot of the position being consumed */
	unsigned nxt = ring + (SXS_FIT_RING - 1) * SXS_FIT_SLOT_BYTES;  /* slot of the position being requested */
	const double *src = row + ((SXS_FIT_RING - 1) % Q) * 6;         /* its place in the row */
	const double *row_end = row + Q * 6;
	int left = 2 * Q - (SXS_FIT_RING - 1);                          /* positions still to request */
#define SXS_FIT_STEP()                                                         \
	do {                                                                       \
		if (left > 0) {                                                        \
			fit_cp16(nxt, src);                                                \
			fit_cp16(nxt + 16, src + 2);                                       \
			fit_cp16(nxt + 32, src + 4);                                       \
		}                                                                      \
		fit_cp_commit();                                                       \
		left--;                                                                \
		src += 6;                                                              \
		if (src == row_end) src = row;                                         \
		nxt += SXS_FIT_SLOT_BYTES;                                             \
		if (nxt == ring_end) nxt = ring;                                       \
		fit_cp_wait<SXS_FIT_RING - 1>();                                       \
		fit_take(cur, scale, &x);                                              \
		cur += SXS_FIT_SLOT_BYTES;                                             \
		if (cur == ring_end) cur = ring;                                       \
	} while (0)
#define SXS_FIT_G(i) (c1_cube * sxs_exp_glibc_core(corr * fs_q[i] * fs_q[i], fs_etab))

	/* ---- pass 1: best scale ---- */
	SXS_FIT_STEP();
	{
		const double G = SXS_FIT_G(0);
		sxs_scale_begin_g(&sr, &ctx, c1, c2, &x, G);
		sxs_scale_node_g(&sr, &ctx, 0, &x, G, 1);
	}
#pragma unroll 1
	for (int i = 1; i < Q; i++) {
		SXS_FIT_STEP();
		sxs_scale_node_g(&sr, &ctx, i, &x, SXS_FIT_G(i), 1);
	}
	const double k = sr.up / sr.down;
	/* ---- pass 2: f and gradient with k frozen ---- */
	SXS_FIT_STEP();
	{
		const double G = SXS_FIT_G(0);
		sxs_grad_begin_g(&gr, &ctx, c1, c2, k, &x, G);
		sxs_grad_node_g(&gr, &ctx, 0, &x, G);
	}
#pragma unroll 1
	for (int i = 1; i < Q; i++) {
		SXS_FIT_STEP();
		sxs_grad_node_g(&gr, &ctx, i, &x, SXS_FIT_G(i));
	}
#undef SXS_FIT_STEP
#undef SXS_FIT_G
	fit_cp_wait<0>();
	out3[0] = gr.score;
	out3[1] = gr.grad0;
	out3[2] = gr.grad1;
}

/* The evaluation with the full exp (range test, libm beyond it) and IEEE quotients, plain loads: taken when
 * |corr q^2| could reach 512 somewhere in the box of c1 (never with a physical q grid; fs_fast) */
__device__ __noinline__ void fit_eval_safe(const struct fit_eval_args *ap, double c1, double c2, double *out3)
{
	struct sxs_fit_ctx ctx;
	ctx.x = ap->row; ctx.stride = 1; ctx.qstride = 6;
	ctx.a = fs_a; ctx.qvals = fs_q; ctx.qnum = ap->qnum; ctx.mult = ap->mult; ctx.scale = ap->scale;
	ctx.rq = fs_rq; ctx.dq = fs_dq; ctx.etab = fs_etab;
	sxs_fit_eval(&ctx, c1, c2, &out3[0], &out3[1], &out3[2]);
}

__global__ void __launch_bounds__(SXS_FIT_THREADS, SXS_FIT_MINBLOCKS)
k_fit(const double *__restrict__ X, long long npts, const double *__restrict__ a, const double *__restrict__ qvals,
      int qnum, double mult, double peak, int rescale, int allow_fast, double *__restrict__ res,
      unsigned long long *__restrict__ ticket)
{
	fit_tables_fill(a, qvals, qnum, mult, allow_fast);
	__syncthreads();
	const int fast = fs_fast;

	struct lq_state st;
	struct sxs_fit_ctx ctx;
	ctx.stride = 1; ctx.qstride = 6; ctx.a = fs_a;
	ctx.qvals = fs_q; ctx.qnum = qnum; ctx.mult = mult;
	ctx.x = X; ctx.scale = 1.0; ctx.rq = fs_rq; ctx.dq = fs_dq; ctx.etab = fs_etab;
	struct fit_eval_args ea;
	ea.qnum = qnum; ea.mult = mult; ea.scale = 1.0; ea.row = X;
	long long p = -1;
	bool drained = false;
	/* what this lane's fit waits for */
	enum { FREE = 0, EVALUATED, WANT_EVAL, WANT_B };
	int mode = FREE;

	for (;;) {
		/* (1) line-search turn (part A of the optimiser, short) for every lane that has a fresh f, g */
		if (mode == EVALUATED) {
			const int r = lq_step_a(&st);
			mode = (r == LQ_NEED_EVAL) ? WANT_EVAL : (r == LQ_NEED_B) ? WANT_B : FREE;
			if (r == LQ_DONE) {
				fit_store(&st, res, p);
			}
		}
		/* (2) free lanes take the next point from the ticket counter */
		if (mode == FREE && !drained) {
			p = (long long)atomicAdd(ticket, 1ull);
			if (p < npts) {
				ctx.x = X + sxs_x_index(p, qnum, 0, 0);
				ctx.scale = 1.0;
				if (rescale) {
					ctx.scale = sxs_fit_rescale(&ctx, peak);
				}
				lq_begin(&st, SXS_C1_DEFAULT, SXS_C2_DEFAULT);
				const int r = lq_step_a(&st); /* projects the start into the box and asks for f, g there */
				mode = (r == LQ_NEED_EVAL) ? WANT_EVAL : (r == LQ_NEED_B) ? WANT_B : FREE;
				if (r == LQ_DONE) {
					fit_store(&st, res, p);
				}
			} else {
				drained = true;
			}
		}
		/* (3) iteration boundary, when enough of the warp's waiting fits wait for it */
		const int n_b = __popc(__ballot_sync(0xffffffffu, mode == WANT_B));
		const int n_e = __popc(__ballot_sync(0xffffffffu, mode == WANT_EVAL));
		if (n_b > 0 && (n_e == 0 || n_b * SXS_FIT_BATCH_DEN >= (n_b + n_e) * SXS_FIT_BATCH_NUM)) {
			if (mode == WANT_B) {
				const int r = lq_step_b(&st, SXS_FIT_PGTOL, SXS_FIT_TOL);
				mode = (r == LQ_NEED_EVAL) ? WANT_EVAL : FREE;
				if (r == LQ_DONE) {
					fit_store(&st, res, p);
				}
			}
		}
		/* (4) the objective */
		if (!__any_sync(0xffffffffu, mode != FREE)) {
			if (__all_sync(0xffffffffu, drained)) {
				break;
			}
			continue;
		}
		if (mode == WANT_EVAL) {
			double fg[3];
			ea.row = ctx.x; ea.scale = ctx.scale;
			if (fast) {
				fit_eval_fast(&ea, st.x[1], st.x[2], fg);
			} else {
				fit_eval_safe(&ea, st.x[1], st.x[2], fg);
			}
			st.f = fg[0]; st.g[1] = fg[1]; st.g[2] = fg[2];
			mode = EVALUATED;
		}
		__syncwarp();
	}
}

int sxs_launch_fit(const double *d_x, long long npts, const double *d_a, const double *d_qvals, int qnum, double mult,
                   double peak, int rescale, double *d_res, unsigned long long *d_ticket, cudaStream_t stream)
{
	if (npts <= 0) {
		return 0;
	}
	if (qnum > SXS_FIT_MAXQ) {
		sxs_cuda_set_error("qnum %d exceeds %d", qnum, SXS_FIT_MAXQ);
		return -1;
	}
	int dev = 0, sms = 148, per_sm = 4;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	/* dynamic shared memory: the lanes' row rings.  Kept small on purpose: the kernel's spills live in L1, which is what
	 * shared memory leaves of 256 KB (with 219 KB of shared memory — G(c1, q) kept for pass 2 plus the optimiser's
	 * matrices parked across the objective — the spill loads went to L2: 42 % of all stall samples, r2p ncu) */
	const size_t shm = sizeof(double) * ((size_t)SXS_FIT_RING * SXS_FIT_THREADS * 6);
	SXS_CK(cudaFuncSetAttribute(k_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fit, SXS_FIT_THREADS, shm);
	if (per_sm < 1) per_sm = 1;
	if (getenv("SXS_FIT_BLOCKS_PER_SM")) { /* tuning only */
		const int v = atoi(getenv("SXS_FIT_BLOCKS_PER_SM"));
		if (v >= 1 && v < per_sm) per_sm = v;
	}
	long long blocks = (long long)sms * per_sm;
	const long long need = (npts + SXS_FIT_THREADS - 1) / SXS_FIT_THREADS;
	if (blocks > need) blocks = need;
	SXS_CK(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned long long), stream));
	k_fit<<<(unsigned)blocks, SXS_FIT_THREADS, shm, stream>>>(d_x, npts, d_a, d_qvals, qnum, mult, peak, rescale,
	                                                           getenv("SXS_FIT_FORCE_SAFE") == NULL /* tests: the fallback objective */, d_res, d_ticket);
	SXS_CK_LAUNCH();
	return 0;
}

/* cross[(p*6 + k)*qnum + q]  ->  X[p*6*qnum + q*6 + k] */
__global__ void k_cross_to_rows(const double *__restrict__ cross, long long npts, int qnum, double *__restrict__ x)
{
	const long long total = npts * 6 * qnum;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const long long p = i / (6 * qnum);
		const int r = (int)(i % (6 * qnum));
		const int q = r / 6, k = r % 6;
		x[sxs_x_index(p, qnum, q, k)] = cross[(p * 6 + k) * qnum + q];
	}
}

extern "C" int sxs_cuda_fit_profiles(int device, const double *cross, long long npts, const double *a,
                                     const double *qvals, int qnum, double mult, double peak, int rescale, double *out)
{
	if (npts <= 0) {
		return 0;
	}
	SXS_CK(cudaSetDevice(device));
	double *d_cross = NULL, *d_x = NULL, *d_a = NULL, *d_q = NULL, *d_res = NULL;
	unsigned long long *d_ticket = NULL;
	const size_t nx = (size_t)npts * 6 * qnum;
	SXS_CK(cudaMalloc(&d_cross, sizeof(double) * nx));
	SXS_CK(cudaMalloc(&d_x, sizeof(double) * nx));
	SXS_CK(cudaMalloc(&d_a, sizeof(double) * 6 * qnum));
	SXS_CK(cudaMalloc(&d_q, sizeof(double) * qnum));
	SXS_CK(cudaMalloc(&d_res, sizeof(double) * 4 * npts));
	SXS_CK(cudaMalloc(&d_ticket, sizeof(unsigned long long)));
	SXS_CK(cudaMemcpy(d_cross, cross, sizeof(double) * nx, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_a, a, sizeof(double) * 6 * qnum, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_q, qvals, sizeof(double) * qnum, cudaMemcpyHostToDevice));
	k_cross_to_rows<<<1184, 256>>>(d_cross, npts, qnum, d_x);
	SXS_CK_LAUNCH();
	int rc = sxs_launch_fit(d_x, npts, d_a, d_q, qnum, mult, peak, rescale, d_res, d_ticket, 0);
	if (rc == 0) {
		SXS_CK(cudaMemcpy(out, d_res, sizeof(double) * 4 * npts, cudaMemcpyDeviceToHost));
	}
	cudaFree(d_cross); cudaFree(d_x); cudaFree(d_a); cudaFree(d_q); cudaFree(d_res); cudaFree(d_ticket);
	return rc;
}

__global__ void k_fit_eval(const double *__restrict__ cross, const double *__restrict__ a,
                           const double *__restrict__ qvals, int qnum, double mult, double c1, double c2,
                           double *__restrict__ out4)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) {
		return;
	}
	struct sxs_fit_ctx ctx;
	ctx.x = cross; /* point-major row x[q*6 + k] */
	ctx.stride = 1;
	ctx.qstride = 6;
	double rq[SXS_FIT_MAXQ], dq[SXS_FIT_MAXQ];
	for (int i = 0; i < qnum; i++) {
		dq[i] = qvals[i] - (i > 0 ? qvals[i - 1] : -1.0);
		rq[i] = 1.0 / dq[i];
	}
	ctx.a = a; ctx.qvals = qvals; ctx.qnum = qnum; ctx.mult = mult; ctx.scale = 1.0; ctx.rq = rq; ctx.dq = dq; ctx.etab = d_exp_tab;
	out4[0] = sxs_fit_best_scale(&ctx, c1, c2);
	sxs_fit_eval(&ctx, c1, c2, &out4[1], &out4[2], &out4[3]);
}

extern "C" int sxs_cuda_fit_eval(int device, const double *cross, const double *a, const double *qvals, int qnum,
                                 double mult, double c1, double c2, double *out4)
{
	SXS_CK(cudaSetDevice(device));
	/* host layout cross[k*qnum + q] -> device layout [(q*6 + k)] */
	double *tmp = (double *)malloc(sizeof(double) * 6 * qnum);
	for (int k = 0; k < 6; k++) {
		for (int q = 0; q < qnum; q++) {
			tmp[q * 6 + k] = cross[k * qnum + q];
		}
	}
	double *d_x = NULL, *d_a = NULL, *d_q = NULL, *d_o = NULL;
	SXS_CK(cudaMalloc(&d_x, sizeof(double) * 6 * qnum));
	SXS_CK(cudaMalloc(&d_a, sizeof(double) * 6 * qnum));
	SXS_CK(cudaMalloc(&d_q, sizeof(double) * qnum));
	SXS_CK(cudaMalloc(&d_o, sizeof(double) * 4));
	SXS_CK(cudaMemcpy(d_x, tmp, sizeof(double) * 6 * qnum, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_a, a, sizeof(double) * 6 * qnum, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_q, qvals, sizeof(double) * qnum, cudaMemcpyHostToDevice));
	free(tmp);
	k_fit_eval<<<1, 32>>>(d_x, d_a, d_q, qnum, mult, c1, c2, d_o);
	SXS_CK_LAUNCH();
	SXS_CK(cudaMemcpy(out4, d_o, sizeof(double) * 4, cudaMemcpyDeviceToHost));
	cudaFree(d_x); cudaFree(d_a); cudaFree(d_q); cudaFree(d_o);
	return 0;
}

__global__ void k_exp_array(const double *__restrict__ x, long long n, double *__restrict__ y)
{
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		y[i] = sxs_exp_glibc(x[i], d_exp_tab);
	}
}

extern "C" int sxs_cuda_exp_array(int device, const double *x, long long n, double *y)
{
	if (n <= 0) {
		return 0;
	}
	SXS_CK(cudaSetDevice(device));
	double *d_x = NULL, *d_y = NULL;
	SXS_CK(cudaMalloc(&d_x, sizeof(double) * n));
	SXS_CK(cudaMalloc(&d_y, sizeof(double) * n));
	SXS_CK(cudaMemcpy(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice));
	k_exp_array<<<592, 256>>>(d_x, n, d_y);
	SXS_CK_LAUNCH();
	SXS_CK(cudaMemcpy(y, d_y, sizeof(double) * n, cudaMemcpyDeviceToHost));
	cudaFree(d_x); cudaFree(d_y);
	return 0;
}

/* ------------------------------------------------------------------ ft rows -> grid indices
 * One thread per ft row: (rotation id, translation) -> (z, b1, g1, a2, b2, g2) as sxs_ft2euler does it
 * (src/index.c:38-75), every value through the three decimals of the Euler text file (src/index.c:114 writes "% .3f",
 * tools/correlate.c:214 reads it back), the z table lookup and the snapping of tools/correlate.c:219-247 /
 * sxs_euler_to_index64.  In this translation unit because the products of the rotation composition must not be fused
 * (-fmad=false, like the host's -ffp-contract=off).
 *
 * The text round trip is arithmetic here: the decimal the C library prints is the integer nearest to x * 1000 taken
 * exactly (ties to even), and what strtod makes of it is the IEEE quotient n / 1000.  x * 1000 = p + e exactly with
 * p = RN(x * 1000), e = fma(x, 1000, -p); only when p lies exactly on k + 1/2 does e decide.
 * acos / sin / cos are CUDA's (<= 1-2 ulp from the host libm's): an angle can differ from the host's before the
 * rounding only in its last bits, i.e. after it only within ~1e-13 of a rounding boundary of the third decimal. */
__device__ __forceinline__ double ft_through_text(double x)
{
	const double p = x * 1000.0;
	const double e = fma(x, 1000.0, -p);
	double n = rint(p);
	const double diff = p - n;
	if (diff == 0.5 && e > 0.0) {
		n += 1.0;
	} else if (diff == -0.5 && e < 0.0) {
		n -= 1.0;
	}
	return n / 1000.0;
}

__device__ __forceinline__ double ft_clamp_unit(double a) { return a < -1.0 ? -1.0 : (a > 1.0 ? 1.0 : a); }

#define SXS_FT_PI 3.14159265358979323846

__global__ void k_ft_rows_to_index(const int *__restrict__ rot_id, const double *__restrict__ trans, long long n,
                                   const double *__restrict__ rots, long long nrot, double rx, double ry, double rz,
                                   const double *__restrict__ zvals, int znum, int L, long long *__restrict__ out)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const int id = rot_id[i];
	if (id < 0 || id >= nrot) {
		out[i] = -2; /* rotation index outside the rotation table */
		return;
	}
	const double vx = trans[3 * i] + rx, vy = trans[3 * i + 1] + ry, vz = trans[3 * i + 2] + rz;
	double z = round(sqrt(vx * vx + vy * vy + vz * vz));
	double b1 = acos(ft_clamp_unit(vz / z));
	double g1 = acos(ft_clamp_unit(-vx / (z * sin(b1))));
	if (vy / (z * sin(b1)) < 0.0) {
		g1 = 2 * SXS_FT_PI - g1;
	}
	/* active z-y-z rotation (0, b1, g1), src/saxs_utils.c:65-79 with alpha = 0 spelled out as the host evaluates it */
	const double ca = cos(0.0), sa = sin(0.0);
	const double cb = cos(b1), sb = sin(b1);
	const double cg = cos(g1), sg = sin(g1);
	double a[9];
	a[0] = cg * cb * ca - sg * sa;  a[1] = -sg * cb * ca - cg * sa; a[2] = sb * ca;
	a[3] = cg * cb * sa + sg * ca;  a[4] = -sg * cb * sa + cg * ca; a[5] = sb * sa;
	a[6] = -cg * sb;                a[7] = sg * sb;                 a[8] = cb;
	const double *b = rots + 9 * (long long)id;
	/* the entries of (rec_rm x rm) the angles need: m13, m23, m31, m32, m33 (src/saxs_utils.c:81-94) */
	const double m13 = a[0] * b[2] + a[1] * b[5] + a[2] * b[8];
	const double m23 = a[3] * b[2] + a[4] * b[5] + a[5] * b[8];
	const double m31 = a[6] * b[0] + a[7] * b[3] + a[8] * b[6];
	const double m32 = a[6] * b[1] + a[7] * b[4] + a[8] * b[7];
	const double m33 = a[6] * b[2] + a[7] * b[5] + a[8] * b[8];
	double b2 = acos(ft_clamp_unit(m33));
	double a2 = acos(ft_clamp_unit(m13 / sin(b2)));
	if (m23 / sin(b2) < 0.0) {
		a2 = 2 * SXS_FT_PI - a2;
	}
	double g2 = acos(ft_clamp_unit(-m31 / sin(b2)));
	if (m32 / sin(b2) < 0.0) {
		g2 = 2 * SXS_FT_PI - g2;
	}
	z = ft_through_text(z); b1 = ft_through_text(b1); g1 = ft_through_text(g1);
	a2 = ft_through_text(a2); b2 = ft_through_text(b2); g2 = ft_through_text(g2);
	/* a translation on the z axis or a ligand z axis on the receptor's gives 0 / 0 above: "nan" in the Euler file, and
	 * (int)round(nan) = INT_MIN in the tool's index (tools/correlate.c:225-240), a negative index that the host route
	 * drops (index_rows.c keeps rows with flat >= 0); the same rows are dropped here */
	if (b1 != b1 || g1 != g1 || a2 != a2 || b2 != b2 || g2 != g2 || z != z) {
		out[i] = -1;
		return;
	}
	long long flat = -1;
	for (int k = 0; k < znum; k++) {
		if (zvals[k] > z - 0.001 && zvals[k] < z + 0.001) {
			const long long nbeta = L + 1, nn = 2 * L + 1;
			const double b_step = SXS_FT_PI / L;
			const double a_step = 2.0 * SXS_FT_PI / nn;
			const double a2r = 2 * SXS_FT_PI - a2;
			const double g2r = 2 * SXS_FT_PI - g2;
			long long f = k * nbeta;
			f = (f + (int)(round(b1 / b_step))) * nbeta;
			f = (f + (int)(round(b2 / b_step))) * nn;
			f = (f + (int)(round(a2r / a_step))) * nn;
			f = (f + (int)(round(g1 / a_step))) * nn;
			f = f + (int)(round(g2r / a_step));
			flat = f;
			break; /* the 1 A table of the tool matches at most one z */
		}
	}
	out[i] = flat;
}

extern "C" int sxs_cuda_ft_rows_to_indices_dev(const int *d_rot_id, const double *d_trans, long long n, const double *d_rots,
                                               long long nrot, const double *ref_lig, const double *d_zvals, int znum, int L,
                                               long long *d_index, void *stream)
{
	if (n <= 0) {
		return 0;
	}
	k_ft_rows_to_index<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_rot_id, d_trans, n, d_rots, nrot, ref_lig[0],
	                                                                                  ref_lig[1], ref_lig[2], d_zvals, znum, L, d_index);
	SXS_CK_LAUNCH();
	return 0;
}

extern "C" int sxs_cuda_ft_rows_to_indices(int device, const int *rot_id, const double *trans, long long n, const double *rots,
                                           long long nrot, const double *ref_lig, const double *zvals, int znum, int L,
                                           long long *index)
{
	if (n <= 0) {
		return 0;
	}
	SXS_CK(cudaSetDevice(device));
	int *d_id = NULL;
	double *d_t = NULL, *d_r = NULL, *d_z = NULL;
	long long *d_o = NULL;
	SXS_CK(cudaMalloc(&d_id, sizeof(int) * n));
	SXS_CK(cudaMalloc(&d_t, sizeof(double) * 3 * n));
	SXS_CK(cudaMalloc(&d_r, sizeof(double) * 9 * nrot));
	SXS_CK(cudaMalloc(&d_z, sizeof(double) * znum));
	SXS_CK(cudaMalloc(&d_o, sizeof(long long) * n));
	SXS_CK(cudaMemcpy(d_id, rot_id, sizeof(int) * n, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_t, trans, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_r, rots, sizeof(double) * 9 * nrot, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_z, zvals, sizeof(double) * znum, cudaMemcpyHostToDevice));
	int rc = sxs_cuda_ft_rows_to_indices_dev(d_id, d_t, n, d_r, nrot, ref_lig, d_z, znum, L, d_o, NULL);
	if (rc == 0) {
		SXS_CK(cudaMemcpy(index, d_o, sizeof(long long) * n, cudaMemcpyDeviceToHost));
	}
	cudaFree(d_id); cudaFree(d_t); cudaFree(d_r); cudaFree(d_z); cudaFree(d_o);
	return rc;
}

/* ------------------------------------------------------------------ self terms */

/* one thread per (k, q); serial over (l, m) in the reference's order */
__global__ void k_pair_const(const double2 *__restrict__ A, const double2 *__restrict__ B, int qnum, int lm_n,
                             double *__restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 6 * qnum) {
		return;
	}
	const int k = i / qnum, q = i % qnum;
	const int c1s[6] = {0, 0, 0, 1, 1, 2};
	const int c2s[6] = {0, 1, 2, 1, 2, 2};
	const double2 *A1 = A + ((size_t)c1s[k] * qnum + q) * lm_n, *A2 = A + ((size_t)c2s[k] * qnum + q) * lm_n;
	const double2 *B1 = B + ((size_t)c1s[k] * qnum + q) * lm_n, *B2 = B + ((size_t)c2s[k] * qnum + q) * lm_n;
	double in = 0.0;
	for (int j = 0; j < lm_n; j++) {
		in += A1[j].x * A2[j].x + A1[j].y * A2[j].y;
		in += B1[j].x * B2[j].x + B1[j].y * B2[j].y;
	}
	out[i] = (c1s[k] != c2s[k]) ? in * 2.0 : in;
}

int sxs_launch_pair_const(const double *d_coefA, const double *d_coefB, int qnum, int L, double *d_out,
                          cudaStream_t stream)
{
	const int n = 6 * qnum;
	k_pair_const<<<(n + 127) / 128, 128, 0, stream>>>((const double2 *)d_coefA, (const double2 *)d_coefB, qnum,
	                                                   (L + 1) * (L + 1), d_out);
	SXS_CK_LAUNCH();
	return 0;
}

__global__ void k_self_terms(const double2 *__restrict__ A, int qnum, int lm_n, double *__restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 6 * qnum) {
		return;
	}
	const int k = i / qnum, q = i % qnum;
	const int c1s[6] = {0, 0, 0, 1, 1, 2};
	const int c2s[6] = {0, 1, 2, 1, 2, 2};
	const double2 *A1 = A + ((size_t)c1s[k] * qnum + q) * lm_n, *A2 = A + ((size_t)c2s[k] * qnum + q) * lm_n;
	double acc = 0.0;
	for (int j = 0; j < lm_n; j++) {
		const double val = A1[j].x * A2[j].x + A1[j].y * A2[j].y;
		acc += val;
	}
	out[i] = (c1s[k] != c2s[k]) ? acc * 2.0 : acc;
}

extern "C" int sxs_cuda_self_terms(int device, const double *coef, int qnum, int L, double *out)
{
	SXS_CK(cudaSetDevice(device));
	const int lm_n = (L + 1) * (L + 1);
	const size_t nc = (size_t)3 * qnum * lm_n * 2;
	double *d_c = NULL, *d_o = NULL;
	SXS_CK(cudaMalloc(&d_c, sizeof(double) * nc));
	SXS_CK(cudaMalloc(&d_o, sizeof(double) * 6 * qnum));
	SXS_CK(cudaMemcpy(d_c, coef, sizeof(double) * nc, cudaMemcpyHostToDevice));
	k_self_terms<<<(6 * qnum + 127) / 128, 128>>>((const double2 *)d_c, qnum, lm_n, d_o);
	SXS_CK_LAUNCH();
	SXS_CK(cudaMemcpy(out, d_o, sizeof(double) * 6 * qnum, cudaMemcpyDeviceToHost));
	cudaFree(d_c); cudaFree(d_o);
	return 0;
}

/* I(q) = sum_lm |V - G D + c2 W|^2, G evaluated at qvals[0] for every q like the reference */
__global__ void k_profile(const double2 *__restrict__ A, int qnum, int lm_n, double mult, double q0, double c1,
                          double c2, double *__restrict__ in)
{
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= qnum) {
		return;
	}
	const double corr = -mult * (c1 * c1 - 1.0);
	const double G = c1 * c1 * c1 * sxs_exp_glibc(corr * q0 * q0, d_exp_tab);
	const double2 *V = A + ((size_t)0 * qnum + q) * lm_n, *D = A + ((size_t)1 * qnum + q) * lm_n,
	              *W = A + ((size_t)2 * qnum + q) * lm_n;
	double acc = 0.0;
	for (int j = 0; j < lm_n; j++) {
		const double a_re = V[j].x - G * D[j].x + c2 * W[j].x;
		const double a_im = V[j].y - G * D[j].y + c2 * W[j].y;
		acc += a_re * a_re + a_im * a_im;
	}
	in[q] = acc;
}

extern "C" int sxs_cuda_profile_from_spf(int device, const double *coef, int qnum, int L, double mult,
                                         const double *qvals, double c1, double c2, double *intensity)
{
	SXS_CK(cudaSetDevice(device));
	const int lm_n = (L + 1) * (L + 1);
	const size_t nc = (size_t)3 * qnum * lm_n * 2;
	double *d_c = NULL, *d_o = NULL;
	SXS_CK(cudaMalloc(&d_c, sizeof(double) * nc));
	SXS_CK(cudaMalloc(&d_o, sizeof(double) * qnum));
	SXS_CK(cudaMemcpy(d_c, coef, sizeof(double) * nc, cudaMemcpyHostToDevice));
	k_profile<<<(qnum + 63) / 64, 64>>>((const double2 *)d_c, qnum, lm_n, mult, qvals[0], c1, c2, d_o);
	SXS_CK_LAUNCH();
	SXS_CK(cudaMemcpy(intensity, d_o, sizeof(double) * qnum, cudaMemcpyDeviceToHost));
	cudaFree(d_c); cudaFree(d_o);
	return 0;
}
