/* K2-K4 orchestration: FFT-SAXS dimer scoring on one B200.
 *
 * Reference computation (src/fftsaxs.c:608-986), per z and per (beta1, beta2) cell:
 *   sum1  S[q][m][m2][l]   = sum_l1 d^l1_{m m2}(b2) conj(B_{l1 m2} T^{|m|}_{l l1}(q z))          (:251-333)
 *   sum2  G[m][m1][m2]     = sum_l  d^l_{m m1}(b1) A_{l m1} S[m][m2][l]   (6 component pairs)     (:416-524)
 *   F[a2][g1][g2]          = Re 3-D DFT of G over (m, m1, m2)                                    (:529-606, fftw)
 *   X_k[q]                 = const_k[q] + 2 F_k                                                 (:52-108)
 *   fit (c1, c2) per listed grid point                                                          (:909)
 *
 * The DFT is separable and linear, so the two gamma transforms are applied to the coefficients
 * BEFORE the l-contraction, once per molecule and independent of z:
 *   At[b1][q][c][m,l][g1] = sum_m1 w^(m1 g1) d^l_{m m1}(b1)      A^c_{l m1}(q)        (k_rotate, receptor)
 *   Bt[b2][q][c][m,l][g2] = sum_m2 w^(m2 g2) d^l_{m m2}(b2) conj(B^c_{l m2}(q))       (k_rotate, ligand)
 *   St[z,b2][q][c][m,l][g2] = sum_l1 conj(T^m_{l l1}(q z)) Bt[b2][q][c][m,l1][g2]      (k_translate_rt)
 * and what is left per listed pose (z, b1, b2, a2, g1, g2) is the alpha transform of a length-(L+1) dot:
 *   F_k = Re sum_m w^(m a2) sum_{l>=|m|} At^{c}[m,l,g1] St^{c'}[m,l,g2]                  (k_cross)
 * Only m >= 0 is evaluated: the m <-> -m terms are complex conjugates of each other (A_{l,-m} =
 * (-1)^{l+m} conj(A_{lm}) for real densities; verified to 6e-13 on the reference's golden inputs), so
 * F = Re C_0 + 2 Re sum_{m>0}.  No N^3 grid is ever materialised; cost per pose is
 * qnum * 9 * (L+1)(L+2)/2 complex MACs (0.49 MFLOP at L=15, Q=50) instead of N^3-sized transforms
 * per cell.  w = exp(-2 pi i/N) comes from the host table built like the reference's direct DFT
 * (truncated pi, src/fftsaxs.c:536-548).
 *
 * Pose list handling (replaces the O(cells*nout) mask scans, src/fftsaxs.c:716-733,850-872,922-936):
 * keys (z, b2, b1, g2 / 8, g1, g2 % 8, a2) are radix-sorted on the device, duplicates collapse into distinct grid
 * points, each point is fitted once and scattered back to every row that named it.
 */
#include <cub/cub.cuh>
#include <math.h>
#include <stdarg.h>

#include "sxs_dev.cuh"

/* ------------------------------------------------------------------ errors */

/* one buffer per host thread: the device threads of a multi-GPU call each keep their own message, and the host
 * layer copies it into the shard job before the thread ends (csrc/host/fftsaxs.c) */
static thread_local char g_err[512] = "no error";

extern "C" void sxs_cuda_set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *sxs_cuda_last_error(void)
{
	return g_err;
}

extern "C" int sxs_cuda_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		return 0;
	}
	return n;
}

/* -------------------------------------------------------------------- plan */

#define SXS_NTIMERS 5   /* 0 sort+points, 1 tmatrix+translate, 2 cross, 3 fit, 4 scatter */
#define SXS_MAX_TIMED 256

struct sxs_cuda_plan {
	int device;
	int L, nb, N, ML, qnum;
	/* tables */
	double *d_qvals, *d_dsymb, *d_dwig;
	double2 *d_tw;
	/* molecules */
	double2 *d_coefA, *d_coefB;
	double *d_const;  /* [6][qnum] */
	double2 *d_At, *d_Bt; /* [beta][q][c][ml][g], rows padded to sxs_row_pad(N) */
	int have_molecules;
	/* experiment */
	double *d_a;
	double mult, peak;
	int have_experiment;
	/* translations */
	double *d_bessel;
	int znum;
	/* workspace (grow-only) */
	double2 *d_T;  size_t cap_T;   /* [zg][q][m][l][l1] */
	double2 *d_St; size_t cap_St;  /* [zg*nb slabs][q][c][ml][g], rows padded like At */
	double *d_X;   size_t cap_X;   /* cross terms handed from K3 to K4, one row per point (sxs_x_index) */
	unsigned long long *d_ticket;
	double *d_res; size_t cap_res; /* [points][4] */
	unsigned long long *d_keys, *d_keys_sorted, *d_pkeys;
	unsigned int *d_rows, *d_rows_sorted, *d_pid;
	size_t cap_rows;
	void *d_cub; size_t cap_cub;
	long long *d_zoff; int cap_zoff;
	int *d_slab_flag; int cap_slab;
	void *d_index_in; size_t cap_index_in;
	double *d_out3; size_t cap_out3;
	long long *h_zoff;
	void *h_pinned; size_t cap_pinned;
	void *h_shard; size_t cap_shard; /* pinned staging of a z shard's compact list and results (sxs_cuda_plan_host_buffer) */
	/* budgets */
	size_t budget_St, budget_X;
	long long stats[5];
	/* optional per-kernel-class device timing (CUDA events on the launching stream) */
	int profiling, events_ready;
	cudaEvent_t ev[SXS_NTIMERS][2 * SXS_MAX_TIMED];
	int ev_used[SXS_NTIMERS];
	double ms[SXS_NTIMERS];
	long long timed_launches[SXS_NTIMERS];
};

static void timer_begin(sxs_cuda_plan *p, int which, cudaStream_t st)
{
	if (p->profiling && p->ev_used[which] < SXS_MAX_TIMED) {
		cudaEventRecord(p->ev[which][2 * p->ev_used[which]], st);
	}
}

static void timer_end(sxs_cuda_plan *p, int which, cudaStream_t st)
{
	if (p->profiling && p->ev_used[which] < SXS_MAX_TIMED) {
		cudaEventRecord(p->ev[which][2 * p->ev_used[which] + 1], st);
		p->ev_used[which]++;
	}
}

/* Sums the recorded intervals (synchronises the last event of each class). */
static void timer_collect(sxs_cuda_plan *p)
{
	for (int w = 0; w < SXS_NTIMERS; w++) {
		p->ms[w] = 0.0;
		p->timed_launches[w] = p->ev_used[w];
		for (int i = 0; i < p->ev_used[w]; i++) {
			float t = 0.f;
			cudaEventSynchronize(p->ev[w][2 * i + 1]);
			cudaEventElapsedTime(&t, p->ev[w][2 * i], p->ev[w][2 * i + 1]);
			p->ms[w] += t;
		}
		p->ev_used[w] = 0;
	}
}

template <typename T>
static int ensure(T **ptr, size_t *cap, size_t need)
{
	if (need <= *cap && *ptr != NULL) {
		return 0;
	}
	if (*ptr != NULL) {
		cudaFree(*ptr);
		*ptr = NULL;
	}
	size_t want = need + need / 8;
	cudaError_t e = cudaMalloc((void **)ptr, want * sizeof(T));
	if (e != cudaSuccess) {
		want = need;
		e = cudaMalloc((void **)ptr, want * sizeof(T));
	}
	if (e != cudaSuccess) {
		*cap = 0;
		sxs_cuda_set_error("cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e));
		return -1;
	}
	*cap = want;
	return 0;
}

static size_t env_gb(const char *name, double dflt_gb)
{
	const char *v = getenv(name);
	double gb = dflt_gb;
	if (v != NULL && *v) {
		gb = atof(v);
	}
	return (size_t)(gb * 1024.0 * 1024.0 * 1024.0);
}

extern "C" sxs_cuda_plan *sxs_cuda_plan_create(int device, int L, int qnum, const double *qvals, const double *dsymb,
                                               const double *dwig, const double *twiddle)
{
	if (cudaSetDevice(device) != cudaSuccess) {
		sxs_cuda_set_error("cudaSetDevice(%d) failed", device);
		return NULL;
	}
	if (L < 1 || L > 127 || qnum < 1 || qnum > 512) {
		sxs_cuda_set_error("unsupported L=%d or qnum=%d", L, qnum);
		return NULL;
	}
	sxs_cuda_plan *p = (sxs_cuda_plan *)calloc(1, sizeof(*p));
	p->device = device;
	p->L = L; p->nb = L + 1; p->N = 2 * L + 1; p->ML = sxs_ml_count(L); p->qnum = qnum;
	const size_t nds = (size_t)p->nb * p->nb * p->nb * p->N;
	const size_t ndw = (size_t)p->nb * p->nb * p->N * p->N;
	const size_t ncoef = (size_t)3 * qnum * p->nb * p->nb;
	const size_t nrot = (size_t)p->nb * qnum * 3 * p->ML * sxs_row_pad(p->N);
	cudaError_t e = cudaSuccess;
#define PC(call) do { if (e == cudaSuccess) e = (call); } while (0)
	PC(cudaMalloc(&p->d_qvals, sizeof(double) * qnum));
	PC(cudaMalloc(&p->d_dsymb, sizeof(double) * nds));
	PC(cudaMalloc(&p->d_dwig, sizeof(double) * ndw));
	PC(cudaMalloc(&p->d_tw, sizeof(double2) * p->N));
	PC(cudaMalloc(&p->d_coefA, sizeof(double2) * ncoef));
	PC(cudaMalloc(&p->d_coefB, sizeof(double2) * ncoef));
	PC(cudaMalloc(&p->d_const, sizeof(double) * 6 * qnum));
	PC(cudaMalloc(&p->d_At, sizeof(double2) * nrot));
	PC(cudaMalloc(&p->d_Bt, sizeof(double2) * nrot));
	PC(cudaMalloc(&p->d_a, sizeof(double) * 6 * qnum));
	PC(cudaMalloc(&p->d_ticket, sizeof(unsigned long long)));
	PC(cudaMemcpy(p->d_qvals, qvals, sizeof(double) * qnum, cudaMemcpyHostToDevice));
	PC(cudaMemcpy(p->d_dsymb, dsymb, sizeof(double) * nds, cudaMemcpyHostToDevice));
	PC(cudaMemcpy(p->d_dwig, dwig, sizeof(double) * ndw, cudaMemcpyHostToDevice));
	PC(cudaMemcpy(p->d_tw, twiddle, sizeof(double2) * p->N, cudaMemcpyHostToDevice));
	PC(cudaMallocHost(&p->h_zoff, sizeof(long long) * 4096));
#undef PC
	if (e != cudaSuccess) {
		sxs_cuda_set_error("plan allocation failed: %s", cudaGetErrorString(e));
		sxs_cuda_plan_destroy(p);
		return NULL;
	}
	p->budget_St = env_gb("SXS_CUDA_ST_GB", 12.0);
	p->budget_X = env_gb("SXS_CUDA_X_GB", 10.0);
	return p;
}

extern "C" void sxs_cuda_plan_destroy(sxs_cuda_plan *p)
{
	if (p == NULL) {
		return;
	}
	cudaSetDevice(p->device);
	void *ptrs[] = {p->d_qvals, p->d_dsymb, p->d_dwig, p->d_tw, p->d_coefA, p->d_coefB, p->d_const, p->d_At, p->d_Bt,
	                p->d_a, p->d_bessel, p->d_T, p->d_St, p->d_X, p->d_res, p->d_keys, p->d_keys_sorted, p->d_pkeys,
	                p->d_rows, p->d_rows_sorted, p->d_pid, p->d_cub, p->d_zoff, p->d_slab_flag, p->d_index_in,
	                p->d_out3, p->d_ticket};
	for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) {
		if (ptrs[i] != NULL) {
			cudaFree(ptrs[i]);
		}
	}
	if (p->h_zoff != NULL) {
		cudaFreeHost(p->h_zoff);
	}
	if (p->h_shard != NULL) {
		cudaFreeHost(p->h_shard);
	}
	if (p->h_pinned != NULL) {
		cudaFreeHost(p->h_pinned);
	}
	free(p);
}

extern "C" int sxs_cuda_plan_stats(const sxs_cuda_plan *p, long long *stats5)
{
	memcpy(stats5, p->stats, sizeof(long long) * 5);
	return 0;
}

extern "C" int sxs_cuda_plan_set_profiling(sxs_cuda_plan *p, int on)
{
	SXS_CK(cudaSetDevice(p->device));
	if (on && !p->events_ready) {
		for (int w = 0; w < SXS_NTIMERS; w++) {
			for (int i = 0; i < 2 * SXS_MAX_TIMED; i++) {
				SXS_CK(cudaEventCreate(&p->ev[w][i]));
			}
		}
		p->events_ready = 1; /* events stay allocated once created */
	}
	p->profiling = on ? 1 : 0;
	for (int w = 0; w < SXS_NTIMERS; w++) p->ev_used[w] = 0;
	return 0;
}

extern "C" int sxs_cuda_plan_kernel_times(sxs_cuda_plan *p, double *ms5, long long *launches5)
{
	SXS_CK(cudaSetDevice(p->device));
	timer_collect(p);
	for (int w = 0; w < SXS_NTIMERS; w++) {
		ms5[w] = p->ms[w];
		launches5[w] = p->timed_launches[w];
	}
	return 0;
}

/* ------------------------------------------------ K2a: rotation + gamma DFT */

/* out[b][q][c][ml][g] = sum_{m1=-l..l} w^(m1 g) d^l_{m m1}(beta_b) coef^c[q][l,m1]   (conj(coef) for the ligand).
 * One thread per output element, g fastest (rows of NP = sxs_row_pad(N) values, the padding is zero): the d and
 * coef operands are warp-uniform broadcasts. */
__global__ void __launch_bounds__(256)
k_rotate(int L, int qnum, const double *__restrict__ dwig, const double2 *__restrict__ coef,
         const double2 *__restrict__ tw, int conjugate, double2 *__restrict__ out)
{
	extern __shared__ double2 s_tw[];
	const int N = 2 * L + 1, NP = sxs_row_pad(N), nb = L + 1, ML = sxs_ml_count(L), lm_n = nb * nb;
	for (int i = threadIdx.x; i < N; i += blockDim.x) {
		s_tw[i] = tw[i];
	}
	__syncthreads();
	const size_t total = (size_t)nb * qnum * 3 * ML * NP;
	for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
		const int g = (int)(e % NP);
		size_t rest = e / NP;
		if (g >= N) {
			out[e] = make_double2(0.0, 0.0);
			continue;
		}
		const int ml = (int)(rest % ML); rest /= ML;
		const int c = (int)(rest % 3); rest /= 3;
		const int q = (int)(rest % qnum);
		const int b = (int)(rest / qnum);
		/* unpack ml -> (m, l): rows m hold l = m..L */
		int m = 0, off = ml;
		while (off >= nb - m) {
			off -= nb - m;
			m++;
		}
		const int l = m + off;
		const double *drow = dwig + (((size_t)b * nb + l) * N + (m + L)) * N + L; /* index by m1 */
		const double2 *crow = coef + ((size_t)c * qnum + q) * lm_n + l * (l + 1);   /* index by m1 */
		double re = 0.0, im = 0.0;
		int k = ((-l * g) % N + N) % N; /* (m1*g) mod N, advanced incrementally */
		for (int m1 = -l; m1 <= l; m1++) {
			const double d = drow[m1];
			double2 a = crow[m1];
			if (conjugate) {
				a.y = -a.y;
			}
			const double2 w = s_tw[k];
			const double pr = a.x * w.x - a.y * w.y;
			const double pi = a.x * w.y + a.y * w.x;
			re += d * pr;
			im += d * pi;
			k += g;
			if (k >= N) {
				k -= N;
			}
		}
		out[e] = make_double2(re, im);
	}
}

/* ------------------------------------------------------ K2b: translation */

/* T^m_{l l1}(q z) = (-1)^m sum_p i^p dsymb[l,m,l1,p] j_p(q z), p = |l-l1| .. l+l1 ascending
 * (fill_t_matrix, src/fftsaxs.c:182-243).  Layout T[((zl*qnum + q)*nb + m)*nb*nb + l*nb + l1]. */
__global__ void __launch_bounds__(256)
k_tmatrix(int L, int qnum, int nz, const int *__restrict__ zlist, const double *__restrict__ dsymb,
          const double *__restrict__ bessel, double2 *__restrict__ T)
{
	const int nb = L + 1, N = 2 * L + 1;
	const size_t per = (size_t)nb * nb * nb;
	const size_t total = (size_t)nz * qnum * per;
	for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
		const int l1 = (int)(e % nb);
		size_t rest = e / nb;
		const int l = (int)(rest % nb); rest /= nb;
		const int m = (int)(rest % nb); rest /= nb;
		const int q = (int)(rest % qnum);
		const int zl = (int)(rest / qnum);
		if (l < m || l1 < l) {
			continue; /* lower triangle is written by its mirror; l < m entries are never read */
		}
		const double sgn = (m % 2) ? -1.0 : 1.0;
		const double *drow = dsymb + ((size_t)(l * (l + 1) + m) * nb + l1) * N;
		const double *brow = bessel + ((size_t)zlist[zl] * qnum + q) * N;
		double re = 0.0, im = 0.0;
		for (int p = l1 - l; p <= l + l1; p++) {
			const double val = __dmul_rn(__dmul_rn(sgn, drow[p]), brow[p]);
			switch (p & 3) {
			case 0: re = __dadd_rn(re, val); break;
			case 1: im = __dadd_rn(im, val); break;
			case 2: re = __dadd_rn(re, -val); break;
			default: im = __dadd_rn(im, -val); break;
			}
		}
		const size_t base = ((size_t)zl * qnum + q) * per + (size_t)m * nb * nb;
		T[base + (size_t)l * nb + l1] = make_double2(re, im);
		T[base + (size_t)l1 * nb + l] = make_double2(re, im);
	}
}

/* K2c.  St[slab][q][c][ml(m,l)][g] = sum_{l1=m..L} conj(T^m_{l l1}) Bt[b2][q][c][ml(m,l1)][g], slab = zl*nb + b2;
 * slabs whose (z, b2) holds no listed pose are skipped.  A block owns one (b2, q, c) and the pair of orders
 * (m, L - m) — together L + 2 rows of N outputs, so that every block has the same number of outputs — keeps
 * the ligand rows Bt[b2][q][c][m, m..L][0..N) of those two orders in shared memory, and walks over the z steps of
 * the launch: per z it stages the two T^m blocks (rows l, columns l1 >= m) and writes the (L + 2) * N
 * translated coefficients.  Bt is read from HBM once per launch instead of once per z (an untiled
 * one-thread-per-output form read 10.5 GB per 64 z at L = 15, profiles/r1c_ncu_traffic.json); l1 ascends. */
__global__ void __launch_bounds__(288)
k_translate_tiled(int L, int qnum, int nz, const int *__restrict__ slab_flag, const double2 *__restrict__ T,
                  const double2 *__restrict__ Bt, double2 *__restrict__ St)
{
	extern __shared__ double2 s_tile[];
	const int nb = L + 1, N = 2 * L + 1, NP = sxs_row_pad(N), ML = sxs_ml_count(L);
	const int ma = blockIdx.x, mb = L - (int)blockIdx.x;
	const int nla = nb - ma, nlb = (mb != ma) ? nb - mb : 0;
	const int q = blockIdx.y / 3, c = blockIdx.y % 3, b2 = blockIdx.z;
	const int nrow = nla + nlb, nout = nrow * N;
	const int mla = sxs_ml_index(L, ma, ma), mlb = sxs_ml_index(L, mb, mb);
	double2 *sB = s_tile;            /* [nrow][N] */
	double2 *sT = s_tile + nout;     /* [nla][nla] then [nlb][nlb] */

	const double2 *bsrc = Bt + (((size_t)b2 * qnum + q) * 3 + c) * ML * NP;
	for (int e = threadIdx.x; e < nla * N; e += blockDim.x) {
		const int r = e / N, g = e - r * N;
		sB[e] = bsrc[(size_t)(mla + r) * NP + g];
	}
	for (int e = threadIdx.x; e < nlb * N; e += blockDim.x) {
		const int r = e / N, g = e - r * N;
		sB[nla * N + e] = bsrc[(size_t)(mlb + r) * NP + g];
	}
	for (int zl = 0; zl < nz; zl++) {
		if (!slab_flag[zl * nb + b2]) {
			continue;
		}
		__syncthreads();
		const double2 *tsrc = T + ((size_t)zl * qnum + q) * nb * nb * nb;
		for (int e = threadIdx.x; e < nla * nla; e += blockDim.x) {
			const int r = e / nla, j = e - r * nla;
			sT[e] = tsrc[((size_t)ma * nb + ma + r) * nb + ma + j];
		}
		for (int e = threadIdx.x; e < nlb * nlb; e += blockDim.x) {
			const int r = e / nlb, j = e - r * nlb;
			sT[nla * nla + e] = tsrc[((size_t)mb * nb + mb + r) * nb + mb + j];
		}
		__syncthreads();
		double2 *dst = St + ((((size_t)zl * nb + b2) * qnum + q) * 3 + c) * ML * NP;
		for (int o = threadIdx.x; o < nout; o += blockDim.x) {
			const int r = o / N, g = o - r * N;
			const double2 *trow, *bcol;
			int n, mlrow;
			if (r < nla) {
				trow = sT + r * nla; bcol = sB + g; n = nla; mlrow = mla + r;
			} else {
				trow = sT + nla * nla + (r - nla) * nlb; bcol = sB + nla * N + g; n = nlb; mlrow = mlb + (r - nla);
			}
			double re = 0.0, im = 0.0;
			for (int j = 0; j < n; j++) {
				const double2 t = trow[j];
				const double2 b = bcol[j * N];
				/* conj(t) * b */
				re += t.x * b.x + t.y * b.y;
				im += t.x * b.y - t.y * b.x;
			}
			dst[(size_t)mlrow * NP + g] = make_double2(re, im);
		}
	}
}

/* Register-tiled form of the same contraction (the one launched).  Same block ownership and staging as
 * k_translate_tiled; a thread now produces a 4 x 2 tile — four consecutive rows l of one order and the two columns
 * g, g + GH (GH = (N + 1) / 2) — so that per l1 it reads two ligand values and four (warp-broadcast) T values from
 * shared memory for eight complex MACs: 0.75 LDS.128 per complex MAC instead of 2, which is what bounded the
 * one-output-per-thread form (ncu: shared-memory pipe, 23 % of HBM peak).  Every output sees the same operations in
 * the same order as before (product, fused multiply-add, add; l1 ascending): St is bit-identical.
 * [B200] 64 z, every slab: L = 15 8.4 -> 6.9 ms, L = 30 160 -> 136 ms (128 registers: 216 ms; 2-row tiles: 160 ms).
 * Six FP64 instructions per complex MAC put the FP64-pipe floor at 61 ms for L = 30; the HBM floor (151 GB written)
 * is 25 ms. */
#ifndef SXS_TR_ROWS
#define SXS_TR_ROWS 4
#endif
#ifndef SXS_TR_MINBLOCKS
#define SXS_TR_MINBLOCKS 2 /* 64 registers: at L = 30 (288 threads) three blocks per SM instead of one */
#endif
#ifndef SXS_TR_UNROLL
#define SXS_TR_UNROLL 1
#endif
#define SXS_PRAGMA_(x) _Pragma(#x)
#define SXS_UNROLL(n) SXS_PRAGMA_(unroll n)
__global__ void __launch_bounds__(512, SXS_TR_MINBLOCKS)
k_translate_rt(int L, int qnum, int nz, const int *__restrict__ slab_flag, const double2 *__restrict__ T,
               const double2 *__restrict__ Bt, double2 *__restrict__ St)
{
	extern __shared__ double2 s_tile[];
	const int nb = L + 1, N = 2 * L + 1, NP = sxs_row_pad(N), ML = sxs_ml_count(L), GH = (N + 1) / 2;
	const int ma = blockIdx.x, mb = L - (int)blockIdx.x;
	const int nla = nb - ma, nlb = (mb != ma) ? nb - mb : 0;
	const int q = blockIdx.y / 3, c = blockIdx.y % 3, b2 = blockIdx.z;
	const int nrow = nla + nlb, nout = nrow * N;
	const int mla = sxs_ml_index(L, ma, ma), mlb = sxs_ml_index(L, mb, mb);
	double2 *sB = s_tile;            /* [nrow][N] */
	double2 *sT = s_tile + nout;     /* [nla][nla] then [nlb][nlb] */

	const double2 *bsrc = Bt + (((size_t)b2 * qnum + q) * 3 + c) * ML * NP;
	for (int e = threadIdx.x; e < nla * N; e += blockDim.x) {
		const int r = e / N, g = e - r * N;
		sB[e] = bsrc[(size_t)(mla + r) * NP + g];
	}
	for (int e = threadIdx.x; e < nlb * N; e += blockDim.x) {
		const int r = e / N, g = e - r * N;
		sB[nla * N + e] = bsrc[(size_t)(mlb + r) * NP + g];
	}
	/* this thread's tile */
	const int tiles_a = (nla + SXS_TR_ROWS - 1) / SXS_TR_ROWS, tiles_b = (nlb + SXS_TR_ROWS - 1) / SXS_TR_ROWS;
	const int tile = threadIdx.x / GH, g0 = threadIdx.x - tile * GH, g1 = g0 + GH;
	const bool active = tile < tiles_a + tiles_b;
	const bool in_a = tile < tiles_a;
	const int n = in_a ? nla : nlb;                               /* rows = columns of this order's T block */
	const int r0 = SXS_TR_ROWS * (in_a ? tile : tile - tiles_a); /* first row of the tile inside its order */
	const double2 *bcol = sB + (in_a ? 0 : nla * N);
	const int mlrow0 = (in_a ? mla : mlb) + r0;
	const bool has_g1 = g1 < N;
	int rr[SXS_TR_ROWS];
#pragma unroll
	for (int k = 0; k < SXS_TR_ROWS; k++) {
		rr[k] = (r0 + k < n) ? r0 + k : n - 1; /* rows past the block repeat the last one and are not stored */
	}

	/* The two T blocks of a z step are copied into one of two shared-memory stages (cp.async) while the step before it
	 * is computed: with a synchronous staging loop per z, 20 % of the stall samples at L = 30 sat on the long scoreboard
	 * (the T table misses L2 behind the stream of St stores) and 21 % on the two barriers per z (r2ai ncu).  One barrier
	 * per z is left: it publishes the landed stage and retires the one the next copy overwrites. */
	const int tsz = nla * nla + nlb * nlb;
	auto t_stage = [&](int zl, double2 *stg) {
		const double2 *tsrc = T + ((size_t)zl * qnum + q) * nb * nb * nb;
		for (int e = threadIdx.x; e < nla * nla; e += blockDim.x) {
			const int r = e / nla, j = e - r * nla;
			const unsigned dsta = (unsigned)__cvta_generic_to_shared(stg + e);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsta), "l"(tsrc + ((size_t)ma * nb + ma + r) * nb + ma + j) : "memory");
		}
		for (int e = threadIdx.x; e < nlb * nlb; e += blockDim.x) {
			const int r = e / nlb, j = e - r * nlb;
			const unsigned dstb = (unsigned)__cvta_generic_to_shared(stg + nla * nla + e);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstb), "l"(tsrc + ((size_t)mb * nb + mb + r) * nb + mb + j) : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	int zl = 0;
	while (zl < nz && !slab_flag[zl * nb + b2]) {
		zl++;
	}
	int stage = 0;
	if (zl < nz) {
		t_stage(zl, sT);
	}
	for (; zl < nz;) {
		int zn = zl + 1;
		while (zn < nz && !slab_flag[zn * nb + b2]) {
			zn++;
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads(); /* this z's blocks (and, first time, the B rows) are in place; everybody is done with the previous z */
		if (zn < nz) {
			t_stage(zn, sT + (stage ^ 1) * tsz);
		}
		const double2 *tblk = sT + stage * tsz + (in_a ? 0 : nla * nla);
		const int zcur = zl;
		zl = zn;
		stage ^= 1;
		if (!active) {
			continue;
		}
		double re[SXS_TR_ROWS][2], im[SXS_TR_ROWS][2];
#pragma unroll
		for (int k = 0; k < SXS_TR_ROWS; k++) {
			re[k][0] = re[k][1] = im[k][0] = im[k][1] = 0.0;
		}
		SXS_UNROLL(SXS_TR_UNROLL)
		for (int j = 0; j < n; j++) {
			const double2 bA = bcol[j * N + g0];
			const double2 bB = has_g1 ? bcol[j * N + g1] : make_double2(0.0, 0.0);
#pragma unroll
			for (int k = 0; k < SXS_TR_ROWS; k++) {
				const double2 t = tblk[rr[k] * n + j];
				/* conj(t) * b, accumulated as acc + fma(t.x, b.x, t.y * b.y) and acc + fma(t.x, b.y, -(t.y * b.x)) */
				re[k][0] = __dadd_rn(re[k][0], __fma_rn(t.x, bA.x, __dmul_rn(t.y, bA.y)));
				im[k][0] = __dadd_rn(im[k][0], __fma_rn(t.x, bA.y, -__dmul_rn(t.y, bA.x)));
				re[k][1] = __dadd_rn(re[k][1], __fma_rn(t.x, bB.x, __dmul_rn(t.y, bB.y)));
				im[k][1] = __dadd_rn(im[k][1], __fma_rn(t.x, bB.y, -__dmul_rn(t.y, bB.x)));
			}
		}
		double2 *dst = St + ((((size_t)zcur * nb + b2) * qnum + q) * 3 + c) * ML * NP;
#pragma unroll
		for (int k = 0; k < SXS_TR_ROWS; k++) {
			if (r0 + k < n) {
				/* streaming stores: St is read back by K3 from HBM anyway (10 GB per group) and must not push T out of L2 */
				__stcs(&dst[(size_t)(mlrow0 + k) * NP + g0], make_double2(re[k][0], im[k][0]));
				if (has_g1) {
					__stcs(&dst[(size_t)(mlrow0 + k) * NP + g1], make_double2(re[k][1], im[k][1]));
				}
			}
		}
	}
}

/* ----------------------------------------------------- pose list handling */

#define SXS_KEY_NONE 0xFFFFFFFFFFFFFFFFull

/* flat index (z,b1,b2,a2,g1,g2 digits, src/index.c:9-29) -> sort key (z, b2, b1, g2 / 8, g1, g2 % 8, a2).
 * Inside a cell the points are ordered by the 128-byte band of their ligand operand first, then by g1: eight
 * consecutive points (a quarter warp of k_cross) then read ligand values of ONE cache line and receptor values of
 * neighbouring g1, i.e. one line as well.  With (g1, g2) order a quarter warp's ligand values spread over the
 * whole row (4 lines).  per_cell = (NP / 8) * N * 8 * N keys per (z, b2, b1). */
template <typename IndexT>
__global__ void k_make_keys(const IndexT *__restrict__ index, long long nout, int nb, int N, int z_lo, int z_hi,
                            unsigned long long *__restrict__ keys, unsigned int *__restrict__ rows)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nout) {
		return;
	}
	long long v = (long long)index[i];
	unsigned long long key = SXS_KEY_NONE;
	if (v >= 0) {
		sxs_pose_digits d;
		d.g2 = (int)(v % N); v /= N;
		d.g1 = (int)(v % N); v /= N;
		d.a2 = (int)(v % N); v /= N;
		d.b2 = (int)(v % nb); v /= nb;
		d.b1 = (int)(v % nb);
		const long long z = v / nb;
		if (z >= z_lo && z < z_hi) {
			d.z = (int)z;
			key = sxs_key_pack(nb, N, d);
		}
	}
	keys[i] = key;
	rows[i] = (unsigned int)i;
}

__global__ void k_mark_heads(const unsigned long long *__restrict__ ks, long long n, unsigned int *__restrict__ head)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const unsigned long long k = ks[i];
	head[i] = (k != SXS_KEY_NONE && (i == 0 || ks[i - 1] != k)) ? 1u : 0u;
}

/* pid_incl = inclusive scan of heads; point id of sorted row i is pid_incl[i]-1 */
__global__ void k_compact_points(const unsigned long long *__restrict__ ks, const unsigned int *__restrict__ pid_incl,
                                 long long n, unsigned long long *__restrict__ pkeys)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const unsigned long long k = ks[i];
	if (k != SXS_KEY_NONE && (i == 0 || ks[i - 1] != k)) {
		pkeys[pid_incl[i] - 1] = k;
	}
}

/* zoff[z] = first point whose z digit >= z (z = 0..znum); zoff[znum+1] = number of valid sorted rows */
__global__ void k_z_offsets(const unsigned long long *__restrict__ pkeys, long long np,
                            const unsigned long long *__restrict__ ks, long long n, int znum,
                            unsigned long long per_z, long long *__restrict__ zoff)
{
	const int z = blockIdx.x * blockDim.x + threadIdx.x;
	if (z <= znum) {
		const unsigned long long target = (unsigned long long)z * per_z;
		long long lo = 0, hi = np;
		while (lo < hi) {
			const long long mid = (lo + hi) >> 1;
			if (pkeys[mid] < target) lo = mid + 1; else hi = mid;
		}
		zoff[z] = lo;
	} else if (z == znum + 1) {
		long long lo = 0, hi = n;
		while (lo < hi) {
			const long long mid = (lo + hi) >> 1;
			if (ks[mid] != SXS_KEY_NONE) lo = mid + 1; else hi = mid;
		}
		zoff[z] = lo;
	}
}

__global__ void k_slab_flags(const unsigned long long *__restrict__ pkeys, long long p0, long long p1, int z0, int nb,
                             unsigned long long per_zb2, int *__restrict__ flag)
{
	const long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= p1) {
		return;
	}
	const long long zb2 = (long long)(pkeys[p] / per_zb2); /* z*nb + b2 */
	flag[zb2 - (long long)z0 * nb] = 1;
}

/* ------------------------------------------------------------ K3: cross terms */

__device__ __forceinline__ void cmac(double2 &acc, const double2 a, const double2 b)
{
	acc.x = fma(a.x, b.x, acc.x);
	acc.x = fma(-a.y, b.y, acc.x);
	acc.y = fma(a.x, b.y, acc.y);
	acc.y = fma(a.y, b.x, acc.y);
}

/* One thread per distinct grid point, one q per blockIdx.y.  Consecutive threads are consecutive in key order
 * (k_make_keys), so a block mostly works inside one cell and a quarter warp inside one band of g2: its ligand
 * operands are 16-byte elements of one 128-byte line, its receptor operands those of a few neighbouring g1 —
 * served by L1/L2.
 * X[sxs_x_index(p - xbase, qnum, q, k)] = const_k[q] + 2 F_k   (fill_const + fill_var, src/fftsaxs.c:52-108), in tiles
 * of 32 points so that a warp of K4 reads every term with one coalesced load. */
/* [B200] 1.12 M points, (g1, g2, a2) order, 496-byte rows: 256x1 39.3 ms, 256x3 (80 regs, spills) 40.9, 128x4 38.4,
 * 128x5 45.1, 64x8 38.4, 512x1 41.1.  Band-major order + 512-byte rows: 128x4 36.1 (L1 data-pipe wavefronts per
 * 16-byte warp load 6.4 -> 4.2, the minimum: one per quarter warp); + next row requested before the MACs of the
 * current one: 128x4 35.7-36.5, 256x2 36.5, 512x1 38.1, 64x8 35.4 (launched), 32x16 35.2. */
#ifndef SXS_CROSS_THREADS
#define SXS_CROSS_THREADS 64
#endif
#ifndef SXS_CROSS_MINBLOCKS
#define SXS_CROSS_MINBLOCKS 8
#endif
/* K > 1: points that differ only in a2 (same cell, g1, g2) share every operand and all nine inner sums; they differ
 * in the phase w^(m a2) applied once per m.  A thread then takes up to K such points (consecutive in key order; the
 * groups are formed per launch by k_group_*): one set of loads and MACs, K phase accumulations.  Each point still sees
 * exactly the operations of the K = 1 form in the same order, so X is bit-identical.  groups[i] = first point of the
 * group (relative to p0) | (members - 1) << SXS_GROUP_SHIFT. */
#define SXS_GROUP_SHIFT 28
template <int K>
__global__ void __launch_bounds__(SXS_CROSS_THREADS, (K == 1) ? SXS_CROSS_MINBLOCKS : (K == 2 ? 6 : 4))
k_cross(int L, int qnum, const unsigned long long *__restrict__ pkeys, long long p0, long long p1, long long xbase, int z0,
        const unsigned int *__restrict__ groups, unsigned int ngroups,
        const double2 *__restrict__ At, const double2 *__restrict__ St, const double2 *__restrict__ tw,
        const double *__restrict__ cst, double *__restrict__ X)
{
	extern __shared__ double2 s_tw[];
	const int N = 2 * L + 1, nb = L + 1, ML = sxs_ml_count(L);
	for (int i = threadIdx.x; i < N; i += blockDim.x) {
		s_tw[i] = tw[i];
	}
	__syncthreads();
	long long p;
	int members = 1;
	if (K == 1) {
		p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
		if (p >= p1) {
			return;
		}
	} else {
		const unsigned int gi = blockIdx.x * blockDim.x + threadIdx.x;
		if (gi >= ngroups) {
			return;
		}
		const unsigned int word = groups[gi];
		p = p0 + (long long)(word & ((1u << SXS_GROUP_SHIFT) - 1u));
		members = (int)(word >> SXS_GROUP_SHIFT) + 1;
	}
	const int q = blockIdx.y;
	const int NP = sxs_row_pad(N), nband = NP / 8;
	/* the digits of sxs_key_unpack, spelled out: this order of the divisions keeps the kernel at 108 / 122 registers */
	unsigned long long key = pkeys[p];
	int a2[K];
	a2[0] = (int)(key % N); key /= N;
	int g2 = (int)(key % 8); key /= 8;
	const int g1 = (int)(key % N); key /= N;
	g2 += 8 * (int)(key % nband); key /= nband;
	const int b1 = (int)(key % nb); key /= nb;
	const int b2 = (int)(key % nb);
	const int z = (int)(key / nb);
	const int slab = (z - z0) * nb + b2;
#pragma unroll
	for (int k = 1; k < K; k++) {
		a2[k] = (k < members) ? (int)(pkeys[p + k] % N) : 0;
	}

	const size_t cstride = (size_t)ML * NP;
	const double2 *a_ptr = At + (((size_t)b1 * qnum + q) * 3) * cstride + g1;
	const double2 *s_ptr = St + (((size_t)slab * qnum + q) * 3) * cstride + g2;

	double f[K][6];
	int ka[K]; /* (m*a2) mod N */
#pragma unroll
	for (int k = 0; k < K; k++) {
		ka[k] = 0;
#pragma unroll
		for (int j = 0; j < 6; j++) {
			f[k][j] = 0.0;
		}
	}
	/* one flat loop over the ML (m, l) rows; the six operands of the next row are requested before the nine complex
	 * MACs of the current one, so that a warp always has a row in flight behind its arithmetic */
	double2 cvv = {0, 0}, cvd = {0, 0}, cvw = {0, 0}, cdd = {0, 0}, cdw = {0, 0}, cww = {0, 0};
	double2 av = __ldg(a_ptr), ad = __ldg(a_ptr + cstride), aw = __ldg(a_ptr + 2 * cstride);
	double2 sv = __ldg(s_ptr), sd = __ldg(s_ptr + cstride), sw = __ldg(s_ptr + 2 * cstride);
	int m = 0, l = 0;
	size_t row = 0;
#pragma unroll 2
	for (int step = 0; step < ML; step++) {
		row += NP;
		const size_t nrow = (step + 1 < ML) ? row : 0;
		const double2 nav = __ldg(a_ptr + nrow), nad = __ldg(a_ptr + cstride + nrow), naw = __ldg(a_ptr + 2 * cstride + nrow);
		const double2 nsv = __ldg(s_ptr + nrow), nsd = __ldg(s_ptr + cstride + nrow), nsw = __ldg(s_ptr + 2 * cstride + nrow);
		cmac(cvv, av, sv);
		cmac(cvd, av, sd); cmac(cvd, ad, sv);
		cmac(cvw, av, sw); cmac(cvw, aw, sv);
		cmac(cdd, ad, sd);
		cmac(cdw, ad, sw); cmac(cdw, aw, sd);
		cmac(cww, aw, sw);
		if (l == L) {
			const double fac = (m == 0) ? 1.0 : 2.0;
#pragma unroll
			for (int k = 0; k < K; k++) {
				const double2 w = s_tw[ka[k]];
				f[k][0] += fac * (w.x * cvv.x - w.y * cvv.y);
				f[k][1] += fac * (w.x * cvd.x - w.y * cvd.y);
				f[k][2] += fac * (w.x * cvw.x - w.y * cvw.y);
				f[k][3] += fac * (w.x * cdd.x - w.y * cdd.y);
				f[k][4] += fac * (w.x * cdw.x - w.y * cdw.y);
				f[k][5] += fac * (w.x * cww.x - w.y * cww.y);
				ka[k] += a2[k];
				if (ka[k] >= N) {
					ka[k] -= N;
				}
			}
			cvv = cvd = cvw = cdd = cdw = cww = make_double2(0.0, 0.0);
			m++;
			l = m;
		} else {
			l++;
		}
		av = nav; ad = nad; aw = naw; sv = nsv; sd = nsd; sw = nsw;
	}
	const double c0 = cst[0 * qnum + q], c1 = cst[1 * qnum + q], c2 = cst[2 * qnum + q], c3 = cst[3 * qnum + q],
	             c4 = cst[4 * qnum + q], c5 = cst[5 * qnum + q];
#pragma unroll
	for (int k = 0; k < K; k++) {
		if (k < members) {
			double2 *xo = reinterpret_cast<double2 *>(X + sxs_x_index(p + k - xbase, qnum, q, 0));
			xo[0] = make_double2(c0 + 2.0 * f[k][0], c1 + 2.0 * f[k][1]);
			xo[1] = make_double2(c2 + 2.0 * f[k][2], c3 + 2.0 * f[k][3]);
			xo[2] = make_double2(c4 + 2.0 * f[k][4], c5 + 2.0 * f[k][5]);
		}
	}
}

/* start[i] = i where point p0+i opens a run of points that differ only in a2, else 0; a running maximum then gives
 * every point the start of its run */
__global__ void k_group_heads(const unsigned long long *__restrict__ pkeys, long long p0, unsigned int cnt,
                              unsigned long long per_run, unsigned int *__restrict__ start)
{
	const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cnt) {
		return;
	}
	const bool head = (i == 0) || (pkeys[p0 + i] / per_run != pkeys[p0 + i - 1] / per_run);
	start[i] = head ? i : 0u;
}

/* lead[i] = 1 where point i opens a group of K inside its run */
__global__ void k_group_leads(const unsigned int *__restrict__ run_start, unsigned int cnt, int K,
                              unsigned int *__restrict__ lead)
{
	const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cnt) {
		return;
	}
	lead[i] = ((i - run_start[i]) % (unsigned)K == 0u) ? 1u : 0u;
}

/* lead_excl = exclusive sum of lead = group number of every leading point; groups[cnt] receives the number of groups */
__global__ void k_group_emit(const unsigned int *__restrict__ run_start, const unsigned int *__restrict__ lead_excl,
                             unsigned int cnt, int K, unsigned int *__restrict__ groups)
{
	const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cnt) {
		return;
	}
	const unsigned int rs = run_start[i];
	const bool lead = (i - rs) % (unsigned)K == 0u;
	if (lead) {
		unsigned int members = 1;
		while (members < (unsigned)K && i + members < cnt && run_start[i + members] == rs) {
			members++;
		}
		groups[lead_excl[i]] = i | ((members - 1u) << SXS_GROUP_SHIFT);
	}
	if (i == cnt - 1) {
		groups[cnt] = lead_excl[i] + (lead ? 1u : 0u);
	}
}

struct sxs_max_op {
	__device__ __forceinline__ unsigned int operator()(unsigned int a, unsigned int b) const { return a > b ? a : b; }
};

/* ---------------------------------------------------------------- scatter */

/* sorted row i (valid rows only) -> point pid_incl[i]-1 -> user arrays at rows_sorted[i] */
__global__ void k_scatter(const unsigned int *__restrict__ rows_sorted, const unsigned int *__restrict__ pid_incl,
                          long long nvalid, const double *__restrict__ res, double *__restrict__ scores,
                          double *__restrict__ c1, double *__restrict__ c2)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nvalid) {
		return;
	}
	const unsigned int row = rows_sorted[i];
	const double *r = res + (size_t)(pid_incl[i] - 1) * 4;
	scores[row] = r[0];
	c1[row] = r[1];
	c2[row] = r[2];
}

/* compact variant for the host path: out3[i*3..] in sorted order, row ids come from rows_sorted */
__global__ void k_gather_sorted(const unsigned int *__restrict__ pid_incl, long long nvalid,
                                const double *__restrict__ res, double *__restrict__ out3)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nvalid) {
		return;
	}
	const double *r = res + (size_t)(pid_incl[i] - 1) * 4;
	out3[i * 3 + 0] = r[0];
	out3[i * 3 + 1] = r[1];
	out3[i * 3 + 2] = r[2];
}

/* cross terms of sorted rows back to list order (stage access for the parity tests) */
__global__ void k_gather_cross(const unsigned int *__restrict__ rows_sorted, const unsigned int *__restrict__ pid_incl,
                               long long nvalid, long long p0, long long p1, long long xbase,
                               const double *__restrict__ X, int qnum, double *__restrict__ cross)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nvalid) {
		return;
	}
	const long long p = (long long)pid_incl[i] - 1;
	if (p < p0 || p >= p1) {
		return;
	}
	const unsigned int row = rows_sorted[i];
	for (int q = 0; q < qnum; q++) {
		for (int k = 0; k < 6; k++) {
			cross[((size_t)row * 6 + k) * qnum + q] = X[sxs_x_index(p - xbase, qnum, q, k)];
		}
	}
}

/* ------------------------------------------------------------ plan set-up */

static unsigned grid_for(size_t total, int threads, unsigned cap = 148u * 16u)
{
	size_t b = (total + threads - 1) / threads;
	if (b > cap) b = cap;
	if (b < 1) b = 1;
	return (unsigned)b;
}

extern "C" int sxs_cuda_plan_set_molecules(sxs_cuda_plan *p, const double *coefA, const double *coefB)
{
	SXS_CK(cudaSetDevice(p->device));
	const size_t ncoef = (size_t)3 * p->qnum * p->nb * p->nb;
	SXS_CK(cudaMemcpy(p->d_coefA, coefA, sizeof(double2) * ncoef, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(p->d_coefB, coefB, sizeof(double2) * ncoef, cudaMemcpyHostToDevice));
	if (sxs_launch_pair_const((const double *)p->d_coefA, (const double *)p->d_coefB, p->qnum, p->L, p->d_const, 0) != 0) {
		return -1;
	}
	const size_t nrot = (size_t)p->nb * p->qnum * 3 * p->ML * sxs_row_pad(p->N);
	const size_t shm = sizeof(double2) * p->N;
	k_rotate<<<grid_for(nrot, 256), 256, shm>>>(p->L, p->qnum, p->d_dwig, p->d_coefA, p->d_tw, 0, p->d_At);
	SXS_CK_LAUNCH();
	k_rotate<<<grid_for(nrot, 256), 256, shm>>>(p->L, p->qnum, p->d_dwig, p->d_coefB, p->d_tw, 1, p->d_Bt);
	SXS_CK_LAUNCH();
	SXS_CK(cudaDeviceSynchronize());
	p->have_molecules = 1;
	return 0;
}

extern "C" int sxs_cuda_plan_set_experiment(sxs_cuda_plan *p, const double *a, double mult, double peak)
{
	SXS_CK(cudaSetDevice(p->device));
	SXS_CK(cudaMemcpy(p->d_a, a, sizeof(double) * 6 * p->qnum, cudaMemcpyHostToDevice));
	p->mult = mult;
	p->peak = peak;
	p->have_experiment = 1;
	return 0;
}

extern "C" int sxs_cuda_plan_set_translations(sxs_cuda_plan *p, const double *bessel, int znum)
{
	SXS_CK(cudaSetDevice(p->device));
	if (znum < 1 || znum > 4000) {
		sxs_cuda_set_error("znum %d out of range", znum);
		return -1;
	}
	const size_t n = (size_t)znum * p->qnum * p->N;
	if (p->d_bessel != NULL && znum != p->znum) {
		cudaFree(p->d_bessel);
		p->d_bessel = NULL;
	}
	if (p->d_bessel == NULL) {
		SXS_CK(cudaMalloc(&p->d_bessel, sizeof(double) * n));
	}
	SXS_CK(cudaMemcpy(p->d_bessel, bessel, sizeof(double) * n, cudaMemcpyHostToDevice));
	p->znum = znum;
	return 0;
}

/* ------------------------------------------------------------ score (core) */

/* T-matrices of `zspan` z steps (d_zlist) and the translation of every flagged (z, b2) slab of the ligand table:
 * K2b + K2c, two launches */
static int launch_translate(sxs_cuda_plan *p, int zspan, const int *d_zlist, cudaStream_t st)
{
	const int L = p->L, Q = p->qnum, N = p->N, nb = p->nb;
	k_tmatrix<<<grid_for((size_t)zspan * Q * nb * nb * nb, 256), 256, 0, st>>>(L, Q, zspan, d_zlist, p->d_dsymb, p->d_bessel, p->d_T);
	SXS_CK_LAUNCH();
	{
		const int npair = (L + 2) / 2, rows = L + 2;
		/* the B rows of the block's two orders + two stages of its T blocks (nla^2 + nlb^2 <= nb^2 + 1 elements each) */
		const size_t shm_t = sizeof(double2) * ((size_t)rows * N + 2 * ((size_t)nb * nb + 1));
		const int GH = (N + 1) / 2;
		int max_tiles = 0;
		for (int ma = 0; ma < npair; ma++) {
			const int mb = L - ma, nla = nb - ma, nlb = (mb != ma) ? nb - mb : 0;
			const int t = (nla + SXS_TR_ROWS - 1) / SXS_TR_ROWS + (nlb + SXS_TR_ROWS - 1) / SXS_TR_ROWS;
			if (t > max_tiles) max_tiles = t;
		}
		const int threads_rt = 32 * ((max_tiles * GH + 31) / 32);
		/* one output per thread: beyond L = 40 (the tiled form would need more than 512 threads), and for the
		 * bit-identity test */
		if (getenv("SXS_TRANSLATE_V1") != NULL || threads_rt > 512) {
			int iters = (rows * N + 287) / 288;
			int threads = 32 * ((rows * N + 32 * iters - 1) / (32 * iters));
			if (threads > 288) threads = 288;
			if (shm_t > 48 * 1024) {
				SXS_CK(cudaFuncSetAttribute(k_translate_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm_t));
			}
			k_translate_tiled<<<dim3(npair, 3 * Q, nb), threads, shm_t, st>>>(L, Q, zspan, p->d_slab_flag, p->d_T, p->d_Bt, p->d_St);
		} else {
			/* 4-row x 2-column tiles: (ceil(nla/4) + ceil(nlb/4)) * GH threads, nla + nlb = L + 2 */
			if (shm_t > 48 * 1024) {
				SXS_CK(cudaFuncSetAttribute(k_translate_rt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm_t));
			}
			k_translate_rt<<<dim3(npair, 3 * Q, nb), threads_rt, shm_t, st>>>(L, Q, zspan, p->d_slab_flag, p->d_T, p->d_Bt, p->d_St);
		}
	}
	SXS_CK_LAUNCH();
	return 0;
}

template <typename IndexT>
static int score_core(sxs_cuda_plan *p, const IndexT *d_index, long long nout, int z_lo, int z_hi, double *d_scores,
                      double *d_c1, double *d_c2, double *d_out3_sorted, double *d_cross_out, cudaStream_t st,
                      long long *nvalid_out)
{
	if (!p->have_molecules || !p->have_experiment || p->d_bessel == NULL) {
		sxs_cuda_set_error("plan is missing molecules, experiment or translations");
		return -1;
	}
	if (nout > 0x7FFFFFF0ll) {
		sxs_cuda_set_error("more than 2^31 rows per call");
		return -1;
	}
	memset(p->stats, 0, sizeof(p->stats));
	if (nvalid_out) *nvalid_out = 0;
	if (nout <= 0) {
		return 0;
	}
	if (z_lo < 0) z_lo = 0;
	if (z_hi > p->znum) z_hi = p->znum;
	if (z_hi <= z_lo) {
		return 0;
	}
	const int L = p->L, nb = p->nb, N = p->N, ML = p->ML, Q = p->qnum, znum = p->znum;
	long long launches = 0;

	/* --- keys, sort, distinct points --- */
	size_t cap;
	cap = p->cap_rows;
	if ((size_t)nout > p->cap_rows || p->d_keys == NULL) {
		size_t c1_ = 0, c2_ = 0, c3_ = 0, c4_ = 0, c5_ = 0, c6_ = 0;
		void *olds[] = {p->d_keys, p->d_keys_sorted, p->d_pkeys, p->d_rows, p->d_rows_sorted, p->d_pid};
		for (int i = 0; i < 6; i++) if (olds[i]) cudaFree(olds[i]);
		p->d_keys = p->d_keys_sorted = p->d_pkeys = NULL;
		p->d_rows = p->d_rows_sorted = p->d_pid = NULL;
		if (ensure(&p->d_keys, &c1_, (size_t)nout) || ensure(&p->d_keys_sorted, &c2_, (size_t)nout) ||
		    ensure(&p->d_pkeys, &c3_, (size_t)nout) || ensure(&p->d_rows, &c4_, (size_t)nout + 1) ||
		    ensure(&p->d_rows_sorted, &c5_, (size_t)nout) || ensure(&p->d_pid, &c6_, (size_t)nout)) {
			p->cap_rows = 0;
			return -1;
		}
		p->cap_rows = (size_t)nout;
	}
	(void)cap;
	const int T256 = 256;
	const unsigned gb = (unsigned)((nout + T256 - 1) / T256);
	timer_begin(p, 0, st);
	k_make_keys<IndexT><<<gb, T256, 0, st>>>(d_index, nout, nb, N, z_lo, z_hi, p->d_keys, p->d_rows);
	SXS_CK_LAUNCH(); launches++;

	unsigned long long max_key = ((unsigned long long)znum * nb * nb) * (unsigned long long)N * N * N;
	int bits = 1;
	while (bits < 64 && (max_key >> bits) != 0) bits++;
	/* the sentinel has every bit set: sorting on all 64 bits keeps it last */
	(void)bits;
	size_t need = 0, need2 = 0;
	cub::DeviceRadixSort::SortPairs(NULL, need, p->d_keys, p->d_keys_sorted, p->d_rows, p->d_rows_sorted, (int)nout, 0, 64, st);
	cub::DeviceScan::InclusiveSum(NULL, need2, p->d_pid, p->d_pid, (int)nout, st);
	if (need2 > need) need = need2;
	{
		unsigned char *tmp = (unsigned char *)p->d_cub;
		size_t c = p->cap_cub;
		if (ensure(&tmp, &c, need + 256)) return -1;
		p->d_cub = tmp; p->cap_cub = c;
	}
	size_t tb = p->cap_cub;
	cub::DeviceRadixSort::SortPairs(p->d_cub, tb, p->d_keys, p->d_keys_sorted, p->d_rows, p->d_rows_sorted, (int)nout, 0, 64, st);
	/* heads -> inclusive scan (reuse d_rows as the head buffer: the unsorted row ids are no longer needed) */
	k_mark_heads<<<gb, T256, 0, st>>>(p->d_keys_sorted, nout, p->d_rows);
	SXS_CK_LAUNCH(); launches++;
	tb = p->cap_cub;
	cub::DeviceScan::InclusiveSum(p->d_cub, tb, p->d_rows, p->d_pid, (int)nout, st);
	k_compact_points<<<gb, T256, 0, st>>>(p->d_keys_sorted, p->d_pid, nout, p->d_pkeys);
	SXS_CK_LAUNCH(); launches++;
	timer_end(p, 0, st);

	/* number of points: last inclusive-scan value (read back together with the z offsets) */
	if (p->d_zoff == NULL || p->cap_zoff < znum + 4) {
		if (p->d_zoff) cudaFree(p->d_zoff);
		SXS_CK(cudaMalloc(&p->d_zoff, sizeof(long long) * (znum + 4)));
		p->cap_zoff = znum + 4;
	}
	if (znum + 4 > 4096) {
		sxs_cuda_set_error("znum too large for the offset buffer");
		return -1;
	}
	/* np is needed by the binary search: read it first (tiny sync) */
	unsigned int np_u = 0;
	SXS_CK(cudaMemcpyAsync(&np_u, p->d_pid + (nout - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
	SXS_CK(cudaStreamSynchronize(st));
	const long long np = (long long)np_u;
	if (np == 0) {
		return 0;
	}
	const unsigned long long per_cell = sxs_keys_per_cell(N); /* keys per (z, b2, b1) */
	const unsigned long long per_zb2 = (unsigned long long)nb * per_cell;
	const unsigned long long per_z = per_zb2 * nb;
	k_z_offsets<<<(znum + 2 + 127) / 128, 128, 0, st>>>(p->d_pkeys, np, p->d_keys_sorted, nout, znum, per_z, p->d_zoff);
	SXS_CK_LAUNCH(); launches++;
	SXS_CK(cudaMemcpyAsync(p->h_zoff, p->d_zoff, sizeof(long long) * (znum + 2), cudaMemcpyDeviceToHost, st));
	SXS_CK(cudaStreamSynchronize(st));
	const long long *zoff = p->h_zoff;
	const long long nvalid = zoff[znum + 1];
	if (nvalid_out) *nvalid_out = nvalid;

	/* --- result buffer for all points --- */
	if (ensure(&p->d_res, &p->cap_res, (size_t)np * 4)) return -1;

	/* --- group sizing --- */
	const size_t slab_elems = (size_t)Q * 3 * ML * sxs_row_pad(N); /* double2 per (z,b2) slab */
	const size_t per_z_bytes = slab_elems * nb * sizeof(double2);
	int zg_max = (int)(p->budget_St / per_z_bytes);
	if (zg_max < 1) zg_max = 1;
	if (zg_max > z_hi - z_lo) zg_max = z_hi - z_lo;
	long long chunk_max = (long long)(p->budget_X / (sizeof(double) * 6 * Q));
	if (chunk_max < 1024) chunk_max = 1024;
	if (chunk_max > np) chunk_max = np;
	/* never size past what this call needs */
	{
		int zused = 0;
		for (int z = z_lo; z < z_hi; z++) zused += (zoff[z + 1] > zoff[z]);
		if (zg_max > zused) zg_max = zused > 0 ? zused : 1;
	}
	if (ensure(&p->d_St, &p->cap_St, slab_elems * nb * (size_t)zg_max)) return -1;
	if (ensure(&p->d_T, &p->cap_T, (size_t)zg_max * Q * nb * nb * nb)) return -1;
	if (ensure(&p->d_X, &p->cap_X, (size_t)chunk_max * 6 * Q)) return -1;
	if (p->d_slab_flag == NULL || p->cap_slab < zg_max * nb + zg_max) {
		if (p->d_slab_flag) cudaFree(p->d_slab_flag);
		SXS_CK(cudaMalloc(&p->d_slab_flag, sizeof(int) * (zg_max * nb + zg_max)));
		p->cap_slab = zg_max * nb + zg_max;
	}
	int *d_zlist = p->d_slab_flag + zg_max * nb;
	int *h_zlist = (int *)malloc(sizeof(int) * zg_max);

	long long nslabs_total = 0, ngroups = 0;
	long long xbase = -1, xend = -1; /* points [xbase, xend) have their cross terms in X and are not fitted yet */
	/* K3 threads take up to K points that differ only in a2: SXS_CROSS_GROUP = 1 (never), 2, 4; unset = by run length */
	int cross_group = 0;
	if (getenv("SXS_CROSS_GROUP") != NULL) {
		cross_group = atoi(getenv("SXS_CROSS_GROUP"));
		if (cross_group != 1 && cross_group != 2 && cross_group != 4) cross_group = 0;
	}
	if (cross_group != 1) {
		size_t n1 = 0, n2 = 0;
		cub::DeviceScan::InclusiveScan(NULL, n1, (unsigned int *)NULL, (unsigned int *)NULL, sxs_max_op(), (int)chunk_max, st);
		cub::DeviceScan::ExclusiveSum(NULL, n2, (unsigned int *)NULL, (unsigned int *)NULL, (int)chunk_max, st);
		if (n2 > n1) n1 = n2;
		if (n1 + 256 > p->cap_cub) {
			unsigned char *tmp = (unsigned char *)p->d_cub;
			size_t c = p->cap_cub;
			if (ensure(&tmp, &c, n1 + 256)) { free(h_zlist); return -1; }
			p->d_cub = tmp; p->cap_cub = c;
		}
	}
	int z = z_lo;
	while (z < z_hi) {
		/* gather up to zg_max z steps that hold points; they need not be contiguous, only ordered */
		int nz = 0;
		int z_first = -1, z_last = -1;
		while (z < z_hi && nz < zg_max) {
			if (zoff[z + 1] > zoff[z]) {
				if (z_first < 0) z_first = z;
				z_last = z;
				nz++;
			}
			z++;
			if (z_first >= 0 && (z - z_first) >= zg_max) break; /* slabs are addressed by z - z_first */
		}
		if (nz == 0) break;
		const int zspan = z_last - z_first + 1;
		for (int i = 0; i < zspan; i++) h_zlist[i] = z_first + i;
		const long long g0 = zoff[z_first], g1 = zoff[z_last + 1];
		ngroups++;

		SXS_CK(cudaMemsetAsync(p->d_slab_flag, 0, sizeof(int) * zspan * nb, st));
		SXS_CK(cudaMemcpyAsync(d_zlist, h_zlist, sizeof(int) * zspan, cudaMemcpyHostToDevice, st));
		timer_begin(p, 1, st);
		k_slab_flags<<<(unsigned)((g1 - g0 + 255) / 256), 256, 0, st>>>(p->d_pkeys, g0, g1, z_first, nb, per_zb2, p->d_slab_flag);
		SXS_CK_LAUNCH(); launches++;
		if (launch_translate(p, zspan, d_zlist, st) != 0) {
			free(h_zlist);
			return -1;
		}
		launches++;
		SXS_CK_LAUNCH(); launches++;
		timer_end(p, 1, st);
		nslabs_total += (long long)zspan * nb;

		/* Cross terms of consecutive z groups pile up in X (point p at X index p - xbase) and are fitted by ONE launch
		 * when X is full or the list ends: a sparse list at large L has many z groups of a few thousand points each,
		 * far fewer than the ~113 000 fits K4 keeps in flight. */
		for (long long c0 = g0; c0 < g1;) {
			if (xbase < 0) {
				xbase = c0;
			}
			if (c0 - xbase >= chunk_max) { /* X is full: fit what it holds */
				timer_begin(p, 3, st);
				if (sxs_launch_fit(p->d_X, c0 - xbase, p->d_a, p->d_qvals, Q, p->mult, p->peak, 1, p->d_res + (size_t)xbase * 4,
				                   p->d_ticket, st) != 0) {
					free(h_zlist);
					return -1;
				}
				launches++;
				timer_end(p, 3, st);
				xbase = c0;
			}
			const long long room = xbase + chunk_max;
			const long long c1e = (room < g1) ? room : g1;
			const long long cnt = c1e - c0;
			timer_begin(p, 2, st);
			/* points that differ only in a2 share all operands: count the runs, then let a thread take up to K points
			 * of a run (d_keys and d_rows are free once the points are compacted) */
			int K = 1;
			unsigned int ngroups_k = 0;
			if (cross_group != 1 && cnt < (1ll << SXS_GROUP_SHIFT) && cnt >= 2) {
				unsigned int *run_start = (unsigned int *)p->d_keys, *lead = run_start + nout, *groups = p->d_rows;
				const unsigned gc = (unsigned)((cnt + 255) / 256);
				k_group_heads<<<gc, 256, 0, st>>>(p->d_pkeys, c0, (unsigned)cnt, (unsigned long long)N, run_start);
				SXS_CK_LAUNCH(); launches++;
				size_t tb2 = p->cap_cub;
				SXS_CK(cub::DeviceScan::InclusiveScan(p->d_cub, tb2, run_start, run_start, sxs_max_op(), (int)cnt, st));
				for (int pass = 0; pass < 2; pass++) {
					const int kk = (pass == 0) ? (1 << 30) : K; /* pass 0 counts the runs */
					k_group_leads<<<gc, 256, 0, st>>>(run_start, (unsigned)cnt, kk, lead);
					SXS_CK_LAUNCH(); launches++;
					tb2 = p->cap_cub;
					SXS_CK(cub::DeviceScan::ExclusiveSum(p->d_cub, tb2, lead, lead, (int)cnt, st));
					k_group_emit<<<gc, 256, 0, st>>>(run_start, lead, (unsigned)cnt, kk, groups);
					SXS_CK_LAUNCH(); launches += 2;
					SXS_CK(cudaMemcpyAsync(&ngroups_k, groups + cnt, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
					SXS_CK(cudaStreamSynchronize(st));
					if (pass == 0) {
						const double per_run = (double)cnt / (double)(ngroups_k > 0 ? ngroups_k : 1);
						K = (cross_group > 1) ? cross_group : (per_run < 1.04 ? 1 : (per_run < 1.8 ? 2 : 4));
						if (K == 1) {
							break;
						}
					}
				}
			}
			const size_t shm_c = sizeof(double2) * N;
			if (K == 1) {
				dim3 grid((unsigned)((cnt + SXS_CROSS_THREADS - 1) / SXS_CROSS_THREADS), Q);
				k_cross<1><<<grid, SXS_CROSS_THREADS, shm_c, st>>>(L, Q, p->d_pkeys, c0, c1e, xbase, z_first, NULL, 0u, p->d_At, p->d_St,
				                                        p->d_tw, p->d_const, p->d_X);
			} else {
				dim3 grid((ngroups_k + SXS_CROSS_THREADS - 1) / SXS_CROSS_THREADS, Q);
				if (K == 2) {
					k_cross<2><<<grid, SXS_CROSS_THREADS, shm_c, st>>>(L, Q, p->d_pkeys, c0, c1e, xbase, z_first, p->d_rows, ngroups_k,
					                                        p->d_At, p->d_St, p->d_tw, p->d_const, p->d_X);
				} else {
					K = 4;
					k_cross<4><<<grid, SXS_CROSS_THREADS, shm_c, st>>>(L, Q, p->d_pkeys, c0, c1e, xbase, z_first, p->d_rows, ngroups_k,
					                                        p->d_At, p->d_St, p->d_tw, p->d_const, p->d_X);
				}
				p->stats[3] += ngroups_k;
			}
			SXS_CK_LAUNCH(); launches++;
			timer_end(p, 2, st);
			if (d_cross_out != NULL) {
				k_gather_cross<<<(unsigned)((nvalid + 255) / 256), 256, 0, st>>>(p->d_rows_sorted, p->d_pid, nvalid, c0, c1e, xbase,
				                                                              p->d_X, Q, d_cross_out);
				SXS_CK_LAUNCH(); launches++;
			}
			xend = c1e;
			c0 = c1e;
		}
	}
	if (xbase >= 0 && xend > xbase) {
		timer_begin(p, 3, st);
		if (sxs_launch_fit(p->d_X, xend - xbase, p->d_a, p->d_qvals, Q, p->mult, p->peak, 1, p->d_res + (size_t)xbase * 4,
		                   p->d_ticket, st) != 0) {
			free(h_zlist);
			return -1;
		}
		launches++;
		timer_end(p, 3, st);
	}
	free(h_zlist);

	timer_begin(p, 4, st);
	if (d_scores != NULL) {
		k_scatter<<<(unsigned)((nvalid + 255) / 256), 256, 0, st>>>(p->d_rows_sorted, p->d_pid, nvalid, p->d_res, d_scores, d_c1, d_c2);
		SXS_CK_LAUNCH(); launches++;
	}
	if (d_out3_sorted != NULL) {
		k_gather_sorted<<<(unsigned)((nvalid + 255) / 256), 256, 0, st>>>(p->d_pid, nvalid, p->d_res, d_out3_sorted);
		SXS_CK_LAUNCH(); launches++;
	}
	timer_end(p, 4, st);
	p->stats[0] = np;
	p->stats[1] = nslabs_total;
	p->stats[2] = launches;
	p->stats[4] = ngroups;
	return 0;
}

/* ----------------------------------------------------------- public score */

extern "C" int sxs_cuda_plan_score_dev_i32(sxs_cuda_plan *p, const int *d_index, long long nout, int z_lo, int z_hi,
                                           double *d_scores, double *d_c1, double *d_c2, void *stream)
{
	SXS_CK(cudaSetDevice(p->device));
	return score_core<int>(p, d_index, nout, z_lo, z_hi, d_scores, d_c1, d_c2, NULL, NULL, (cudaStream_t)stream, NULL);
}

extern "C" int sxs_cuda_plan_score_dev_i64(sxs_cuda_plan *p, const long long *d_index, long long nout, int z_lo,
                                           int z_hi, double *d_scores, double *d_c1, double *d_c2, void *stream)
{
	SXS_CK(cudaSetDevice(p->device));
	return score_core<long long>(p, d_index, nout, z_lo, z_hi, d_scores, d_c1, d_c2, NULL, NULL, (cudaStream_t)stream, NULL);
}

/* ------------------------------------------------------------ K3, dense form (every grid point of a cell)
 * The list form above evaluates F at the poses a list names; a dense scan needs all N^3 points of a cell, and there the
 * inner sums are shared: for a pair (g1, g2) the nine sums over l of one m serve all N values of a2.
 *
 *   C_k^m(g1, g2) = sum_{l >= m} At^c[m,l,g1] St^c'[m,l,g2]          (9 products -> 6 terms; 4 lanes split l, butterfly sum)
 *   F_k(a2)       = Re sum_m fac_m w^(m a2) C_k^m                     (fac_0 = 1, fac_m = 2)
 *                 = P(a2) - S(a2),  F_k(N - a2) = P(a2) + S(a2),  P = sum fac Re w^(m a2) Re C,  S = sum fac Im w^(m a2) Im C
 *
 * so the alpha transform costs half of a plain one-sided DFT.  Per cell and q that is 961 x ~7 900 DFMA = 15 MFLOP
 * against 29 791 x 1 512 x 2 = 90 MFLOP for the list form with four a2 per thread (and 303 MFLOP point by point).
 * The summation order differs from the list form, so the two agree to rounding, not to the bit.
 *
 * grid (pair tiles of TPB / LP, b2, q); LP (4 or 8) lanes per pair; lane j of a pair takes l = m + j, m + j + LP, ...
 * and the a2 pairs {j + 1, j + 1 + LP, ...} of 1 .. (N-1)/2 (lane 0 also a2 = 0).
 * X[sxs_x_index(b2 * N^3 + (a2 N + g1) N + g2, qnum, q, k)] = const_k[q] + 2 F_k. */
/* LP lanes share a pair (they split l and a2), JMAX = a2 pairs per lane = ceil(L / LP): template parameters so that
 * P, S stay in registers (with one L = 40 bound for every L the kernel spilled 992 bytes per thread and ran at the
 * list form's speed).  TPB threads = TPB / LP pairs per block. */
template <int LP, int SXS_DENSE_JMAX, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
k_cross_dense(int L, int qnum, int b1, int slab0, const double2 *__restrict__ At, const double2 *__restrict__ St,
              const double2 *__restrict__ tw, const double *__restrict__ cst, double *__restrict__ X)
{
	/* shared memory: twiddles [N] | two stages of { S rows of one m: [3][L + 1][NP], A entries of one m: [3][L + 1][GA] }.
	 * The operands of m + 1 are copied (cp.async, whole 16-byte elements, coalesced rows) while m is computed: with
	 * the lanes loading them straight from global memory half of all stall samples sat on the long scoreboard and the
	 * L1 pipe was 74 % busy (r2v ncu, 47 ms per z); each S row is now fetched once per block instead of once per lane. */
	extern __shared__ double2 s_dense[];
	const int N = 2 * L + 1, ML = sxs_ml_count(L), NP = sxs_row_pad(N), H = (N - 1) / 2;
	const int GA = (TPB / LP + N - 1) / N + 1; /* g1 values the block's TPB / LP consecutive pairs can span */
	const int nlmax = L + 1;
	/* rows of the S stage NP + 2 (LP = 4) or NP + 1 (LP = 8) elements apart, so that the rows a quarter warp reads (the
	 * LP lanes of a pair read the same g2 of LP consecutive rows) fall on disjoint banks; measured: no effect on the
	 * kernel's time (41.2 -> 42.9 ms per z, r2z), the 105 M bank conflicts per launch of the unpadded form are not what
	 * bounds it */
	const int NPS = NP + (LP == 4 ? 2 : 1);
	double2 *s_tw = s_dense;
	const int stage_elems = 3 * nlmax * NPS + 3 * nlmax * GA;
	double2 *stage0 = s_dense + ((N + 1) & ~1);
	for (int i = threadIdx.x; i < N; i += blockDim.x) {
		s_tw[i] = tw[i];
	}
	const int b2 = blockIdx.y, q = blockIdx.z;
	const int lane4 = threadIdx.x % LP;
	const int pair0 = blockIdx.x * (TPB / LP);
	const int pair = pair0 + (threadIdx.x / LP);
	const bool live = pair < N * N;
	const int g1base = min(pair0 / N, N - 1);
	const int g1 = live ? pair / N : g1base, g2 = live ? pair % N : 0;
	const int ga = g1 - g1base; /* 0 .. GA-1 */
	const size_t cstride = (size_t)ML * NP;
	const double2 *a_base = At + (((size_t)b1 * qnum + q) * 3) * cstride;
	const double2 *s_base = St + (((size_t)(slab0 + b2) * qnum + q) * 3) * cstride;

	/* copy of the operands of one m into a stage: S rows row0 .. row0 + nl - 1 (contiguous, NP elements each, padding
	 * included) of the three components, and the A entries g1base .. g1base + GA - 1 of the same rows */
	auto stage_fill = [&](double2 *stg, int m, int row0) {
		const int nl = L + 1 - m;
		const int s_cnt = nl * NP;
		for (int i = threadIdx.x; i < 3 * s_cnt; i += TPB) {
			const int c = i / s_cnt, r = i - c * s_cnt;
			const int rl = r / NP, col = r - rl * NP;
			const unsigned dst = (unsigned)__cvta_generic_to_shared(stg + c * nlmax * NPS + rl * NPS + col);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(s_base + c * cstride + (size_t)row0 * NP + r) : "memory");
		}
		const int a_cnt = nl * GA;
		for (int i = threadIdx.x; i < 3 * a_cnt; i += TPB) {
			const int c = i / a_cnt, r = i - c * a_cnt;
			const int rl = r / GA, gi = r - rl * GA;
			const int gg = min(g1base + gi, NP - 1);
			const unsigned dst = (unsigned)__cvta_generic_to_shared(stg + 3 * nlmax * NPS + c * nlmax * GA + r);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(a_base + c * cstride + (size_t)(row0 + rl) * NP + gg) : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};

	double P[SXS_DENSE_JMAX][6], S[SXS_DENSE_JMAX][6], F0[6];
	int kk[SXS_DENSE_JMAX]; /* (m * a2) mod N for this lane's a2 values */
#pragma unroll
	for (int j = 0; j < SXS_DENSE_JMAX; j++) {
		kk[j] = 0;
#pragma unroll
		for (int c = 0; c < 6; c++) {
			P[j][c] = 0.0; S[j][c] = 0.0;
		}
	}
#pragma unroll
	for (int c = 0; c < 6; c++) {
		F0[c] = 0.0;
	}
	int row0 = 0; /* packed index of (m, m) */
	stage_fill(stage0, 0, 0);
	for (int m = 0; m <= L; m++) {
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads(); /* the copies of m have landed for every thread, and everybody is done with m - 1 */
		const double2 *stg = stage0 + (m & 1) * stage_elems;
		if (m < L) {
			stage_fill(stage0 + ((m + 1) & 1) * stage_elems, m + 1, row0 + L + 1 - m);
		}
		const double2 *sS = stg + g2, *sA = stg + 3 * nlmax * NPS + ga;
		double2 C[6];
#pragma unroll
		for (int c = 0; c < 6; c++) {
			C[c] = make_double2(0.0, 0.0);
		}
		for (int r = lane4; r <= L - m; r += LP) {
			const double2 av = sA[r * GA], ad = sA[nlmax * GA + r * GA], aw = sA[2 * nlmax * GA + r * GA];
			const double2 sv = sS[r * NPS], sd = sS[nlmax * NPS + r * NPS], sw = sS[2 * nlmax * NPS + r * NPS];
			cmac(C[0], av, sv);
			cmac(C[1], av, sd); cmac(C[1], ad, sv);
			cmac(C[2], av, sw); cmac(C[2], aw, sv);
			cmac(C[3], ad, sd);
			cmac(C[4], ad, sw); cmac(C[4], aw, sd);
			cmac(C[5], aw, sw);
		}
		row0 += L + 1 - m;
#pragma unroll
		for (int c = 0; c < 6; c++) {
			C[c].x += __shfl_xor_sync(0xffffffffu, C[c].x, 1);
			C[c].y += __shfl_xor_sync(0xffffffffu, C[c].y, 1);
			C[c].x += __shfl_xor_sync(0xffffffffu, C[c].x, 2);
			C[c].y += __shfl_xor_sync(0xffffffffu, C[c].y, 2);
			if (LP == 8) {
				C[c].x += __shfl_xor_sync(0xffffffffu, C[c].x, 4);
				C[c].y += __shfl_xor_sync(0xffffffffu, C[c].y, 4);
			}
		}
		const double fac = (m == 0) ? 1.0 : 2.0;
		if (lane4 == 0) {
#pragma unroll
			for (int c = 0; c < 6; c++) {
				F0[c] += fac * C[c].x;
			}
		}
#pragma unroll
		for (int j = 0; j < SXS_DENSE_JMAX; j++) {
			const int a2 = 1 + lane4 + LP * j;
			if (a2 <= H) {
				const double2 w = s_tw[kk[j]];
				const double wx = fac * w.x, wy = fac * w.y;
#pragma unroll
				for (int c = 0; c < 6; c++) {
					P[j][c] = fma(wx, C[c].x, P[j][c]);
					S[j][c] = fma(wy, C[c].y, S[j][c]);
				}
				kk[j] += a2;
				if (kk[j] >= N) {
					kk[j] -= N;
				}
			}
		}
	}
	if (!live) {
		return;
	}
	double c6[6];
#pragma unroll
	for (int c = 0; c < 6; c++) {
		c6[c] = cst[c * qnum + q];
	}
	const long long cell_base = (long long)b2 * N * N * N + (long long)g1 * N + g2;
	if (lane4 == 0) {
		double2 *xo = reinterpret_cast<double2 *>(X + sxs_x_index(cell_base, qnum, q, 0));
		xo[0] = make_double2(c6[0] + 2.0 * F0[0], c6[1] + 2.0 * F0[1]);
		xo[1] = make_double2(c6[2] + 2.0 * F0[2], c6[3] + 2.0 * F0[3]);
		xo[2] = make_double2(c6[4] + 2.0 * F0[4], c6[5] + 2.0 * F0[5]);
	}
#pragma unroll
	for (int j = 0; j < SXS_DENSE_JMAX; j++) {
		const int a2 = 1 + lane4 + LP * j;
		if (a2 <= H) {
			double2 *xa = reinterpret_cast<double2 *>(X + sxs_x_index(cell_base + (long long)a2 * N * N, qnum, q, 0));
			double2 *xb = reinterpret_cast<double2 *>(X + sxs_x_index(cell_base + (long long)(N - a2) * N * N, qnum, q, 0));
			xa[0] = make_double2(c6[0] + 2.0 * (P[j][0] - S[j][0]), c6[1] + 2.0 * (P[j][1] - S[j][1]));
			xa[1] = make_double2(c6[2] + 2.0 * (P[j][2] - S[j][2]), c6[3] + 2.0 * (P[j][3] - S[j][3]));
			xa[2] = make_double2(c6[4] + 2.0 * (P[j][4] - S[j][4]), c6[5] + 2.0 * (P[j][5] - S[j][5]));
			xb[0] = make_double2(c6[0] + 2.0 * (P[j][0] + S[j][0]), c6[1] + 2.0 * (P[j][1] + S[j][1]));
			xb[1] = make_double2(c6[2] + 2.0 * (P[j][2] + S[j][2]), c6[3] + 2.0 * (P[j][3] + S[j][3]));
			xb[2] = make_double2(c6[4] + 2.0 * (P[j][4] + S[j][4]), c6[5] + 2.0 * (P[j][5] + S[j][5]));
		}
	}
}

/* res[p*4 + 0..2] -> three arrays */
__global__ void k_res_split(const double *__restrict__ res, long long n, double *__restrict__ s, double *__restrict__ c1,
                            double *__restrict__ c2)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		s[i] = res[i * 4 + 0];
		c1[i] = res[i * 4 + 1];
		c2[i] = res[i * 4 + 2];
	}
}

/* ------------------------------------------------------------ dense scan with top-k */

__global__ void k_scan_indices(long long base, long long n, long long *__restrict__ idx)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		idx[i] = base + i;
	}
}

/* merge keys: [0, k) the running best, [k, k + n) the new chunk; NaN sorts last */
__global__ void k_scan_keys(const double *__restrict__ best, int k, const double *__restrict__ chunk, long long n,
                            double *__restrict__ keys, unsigned int *__restrict__ vals)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= k + n) {
		return;
	}
	double v = i < k ? best[i] : chunk[i - k];
	if (!(v == v)) {
		v = INFINITY;
	}
	keys[i] = v;
	vals[i] = (unsigned int)i;
}

__global__ void k_scan_take(const unsigned int *__restrict__ vals_sorted, const double *__restrict__ keys_sorted, int k,
                            const long long *__restrict__ old_idx, const double *__restrict__ old3,
                            const long long *__restrict__ new_idx, const double *__restrict__ s,
                            const double *__restrict__ c1, const double *__restrict__ c2, long long *__restrict__ out_idx,
                            double *__restrict__ out3)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= k) {
		return;
	}
	const unsigned int src = vals_sorted[j];
	if (src < (unsigned int)k) {
		out_idx[j] = old_idx[src];
		out3[j] = old3[src]; out3[k + j] = old3[k + src]; out3[2 * k + j] = old3[2 * k + src];
	} else {
		const unsigned int r = src - (unsigned int)k;
		out_idx[j] = isinf(keys_sorted[j]) ? -1 : new_idx[r];
		out3[j] = keys_sorted[j]; out3[k + j] = c1[r]; out3[2 * k + j] = c2[r];
	}
}

/* Every grid point (b1, b2, a2, g1, g2) of the z steps [z_lo, z_hi) is scored, one (z, b1) row of cells per pass,
 * and the k points of lowest chi are kept.  The reference's skip = 0 mode computes the same dense set per cell
 * (src/fftsaxs.c:867-872) but can only report rows of a list; this returns the best rows themselves.
 * Out: flat 64-bit indices (-1 when fewer than k points exist), chi ascending, c1, c2. */
extern "C" int sxs_cuda_plan_scan_topk(sxs_cuda_plan *p, int z_lo, int z_hi, int k, long long *index, double *scores,
                                       double *c1, double *c2)
{
	SXS_CK(cudaSetDevice(p->device));
	if (k < 1 || k > (1 << 24)) {
		sxs_cuda_set_error("scan_topk: k out of range");
		return -1;
	}
	if (z_lo < 0) z_lo = 0;
	if (z_hi > p->znum) z_hi = p->znum;
	const long long nb = p->nb, N = p->N;
	const long long per_row = nb * N * N * N; /* points of one (z, b1): all b2, a2, g1, g2 */
	cudaStream_t st = 0;
	long long *d_idx = NULL, *d_best_idx[2] = {NULL, NULL};
	double *d_out = NULL, *d_best3[2] = {NULL, NULL}, *d_keys = NULL, *d_keys_sorted = NULL;
	unsigned int *d_vals = NULL, *d_vals_sorted = NULL;
	void *d_tmp = NULL;
	size_t tmp_bytes = 0;
	const long long m = per_row + k;
	int rc = -1;
	cudaError_t e = cudaSuccess;
#define TK(call) do { if (e == cudaSuccess) e = (call); } while (0)
	TK(cudaMalloc(&d_idx, sizeof(long long) * per_row));
	TK(cudaMalloc(&d_out, sizeof(double) * 3 * per_row));
	TK(cudaMalloc(&d_keys, sizeof(double) * m));
	TK(cudaMalloc(&d_keys_sorted, sizeof(double) * m));
	TK(cudaMalloc(&d_vals, sizeof(unsigned int) * m));
	TK(cudaMalloc(&d_vals_sorted, sizeof(unsigned int) * m));
	for (int b = 0; b < 2; b++) {
		TK(cudaMalloc(&d_best_idx[b], sizeof(long long) * k));
		TK(cudaMalloc(&d_best3[b], sizeof(double) * 3 * k));
	}
	if (e == cudaSuccess) {
		e = cub::DeviceRadixSort::SortPairs(NULL, tmp_bytes, d_keys, d_keys_sorted, d_vals, d_vals_sorted, (int)m);
	}
	TK(cudaMalloc(&d_tmp, tmp_bytes));
	if (e == cudaSuccess) {
		/* running best starts as k empty slots: chi = +inf, index -1 */
		double *h3 = (double *)malloc(sizeof(double) * 3 * k);
		long long *hi = (long long *)malloc(sizeof(long long) * k);
		for (int j = 0; j < k; j++) { h3[j] = INFINITY; h3[k + j] = 0.0; h3[2 * k + j] = 0.0; hi[j] = -1; }
		TK(cudaMemcpy(d_best3[0], h3, sizeof(double) * 3 * k, cudaMemcpyHostToDevice));
		TK(cudaMemcpy(d_best_idx[0], hi, sizeof(long long) * k, cudaMemcpyHostToDevice));
		free(h3); free(hi);
	}
	int cur = 0;
	long long launches = 0, points = 0;
	if (e == cudaSuccess) {
		rc = 0;
		/* SXS_SCAN_LIST=1: every point goes through the list path (sort, distinct points, k_cross<4>) as in round 1 */
		const int dense = getenv("SXS_SCAN_LIST") == NULL;
		const int L = p->L, Q = p->qnum;
		if (dense) {
			const size_t slab_elems = (size_t)Q * 3 * sxs_ml_count(L) * sxs_row_pad((int)N);
			if (ensure(&p->d_St, &p->cap_St, slab_elems * nb)) rc = -1;
			if (rc == 0 && ensure(&p->d_T, &p->cap_T, (size_t)Q * nb * nb * nb)) rc = -1;
			if (rc == 0 && ensure(&p->d_X, &p->cap_X, (size_t)per_row * 6 * Q)) rc = -1;
			if (rc == 0 && ensure(&p->d_res, &p->cap_res, (size_t)per_row * 4)) rc = -1;
			if (rc == 0 && (p->d_slab_flag == NULL || p->cap_slab < nb + 1)) {
				if (p->d_slab_flag) cudaFree(p->d_slab_flag);
				p->d_slab_flag = NULL;
				if (cudaMalloc(&p->d_slab_flag, sizeof(int) * (nb + 1)) != cudaSuccess) rc = -1;
				p->cap_slab = (int)nb + 1;
			}
		}
		for (int z = z_lo; z < z_hi && rc == 0; z++) {
			if (dense) {
				/* every (z, b2) slab of this z is translated once and serves all b1 */
				int *h_flag = (int *)malloc(sizeof(int) * (nb + 1));
				for (int i = 0; i < nb; i++) h_flag[i] = 1;
				h_flag[nb] = z;
				e = cudaMemcpyAsync(p->d_slab_flag, h_flag, sizeof(int) * (nb + 1), cudaMemcpyHostToDevice, st);
				if (e == cudaSuccess) e = cudaStreamSynchronize(st);
				free(h_flag);
				if (e != cudaSuccess) { rc = -1; break; }
				timer_begin(p, 1, st);
				rc = launch_translate(p, 1, p->d_slab_flag + nb, st);
				timer_end(p, 1, st);
				launches += 2;
			}
			for (int b1 = 0; b1 < nb && rc == 0; b1++) {
				const long long base = ((long long)z * nb + b1) * per_row;
				k_scan_indices<<<(unsigned)((per_row + 255) / 256), 256, 0, st>>>(base, per_row, d_idx);
				if (dense) {
					timer_begin(p, 2, st);
					{
						const int NPd = sxs_row_pad((int)N);
						size_t sh = 0; /* set per instantiation: the A stage depends on the pairs per block */
						const unsigned npair = (unsigned)(N * N);
#define SXS_DENSE_LAUNCH(LP, J, TPB, MINB)                                                                           \
	do {                                                                                                             \
		const int ga_ = (int)(((TPB / LP) + N - 1) / N + 1);                                                         \
		sh = sizeof(double2) * ((size_t)((N + 1) & ~1) + 2 * (size_t)(3 * (L + 1) * (NPd + 2) + 3 * (L + 1) * ga_)); \
		if (sh > 48 * 1024) {                                                                                        \
			cudaFuncSetAttribute(k_cross_dense<LP, J, TPB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); \
		}                                                                                                            \
		k_cross_dense<LP, J, TPB, MINB><<<dim3((npair + (TPB / LP) - 1) / (TPB / LP), (unsigned)nb, (unsigned)Q), TPB, sh, st>>>( \
		    L, Q, b1, 0, p->d_At, p->d_St, p->d_tw, p->d_const, p->d_X);                                              \
	} while (0)
						const int variant = getenv("SXS_DENSE_VARIANT") ? atoi(getenv("SXS_DENSE_VARIANT")) : 0; /* tuning */
						if (L <= 8) SXS_DENSE_LAUNCH(4, 2, 256, 2);
						else if (L <= 16 && variant == 1) SXS_DENSE_LAUNCH(4, 4, 256, 1);
						else if (L <= 16 && variant == 2) SXS_DENSE_LAUNCH(8, 2, 256, 2);
						else if (L <= 16) SXS_DENSE_LAUNCH(4, 4, 192, 2);
						else if (L <= 32) SXS_DENSE_LAUNCH(8, 4, 192, 2);
						else if (L <= 48) SXS_DENSE_LAUNCH(8, 6, 256, 1);
						else { sxs_cuda_set_error("dense scan: L = %d beyond 48", L); rc = -1; break; }
#undef SXS_DENSE_LAUNCH
					}
					if (cudaGetLastError() != cudaSuccess) { rc = -1; break; }
					timer_end(p, 2, st);
					timer_begin(p, 3, st);
					rc = sxs_launch_fit(p->d_X, per_row, p->d_a, p->d_qvals, Q, p->mult, p->peak, 1, p->d_res, p->d_ticket, st);
					timer_end(p, 3, st);
					if (rc != 0) break;
					k_res_split<<<(unsigned)((per_row + 255) / 256), 256, 0, st>>>(p->d_res, per_row, d_out, d_out + per_row, d_out + 2 * per_row);
					launches += 4;
				} else {
					rc = score_core<long long>(p, d_idx, per_row, z, z + 1, d_out, d_out + per_row, d_out + 2 * per_row, NULL,
					                           NULL, st, NULL);
					if (rc != 0) break;
					launches += p->stats[2] + 4;
				}
				points += per_row;
				k_scan_keys<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(d_best3[cur], k, d_out, per_row, d_keys, d_vals);
				e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys_sorted, d_vals, d_vals_sorted, (int)m, 0,
				                                    64, st);
				if (e != cudaSuccess) { rc = -1; break; }
				k_scan_take<<<(k + 255) / 256, 256, 0, st>>>(d_vals_sorted, d_keys_sorted, k, d_best_idx[cur], d_best3[cur],
				                                            d_idx, d_out, d_out + per_row, d_out + 2 * per_row,
				                                            d_best_idx[cur ^ 1], d_best3[cur ^ 1]);
				cur ^= 1;
			}
		}
		if (rc == 0) {
			TK(cudaStreamSynchronize(st));
			TK(cudaGetLastError());
			TK(cudaMemcpy(index, d_best_idx[cur], sizeof(long long) * k, cudaMemcpyDeviceToHost));
			TK(cudaMemcpy(scores, d_best3[cur], sizeof(double) * k, cudaMemcpyDeviceToHost));
			TK(cudaMemcpy(c1, d_best3[cur] + k, sizeof(double) * k, cudaMemcpyDeviceToHost));
			TK(cudaMemcpy(c2, d_best3[cur] + 2 * k, sizeof(double) * k, cudaMemcpyDeviceToHost));
		}
	}
#undef TK
	if (e != cudaSuccess) {
		sxs_cuda_set_error("scan_topk: %s", cudaGetErrorString(e));
		rc = -1;
	}
	cudaFree(d_idx); cudaFree(d_out); cudaFree(d_keys); cudaFree(d_keys_sorted); cudaFree(d_vals); cudaFree(d_vals_sorted);
	cudaFree(d_tmp);
	for (int b = 0; b < 2; b++) { cudaFree(d_best_idx[b]); cudaFree(d_best3[b]); }
	if (rc == 0) {
		p->stats[0] = points;
		p->stats[2] = launches;
	}
	return rc;
}

template <typename IndexT>
static int score_host(sxs_cuda_plan *p, const IndexT *index, long long nout, int z_lo, int z_hi, double *scores,
                      double *c1, double *c2)
{
	SXS_CK(cudaSetDevice(p->device));
	if (nout <= 0) {
		return 0;
	}
	{
		unsigned char *tmp = (unsigned char *)p->d_index_in;
		size_t c = p->cap_index_in;
		if (ensure(&tmp, &c, sizeof(IndexT) * (size_t)nout)) return -1;
		p->d_index_in = tmp; p->cap_index_in = c;
	}
	if (ensure(&p->d_out3, &p->cap_out3, (size_t)nout * 3)) return -1;
	SXS_CK(cudaMemcpy(p->d_index_in, index, sizeof(IndexT) * (size_t)nout, cudaMemcpyHostToDevice));
	/* results are scattered to list order on the device: d_out3 = [scores | c1 | c2], nout each */
	double *d_s = p->d_out3, *d_c1 = p->d_out3 + nout, *d_c2 = p->d_out3 + 2 * nout;
	long long nvalid = 0;
	if (score_core<IndexT>(p, (const IndexT *)p->d_index_in, nout, z_lo, z_hi, d_s, d_c1, d_c2, NULL, NULL, 0, &nvalid) != 0) {
		return -1;
	}
	SXS_CK(cudaDeviceSynchronize());
	if (nvalid == 0) {
		return 0;
	}
	if (nvalid == nout) {
		/* every row was scored: three straight copies into the caller's arrays */
		SXS_CK(cudaMemcpy(scores, d_s, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToHost));
		SXS_CK(cudaMemcpy(c1, d_c1, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToHost));
		SXS_CK(cudaMemcpy(c2, d_c2, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToHost));
		return 0;
	}
	/* Partial coverage (a z shard of a multi-GPU call, or rows off the z table): unscored rows must keep the
	 * caller's values and other shards write the same arrays concurrently, so only this call's rows are touched. */
	const size_t need = sizeof(double) * 3 * (size_t)nout + sizeof(unsigned int) * (size_t)nvalid;
	if (need > p->cap_pinned) {
		if (p->h_pinned) cudaFreeHost(p->h_pinned);
		p->h_pinned = NULL;
		p->cap_pinned = 0;
		SXS_CK(cudaMallocHost(&p->h_pinned, need));
		p->cap_pinned = need;
	}
	double *h3 = (double *)p->h_pinned;
	unsigned int *hrows = (unsigned int *)(h3 + 3 * (size_t)nout);
	SXS_CK(cudaMemcpy(h3, p->d_out3, sizeof(double) * 3 * (size_t)nout, cudaMemcpyDeviceToHost));
	SXS_CK(cudaMemcpy(hrows, p->d_rows_sorted, sizeof(unsigned int) * (size_t)nvalid, cudaMemcpyDeviceToHost));
	for (long long i = 0; i < nvalid; i++) {
		const unsigned int r = hrows[i];
		scores[r] = h3[r];
		c1[r] = h3[(size_t)nout + r];
		c2[r] = h3[2 * (size_t)nout + r];
	}
	return 0;
}

/* pinned host memory owned by the plan (grow-only, freed with it): the host layer stages a shard's compact index list
 * and its three result columns here, so that the copies of sxs_cuda_plan_score_* run at the pinned rate */
extern "C" void *sxs_cuda_plan_host_buffer(sxs_cuda_plan *p, size_t bytes)
{
	if (cudaSetDevice(p->device) != cudaSuccess) {
		return NULL;
	}
	if (bytes > p->cap_shard) {
		if (p->h_shard) cudaFreeHost(p->h_shard);
		p->h_shard = NULL;
		p->cap_shard = 0;
		if (cudaHostAlloc(&p->h_shard, bytes + bytes / 8, cudaHostAllocPortable) != cudaSuccess) {
			sxs_cuda_set_error("cudaHostAlloc of %zu bytes failed", bytes);
			p->h_shard = NULL;
			return NULL;
		}
		p->cap_shard = bytes + bytes / 8;
	}
	return p->h_shard;
}

extern "C" int sxs_cuda_plan_score_i32(sxs_cuda_plan *p, const int *index, long long nout, int z_lo, int z_hi,
                                       double *scores, double *c1, double *c2)
{
	return score_host<int>(p, index, nout, z_lo, z_hi, scores, c1, c2);
}

extern "C" int sxs_cuda_plan_score_i64(sxs_cuda_plan *p, const long long *index, long long nout, int z_lo, int z_hi,
                                       double *scores, double *c1, double *c2)
{
	return score_host<long long>(p, index, nout, z_lo, z_hi, scores, c1, c2);
}

extern "C" int sxs_cuda_plan_cross_terms_i32(sxs_cuda_plan *p, const int *index, long long nout, double *cross)
{
	SXS_CK(cudaSetDevice(p->device));
	if (nout <= 0) {
		return 0;
	}
	int *d_idx = NULL;
	double *d_cross = NULL;
	const size_t nc = (size_t)nout * 6 * p->qnum;
	SXS_CK(cudaMalloc(&d_idx, sizeof(int) * (size_t)nout));
	SXS_CK(cudaMalloc(&d_cross, sizeof(double) * nc));
	SXS_CK(cudaMemset(d_cross, 0, sizeof(double) * nc));
	SXS_CK(cudaMemcpy(d_idx, index, sizeof(int) * (size_t)nout, cudaMemcpyHostToDevice));
	int rc = score_core<int>(p, d_idx, nout, 0, p->znum, NULL, NULL, NULL, NULL, d_cross, 0, NULL);
	if (rc == 0) {
		SXS_CK(cudaDeviceSynchronize());
		SXS_CK(cudaMemcpy(cross, d_cross, sizeof(double) * nc, cudaMemcpyDeviceToHost));
	}
	cudaFree(d_idx);
	cudaFree(d_cross);
	return rc;
}

/* histogram of objective evaluations per fit over the points of the last score call (bin 63 = 63 or more) */
__global__ void k_nfg_hist(const double *__restrict__ res, long long np, unsigned long long *__restrict__ hist)
{
	const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= np) {
		return;
	}
	int n = (int)res[p * 4 + 3];
	if (n > 63) n = 63;
	if (n < 0) n = 0;
	atomicAdd(&hist[n], 1ull);
}

extern "C" int sxs_cuda_plan_fit_evaluations(sxs_cuda_plan *p, long long *hist64)
{
	SXS_CK(cudaSetDevice(p->device));
	const long long np = p->stats[0];
	memset(hist64, 0, sizeof(long long) * 64);
	if (np <= 0 || p->d_res == NULL) {
		return 0;
	}
	unsigned long long *d_h = NULL;
	SXS_CK(cudaMalloc(&d_h, sizeof(unsigned long long) * 64));
	SXS_CK(cudaMemset(d_h, 0, sizeof(unsigned long long) * 64));
	k_nfg_hist<<<(unsigned)((np + 255) / 256), 256>>>(p->d_res, np, d_h);
	SXS_CK_LAUNCH();
	SXS_CK(cudaMemcpy(hist64, d_h, sizeof(long long) * 64, cudaMemcpyDeviceToHost));
	cudaFree(d_h);
	return 0;
}

/* ------------------------------------------------------------ calibration */

/* FP64 FMA rate of this GPU: 8 independent DFMA chains per thread, full occupancy.  MEASURED_PEAKS.json
 * carries no FP64 figure, so bench.py calibrates the roofline denominator of K3/K4 with this. */
__global__ void __launch_bounds__(256)
k_dfma_peak(double *out, int iters, double seed)
{
	double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double b = 1.0000001, c = 1e-9;
#pragma unroll 1
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
			a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
		}
	}
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int sxs_cuda_fp64_peak(int device, double *tflops)
{
	SXS_CK(cudaSetDevice(device));
	cudaDeviceProp prop;
	SXS_CK(cudaGetDeviceProperties(&prop, device));
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
	double *d = NULL;
	SXS_CK(cudaMalloc(&d, sizeof(double) * blocks * threads));
	cudaEvent_t e0, e1;
	SXS_CK(cudaEventCreate(&e0));
	SXS_CK(cudaEventCreate(&e1));
	double best = 0.0;
	for (int rep = 0; rep < 5; rep++) {
		SXS_CK(cudaEventRecord(e0));
		k_dfma_peak<<<blocks, threads>>>(d, iters, 1.0 + rep);
		SXS_CK(cudaEventRecord(e1));
		SXS_CK(cudaEventSynchronize(e1));
		float ms = 0.f;
		SXS_CK(cudaEventElapsedTime(&ms, e0, e1));
		const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
		const double tf = fl / (ms * 1e-3) / 1e12;
		if (rep > 0 && tf > best) best = tf;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d);
	*tflops = best;
	return 0;
}
