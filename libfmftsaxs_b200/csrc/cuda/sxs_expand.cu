/* K1 — spherical-harmonic amplitude expansion on the GPU.
 *
 *   A^c_lm(q) = 4 pi i^l  sum_j f^c_j  j_l(q r_j)  Y*_lm(theta_j, phi_j),   c = vacuum, dummy, water
 *
 * replaces the atom loop of atom_grp2spf_inplace (src/pdb2spf.c:63-116) and its 4 pi i^l pass
 * (:118-147).  Two kernels:
 *   k_ylm     one thread per atom: associated Legendre recurrences (src/borrowed.c:40-65), Y_lm norms,
 *             cos/sin(m phi) -> Ystar[lm][atom]   (layout: atom fastest, so K1's reads coalesce)
 *   k_expand  one block per (q, l) band.  Phase 1: every thread evaluates the reference-exact Bessel
 *             series j_l(q r_j) for its atoms and parks it in shared memory (atoms are staged in
 *             chunks of up to SXS_ATOM_CHUNK).  Phase 2: one warp per m sums f_j * (j_l * Y*_lm) over the
 *             chunk with lanes striding over atoms, finishes with a warp-shuffle tree, and accumulates
 *             the 2l+1 band results in shared memory across chunks.
 * Roofline: FP64 pipe — the Bessel series costs two IEEE divisions per term and dominates; the
 * accumulation is natoms*qnum*(L+1)^2*6 FMAs.  HBM traffic is the Ystar table once per q (L2-resident).
 */
#include <math.h>

#include "sxs_dev.cuh"

#define SXS_ATOM_CHUNK 8192
#define SXS_EXPAND_THREADS 512

/* Truncated ascending series of the reference (src/sfbessel.c:40-60), same operations, same order,
 * round-to-nearest intrinsics so that nothing is contracted.  inv_dfact = 1/(2l+1)!!. */
__device__ __forceinline__ double sxs_sbessel_dev(int l, double x, double inv_dfact)
{
	if (!(x > 0.0)) {
		return l == 0 ? 1.0 : 0.0;
	}
	const double tlp1 = 2.0 * (double)l + 1.0;
	double term = inv_dfact;
	double sum = term;
	const double nx2 = __dmul_rn(__dmul_rn(-1.0, x), x);
	for (int k = 1; fabs(__ddiv_rn(term, sum)) > 0.00001; k++) {
		const double tk = 2.0 * (double)k;
		const double den = __dmul_rn(tk, __dadd_rn(tk, tlp1));
		term = __dmul_rn(term, __ddiv_rn(nx2, den));
		sum = __dadd_rn(sum, term);
	}
	return __dmul_rn(sum, pow(x, (double)l));
}

/* ynorm index of generate_spherical_norm(N): N*(m+N-1)+l */
__device__ __forceinline__ int ynorm_index(int N, int l, int m) { return N * (m + N - 1) + l; }

__global__ void k_ylm(int natoms, int L, const double *__restrict__ cos_theta, const double *__restrict__ phi,
                      const double *__restrict__ ynorm, double2 *__restrict__ ystar)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= natoms) {
		return;
	}
	const int N = L + 1;
	const double x = cos_theta[j];
	const double fi = phi[j];
	const double y = sqrt(__dsub_rn(1.0, __dmul_rn(x, x)));
	const size_t stride = (size_t)natoms;

	double pmm = 1.0; /* P_m^m, starts at P_0^0 */
	for (int m = 0; m <= L; m++) {
		if (m > 0) {
			/* P_m^m = -(2m-1) y P_{m-1}^{m-1} */
			pmm = __dmul_rn(__dmul_rn(__dmul_rn(-1.0, __dsub_rn(2.0 * m, 1.0)), y), pmm);
		}
		const double cp = cos(__dmul_rn(fi, (double)m));          /* cos(phi*m) */
		const double sn = sin(__dmul_rn(-fi, (double)m));         /* sin(-phi*m): conjugate, +m */
		const double cpn = cos(__dmul_rn(fi, (double)(-m)));      /* cos(phi*(-m)) */
		const double snn = sin(__dmul_rn(-fi, (double)(-m)));     /* sin(-phi*(-m)) */

		double p2 = 0.0;  /* P_{l-2}^m */
		double p1 = pmm;  /* P_{l-1}^m */
		for (int l = m; l <= L; l++) {
			double p;
			if (l == m) {
				p = pmm;
			} else if (l == m + 1) {
				/* P_{m+1}^m = (2m+1) x P_m^m; for m = 0 this is the reference's P_1^0 = x */
				p = (m == 0) ? x : __dmul_rn(__dmul_rn(__dadd_rn(2.0 * m, 1.0), x), pmm);
			} else {
				/* (2l-1)/(l-m) x P_{l-1} - (l+m-1)/(l-m) P_{l-2} */
				const double lm = (double)(l - m);
				const double t1 = __dmul_rn(__dmul_rn(__ddiv_rn(__dsub_rn(2.0 * l, 1.0), lm), x), p1);
				const double t2 = __dmul_rn(__ddiv_rn(__dsub_rn((double)(l + m), 1.0), lm), p2);
				p = __dsub_rn(t1, t2);
			}
			p2 = p1;
			p1 = p;

			const double hp = __dmul_rn(ynorm[ynorm_index(N, l, m)], p);
			ystar[(size_t)(l * (l + 1) + m) * stride + j] = make_double2(__dmul_rn(hp, cp), __dmul_rn(hp, sn));
			if (m > 0) {
				const double hn = __dmul_rn(ynorm[ynorm_index(N, l, -m)], p);
				ystar[(size_t)(l * (l + 1) - m) * stride + j] = make_double2(__dmul_rn(hn, cpn), __dmul_rn(hn, snn));
			}
		}
	}
}

__global__ void __launch_bounds__(SXS_EXPAND_THREADS)
k_expand(int natoms, int L, int qnum, const double *__restrict__ r, const double *__restrict__ fv,
         const double *__restrict__ fd, const double *__restrict__ fw, const double *__restrict__ qvals,
         const double *__restrict__ inv_dfact, const double2 *__restrict__ ystar, double four_pi,
         double2 *__restrict__ coef)
{
	extern __shared__ double smem[];
	double *s_bes = smem;                        /* [chunk] */
	double *s_acc = smem + SXS_ATOM_CHUNK;       /* [2l+1][6] running band sums */

	const int q = blockIdx.x;
	const int l = blockIdx.y;
	const double qv = qvals[q];
	const double idf = inv_dfact[l];
	const int nm = 2 * l + 1;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

	for (int i = threadIdx.x; i < nm * 6; i += blockDim.x) {
		s_acc[i] = 0.0;
	}

	for (int base = 0; base < natoms; base += SXS_ATOM_CHUNK) {
		const int cnt = min(SXS_ATOM_CHUNK, natoms - base);
		__syncthreads();
		for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
			s_bes[i] = sxs_sbessel_dev(l, __dmul_rn(qv, r[base + i]), idf);
		}
		__syncthreads();
		for (int mi = warp; mi < nm; mi += nwarps) {
			const int m = mi - l;
			const double2 *yrow = ystar + (size_t)(l * (l + 1) + m) * natoms + base;
			double vr = 0, vi = 0, dr = 0, di = 0, wr = 0, wi = 0;
			for (int i = lane; i < cnt; i += 32) {
				const double2 yv = yrow[i];
				const double b = s_bes[i];
				const double tr = __dmul_rn(b, yv.x), ti = __dmul_rn(b, yv.y);
				const double a = fv[base + i], d = fd[base + i], w = fw[base + i];
				vr += a * tr; vi += a * ti;
				dr += d * tr; di += d * ti;
				wr += w * tr; wi += w * ti;
			}
#pragma unroll
			for (int off = 16; off > 0; off >>= 1) {
				vr += __shfl_down_sync(0xffffffffu, vr, off);
				vi += __shfl_down_sync(0xffffffffu, vi, off);
				dr += __shfl_down_sync(0xffffffffu, dr, off);
				di += __shfl_down_sync(0xffffffffu, di, off);
				wr += __shfl_down_sync(0xffffffffu, wr, off);
				wi += __shfl_down_sync(0xffffffffu, wi, off);
			}
			if (lane == 0) {
				double *acc = s_acc + mi * 6;
				acc[0] += vr; acc[1] += vi; acc[2] += dr; acc[3] += di; acc[4] += wr; acc[5] += wi;
			}
		}
	}
	__syncthreads();

	/* times 4 pi i^l (src/pdb2spf.c:55-57,118-147) */
	const double cre_t[4] = {four_pi, 0.0, -four_pi, 0.0};
	const double cim_t[4] = {0.0, four_pi, 0.0, -four_pi};
	const double cre = cre_t[l % 4], cim = cim_t[l % 4];
	const int lm_n = (L + 1) * (L + 1);
	for (int i = threadIdx.x; i < nm * 3; i += blockDim.x) {
		const int mi = i / 3, c = i % 3;
		const double re = s_acc[mi * 6 + 2 * c], im = s_acc[mi * 6 + 2 * c + 1];
		const int lm = l * (l + 1) + (mi - l);
		coef[((size_t)c * qnum + q) * lm_n + lm] =
		    make_double2(__dsub_rn(__dmul_rn(cre, re), __dmul_rn(cim, im)), __dadd_rn(__dmul_rn(cre, im), __dmul_rn(cim, re)));
	}
}

extern "C" int sxs_cuda_expand(int device, int natoms, const double *r, const double *cos_theta, const double *phi,
                               const double *ff_vacuum, const double *ff_dummy, const double *ff_water,
                               const double *qvals, int qnum, int L, const double *ynorm, const double *inv_dfact,
                               double four_pi, double *coef)
{
	if (natoms <= 0 || qnum <= 0 || L < 0) {
		sxs_cuda_set_error("sxs_cuda_expand: bad sizes");
		return -1;
	}
	SXS_CK(cudaSetDevice(device));
	const int nb = L + 1, lm_n = nb * nb;
	const size_t na = (size_t)natoms;
	double *d_atoms = NULL, *d_q = NULL, *d_yn = NULL, *d_idf = NULL;
	double2 *d_ystar = NULL, *d_coef = NULL;
	SXS_CK(cudaMalloc(&d_atoms, sizeof(double) * na * 6));
	SXS_CK(cudaMalloc(&d_q, sizeof(double) * qnum));
	SXS_CK(cudaMalloc(&d_yn, sizeof(double) * nb * (2 * nb - 1)));
	SXS_CK(cudaMalloc(&d_idf, sizeof(double) * (2 * L + 1)));
	SXS_CK(cudaMalloc(&d_ystar, sizeof(double2) * na * lm_n));
	SXS_CK(cudaMalloc(&d_coef, sizeof(double2) * (size_t)3 * qnum * lm_n));
	const double *src[6] = {r, cos_theta, phi, ff_vacuum, ff_dummy, ff_water};
	for (int i = 0; i < 6; i++) {
		SXS_CK(cudaMemcpy(d_atoms + i * na, src[i], sizeof(double) * na, cudaMemcpyHostToDevice));
	}
	SXS_CK(cudaMemcpy(d_q, qvals, sizeof(double) * qnum, cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_yn, ynorm, sizeof(double) * nb * (2 * nb - 1), cudaMemcpyHostToDevice));
	SXS_CK(cudaMemcpy(d_idf, inv_dfact, sizeof(double) * (2 * L + 1), cudaMemcpyHostToDevice));

	k_ylm<<<(natoms + 127) / 128, 128>>>(natoms, L, d_atoms + na, d_atoms + 2 * na, d_yn, d_ystar);
	SXS_CK_LAUNCH();

	const size_t shmem = sizeof(double) * (SXS_ATOM_CHUNK + (size_t)(2 * L + 1) * 6);
	SXS_CK(cudaFuncSetAttribute(k_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
	dim3 grid(qnum, nb);
	k_expand<<<grid, SXS_EXPAND_THREADS, shmem>>>(natoms, L, qnum, d_atoms, d_atoms + 3 * na, d_atoms + 4 * na,
	                                              d_atoms + 5 * na, d_q, d_idf, d_ystar, four_pi, d_coef);
	SXS_CK_LAUNCH();
	SXS_CK(cudaMemcpy(coef, d_coef, sizeof(double2) * (size_t)3 * qnum * lm_n, cudaMemcpyDeviceToHost));
	cudaFree(d_atoms); cudaFree(d_q); cudaFree(d_yn); cudaFree(d_idf); cudaFree(d_ystar); cudaFree(d_coef);
	return 0;
}
