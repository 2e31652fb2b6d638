/* exp() as the reference's libm computes it.
 *
 * The objective of the (c1, c2) fit calls exp() once or twice per q node (src/min_saxs.c:24,40,273,283).  The
 * minimiser stops mid-convergence, so a last-bit difference in exp() can move a line search by one evaluation
 * (0.3 % of the fits with CUDA's exp(), measured in round 1).  glibc >= 2.28 — what the reference links on any
 * current Linux — evaluates double-precision exp with Szabolcs Nagy's table-driven algorithm (published with the
 * ARM optimized routines): x = k ln2/128 + r, exp(x) = 2^(k/128) (1 + tail_k + r + r^2 (C2 + r C3) + r^4 (C4 + r C5)).
 * On x86-64 hosts with FMA (every current server CPU) libm selects the FMA build of that routine; the operation
 * sequence below is the one that build executes (read from its disassembly: which products are fused matters to the
 * last bit), restated with explicit fma()/IEEE operations so that neither nvcc nor gcc may re-associate it.
 *
 * Valid for 2^-54 <= |x| < 512 plus the tiny-|x| branch (1 + x); the fit's arguments are |x| < 0.02.  Outside that
 * range the caller's libm/CUDA exp is used (never reached on this path).  tests/test_cpu_host.py compares this
 * restatement with the host libm bit for bit over 20 M arguments; tests/test_gpu_parity.py does the same for the
 * device build.
 */
#ifndef SXS_EXP_GLIBC_H
#define SXS_EXP_GLIBC_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "exp_table.h"

#ifndef SXS_HD
#ifdef __CUDACC__
#define SXS_HD __host__ __device__ __forceinline__
#else
#define SXS_HD static inline
#endif
#endif

/* N / ln2, ln2 / N split in two, and the rounding shift 1.5 * 2^52 */
#define SXS_EXP_INVLN2N 0x1.71547652b82fep+7
#define SXS_EXP_SHIFT 0x1.8p52
#define SXS_EXP_NEGLN2HIN (-0x1.62e42fefa0000p-8)
#define SXS_EXP_NEGLN2LON (-0x1.cf79abc9e3b3ap-47)
#define SXS_EXP_C2 0x1.ffffffffffdbdp-2
#define SXS_EXP_C3 0x1.555555555543cp-3
#define SXS_EXP_C4 0x1.55555cf172b91p-5
#define SXS_EXP_C5 0x1.1111167a4d017p-7

#if defined(__CUDA_ARCH__)
#define SXS_EXP_FMA(a, b, c) __fma_rn(a, b, c)
#define SXS_EXP_MUL(a, b) __dmul_rn(a, b)
#define SXS_EXP_ADD(a, b) __dadd_rn(a, b)
#define SXS_EXP_BITS(d) ((uint64_t)__double_as_longlong(d))
#define SXS_EXP_DBL(u) __longlong_as_double((long long)(u))
#else
#define SXS_EXP_FMA(a, b, c) fma(a, b, c)
#define SXS_EXP_MUL(a, b) ((a) * (b))
#define SXS_EXP_ADD(a, b) ((a) + (b))
static inline uint64_t sxs_exp_bits_(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double sxs_exp_dbl_(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
#define SXS_EXP_BITS(d) sxs_exp_bits_(d)
#define SXS_EXP_DBL(u) sxs_exp_dbl_(u)
#endif

/* |x| >= 512, inf, nan: never reached by the fit (|x| < 0.02); a real call on the device so that the math library's
 * exp is not expanded inside the objective's loops (instruction-cache footprint) */
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ double sxs_exp_out_of_range(double x) { return exp(x); }
#else
static inline double sxs_exp_out_of_range(double x) { return exp(x); }
#endif

/* The routine past its range test: valid for |x| < 512.  For |x| < 2^-54 (where the full routine returns 1 + x to keep
 * the exception flags clean) it also returns 1.0: k = 0, r = x, and scale + scale * (x + ...) rounds to 1 — so a caller
 * that knows |x| < 512 may call it for every argument and has no branch at all.
 * tab: the SXS_EXP_TABLE_ENTRIES words of exp_table.h (shared memory on the device) */
SXS_HD double sxs_exp_glibc_core(double x, const uint64_t *tab)
{
	double kd = SXS_EXP_FMA(x, SXS_EXP_INVLN2N, SXS_EXP_SHIFT);
	const uint64_t ki = SXS_EXP_BITS(kd);
	kd = SXS_EXP_ADD(kd, -SXS_EXP_SHIFT);
	double r = SXS_EXP_FMA(kd, SXS_EXP_NEGLN2HIN, x);
	r = SXS_EXP_FMA(kd, SXS_EXP_NEGLN2LON, r);
	const uint32_t idx = 2u * (uint32_t)(ki & 127u);
	const double tail = SXS_EXP_DBL(tab[idx]);
	const uint64_t sbits = tab[idx + 1] + (ki << 45);
	const double p23 = SXS_EXP_FMA(r, SXS_EXP_C3, SXS_EXP_C2);
	const double tr = SXS_EXP_ADD(r, tail);
	const double r2 = SXS_EXP_MUL(r, r);
	const double p45 = SXS_EXP_FMA(r, SXS_EXP_C5, SXS_EXP_C4);
	const double s1 = SXS_EXP_FMA(p23, r2, tr);
	const double r4 = SXS_EXP_MUL(r2, r2);
	const double tmp = SXS_EXP_FMA(r4, p45, s1);
	const double scale = SXS_EXP_DBL(sbits);
	return SXS_EXP_FMA(scale, tmp, scale);
}

SXS_HD double sxs_exp_glibc(double x, const uint64_t *tab)
{
	const uint32_t abstop = (uint32_t)(SXS_EXP_BITS(x) >> 52) & 0x7ff;
	if (abstop - 0x3c9u >= 0x3fu) {
		if (abstop < 0x3c9u) {
			return SXS_EXP_ADD(1.0, x); /* |x| < 2^-54, including +-0 */
		}
		return sxs_exp_out_of_range(x);
	}
	return sxs_exp_glibc_core(x, tab);
}

#endif /* SXS_EXP_GLIBC_H */
