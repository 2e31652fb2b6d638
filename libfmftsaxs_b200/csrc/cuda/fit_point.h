/* One complete (c1, c2) fit for one grid point: peak rescale (src/min_saxs.c:170-188), bounded
 * minimisation (src/min_saxs.c:196-259) and the outputs the reference keeps (:250-255).
 * Box and start: src/define.h:28-34.  factr/pgtol: src/min_saxs.c:224-225. */
#ifndef SXS_FIT_POINT_H
#define SXS_FIT_POINT_H

#include "fit_eval.h"
#ifdef SXS_FIT_GENERIC_LBFGSB
#include "lbfgsb_n2m3.h" /* the memory-resident restatement with the vendored code's loop structure (host comparisons) */
#endif

#define SXS_C1_LOWER 0.96
#define SXS_C1_UPPER 1.04
#define SXS_C2_LOWER (-2.00)
#define SXS_C2_UPPER 4.00
#define SXS_C1_DEFAULT 1.0
#define SXS_C2_DEFAULT 0.0
#define LQ_L1 SXS_C1_LOWER
#define LQ_U1 SXS_C1_UPPER
#define LQ_L2 SXS_C2_LOWER
#define LQ_U2 SXS_C2_UPPER
#include "lbfgsb_lean.h"
/* factr * epsmch (lbfgsb.c mainlb) and pgtol, src/min_saxs.c:224-225 */
#define SXS_FIT_TOL (1e+7 * LQ_EPSMCH)
#define SXS_FIT_PGTOL 1e-5

/* peak / I(0) of sxs_fit_params, in the reference's summation order (src/min_saxs.c:170-179) */
SXS_HD double sxs_fit_rescale(const struct sxs_fit_ctx *ctx_unit, double peak)
{
	const double i0 = SXS_X(ctx_unit, 0, SXS_VV) + SXS_X(ctx_unit, 0, SXS_DD) + SXS_X(ctx_unit, 0, SXS_WW) +
	                  SXS_X(ctx_unit, 0, SXS_VW) - SXS_X(ctx_unit, 0, SXS_VD) - SXS_X(ctx_unit, 0, SXS_DW);
	return peak / i0;
}

/* Serial form (one fit start to finish); the CUDA kernel interleaves lb_step() and the evaluation
 * across the lanes of a warp instead, see k_fit in sxs_exact.cu. */
SXS_HD void sxs_fit_point_ex(const double *x, long stride, long qstride, const double *a, const double *qvals, int qnum,
                             double mult, double peak, const uint64_t *etab, double *score, double *c1, double *c2, int *nfg)
{
	struct sxs_fit_ctx ctx;
	ctx.x = x;
	ctx.stride = stride;
	ctx.qstride = qstride;
	ctx.a = a;
	ctx.qvals = qvals;
	ctx.qnum = qnum;
	ctx.mult = mult;
	ctx.rq = NULL; /* the serial form divides in the loop */
	ctx.dq = NULL;
	ctx.etab = etab; /* NULL on the host: libm's exp itself */
	ctx.scale = 1.0;
	ctx.scale = sxs_fit_rescale(&ctx, peak);

#ifdef SXS_FIT_GENERIC_LBFGSB
	struct lb_state st;
	lb_begin(&st, SXS_C1_DEFAULT, SXS_C2_DEFAULT, SXS_C1_LOWER, SXS_C1_UPPER, SXS_C2_LOWER, SXS_C2_UPPER, 1e+7);
	while (lb_step(&st, 1e-5) == LB_NEED_EVAL) {
		SXS_FIT_EVAL(&ctx, st.x[1], st.x[2], &st.f, &st.g[1], &st.g[2]);
	}
#else
	struct lq_state st;
	lq_begin(&st, SXS_C1_DEFAULT, SXS_C2_DEFAULT);
	while (lq_step(&st, SXS_FIT_PGTOL, SXS_FIT_TOL) == LQ_NEED_EVAL) {
		SXS_FIT_EVAL(&ctx, st.x[1], st.x[2], &st.f, &st.g[1], &st.g[2]);
	}
#endif
	*score = sqrt(st.f);
	*c1 = st.x[1];
	*c2 = st.x[2];
	*nfg = st.nfgv;
}

SXS_HD void sxs_fit_point(const double *x, long stride, long qstride, const double *a, const double *qvals, int qnum,
                          double mult, double peak, double *score, double *c1, double *c2, int *nfg)
{
	sxs_fit_point_ex(x, stride, qstride, a, qvals, qnum, mult, peak, NULL, score, c1, c2, nfg);
}

#endif
