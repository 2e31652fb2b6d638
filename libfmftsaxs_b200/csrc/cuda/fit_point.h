/* One complete (c1, c2) fit for one grid point: peak rescale (src/min_saxs.c:170-188), bounded
 * minimisation (src/min_saxs.c:196-259) and the outputs the reference keeps (:250-255).
 * Box and start: src/define.h:28-34.  factr/pgtol: src/min_saxs.c:224-225. */
#ifndef SXS_FIT_POINT_H
#define SXS_FIT_POINT_H

#include "fit_eval.h"
#include "lbfgsb_n2m3.h"

#define SXS_C1_LOWER 0.96
#define SXS_C1_UPPER 1.04
#define SXS_C2_LOWER (-2.00)
#define SXS_C2_UPPER 4.00
#define SXS_C1_DEFAULT 1.0
#define SXS_C2_DEFAULT 0.0

SXS_HD void sxs_fit_point(const double *x, long stride, const double *a, const double *qvals, int qnum,
                          double mult, double peak, double *score, double *c1, double *c2, int *nfg)
{
	struct sxs_fit_ctx ctx;
	ctx.x = x;
	ctx.stride = stride;
	ctx.a = a;
	ctx.qvals = qvals;
	ctx.qnum = qnum;
	ctx.mult = mult;
	ctx.scale = 1.0;
	double i0 = SXS_X(&ctx, 0, SXS_VV) + SXS_X(&ctx, 0, SXS_DD) + SXS_X(&ctx, 0, SXS_WW) + SXS_X(&ctx, 0, SXS_VW) -
	            SXS_X(&ctx, 0, SXS_VD) - SXS_X(&ctx, 0, SXS_DW);
	ctx.scale = peak / i0;

	struct lb_state st;
	st.x[1] = SXS_C1_DEFAULT; st.x[2] = SXS_C2_DEFAULT;
	st.l[1] = SXS_C1_LOWER;   st.l[2] = SXS_C2_LOWER;
	st.u[1] = SXS_C1_UPPER;   st.u[2] = SXS_C2_UPPER;
	st.g[1] = 0.0; st.g[2] = 0.0;
	st.f = 0.0;
	/* the reference zeroes its whole workspace before every fit (src/min_saxs.c:217-221) */
	for (int i = 0; i <= LB_N; i++) {
		for (int j = 0; j <= LB_M; j++) { st.ws[i][j] = 0.0; st.wy[i][j] = 0.0; }
		st.z[i] = st.r[i] = st.d[i] = st.t[i] = st.xp[i] = 0.0;
		st.index[i] = st.iwhere[i] = st.indx2[i] = 0;
	}
	for (int i = 0; i <= LB_M; i++) {
		for (int j = 0; j <= LB_M; j++) { st.sy[i][j] = 0.0; st.ss[i][j] = 0.0; st.wt[i][j] = 0.0; }
	}
	for (int i = 0; i <= LB_M2; i++) {
		for (int j = 0; j <= LB_M2; j++) { st.wn[i][j] = 0.0; st.wn1[i][j] = 0.0; }
	}
	for (int i = 0; i <= 8 * LB_M; i++) { st.wa[i] = 0.0; }
	st.brackt = 0; st.stage = 0; st.ls_task = LS_START;
	st.ginit = st.gtest = st.gx = st.gy = st.finit = st.fx = st.fy = 0.0;
	st.stx = st.sty = st.stmin = st.stmax = st.width = st.width1 = 0.0;

	LB_MINIMIZE(&st, LB_EVAL(&ctx, &st), 1e+7, 1e-5);

	*score = sqrt(st.f);
	*c1 = st.x[1];
	*c2 = st.x[2];
	*nfg = st.nfgv;
}

#endif
