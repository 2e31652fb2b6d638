/* TEST INFRASTRUCTURE — the objective of the (c1, c2) fit and the fit driver for the oracle port:
 * sxs_best_scale (src/min_saxs.c:261-319), gradient (src/min_saxs.c:3-105), sxs_fit_params' peak rescale
 * (:170-188) and the sxs_lbfgs_fitting loop (:196-259).  Written against the port's own copy of the optimiser. */
#ifndef ORACLE_PORT_FIT_H
#define ORACLE_PORT_FIT_H
#include <math.h>
#include "lbfgsb_port.h"

struct port_ctx {
	const double *x; long stride, qstride; const double *a, *q; int qnum; double mult, scale;
};
#define PX(c, i, k) ((c)->x[(long)(i) * (c)->qstride + (long)(k) * (c)->stride] * (c)->scale)

static inline double port_best_scale(const struct port_ctx *c, double c1, double c2)
{
	const double *a = c->a, *q = c->q;
	double corr = -c->mult * (c1 * c1 - 1.0);
	double G = c1 * c1 * c1 * exp(corr * q[0] * q[0]);
	double in_prev = PX(c, 0, 0) - G * PX(c, 0, 1) + c2 * PX(c, 0, 2) + G * G * PX(c, 0, 3) - G * c2 * PX(c, 0, 4) + c2 * c2 * PX(c, 0, 5);
	double c1_cube = c1 * c1 * c1, q_prev = -1.0, up = 0.0, down = 0.0;
	for (int i = 0; i < c->qnum; i++) {
		double q_cur = q[i];
		G = c1_cube * exp(corr * q_cur * q_cur);
		double in = PX(c, i, 0) - G * PX(c, i, 1) + c2 * PX(c, i, 2) + G * G * PX(c, i, 3) - G * c2 * PX(c, i, 4) + c2 * c2 * PX(c, i, 5);
		double tan = (in - in_prev) / (q_cur - q_prev);
		double buf = in - tan * q_cur;
		up += buf * a[i * 6 + 1] + tan * a[i * 6 + 2];
		down += buf * buf * a[i * 6 + 3] + 2.0 * tan * buf * a[i * 6 + 4] + tan * tan * a[i * 6 + 5];
		in_prev = in;
		q_prev = q_cur;
	}
	return up / down;
}

static inline void port_eval(const struct port_ctx *c, double c1, double c2, double *f, double *g0, double *g1)
{
	const double k = port_best_scale(c, c1, c2);
	const double *a = c->a, *q = c->q;
	const double mult = c->mult;
	double grad0 = 0.0, grad1 = 0.0;
	double corr = -mult * (c1 * c1 - 1.0);
	double G = c1 * c1 * c1 * exp(corr * q[0] * q[0]);
	double G_der = G * (3.0 / c1 - 2.0 * c1 * mult * q[0] * q[0]);
	double in_prev = PX(c, 0, 0) - G * PX(c, 0, 1) + c2 * PX(c, 0, 2) + G * G * PX(c, 0, 3) - G * c2 * PX(c, 0, 4) + c2 * c2 * PX(c, 0, 5);
	double d1_prev = -G_der * PX(c, 0, 1) + 2.0 * G * G_der * PX(c, 0, 3) - G_der * c2 * PX(c, 0, 4);
	double d2_prev = PX(c, 0, 2) - G * PX(c, 0, 4) + 2.0 * c2 * PX(c, 0, 5);
	double q_prev = -1.0, score = 0.0, c1_cube = c1 * c1 * c1;
	for (int i = 0; i < c->qnum; i++) {
		double q_cur = q[i];
		G = c1_cube * exp(corr * q_cur * q_cur);
		G_der = G * (3.0 / c1 - 2.0 * c1 * mult * q_cur * q_cur);
		double VV = PX(c, i, 0), VD = PX(c, i, 1), VW = PX(c, i, 2), DD = PX(c, i, 3), DW = PX(c, i, 4), WW = PX(c, i, 5);
		double in = VV - G * VD + c2 * VW + G * G * DD - G * c2 * DW + c2 * c2 * WW;
		double d1 = G_der * (-VD + 2.0 * G * DD - c2 * DW);
		double d2 = VW - G * DW + 2.0 * c2 * WW;
		double buf = 1.0 / (q_cur - q_prev);
		double tan = (in - in_prev) * buf;
		double t1 = (d1 - d1_prev) * buf;
		double t2 = (d2 - d2_prev) * buf;
		grad0 += 2.0 * k * (-(d1 - t1 * q_cur) * a[i * 6 + 1] - t1 * a[i * 6 + 2] +
		                    k * ((in - tan * q_cur) * (d1 - t1 * q_cur) * a[i * 6 + 3] +
		                         (in * t1 + d1 * tan - 2.0 * tan * t1 * q_cur) * a[i * 6 + 4] + tan * t1 * a[i * 6 + 5]));
		grad1 += 2.0 * k * (-(d2 - t2 * q_cur) * a[i * 6 + 1] - t2 * a[i * 6 + 2] +
		                    k * ((in - tan * q_cur) * (d2 - t2 * q_cur) * a[i * 6 + 3] +
		                         (in * t2 + d2 * tan - 2.0 * tan * t2 * q_cur) * a[i * 6 + 4] + tan * t2 * a[i * 6 + 5]));
		buf = in - tan * q_cur;
		score += a[i * 6] + k * (-2.0 * buf * a[i * 6 + 1] - 2.0 * tan * a[i * 6 + 2] +
		                         k * (buf * buf * a[i * 6 + 3] + 2.0 * buf * tan * a[i * 6 + 4] + tan * tan * a[i * 6 + 5]));
		in_prev = in; d1_prev = d1; d2_prev = d2; q_prev = q_cur;
	}
	*g0 = grad0; *g1 = grad1; *f = score;
}

static inline void port_fit_point(const double *x, long stride, long qstride, const double *a, const double *qvals,
                                  int qnum, double mult, double peak, double *score, double *c1, double *c2, int *nfg)
{
	struct port_ctx c = {x, stride, qstride, a, qvals, qnum, mult, 1.0};
	double i0 = PX(&c, 0, 0) + PX(&c, 0, 3) + PX(&c, 0, 5) + PX(&c, 0, 2) - PX(&c, 0, 1) - PX(&c, 0, 4); /* :170-177 */
	c.scale = peak / i0;
	struct lb_state st;
	lb_begin(&st, 1.0, 0.0, 0.96, 1.04, -2.00, 4.00, 1e+7);   /* src/define.h:28-34, src/min_saxs.c:224 */
	while (lb_step(&st, 1e-5) == LB_NEED_EVAL) {
		port_eval(&c, st.x[1], st.x[2], &st.f, &st.g[1], &st.g[2]);
	}
	*score = sqrt(st.f); *c1 = st.x[1]; *c2 = st.x[2]; *nfg = st.nfgv;
}
#endif
