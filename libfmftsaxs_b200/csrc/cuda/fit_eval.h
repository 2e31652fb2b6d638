/* chi^2(c1, c2) objective of the per-conformation fit and its analytic gradient.
 *
 * Follows src/min_saxs.c operation for operation: sxs_best_scale (:261-319) gives the optimal linear
 * scale k for the trial (c1, c2), gradient() (:3-105) then accumulates f and df/dc with k frozen.
 * Both walk the q nodes once with a piecewise-linear model between nodes (q_{-1} = -1).
 * The six cross terms are read through two strides (between terms, between q nodes) so that the same
 * code serves a point-major row (stride 1, qstride 6) and a point-minor column; the peak rescale of sxs_fit_params (:170-188) is applied
 * on the fly as x*scale, which rounds exactly like the reference's in-place `*= scale`.
 *
 * Compiled with -fmad=false (see lbfgsb_n2m3.h); plain C so the CPU tests can include it.
 */
#ifndef SXS_FIT_EVAL_H
#define SXS_FIT_EVAL_H

#include <math.h>
#include <stddef.h>

#include "exp_glibc.h"

#ifndef SXS_HD
#ifdef __CUDACC__
#define SXS_HD __host__ __device__ __forceinline__
#else
#define SXS_HD static inline
#endif
#endif

struct sxs_fit_ctx {
	const double *x;      /* cross terms of this point: x[q*qstride + c*stride] */
	long stride;          /* between the six terms of one q */
	long qstride;         /* between q nodes */
	const double *a;      /* compressed experiment, a[q*6 + 0..5] (src/min_saxs.c:353-389) */
	const double *qvals;
	int qnum;
	double mult;          /* (4pi/3)^(3/2) rm^2 / (16 pi), src/min_saxs.c:121 */
	double scale;         /* peak / I(0) rescale, src/min_saxs.c:170-179 */
	const double *rq;     /* optional table rq[i] = 1/(q_i - q_{i-1}), q_{-1} = -1 (same IEEE quotient as in the loop) */
	const double *dq;     /* optional table dq[i] = q_i - q_{i-1}, q_{-1} = -1 (the difference the loops form) */
	const uint64_t *etab; /* 2^(k/128) table of exp_glibc.h; NULL on the host = call libm's exp itself */
};

/* exp() of the objective: the reference's libm algorithm, bit for bit (exp_glibc.h) */
SXS_HD double sxs_fit_exp(const struct sxs_fit_ctx *ctx, double x)
{
#ifdef __CUDA_ARCH__
	return sxs_exp_glibc(x, ctx->etab);
#else
	return ctx->etab ? sxs_exp_glibc(x, ctx->etab) : exp(x);
#endif
}

enum { SXS_VV = 0, SXS_VD, SXS_VW, SXS_DD, SXS_DW, SXS_WW };

#ifndef SXS_XLD1
#define SXS_XLD1(p) (*(p))
#endif
#define SXS_X(ctx, q, c) (SXS_XLD1(&(ctx)->x[(long)(q) * (ctx)->qstride + (long)(c) * (ctx)->stride]) * (ctx)->scale)

/* The six scaled terms of node q.  With SXS_ROWMAJOR_VEC (CUDA build, point-major rows x[q*6 + c],
 * 16-byte aligned) they come in as three 16-byte loads; the values and their use are identical. */
#if defined(SXS_ROWMAJOR_VEC) && defined(__CUDA_ARCH__)
#ifndef SXS_XLD
#define SXS_XLD(p) (*(p))
#endif
#define SXS_LOAD6(ctx, q, vv, vd, vw, dd, dw, ww)                                       \
	do {                                                                               \
		const double2 *r_ = reinterpret_cast<const double2 *>((ctx)->x + (long)(q) * 6); \
		const double2 p0_ = SXS_XLD(r_), p1_ = SXS_XLD(r_ + 1), p2_ = SXS_XLD(r_ + 2);   \
		vv = p0_.x * (ctx)->scale; vd = p0_.y * (ctx)->scale;                            \
		vw = p1_.x * (ctx)->scale; dd = p1_.y * (ctx)->scale;                            \
		dw = p2_.x * (ctx)->scale; ww = p2_.y * (ctx)->scale;                            \
	} while (0)
#else
#define SXS_LOAD6(ctx, q, vv, vd, vw, dd, dw, ww)                                       \
	do {                                                                               \
		vv = SXS_X(ctx, q, SXS_VV); vd = SXS_X(ctx, q, SXS_VD); vw = SXS_X(ctx, q, SXS_VW); \
		dd = SXS_X(ctx, q, SXS_DD); dw = SXS_X(ctx, q, SXS_DW); ww = SXS_X(ctx, q, SXS_WW); \
	} while (0)
#endif

/* ---- the objective node by node ------------------------------------------------------------------------------
 * Both passes of the reference walk the q nodes with a handful of running values.  They are written here as
 * "begin at node 0" + "advance by one node" steps on the six scaled cross terms of that node, so that the serial form
 * below (pointer + strides) and the kernel's streamed form (terms arriving through a shared-memory ring) execute the
 * same expressions in the same order. */
struct sxs_six {
	double vv, vd, vw, dd, dw, ww;
};

/* pass 1, src/min_saxs.c:261-319 */
struct sxs_scale_run {
	double corr, c1_cube, c2;
	double in_prev, q_prev, up, down;
};

SXS_HD void sxs_scale_begin(struct sxs_scale_run *r, const struct sxs_fit_ctx *ctx, double c1, double c2,
                            const struct sxs_six *x0)
{
	const double *q = ctx->qvals;
	r->corr = -ctx->mult * (c1 * c1 - 1.0);
	const double G = c1 * c1 * c1 * sxs_fit_exp(ctx, r->corr * q[0] * q[0]);
	r->in_prev = x0->vv - G * x0->vd + c2 * x0->vw + G * G * x0->dd - G * c2 * x0->dw + c2 * c2 * x0->ww;
	r->c1_cube = c1 * c1 * c1;
	r->c2 = c2;
	r->q_prev = -1.0;
	r->up = 0.0;
	r->down = 0.0;
}

/* a / d through r = RN(1 / d): two residual corrections with fused multiply-adds.  After the first, q is a faithful
 * quotient (its error is below one ulp whatever the two roundings of a * r did); the second is then Markstein's
 * correction step, which returns RN(a / d) exactly when r is the correctly rounded reciprocal (it is: the tables hold
 * the IEEE quotient 1 / d).  Holds for a = 0 and for normal a, d, a / d (no overflow, no denormals) — the fit's
 * (in - in_prev) / (q_i - q_{i-1}); tests/test_cpu_host.py checks it against the IEEE quotient on 10^8 operands.
 * No branch, 5 instructions against ~13 + a slow-path test for the division nvcc emits. */
SXS_HD double sxs_div_by_recip(double a, double d, double r)
{
	double q = SXS_EXP_MUL(a, r);
	double e = SXS_EXP_FMA(-d, q, a);
	q = SXS_EXP_FMA(e, r, q);
	e = SXS_EXP_FMA(-d, q, a);
	return SXS_EXP_FMA(e, r, q);
}

/* G(c1, q_i) = c1^3 exp(corr q_i^2) is the same number in both passes (and in begin() for node 0): the node steps take
 * it as a parameter so that a caller may form it once.  recip: (in - in_prev) / dq[i] through sxs_div_by_recip. */
SXS_HD void sxs_scale_begin_g(struct sxs_scale_run *r, const struct sxs_fit_ctx *ctx, double c1, double c2,
                              const struct sxs_six *x0, double G)
{
	r->corr = -ctx->mult * (c1 * c1 - 1.0);
	r->in_prev = x0->vv - G * x0->vd + c2 * x0->vw + G * G * x0->dd - G * c2 * x0->dw + c2 * c2 * x0->ww;
	r->c1_cube = c1 * c1 * c1;
	r->c2 = c2;
	r->q_prev = -1.0;
	r->up = 0.0;
	r->down = 0.0;
}

SXS_HD void sxs_scale_node_g(struct sxs_scale_run *r, const struct sxs_fit_ctx *ctx, int i, const struct sxs_six *x,
                             double G, const int recip)
{
	const double *a = ctx->a;
	const double c2 = r->c2;
	const double q_cur = ctx->qvals[i];
	const double in = x->vv - G * x->vd + c2 * x->vw + G * G * x->dd - G * c2 * x->dw + c2 * c2 * x->ww;
	double tan;
	if (recip) {
		tan = sxs_div_by_recip(in - r->in_prev, ctx->dq[i], ctx->rq[i]);
	} else {
#if defined(__CUDA_ARCH__)
		tan = (in - r->in_prev) / ctx->dq[i];
#else
		tan = (in - r->in_prev) / (ctx->dq ? ctx->dq[i] : q_cur - r->q_prev);
#endif
	}
	const double buf = in - tan * q_cur;

	r->up += buf * a[i * 6 + 1] + tan * a[i * 6 + 2];
	r->down += buf * buf * a[i * 6 + 3] + 2.0 * tan * buf * a[i * 6 + 4] + tan * tan * a[i * 6 + 5];

	r->in_prev = in;
	r->q_prev = q_cur;
}

SXS_HD void sxs_scale_node(struct sxs_scale_run *r, const struct sxs_fit_ctx *ctx, int i, const struct sxs_six *x)
{
	const double q_cur = ctx->qvals[i];
	const double G = r->c1_cube * sxs_fit_exp(ctx, r->corr * q_cur * q_cur);
	sxs_scale_node_g(r, ctx, i, x, G, 0);
}

/* pass 2, src/min_saxs.c:3-105 with k frozen */
struct sxs_grad_run {
	double corr, c1, c1_cube, c2, k;
	double three_over_c1, two_c1_mult; /* the node-independent factors of dG/dc1 = G (3/c1 - 2 c1 mult q^2) */
	double in_prev, in_der_c1_prev, in_der_c2_prev, q_prev;
	double grad0, grad1, score;
};

SXS_HD void sxs_grad_begin(struct sxs_grad_run *r, const struct sxs_fit_ctx *ctx, double c1, double c2, double k,
                           const struct sxs_six *x0)
{
	const double *q = ctx->qvals;
	const double mult = ctx->mult;
	r->corr = -mult * (c1 * c1 - 1.0);
	const double G = c1 * c1 * c1 * sxs_fit_exp(ctx, r->corr * q[0] * q[0]);
	r->three_over_c1 = 3.0 / c1;
	r->two_c1_mult = 2.0 * c1 * mult;
	const double G_der = G * (r->three_over_c1 - r->two_c1_mult * q[0] * q[0]);
	r->in_prev = x0->vv - G * x0->vd + c2 * x0->vw + G * G * x0->dd - G * c2 * x0->dw + c2 * c2 * x0->ww;
	r->in_der_c1_prev = -G_der * x0->vd + 2.0 * G * G_der * x0->dd - G_der * c2 * x0->dw;
	r->in_der_c2_prev = x0->vw - G * x0->dw + 2.0 * c2 * x0->ww;
	r->c1 = c1;
	r->c1_cube = c1 * c1 * c1;
	r->c2 = c2;
	r->k = k;
	r->q_prev = -1.0;
	r->grad0 = 0.0;
	r->grad1 = 0.0;
	r->score = 0.0;
}

SXS_HD void sxs_grad_node_g(struct sxs_grad_run *r, const struct sxs_fit_ctx *ctx, int i, const struct sxs_six *x, double G)
{
	const double *a = ctx->a;
	const double c2 = r->c2, k = r->k;
	const double q_cur = ctx->qvals[i];
	const double G_der = G * (r->three_over_c1 - r->two_c1_mult * q_cur * q_cur);

	const double in = x->vv - G * x->vd + c2 * x->vw + G * G * x->dd - G * c2 * x->dw + c2 * c2 * x->ww;
	const double in_der_c1 = G_der * (-x->vd + 2.0 * G * x->dd - c2 * x->dw);
	const double in_der_c2 = x->vw - G * x->dw + 2.0 * c2 * x->ww;

	/* 1 / (q_i - q_{i-1}): from the table when there is one (the same IEEE quotient, computed once per block) */
#if defined(__CUDA_ARCH__)
	double buf = ctx->rq[i];
#else
	double buf = ctx->rq ? ctx->rq[i] : 1.0 / (q_cur - r->q_prev);
#endif
	const double tan = (in - r->in_prev) * buf;
	const double tan_c1_der = (in_der_c1 - r->in_der_c1_prev) * buf;
	const double tan_c2_der = (in_der_c2 - r->in_der_c2_prev) * buf;

	const double a1 = a[i * 6 + 1], a2 = a[i * 6 + 2], a3 = a[i * 6 + 3], a4 = a[i * 6 + 4], a5 = a[i * 6 + 5];

	r->grad0 += 2.0 * k * (-(in_der_c1 - tan_c1_der * q_cur) * a1 - tan_c1_der * a2 +
	                       k * ((in - tan * q_cur) * (in_der_c1 - tan_c1_der * q_cur) * a3 +
	                            (in * tan_c1_der + in_der_c1 * tan - 2.0 * tan * tan_c1_der * q_cur) * a4 +
	                            tan * tan_c1_der * a5));

	r->grad1 += 2.0 * k * (-(in_der_c2 - tan_c2_der * q_cur) * a1 - tan_c2_der * a2 +
	                       k * ((in - tan * q_cur) * (in_der_c2 - tan_c2_der * q_cur) * a3 +
	                            (in * tan_c2_der + in_der_c2 * tan - 2.0 * tan * tan_c2_der * q_cur) * a4 +
	                            tan * tan_c2_der * a5));

	buf = in - tan * q_cur;
	r->score += a[i * 6] + k * (-2.0 * buf * a1 - 2.0 * tan * a2 +
	                            k * (buf * buf * a3 + 2.0 * buf * tan * a4 + tan * tan * a5));

	r->in_prev = in;
	r->in_der_c1_prev = in_der_c1;
	r->in_der_c2_prev = in_der_c2;
	r->q_prev = q_cur;
}

SXS_HD void sxs_grad_node(struct sxs_grad_run *r, const struct sxs_fit_ctx *ctx, int i, const struct sxs_six *x)
{
	const double q_cur = ctx->qvals[i];
	const double G = r->c1_cube * sxs_fit_exp(ctx, r->corr * q_cur * q_cur);
	sxs_grad_node_g(r, ctx, i, x, G);
}

SXS_HD void sxs_grad_begin_g(struct sxs_grad_run *r, const struct sxs_fit_ctx *ctx, double c1, double c2, double k,
                             const struct sxs_six *x0, double G)
{
	const double *q = ctx->qvals;
	const double mult = ctx->mult;
	r->corr = -mult * (c1 * c1 - 1.0);
	r->three_over_c1 = 3.0 / c1;
	r->two_c1_mult = 2.0 * c1 * mult;
	const double G_der = G * (r->three_over_c1 - r->two_c1_mult * q[0] * q[0]);
	r->in_prev = x0->vv - G * x0->vd + c2 * x0->vw + G * G * x0->dd - G * c2 * x0->dw + c2 * c2 * x0->ww;
	r->in_der_c1_prev = -G_der * x0->vd + 2.0 * G * G_der * x0->dd - G_der * c2 * x0->dw;
	r->in_der_c2_prev = x0->vw - G * x0->dw + 2.0 * c2 * x0->ww;
	r->c1 = c1;
	r->c1_cube = c1 * c1 * c1;
	r->c2 = c2;
	r->k = k;
	r->q_prev = -1.0;
	r->grad0 = 0.0;
	r->grad1 = 0.0;
	r->score = 0.0;
}

/* src/min_saxs.c:261-319 */
SXS_HD double sxs_fit_best_scale(const struct sxs_fit_ctx *ctx, double c1, double c2)
{
	struct sxs_six x;
	struct sxs_scale_run r;
	SXS_LOAD6(ctx, 0, x.vv, x.vd, x.vw, x.dd, x.dw, x.ww);
	sxs_scale_begin(&r, ctx, c1, c2, &x);
	for (int i = 0; i < ctx->qnum; i++) {
		SXS_LOAD6(ctx, i, x.vv, x.vd, x.vw, x.dd, x.dw, x.ww);
		sxs_scale_node(&r, ctx, i, &x);
	}
	return r.up / r.down;
}

/* src/min_saxs.c:3-105 with k = sxs_best_scale(c1, c2) as the driver loop sets it (:233-236). */
SXS_HD void sxs_fit_eval(const struct sxs_fit_ctx *ctx, double c1, double c2, double *f, double *g0, double *g1)
{
	const double k = sxs_fit_best_scale(ctx, c1, c2);
	struct sxs_six x;
	struct sxs_grad_run r;
	SXS_LOAD6(ctx, 0, x.vv, x.vd, x.vw, x.dd, x.dw, x.ww);
	sxs_grad_begin(&r, ctx, c1, c2, k, &x);
	for (int i = 0; i < ctx->qnum; i++) {
		SXS_LOAD6(ctx, i, x.vv, x.vd, x.vw, x.dd, x.dw, x.ww);
		sxs_grad_node(&r, ctx, i, &x);
	}
	*g0 = r.grad0;
	*g1 = r.grad1;
	*f = r.score;
}

/* The evaluation as the kernel's streamed objective performs it (sxs_exact.cu:fit_eval_fast): G formed once per node
 * with the branch-free exp core and kept for pass 2 (stash[0 .. qnum-1]), the pass-1 quotient through the reciprocal
 * table.  Same numbers as sxs_fit_eval bit for bit whenever |corr q^2| < 512 (the exp core's domain; the launcher
 * checks the bound over the box of c1).  Needs ctx->rq, ctx->dq, ctx->etab. */
SXS_HD void sxs_fit_eval_fast(const struct sxs_fit_ctx *ctx, double c1, double c2, double *stash, double *f, double *g0,
                              double *g1)
{
	const double *q = ctx->qvals;
	const double corr = -ctx->mult * (c1 * c1 - 1.0);
	const double c1_cube = c1 * c1 * c1;
	struct sxs_six x;
	struct sxs_scale_run sr;
	for (int i = 0; i < ctx->qnum; i++) {
		const double G = c1_cube * sxs_exp_glibc_core(corr * q[i] * q[i], ctx->etab);
		stash[i] = G;
		SXS_LOAD6(ctx, i, x.vv, x.vd, x.vw, x.dd, x.dw, x.ww);
		if (i == 0) {
			sxs_scale_begin_g(&sr, ctx, c1, c2, &x, G);
		}
		sxs_scale_node_g(&sr, ctx, i, &x, G, 1);
	}
	const double k = sr.up / sr.down;
	struct sxs_grad_run gr;
	for (int i = 0; i < ctx->qnum; i++) {
		SXS_LOAD6(ctx, i, x.vv, x.vd, x.vw, x.dd, x.dw, x.ww);
		if (i == 0) {
			sxs_grad_begin_g(&gr, ctx, c1, c2, k, &x, stash[0]);
		}
		sxs_grad_node_g(&gr, ctx, i, &x, stash[i]);
	}
	*g0 = gr.grad0;
	*g1 = gr.grad1;
	*f = gr.score;
}

/* the evaluation the serial fit uses.  (Round 1 also had a one-pass form that accumulated the coefficients of the
 * polynomial in k instead of walking q twice: algebraically identical, 27 % cheaper, but its rounding left four times as
 * many rows beyond 1e-6 in c2 as the reference's own FMA build does: removed.) */
#define SXS_FIT_EVAL(ctx, c1, c2, f, g0, g1) sxs_fit_eval(ctx, c1, c2, f, g0, g1)

#endif /* SXS_FIT_EVAL_H */
