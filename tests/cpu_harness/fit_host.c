/* TEST INFRASTRUCTURE: compiles the product's host/device fit headers for the CPU so that the
 * optimiser's iterates can be compared with the reference without a GPU.  Never shipped. */
#include "fit_point.h"

static const uint64_t exp_tab[SXS_EXP_TABLE_ENTRIES] = SXS_EXP_TABLE_INIT;

/* x[npts][6][qnum] (same layout as oracle/ref_harness.c:ref_fit) -> out[npts][4].
 * table_exp = 0: exp() is the host libm's; 1: the restatement of exp_glibc.h (what the device runs) */
void cpu_fit_points(const double *x, int npts, const double *a, const double *qvals, int qnum, double mult,
                    double peak, int table_exp, double *out)
{
	for (int p = 0; p < npts; p++) {
		/* re-pack to the point-major device row x[q*6 + c] */
		double buf[6 * 512];
		for (int c = 0; c < 6; c++)
			for (int q = 0; q < qnum; q++)
				buf[q * 6 + c] = x[((long)p * 6 + c) * qnum + q];
		double s, c1, c2;
		int nfg;
		sxs_fit_point_ex(buf, 1, 6, a, qvals, qnum, mult, peak, table_exp ? exp_tab : NULL, &s, &c1, &c2, &nfg);
		out[4 * p] = s; out[4 * p + 1] = c1; out[4 * p + 2] = c2; out[4 * p + 3] = nfg;
	}
}

/* exp_glibc.h against the host libm on n pseudo-random arguments of four magnitudes; returns the mismatches */
long cpu_exp_mismatches(long n)
{
	unsigned long long s = 88172645463325252ull;
	long bad = 0;
	for (long i = 0; i < n; i++) {
		s ^= s << 13; s ^= s >> 7; s ^= s << 17;
		const double u = (double)(s >> 11) / 9007199254740992.0 - 0.5;
		const int k = (int)(i % 4);
		const double x = k == 0 ? u * 0.04 : k == 1 ? u * 1000.0 : k == 2 ? u * 2.0 : u * 1e-15 * (double)(i % 1000);
		const double want = exp(x), got = sxs_exp_glibc(x, exp_tab);
		bad += memcmp(&want, &got, sizeof want) != 0;
	}
	return bad;
}

/* the same fits with the evaluation in the kernel's fast form (sxs_fit_eval_fast: exp core without its range test, G kept
 * for pass 2, quotient through the reciprocal table) */
void cpu_fit_points_fast(const double *x, int npts, const double *a, const double *qvals, int qnum, double mult,
                         double peak, double *out)
{
	double rq[512], dq[512], stash[512];
	for (int i = 0; i < qnum; i++) {
		dq[i] = qvals[i] - (i > 0 ? qvals[i - 1] : -1.0);
		rq[i] = 1.0 / dq[i];
	}
	for (int p = 0; p < npts; p++) {
		double buf[6 * 512];
		for (int c = 0; c < 6; c++)
			for (int q = 0; q < qnum; q++)
				buf[q * 6 + c] = x[((long)p * 6 + c) * qnum + q];
		struct sxs_fit_ctx ctx;
		ctx.x = buf; ctx.stride = 1; ctx.qstride = 6; ctx.a = a; ctx.qvals = qvals; ctx.qnum = qnum; ctx.mult = mult;
		ctx.rq = rq; ctx.dq = dq; ctx.etab = exp_tab; ctx.scale = 1.0;
		ctx.scale = sxs_fit_rescale(&ctx, peak);
		struct lq_state st;
		lq_begin(&st, SXS_C1_DEFAULT, SXS_C2_DEFAULT);
		while (lq_step(&st, SXS_FIT_PGTOL, SXS_FIT_TOL) == LQ_NEED_EVAL) {
			sxs_fit_eval_fast(&ctx, st.x[1], st.x[2], stash, &st.f, &st.g[1], &st.g[2]);
		}
		out[4 * p] = sqrt(st.f); out[4 * p + 1] = st.x[1]; out[4 * p + 2] = st.x[2]; out[4 * p + 3] = st.nfgv;
	}
}

/* exp core (no range test) against the host libm: the fit's range, |x| < 500, tiny and zero arguments */
long cpu_exp_core_mismatches(long n)
{
	unsigned long long s = 1234567890123456789ull;
	long bad = 0;
	for (long i = 0; i < n; i++) {
		s ^= s << 13; s ^= s >> 7; s ^= s << 17;
		const double u = (double)(s >> 11) / 9007199254740992.0 - 0.5;
		const int k = (int)(i % 6);
		double x = k == 0 ? u * 0.04 : k == 1 ? u * 1000.0 : k == 2 ? u * 2.0 : k == 3 ? u * 1e-15 * (double)(i % 1000)
		         : k == 4 ? ldexp(u, -60 - (int)(i % 960)) : (i % 12 == 5 ? 0.0 : -0.0);
		const double want = exp(x), got = sxs_exp_glibc_core(x, exp_tab);
		bad += memcmp(&want, &got, sizeof want) != 0;
	}
	return bad;
}

/* sxs_div_by_recip against the IEEE quotient: n numerators per node spacing of the grid, of the magnitudes the fit sees
 * (differences of intensities: 0, a few ulps of 1e3..1e6, up to 1e7), plus adversarial mantissas near powers of two */
long cpu_div_mismatches(const double *qvals, int qnum, long n)
{
	unsigned long long s = 7777777777777ull;
	long bad = 0;
	for (int i = 0; i < qnum; i++) {
		const double d = qvals[i] - (i > 0 ? qvals[i - 1] : -1.0), r = 1.0 / d;
		for (long j = 0; j < n; j++) {
			s ^= s << 13; s ^= s >> 7; s ^= s << 17;
			const double u = (double)(s >> 11) / 9007199254740992.0 - 0.5;
			const int k = (int)(j % 5);
			double a = k == 0 ? u * 2e7 : k == 1 ? u * 1e-6 : k == 2 ? ldexp(1.0 + ldexp((double)(s & 1023), -52), (int)(s >> 54) - 20)
			         : k == 3 ? ldexp(2.0 - ldexp((double)(1 + (s & 1023)), -52), (int)(s >> 54) - 20) : (j % 10 == 4 ? 0.0 : u * 1e3);
			const double want = a / d, got = sxs_div_by_recip(a, d, r);
			bad += memcmp(&want, &got, sizeof want) != 0 && !(want == 0.0 && got == 0.0);
		}
	}
	return bad;
}
