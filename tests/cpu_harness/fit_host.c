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
