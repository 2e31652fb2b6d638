/* TEST INFRASTRUCTURE: compiles the product's host/device fit headers for the CPU so that the
 * optimiser's iterates can be compared with the reference without a GPU.  Never shipped. */
#include "fit_point.h"

/* x[npts][6][qnum] (same layout as oracle/ref_harness.c:ref_fit) -> out[npts][4] */
void cpu_fit_points(const double *x, int npts, const double *a, const double *qvals, int qnum, double mult,
                    double peak, double *out)
{
	for (int p = 0; p < npts; p++) {
		/* re-pack to the point-major device row x[q*6 + c] */
		double buf[6 * 512];
		for (int c = 0; c < 6; c++)
			for (int q = 0; q < qnum; q++)
				buf[q * 6 + c] = x[((long)p * 6 + c) * qnum + q];
		double s, c1, c2;
		int nfg;
		sxs_fit_point(buf, 1, 6, a, qvals, qnum, mult, peak, &s, &c1, &c2, &nfg);
		out[4 * p] = s; out[4 * p + 1] = c1; out[4 * p + 2] = c2; out[4 * p + 3] = nfg;
	}
}
