/* Spherical Bessel function j_l(x), reference-exact.
 *
 * The reference (src/sfbessel.c:40-60) sums the ascending series
 *     j_l(x) = x^l * sum_k (-x^2/2)^k / (k! (2l+2k+1)!!)
 * term by term and stops as soon as |term/sum| <= 1e-5.  For x beyond ~30 that truncated sum is
 * dominated by rounding noise, and the scoring path feeds it x = q*z up to 40, so matching the
 * reference means repeating its IEEE operations in its order — not computing a better j_l.
 * This file is built with -ffp-contract=off; the CUDA twin (sxs_sbessel_dev in sxs_kernels.cuh)
 * spells the same sequence with __dmul_rn/__ddiv_rn/__dadd_rn.
 */
#include "sfbessel.h"

/* (2l+1)!! accumulated upwards in odd steps, like doublefact() in src/sfbessel.c:22-37. */
double sxs_odd_double_factorial(int n)
{
	double res = 1.0;
	for (int i = 1; i <= n; i += 2) {
		res *= (double)i;
	}
	return res;
}

double sxs_sbessel(int l, double x)
{
	if (!(x > 0.0)) {
		return l == 0 ? 1.0 : 0.0;
	}
	const double tlp1 = 2.0 * (double)l + 1.0;
	double term = 1.0 / sxs_odd_double_factorial((int)tlp1);
	double sum = term;
	for (int k = 1; fabs(term / sum) > 0.00001; k++) {
		term *= (-1.0) * x * x / (2.0 * (double)k * (2.0 * (double)k + tlp1));
		sum += term;
	}
	return sum * pow(x, l);
}
