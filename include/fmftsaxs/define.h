/* Compile-time constants of the scoring path; values as src/define.h:11-34 of the reference. */
#ifndef FMFTSAXS_DEFINE_H
#define FMFTSAXS_DEFINE_H

#define SQRTPI 1.77245385091

/* The reference builds with -std=c11, under which glibc does not expose M_PI, so its own
 * 11-digit fallback is what every formula of the path actually uses (src/define.h:13-15).
 * SXS_PI is that value under a name that cannot be shadowed by <math.h>. */
#define SXS_PI 3.14159265358
#ifndef M_PI
#define M_PI SXS_PI
#endif

#define QMAX 0.5
#define QNUM 50

#define REL_ERR 0.05

#define MINIMIZER_ITERMAX 1000
#define VERBOSE_LBFGS -1

#define C1_LOWER 0.96
#define C1_UPPER 1.04
#define C2_LOWER -2.00
#define C2_UPPER 4.00

#define C1_DEFAULT 1.0
#define C2_DEFAULT 0.0

#endif
