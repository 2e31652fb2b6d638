/* TEST INFRASTRUCTURE (oracle build only).
 * Minimal fftw3.h: FFTW3 is not installed.  Provides the five calls of
 * /root/reference/src/fftsaxs.c:678-683,900,974-976 over a plain separable DFT
 * (fftw_shim.c).  Only rank 3, equal sizes, unit stride, out-of-place, FORWARD. */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H

#include <complex.h>
#include <stddef.h>

typedef double _Complex fftw_complex;
typedef struct oracle_fftw_plan *fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_PATIENT (1U << 5)

void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed, int istride, int idist,
                             fftw_complex *out, const int *onembed, int ostride, int odist,
                             int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#endif
