/* TEST INFRASTRUCTURE (oracle build only).
 * Minimal gmp.h: the image ships the GMP 6.3 runtime (/lib/x86_64-linux-gnu/libgmp.so.10)
 * but not its header.  Declares the mpf ABI struct and the 11 mpf entry points that
 * /root/reference/src/borrowed.c:73-222 and src/fftsaxs.c:138-174 call. */
#ifndef ORACLE_SHIM_GMP_H
#define ORACLE_SHIM_GMP_H

typedef unsigned long int mp_limb_t;
typedef long int mp_exp_t;

typedef struct {
	int _mp_prec;
	int _mp_size;
	mp_exp_t _mp_exp;
	mp_limb_t *_mp_d;
} __mpf_struct;

typedef __mpf_struct mpf_t[1];
typedef __mpf_struct *mpf_ptr;
typedef const __mpf_struct *mpf_srcptr;

#define mpf_init __gmpf_init
#define mpf_init_set_d __gmpf_init_set_d
#define mpf_mul_ui __gmpf_mul_ui
#define mpf_mul __gmpf_mul
#define mpf_div __gmpf_div
#define mpf_add __gmpf_add
#define mpf_set_d __gmpf_set_d
#define mpf_sqrt __gmpf_sqrt
#define mpf_set __gmpf_set
#define mpf_clear __gmpf_clear
#define mpf_get_d __gmpf_get_d

void mpf_init(mpf_ptr);
void mpf_init_set_d(mpf_ptr, double);
void mpf_mul_ui(mpf_ptr, mpf_srcptr, unsigned long int);
void mpf_mul(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_div(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_add(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_set_d(mpf_ptr, double);
void mpf_sqrt(mpf_ptr, mpf_srcptr);
void mpf_set(mpf_ptr, mpf_srcptr);
void mpf_clear(mpf_ptr);
double mpf_get_d(mpf_srcptr);

#endif
