/* gmp.h -- hand-written declarations of the GMP 6 entry points the reference's host sources use, so that they compile
 * in an image that ships the GMP runtime (libgmp.so.10) but no development header.  TEST INFRASTRUCTURE (oracle/): used
 * only to build oracle/_ref/libref_host.so from the reference sources where they lie.  Types and names follow the
 * documented GMP 6 ABI (GMP manual, "Internals"): nothing here is copied from GMP's own header. */
#ifndef FS_ORACLE_GMP_SHIM_H
#define FS_ORACLE_GMP_SHIM_H
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned long mp_limb_t;
typedef long mp_exp_t;
typedef long mp_size_t;
typedef unsigned long mp_bitcnt_t;
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; } __mpz_struct;
typedef struct { int _mp_prec; int _mp_size; mp_exp_t _mp_exp; mp_limb_t *_mp_d; } __mpf_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpf_struct mpf_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;
typedef __mpf_struct *mpf_ptr;
typedef const __mpf_struct *mpf_srcptr;
#define __GNU_MP_VERSION 6
#define __GNU_MP_VERSION_MINOR 3
#define __GNU_MP_VERSION_PATCHLEVEL 0
#define GMP_LIMB_BITS 64
#define GMP_NUMB_BITS 64

#define FS_GMP(name) __g##name
#define mpf_init __gmpf_init
#define mpf_init2 __gmpf_init2
#define mpf_clear __gmpf_clear
#define mpf_set __gmpf_set
#define mpf_set_d __gmpf_set_d
#define mpf_set_ui __gmpf_set_ui
#define mpf_set_si __gmpf_set_si
#define mpf_set_str __gmpf_set_str
#define mpf_set_z __gmpf_set_z
#define mpf_set_prec __gmpf_set_prec
#define mpf_set_prec_raw __gmpf_set_prec_raw
#define mpf_get_prec __gmpf_get_prec
#define mpf_get_d __gmpf_get_d
#define mpf_get_d_2exp __gmpf_get_d_2exp
#define mpf_get_si __gmpf_get_si
#define mpf_get_ui __gmpf_get_ui
#define mpf_get_str __gmpf_get_str
#define mpf_add __gmpf_add
#define mpf_add_ui __gmpf_add_ui
#define mpf_sub __gmpf_sub
#define mpf_sub_ui __gmpf_sub_ui
#define mpf_mul __gmpf_mul
#define mpf_mul_ui __gmpf_mul_ui
#define mpf_mul_2exp __gmpf_mul_2exp
#define mpf_div __gmpf_div
#define mpf_div_ui __gmpf_div_ui
#define mpf_div_2exp __gmpf_div_2exp
#define mpf_neg __gmpf_neg
#define mpf_abs __gmpf_abs
#define mpf_sqrt __gmpf_sqrt
#define mpf_pow_ui __gmpf_pow_ui
#define mpf_cmp __gmpf_cmp
#define mpf_cmp_ui __gmpf_cmp_ui
#define mpf_cmp_d __gmpf_cmp_d
#define mpf_set_default_prec __gmpf_set_default_prec
#define mpf_get_default_prec __gmpf_get_default_prec
#define mpf_swap __gmpf_swap
#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_import __gmpz_import
#define mpz_export __gmpz_export
#define mpz_neg __gmpz_neg
#define mpz_set_ui __gmpz_set_ui
#define mpz_sizeinbase __gmpz_sizeinbase
#define mp_set_memory_functions __gmp_set_memory_functions
#define mp_get_memory_functions __gmp_get_memory_functions
#define gmp_snprintf __gmp_snprintf
#define gmp_asprintf __gmp_asprintf
#define gmp_printf __gmp_printf
#define gmp_sprintf __gmp_sprintf

void mpf_init(mpf_ptr);
void mpf_init2(mpf_ptr, mp_bitcnt_t);
void mpf_clear(mpf_ptr);
void mpf_set(mpf_ptr, mpf_srcptr);
void mpf_set_d(mpf_ptr, double);
void mpf_set_ui(mpf_ptr, unsigned long);
void mpf_set_si(mpf_ptr, long);
int mpf_set_str(mpf_ptr, const char *, int);
void mpf_set_z(mpf_ptr, mpz_srcptr);
void mpf_set_prec(mpf_ptr, mp_bitcnt_t);
void mpf_set_prec_raw(mpf_ptr, mp_bitcnt_t);
mp_bitcnt_t mpf_get_prec(mpf_srcptr);
double mpf_get_d(mpf_srcptr);
double mpf_get_d_2exp(long *, mpf_srcptr);
long mpf_get_si(mpf_srcptr);
unsigned long mpf_get_ui(mpf_srcptr);
char *mpf_get_str(char *, mp_exp_t *, int, size_t, mpf_srcptr);
void mpf_add(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_add_ui(mpf_ptr, mpf_srcptr, unsigned long);
void mpf_sub(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_sub_ui(mpf_ptr, mpf_srcptr, unsigned long);
void mpf_mul(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_mul_ui(mpf_ptr, mpf_srcptr, unsigned long);
void mpf_mul_2exp(mpf_ptr, mpf_srcptr, mp_bitcnt_t);
void mpf_div(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_div_ui(mpf_ptr, mpf_srcptr, unsigned long);
void mpf_div_2exp(mpf_ptr, mpf_srcptr, mp_bitcnt_t);
void mpf_neg(mpf_ptr, mpf_srcptr);
void mpf_abs(mpf_ptr, mpf_srcptr);
void mpf_sqrt(mpf_ptr, mpf_srcptr);
void mpf_pow_ui(mpf_ptr, mpf_srcptr, unsigned long);
int mpf_cmp(mpf_srcptr, mpf_srcptr);
int mpf_cmp_ui(mpf_srcptr, unsigned long);
int mpf_cmp_d(mpf_srcptr, double);
void mpf_set_default_prec(mp_bitcnt_t);
mp_bitcnt_t mpf_get_default_prec(void);
void mpf_swap(mpf_ptr, mpf_ptr);
void mpz_init(mpz_ptr);
void mpz_clear(mpz_ptr);
void mpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void *mpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
void mpz_neg(mpz_ptr, mpz_srcptr);
void mpz_set_ui(mpz_ptr, unsigned long);
size_t mpz_sizeinbase(mpz_srcptr, int);
void mp_set_memory_functions(void *(*)(size_t), void *(*)(void *, size_t, size_t), void (*)(void *, size_t));
void mp_get_memory_functions(void *(**)(size_t), void *(**)(void *, size_t, size_t), void (**)(void *, size_t));
int gmp_snprintf(char *, size_t, const char *, ...);
int gmp_asprintf(char **, const char *, ...);
int gmp_printf(const char *, ...);
int gmp_sprintf(char *, const char *, ...);
#ifdef __cplusplus
}
#endif
#define mpf_sgn(F) ((F)->_mp_size < 0 ? -1 : (F)->_mp_size > 0)
#define mpz_sgn(Z) ((Z)->_mp_size < 0 ? -1 : (Z)->_mp_size > 0)
#endif
