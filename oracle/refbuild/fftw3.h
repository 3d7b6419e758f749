/*
 * fftw3.h -- TEST INFRASTRUCTURE ONLY: stand-in for the FFTW3 public header.
 *
 * The reference includes <fftw3.h> from include/nfft3.h:24 and include/infft.h:51 and calls
 * FFTW only through the handful of entry points below (kernel/nfft/nfft.c:6022-6031, 5516,
 * 5557, 6221-6225; kernel/util/malloc.c:37,54).  FFTW3 itself is not vendored in the
 * reference and is not installed in this image, so oracle/refbuild/fftw_shim.c implements
 * exactly this subset on top of oracle/cpu_fft.c.  Flag values follow the published FFTW3
 * API (FFTW_MEASURE 0, FFTW_DESTROY_INPUT 1<<0, FFTW_ESTIMATE 1<<6, FORWARD -1, BACKWARD +1).
 * This is not FFTW; timings of the F step obtained through it are labelled "FFT != FFTW".
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

#define FFTW_CONCAT(prefix, name) prefix ## name
#define FFTW_MANGLE_DOUBLE(name) FFTW_CONCAT(fftw_, name)
#define FFTW_MANGLE_FLOAT(name) FFTW_CONCAT(fftwf_, name)
#define FFTW_MANGLE_LONG_DOUBLE(name) FFTW_CONCAT(fftwl_, name)

typedef enum
{
  FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3, FFTW_REDFT01 = 4,
  FFTW_REDFT10 = 5, FFTW_REDFT11 = 6, FFTW_RODFT00 = 7, FFTW_RODFT01 = 8, FFTW_RODFT10 = 9,
  FFTW_RODFT11 = 10
} fftw_r2r_kind_do_not_use_me;

#define ORACLE_FFTW_DECLARE(X, R, C) \
  typedef R C[2]; \
  typedef struct X(plan_s) *X(plan); \
  typedef fftw_r2r_kind_do_not_use_me X(r2r_kind); \
  void *X(malloc)(size_t n); \
  void X(free)(void *p); \
  X(plan) X(plan_dft)(int rank, const int *n, C *in, C *out, int sign, unsigned flags); \
  void X(execute)(const X(plan) p); \
  void X(destroy_plan)(X(plan) p); \
  int X(init_threads)(void); \
  void X(plan_with_nthreads)(int nthreads); \
  void X(cleanup)(void); \
  void X(cleanup_threads)(void);

/* When <complex.h> was included first, FFTW's convention is that fftw_complex is the native
 * C99 complex type; the reference relies on that (include/infft.h:29-31 before :51). */
#if defined(_Complex_I) && defined(complex) && defined(I)
#define ORACLE_FFTW_DECLARE_C99(X, R, C) \
  typedef R _Complex C; \
  typedef struct X(plan_s) *X(plan); \
  typedef fftw_r2r_kind_do_not_use_me X(r2r_kind); \
  void *X(malloc)(size_t n); \
  void X(free)(void *p); \
  X(plan) X(plan_dft)(int rank, const int *n, C *in, C *out, int sign, unsigned flags); \
  void X(execute)(const X(plan) p); \
  void X(destroy_plan)(X(plan) p); \
  int X(init_threads)(void); \
  void X(plan_with_nthreads)(int nthreads); \
  void X(cleanup)(void); \
  void X(cleanup_threads)(void);
ORACLE_FFTW_DECLARE_C99(FFTW_MANGLE_DOUBLE, double, fftw_complex)
ORACLE_FFTW_DECLARE_C99(FFTW_MANGLE_FLOAT, float, fftwf_complex)
ORACLE_FFTW_DECLARE_C99(FFTW_MANGLE_LONG_DOUBLE, long double, fftwl_complex)
#else
ORACLE_FFTW_DECLARE(FFTW_MANGLE_DOUBLE, double, fftw_complex)
ORACLE_FFTW_DECLARE(FFTW_MANGLE_FLOAT, float, fftwf_complex)
ORACLE_FFTW_DECLARE(FFTW_MANGLE_LONG_DOUBLE, long double, fftwl_complex)
#endif

#ifdef __cplusplus
}
#endif
#endif
