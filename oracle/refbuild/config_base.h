/*
 * config_base.h -- TEST INFRASTRUCTURE ONLY: hand-written replacement for the config.h that
 * the reference's autotools build would generate (autoconf/automake/libtool are absent from
 * this image, so ./bootstrap.sh && ./configure cannot run).  The Makefile in this directory
 * concatenates this file with one "#define HAVE_DECL_xxx 1" line per HAVE_DECL_ macro that
 * include/infft.h tests, and writes the result to oracle/_ref/inc/config.h.
 *
 * Choices mirror a default Linux/gcc configure run with the Kaiser-Bessel window
 * (configure.ac:239-260 default) and --enable-openmp.
 */
#define HAVE_COMPLEX_H 1
#define HAVE_STDINT_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_STDDEF_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_MATH_H 1
#define HAVE_ALLOCA 1
#define HAVE_ALLOCA_H 1
#define HAVE_TIME_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_UNISTD_H 1
#define HAVE_CLOCK_GETTIME 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_DRAND48 1
#define HAVE_SRAND48 1
#define HAVE_MEMALIGN 1
#define HAVE_POSIX_MEMALIGN 1
#define SIZEOF_PTRDIFF_T 8
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_LONG_LONG 8
#ifdef ORACLE_REF_GAUSSIAN   /* --with-window=gaussian (configure.ac:239-260) */
#define GAUSSIAN 1
#define WINDOW_NAME gaussian
#else
#define KAISER_BESSEL 1
#define WINDOW_NAME kaiserbessel
#endif
#define NFFT_VERSION_MAJOR 3
#define NFFT_VERSION_MINOR 5
#define NFFT_VERSION_PATCH 4
#define PACKAGE "nfft"
#define PACKAGE_STRING "nfft 3.5.4alpha"
#define PACKAGE_VERSION "3.5.4alpha"
#define ABS_SRCDIR "/root/reference"
#ifdef ORACLE_REF_SINGLE
#define NFFT_SINGLE 1
#define NFFT_PRECISION_SINGLE 1
#else
#define NFFT_PRECISION_DOUBLE 1
#endif
