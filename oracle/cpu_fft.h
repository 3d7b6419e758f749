/*
 * cpu_fft.h -- TEST INFRASTRUCTURE ONLY (oracle/): a small, accurate CPU complex DFT.
 *
 * The reference (NFFT3) delegates its F step to FFTW3 (kernel/nfft/nfft.c:6030-6031 plan_dft,
 * 5516/5557 fftw_execute), which is an external dependency with no version pin
 * (configure.ac:321-333) and is absent from this image.  FFTW computes the standard
 * unnormalised d-dimensional DFT
 *     out[k] = sum_l in[l] * exp(sign * 2*pi*i * <k,l>/n),   sign = -1 (FORWARD) / +1 (BACKWARD)
 * so any correct fp64 DFT substitutes.  This file is that substitute; it is used
 *   - by oracle/refbuild/fftw_shim.c to stand in for libfftw3 when the reference sources are
 *     compiled into oracle/_ref/, and
 *   - by oracle/nfft_oracle.c (the CPU restatement of the hot path).
 * Nothing in the product (nfft_b200/) may include or link this.
 */
#ifndef ORACLE_CPU_FFT_H
#define ORACLE_CPU_FFT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cpu_fft_plan_s cpu_fft_plan;

/* rank-d in-place transform over interleaved (re,im) doubles, row-major, last dim fastest. */
cpu_fft_plan *cpu_fft_plan_create(int rank, const long *n, int sign);
void cpu_fft_execute(const cpu_fft_plan *p, double *data);
void cpu_fft_plan_destroy(cpu_fft_plan *p);

#ifdef __cplusplus
}
#endif
#endif
