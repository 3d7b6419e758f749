/*
 * fftw_shim.c -- TEST INFRASTRUCTURE ONLY: the FFTW3 entry points the reference's NFFT path
 * calls (see fftw3.h in this directory for the call-site list), implemented on
 * oracle/cpu_fft.c.  The single-precision twin computes in double and rounds once, so the
 * F step of the nfftf_ reference build is the exact DFT rounded to float.
 */
#include <complex.h>
#include <stdlib.h>
#include <string.h>
#include "fftw3.h"
#include "../cpu_fft.h"

struct fftw_plan_s { cpu_fft_plan *p; double *in, *out; long total; };
struct fftwf_plan_s { cpu_fft_plan *p; float *in, *out; long total; double *work; };
struct fftwl_plan_s { int unused; };

static void *aligned_malloc64(size_t n)
{
  void *p = NULL;
  if (n == 0) n = 1;
  if (posix_memalign(&p, 64, n) != 0) return NULL;
  return p;
}

void *fftw_malloc(size_t n) { return aligned_malloc64(n); }
void fftw_free(void *p) { free(p); }
void *fftwf_malloc(size_t n) { return aligned_malloc64(n); }
void fftwf_free(void *p) { free(p); }

int fftw_init_threads(void) { return 1; }
int fftwf_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int nthreads) { (void) nthreads; }
void fftwf_plan_with_nthreads(int nthreads) { (void) nthreads; }
void fftw_cleanup(void) {}
void fftwf_cleanup(void) {}
void fftw_cleanup_threads(void) {}
void fftwf_cleanup_threads(void) {}

static cpu_fft_plan *make_plan(int rank, const int *n, int sign, long *total)
{
  long ln[16];
  int t;
  *total = 1;
  for (t = 0; t < rank; t++) { ln[t] = n[t]; *total *= n[t]; }
  return cpu_fft_plan_create(rank, ln, sign);
}

fftw_plan fftw_plan_dft(int rank, const int *n, fftw_complex *in, fftw_complex *out,
    int sign, unsigned flags)
{
  fftw_plan pl = (fftw_plan) malloc(sizeof(*pl));
  (void) flags;
  pl->p = make_plan(rank, n, sign, &pl->total);
  pl->in = (double*) in;
  pl->out = (double*) out;
  return pl;
}

void fftw_execute(const fftw_plan pl)
{
  if (pl->in != pl->out)
    memcpy(pl->out, pl->in, sizeof(double) * 2 * (size_t) pl->total);
  cpu_fft_execute(pl->p, pl->out);
}

void fftw_destroy_plan(fftw_plan pl)
{
  if (!pl) return;
  cpu_fft_plan_destroy(pl->p);
  free(pl);
}

fftwf_plan fftwf_plan_dft(int rank, const int *n, fftwf_complex *in, fftwf_complex *out,
    int sign, unsigned flags)
{
  fftwf_plan pl = (fftwf_plan) malloc(sizeof(*pl));
  (void) flags;
  pl->p = make_plan(rank, n, sign, &pl->total);
  pl->in = (float*) in;
  pl->out = (float*) out;
  pl->work = (double*) aligned_malloc64(sizeof(double) * 2 * (size_t) pl->total);
  return pl;
}

void fftwf_execute(const fftwf_plan pl)
{
  long i;
  for (i = 0; i < 2 * pl->total; i++) pl->work[i] = (double) pl->in[i];
  cpu_fft_execute(pl->p, pl->work);
  for (i = 0; i < 2 * pl->total; i++) pl->out[i] = (float) pl->work[i];
}

void fftwf_destroy_plan(fftwf_plan pl)
{
  if (!pl) return;
  cpu_fft_plan_destroy(pl->p);
  free(pl->work);
  free(pl);
}
