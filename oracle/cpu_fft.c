/*
 * cpu_fft.c -- TEST INFRASTRUCTURE ONLY (oracle/). See cpu_fft.h.
 *
 * Row-column algorithm: for every axis, transform all 1-D lines in place.
 *   power-of-two length : iterative radix-2 decimation in time, bit-reversed input
 *   any other length    : recursive mixed-radix Cooley-Tukey (smallest prime factor first,
 *                         naive DFT at prime lengths)
 * Twiddles are tabulated once per axis in long double and rounded to double, so that the
 * transform error stays at a few ulp * log2(n).
 * Lines are processed in bundles of LINE_BUNDLE neighbours so that strided axes still
 * touch whole cache lines; bundles are distributed over OpenMP threads.
 */
#include "cpu_fft.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LINE_BUNDLE 8

typedef struct
{
  long len;
  int is_pow2;
  int log2len;
  double *tw;      /* len entries (re,im): exp(sign*2*pi*i*k/len) */
  long *bitrev;    /* len entries when is_pow2 */
} axis_plan;

struct cpu_fft_plan_s
{
  int rank;
  long *n;
  long total;
  int sign;
  axis_plan *axis;
};

static int ilog2_exact(long v)
{
  int l = 0;
  while ((1L << l) < v) l++;
  return ((1L << l) == v) ? l : -1;
}

static void axis_plan_init(axis_plan *a, long len, int sign)
{
  const long double two_pi = 6.283185307179586476925286766559005768394L;
  long k;
  a->len = len;
  a->log2len = ilog2_exact(len);
  a->is_pow2 = (a->log2len >= 0);
  a->tw = (double*) malloc(sizeof(double) * 2 * (size_t) len);
  for (k = 0; k < len; k++)
  {
    long double ang = two_pi * (long double) k / (long double) len;
    a->tw[2 * k] = (double) cosl(ang);
    a->tw[2 * k + 1] = (double) (sign * sinl(ang));
  }
  a->bitrev = NULL;
  if (a->is_pow2)
  {
    a->bitrev = (long*) malloc(sizeof(long) * (size_t) len);
    for (k = 0; k < len; k++)
    {
      long r = 0, v = k;
      int b;
      for (b = 0; b < a->log2len; b++) { r = (r << 1) | (v & 1); v >>= 1; }
      a->bitrev[k] = r;
    }
  }
}

cpu_fft_plan *cpu_fft_plan_create(int rank, const long *n, int sign)
{
  cpu_fft_plan *p = (cpu_fft_plan*) malloc(sizeof(*p));
  int t;
  p->rank = rank;
  p->sign = sign;
  p->n = (long*) malloc(sizeof(long) * (size_t) rank);
  p->axis = (axis_plan*) malloc(sizeof(axis_plan) * (size_t) rank);
  p->total = 1;
  for (t = 0; t < rank; t++)
  {
    p->n[t] = n[t];
    p->total *= n[t];
    axis_plan_init(&p->axis[t], n[t], sign);
  }
  return p;
}

void cpu_fft_plan_destroy(cpu_fft_plan *p)
{
  int t;
  if (!p) return;
  for (t = 0; t < p->rank; t++)
  {
    free(p->axis[t].tw);
    free(p->axis[t].bitrev);
  }
  free(p->axis);
  free(p->n);
  free(p);
}

/* in-place radix-2 on a contiguous line; buf holds the line in natural order on entry */
static void line_pow2(const axis_plan *a, double *buf, double *tmp)
{
  const long len = a->len;
  long i, half;
  for (i = 0; i < len; i++)
  {
    const long r = a->bitrev[i];
    tmp[2 * r] = buf[2 * i];
    tmp[2 * r + 1] = buf[2 * i + 1];
  }
  for (half = 1; half < len; half <<= 1)
  {
    const long step = len / (2 * half);
    long blk, j;
    for (blk = 0; blk < len; blk += 2 * half)
    {
      for (j = 0; j < half; j++)
      {
        const double wr = a->tw[2 * j * step], wi = a->tw[2 * j * step + 1];
        double *lo = tmp + 2 * (blk + j), *hi = tmp + 2 * (blk + j + half);
        const double tr = hi[0] * wr - hi[1] * wi;
        const double ti = hi[0] * wi + hi[1] * wr;
        hi[0] = lo[0] - tr; hi[1] = lo[1] - ti;
        lo[0] += tr; lo[1] += ti;
      }
    }
  }
  memcpy(buf, tmp, sizeof(double) * 2 * (size_t) len);
}

static long smallest_factor(long v)
{
  long p;
  for (p = 2; p * p <= v; p++)
    if (v % p == 0) return p;
  return v;
}

/* out[0..len) = DFT of in[0], in[istride], ...; W_len = W_full^twstride */
static void mixed_rec(const axis_plan *a, const double *in, long istride, double *out,
    long len, long twstride)
{
  const long full = a->len;
  if (len == 1)
  {
    out[0] = in[0]; out[1] = in[1];
    return;
  }
  {
    const long p = smallest_factor(len);
    const long m = len / p;
    long r, k, q;
    double *acc;
    for (r = 0; r < p; r++)
      mixed_rec(a, in + 2 * r * istride, istride * p, out + 2 * r * m, m, twstride * p);
    acc = (double*) malloc(sizeof(double) * 2 * (size_t) p);
    for (k = 0; k < m; k++)
    {
      for (q = 0; q < p; q++)
      {
        double sr = 0.0, si = 0.0;
        const long kk = k + q * m;
        for (r = 0; r < p; r++)
        {
          const long e = ((r * kk) % len) * twstride % full;
          const double wr = a->tw[2 * e], wi = a->tw[2 * e + 1];
          const double xr = out[2 * (r * m + k)], xi = out[2 * (r * m + k) + 1];
          sr += xr * wr - xi * wi;
          si += xr * wi + xi * wr;
        }
        acc[2 * q] = sr; acc[2 * q + 1] = si;
      }
      for (q = 0; q < p; q++)
      {
        out[2 * (k + q * m)] = acc[2 * q];
        out[2 * (k + q * m) + 1] = acc[2 * q + 1];
      }
    }
    free(acc);
  }
}

static void line_any(const axis_plan *a, double *buf, double *tmp)
{
  if (a->len == 1) return;
  if (a->is_pow2)
    line_pow2(a, buf, tmp);
  else
  {
    mixed_rec(a, buf, 1, tmp, a->len, 1);
    memcpy(buf, tmp, sizeof(double) * 2 * (size_t) a->len);
  }
}

void cpu_fft_execute(const cpu_fft_plan *p, double *data)
{
  int t;
  for (t = 0; t < p->rank; t++)
  {
    const axis_plan *a = &p->axis[t];
    const long len = a->len;
    long inner = 1, outer, nbundles_inner, nb, t2;
    if (len == 1) continue;
    for (t2 = t + 1; t2 < p->rank; t2++) inner *= p->n[t2];
    outer = p->total / (len * inner);
    nbundles_inner = (inner + LINE_BUNDLE - 1) / LINE_BUNDLE;
    nb = outer * nbundles_inner;
#ifdef _OPENMP
    #pragma omp parallel
#endif
    {
      double *buf = (double*) malloc(sizeof(double) * 2 * (size_t) len * LINE_BUNDLE);
      double *tmp = (double*) malloc(sizeof(double) * 2 * (size_t) len);
      long b;
#ifdef _OPENMP
      #pragma omp for schedule(static)
#endif
      for (b = 0; b < nb; b++)
      {
        const long o = b / nbundles_inner;
        const long i0 = (b % nbundles_inner) * LINE_BUNDLE;
        const long cnt = (inner - i0 < LINE_BUNDLE) ? (inner - i0) : LINE_BUNDLE;
        double *base = data + 2 * (o * len * inner + i0);
        long k, c;
        for (k = 0; k < len; k++)
          for (c = 0; c < cnt; c++)
          {
            buf[2 * (c * len + k)] = base[2 * (k * inner + c)];
            buf[2 * (c * len + k) + 1] = base[2 * (k * inner + c) + 1];
          }
        for (c = 0; c < cnt; c++)
          line_any(a, buf + 2 * c * len, tmp);
        for (k = 0; k < len; k++)
          for (c = 0; c < cnt; c++)
          {
            base[2 * (k * inner + c)] = buf[2 * (c * len + k)];
            base[2 * (k * inner + c) + 1] = buf[2 * (c * len + k) + 1];
          }
      }
      free(buf);
      free(tmp);
    }
  }
}
