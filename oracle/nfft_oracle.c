/*
 * nfft_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the NFFT3 hot path
 * (nfft_init_guru -> nfft_precompute_one_psi -> nfft_trafo / nfft_adjoint) with the
 * Kaiser-Bessel window, written from the algorithm, one function per reference stage.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference leg may
 * load this.  The product (nfft_b200/) never links or calls it.
 *
 * PARITY PIN: this restatement is checked (tests/test_oracle.py) against
 *   (1) the reference's own known-answer fixtures tests/data/nfft_*.txt (exact NDFT, 64
 *       digits; committed as tests/golden/ndft_fixtures.npz) with the reference's own bound
 *       (tests/nfft.c:217-284),
 *   (2) the bessel_i0 table of tests/bessel.c:28-133 (tests/golden/bessel_i0.npz),
 *   (3) outputs of the reference itself: oracle/_ref/libnfft3_ref.so, compiled from the
 *       reference's unmodified kernel/nfft/nfft.c + kernel/util by oracle/refbuild/Makefile,
 *       both live (when oracle/_ref is present) and through tests/golden/ref_outputs.npz.
 *
 * The file is compiled twice: R = double (symbols oracle_*) and, with -DORACLE_SINGLE,
 * R = float (symbols oraclef_*), mirroring the reference's nfft_/nfftf_ name mangling
 * (include/infft.h:68-98).  In the float build the arithmetic of the window, the D step and
 * the tap sums is carried out in float exactly where the reference's is; the DFT itself is
 * done in double and rounded once (like oracle/refbuild/fftw_shim.c) and I0 is evaluated in
 * double and rounded once (the reference uses a float rational approximation,
 * kernel/util/bessel_i0.c:175-211; the two agree to a few float ulp).
 *
 * Layout conventions (all row-major, last dimension fastest; complex = interleaved re,im):
 *   x[j*d+t] in [-1/2,1/2);  f[j];  f_hat[k], k_t+N_t/2 as index;  g[l], l_t in [0,n_t).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "cpu_fft.h"

#ifdef ORACLE_SINGLE
typedef float R;
#define X(name) oraclef_ ## name
#define SQRT sqrtf
#define SINH sinhf
#define SIN sinf
#define FLOOR floorf
#define LRINT lrintf
#else
typedef double R;
#define X(name) oracle_ ## name
#define SQRT sqrt
#define SINH sinh
#define SIN sin
#define FLOOR floor
#define LRINT lrint
#endif

#define ORACLE_MAX_D 8
#define KPI ((R) 3.1415926535897932384626433832795028841971693993751L)
#define K2PI ((R) 6.2831853071795864769252867665590057683943387987502L)

typedef long INT; /* NFFT_INT == ptrdiff_t, include/nfft3.h:51 */

/* ------------------------------------------------------------------------------------------
 * I0(x), modified Bessel function of the first kind, order 0.
 * Role in the reference: kernel/util/bessel_i0.c:300-338 (Chebyshev/rational approximations
 * with tabulated coefficients).  Restated here from the defining power series
 *     I0(x) = sum_k ((x/2)^2)^k / (k!)^2,
 * all terms positive, summed in long double until the term drops below 2^-70 of the sum, so
 * the double result is correctly rounded up to ~1 ulp for the argument range the window uses
 * (x = m*sqrt(b^2 - ...) <= m*b < 2*pi*m).
 * ---------------------------------------------------------------------------------------- */
double oracle_bessel_i0(double x);
#ifndef ORACLE_SINGLE
double oracle_bessel_i0(double x)
{
  const long double q = 0.25L * (long double) x * (long double) x;
  long double term = 1.0L, sum = 1.0L;
  int k;
  for (k = 1; k < 2000; k++)
  {
    term *= q / ((long double) k * (long double) k);
    sum += term;
    if (term < sum * 0x1p-70L) break;
  }
  return (double) sum;
}
#endif

/* b_t = pi*(2 - 1/sigma_t), sigma_t = n_t/N_t          (include/infft.h:216-222, nfft.c:5961-5964) */
R X(shape_b)(INT N, INT n)
{
  const R sigma = ((R) n) / ((R) N);
  return KPI * (((R) 2.0) - ((R) 1.0) / sigma);
}

/* phi_hat_t(k) = I0(m*sqrt(b^2 - (2*pi*k/n)^2))         (include/infft.h:208) */
R X(phi_hut)(INT n, INT m, R b, INT k)
{
  const R w = K2PI * (R) k / (R) n;
  return (R) oracle_bessel_i0((double) ((R) m * SQRT(b * b - w * w)));
}

/* c_phi_inv[k+N/2] = 1/phi_hat(k), k = -N/2 .. N/2-1 (..N-1-N/2 for odd N) (nfft.c:5754-5770) */
void X(c_phi_inv)(INT N, INT n, INT m, R *out)
{
  const R b = X(shape_b)(N, n);
  INT ks;
  for (ks = 0; ks < N; ks++)
    out[ks] = ((R) 1.0) / X(phi_hut)(n, m, b, ks - N / 2);
}

/* phi_t(y), y = x - l/n, s = m^2 - (n*y)^2:  s>0 sinh(b*sqrt(s))/(pi*sqrt(s)) ; s<0 the sin
 * branch ; s==0 b/pi.  Not truncated outside |n*y|<=m.  (include/infft.h:209-215) */
R X(phi)(INT n, INT m, R b, R y)
{
  const R mm = (R) m * (R) m;
  const R ny2 = y * (R) n * y * (R) n;
  const R s = mm - ny2;
  if (s > (R) 0.0)
    return SINH(b * SQRT(s)) / (KPI * SQRT(s));
  if (s < (R) 0.0)
    return SIN(b * SQRT(ny2 - mm)) / (KPI * SQRT(ny2 - mm));
  return b / KPI;
}

/* ------------------------------------------------------------------------------------------
 * Node sort (NFFT_SORT_NODES).  Reference: sort0, nfft.c:75-109 builds
 *   key_j = row-major linearisation of u_t = ((floor(n_t*x_jt - m) mod n_t) + n_t) mod n_t
 * and calls nfft_sort_node_indices_radix_lsdf (kernel/util/sort.c:91-167), an LSD radix sort
 * over all significant key bits, i.e. a STABLE sort by key (ties keep ascending j).  Restated
 * as key construction + stable merge sort; output layout index_x[2k]=key, index_x[2k+1]=j.
 * n*x - m is evaluated in R with one rounding per operation (no FMA contraction).
 * ---------------------------------------------------------------------------------------- */
static INT X(node_key)(int d, const INT *n, INT m, const R *xj)
{
  INT key = 0;
  int t;
  for (t = 0; t < d; t++)
  {
    volatile R prod = (R) n[t] * xj[t];          /* volatile: forbid fma(n,x,-m) */
    const INT help = (INT) LRINT(FLOOR(prod - (R) m));
    const INT u = (help % n[t] + n[t]) % n[t];
    key += u;
    if (t + 1 < d) key *= n[t + 1];
  }
  return key;
}

static void merge_sort_pairs(INT *a, INT *tmp, INT cnt)
{
  INT width;
  INT *src = a, *dst = tmp;
  for (width = 1; width < cnt; width *= 2)
  {
    INT lo;
    for (lo = 0; lo < cnt; lo += 2 * width)
    {
      const INT mid = (lo + width < cnt) ? lo + width : cnt;
      const INT hi = (lo + 2 * width < cnt) ? lo + 2 * width : cnt;
      INT i = lo, j = mid, k = lo;
      while (i < mid && j < hi)
      {
        if (src[2 * j] < src[2 * i]) { dst[2 * k] = src[2 * j]; dst[2 * k + 1] = src[2 * j + 1]; j++; }
        else { dst[2 * k] = src[2 * i]; dst[2 * k + 1] = src[2 * i + 1]; i++; }
        k++;
      }
      while (i < mid) { dst[2 * k] = src[2 * i]; dst[2 * k + 1] = src[2 * i + 1]; i++; k++; }
      while (j < hi) { dst[2 * k] = src[2 * j]; dst[2 * k + 1] = src[2 * j + 1]; j++; k++; }
    }
    { INT *s = src; src = dst; dst = s; }
  }
  if (src != a) memcpy(a, src, sizeof(INT) * 2 * (size_t) cnt);
}

void X(sort_nodes)(int d, const INT *n, INT m, INT M, const R *x, INT *index_x)
{
  INT j;
  INT *tmp = (INT*) malloc(sizeof(INT) * 2 * (size_t) (M > 0 ? M : 1));
  for (j = 0; j < M; j++)
  {
    index_x[2 * j] = X(node_key)(d, n, m, x + j * d);
    index_x[2 * j + 1] = j;
  }
  merge_sort_pairs(index_x, tmp, M);
  free(tmp);
}

/* ------------------------------------------------------------------------------------------
 * D step (trafo) and D^T step (adjoint): deconvolution + fftshift + zero padding.
 *   g_hat := 0;  g_hat[kappa(k)] = f_hat[k+N/2] * prod_t c_t[k_t+N_t/2]
 *   kappa_t = k_t (k_t >= 0) | n_t + k_t (k_t < 0)
 * Reference: nfft.c:2793-2831 (1-D), 3833-3897 (2-D), 5415-5513 (3-D), generic 440-518;
 * transposed 2874-2903, 3941-3997, 5560-5650, 535-613.
 * ---------------------------------------------------------------------------------------- */
static void X(stage_D_impl)(int d, const INT *N, const INT *n, INT m, const R *f_hat, R *g_hat,
    int transposed, R *f_hat_out)
{
  R *c[ORACLE_MAX_D];
  INT N_total = 1, n_total = 1, ks[ORACLE_MAX_D], kl;
  int t;
  for (t = 0; t < d; t++)
  {
    N_total *= N[t];
    n_total *= n[t];
    c[t] = (R*) malloc(sizeof(R) * (size_t) N[t]);
    X(c_phi_inv)(N[t], n[t], m, c[t]);
  }
  if (!transposed)
    memset(g_hat, 0, sizeof(R) * 2 * (size_t) n_total);
  for (kl = 0; kl < N_total; kl++)
  {
    INT rem = kl, gi = 0;
    R w = (R) 1.0;
    for (t = d - 1; t >= 0; t--) { ks[t] = rem % N[t]; rem /= N[t]; }
    for (t = 0; t < d; t++)
    {
      const INT k = ks[t] - N[t] / 2;
      gi = gi * n[t] + (k >= 0 ? k : n[t] + k);
      w *= c[t][ks[t]];
    }
    if (!transposed)
    {
      g_hat[2 * gi] = f_hat[2 * kl] * w;
      g_hat[2 * gi + 1] = f_hat[2 * kl + 1] * w;
    }
    else
    {
      f_hat_out[2 * kl] = g_hat[2 * gi] * w;
      f_hat_out[2 * kl + 1] = g_hat[2 * gi + 1] * w;
    }
  }
  for (t = 0; t < d; t++) free(c[t]);
}

void X(stage_D)(int d, const INT *N, const INT *n, INT m, const R *f_hat, R *g_hat)
{
  X(stage_D_impl)(d, N, n, m, f_hat, g_hat, 0, NULL);
}

void X(stage_DT)(int d, const INT *N, const INT *n, INT m, const R *g_hat, R *f_hat)
{
  X(stage_D_impl)(d, N, n, m, NULL, (R*) g_hat, 1, f_hat);
}

/* F step: unnormalised d-dim DFT, sign -1 for trafo (FFTW_FORWARD), +1 for adjoint
 * (FFTW_BACKWARD): nfft.c:6030-6031, executed at 5516 / 5557.  In place. */
void X(stage_F)(int d, const INT *n, int sign, R *g)
{
  cpu_fft_plan *p = cpu_fft_plan_create(d, n, sign);
#ifdef ORACLE_SINGLE
  INT total = 1, i;
  int t;
  double *w;
  for (t = 0; t < d; t++) total *= n[t];
  w = (double*) malloc(sizeof(double) * 2 * (size_t) total);
  for (i = 0; i < 2 * total; i++) w[i] = (double) g[i];
  cpu_fft_execute(p, w);
  for (i = 0; i < 2 * total; i++) g[i] = (R) w[i];
  free(w);
#else
  cpu_fft_execute(p, g);
#endif
  cpu_fft_plan_destroy(p);
}

/* per node and dimension: u = floor(x*n) - m (nfft.c:324-332) and the 2m+2 window values
 * psi[l] = phi(x - (u+l)/n), l = 0..2m+1 (nfft.c:4896-4908, 5838-5840) */
static void X(node_window)(INT n, INT m, R b, R xj, INT *u, R *psi)
{
  const INT c = (INT) LRINT(FLOOR(xj * (R) n));
  INT l;
  *u = c - m;
  for (l = 0; l <= 2 * m + 1; l++)
    psi[l] = X(phi)(n, m, b, xj - ((R) (*u + l)) / (R) n);
}

/* B step (trafo): f_j = sum_{l} prod_t psi_t[l_t] * g[(u+l) mod n]
 * Reference: nfft.c:4020-4265 + 4687-4914 (3-D), 2927-3004/3221-3410 (2-D), 2131-2153/2283-2445
 * (1-D), generic 1172-1278.  Tap product order (psi0*psi1)*psi2 * g as at nfft.c:4048.
 * B^T step (adjoint): g := 0; g[(u+l) mod n] += prod_t psi_t[l_t] * f_j   (nfft.c:5137,
 * 4393-4436, generic 1974-2098).
 * `index_x` (may be NULL) gives the traversal permutation index_x[2k+1]; it only changes the
 * summation order of B^T.  b[t] is the window shape parameter of dimension t. */
static void X(stage_B_impl)(int d, const INT *n, INT m, const R *b, INT M, const R *x, R *g,
    R *f, const INT *index_x, int transposed)
{
  const INT taps1 = 2 * m + 2;
  INT taps = 1, n_total = 1, k;
  int t;
  for (t = 0; t < d; t++) { taps *= taps1; n_total *= n[t]; }
  if (transposed)
    memset(g, 0, sizeof(R) * 2 * (size_t) n_total);
#ifdef _OPENMP
  #pragma omp parallel for if(!transposed) schedule(static)
#endif
  for (k = 0; k < M; k++)
  {
    const INT j = index_x ? index_x[2 * k + 1] : k;
    INT u[ORACLE_MAX_D], l[ORACLE_MAX_D], tap;
    R *psi = (R*) malloc(sizeof(R) * (size_t) (d * taps1));
    R accr = (R) 0.0, acci = (R) 0.0;
    int tt;
    for (tt = 0; tt < d; tt++)
      X(node_window)(n[tt], m, b[tt], x[j * d + tt], &u[tt], psi + tt * taps1);
    for (tap = 0; tap < taps; tap++)
    {
      INT rem = tap, gi = 0;
      R w = (R) 1.0;
      for (tt = d - 1; tt >= 0; tt--) { l[tt] = rem % taps1; rem /= taps1; }
      for (tt = 0; tt < d; tt++)
      {
        const INT idx = (((u[tt] + l[tt]) % n[tt]) + n[tt]) % n[tt];
        gi = gi * n[tt] + idx;
        w = (tt == 0) ? psi[l[0]] : w * psi[tt * taps1 + l[tt]];
      }
      if (!transposed)
      {
        accr += w * g[2 * gi];
        acci += w * g[2 * gi + 1];
      }
      else
      {
        g[2 * gi] += w * f[2 * j];
        g[2 * gi + 1] += w * f[2 * j + 1];
      }
    }
    if (!transposed) { f[2 * j] = accr; f[2 * j + 1] = acci; }
    free(psi);
  }
}

static void X(all_b)(int d, const INT *N, const INT *n, R *b)
{
  int t;
  for (t = 0; t < d; t++) b[t] = X(shape_b)(N[t], n[t]);
}

void X(stage_B)(int d, const INT *N, const INT *n, INT m, INT M, const R *x, const R *g, R *f,
    const INT *index_x)
{
  R b[ORACLE_MAX_D];
  X(all_b)(d, N, n, b);
  X(stage_B_impl)(d, n, m, b, M, x, (R*) g, f, index_x, 0);
}

void X(stage_BT)(int d, const INT *N, const INT *n, INT m, INT M, const R *x, const R *f, R *g,
    const INT *index_x)
{
  R b[ORACLE_MAX_D];
  X(all_b)(d, N, n, b);
  X(stage_B_impl)(d, n, m, b, M, x, g, (R*) f, index_x, 1);
}

/* per-node window table as nfft_precompute_psi lays it out (nfft.c:5819-5844):
 * psi[(j*d+t)*(2m+2)+l] = phi_t(x_jt - (u_jt+l)/n_t) */
void X(precompute_psi)(int d, const INT *N, const INT *n, INT m, INT M, const R *x, R *psi)
{
  R b[ORACLE_MAX_D];
  INT j, u;
  int t;
  X(all_b)(d, N, n, b);
  for (j = 0; j < M; j++)
    for (t = 0; t < d; t++)
      X(node_window)(n[t], m, b[t], x[j * d + t], &u, psi + (j * d + t) * (2 * m + 2));
}

/* exact NDFT, nfft.c:145-205:  f_j = sum_k f_hat_k exp(-2 pi i k x_j), k in [-N/2, N/2)^d.
 * The phase is accumulated per dimension exactly like the reference's Omega[] recurrence
 * (omega = sum_t k_t * (2 pi x_jt)), then one complex exponential per term. */
void X(trafo_direct)(int d, const INT *N, INT M, const R *x, const R *f_hat, R *f)
{
  INT N_total = 1, j;
  int t;
  for (t = 0; t < d; t++) N_total *= N[t];
#ifdef _OPENMP
  #pragma omp parallel for schedule(static)
#endif
  for (j = 0; j < M; j++)
  {
    INT kl, ks[ORACLE_MAX_D];
    R sr = (R) 0.0, si = (R) 0.0;
    int tt;
    for (kl = 0; kl < N_total; kl++)
    {
      INT rem = kl;
      R omega = (R) 0.0;
      for (tt = d - 1; tt >= 0; tt--) { ks[tt] = rem % N[tt]; rem /= N[tt]; }
      for (tt = 0; tt < d; tt++)
        omega = ((R) (ks[tt] - N[tt] / 2)) * (K2PI * x[j * d + tt]) + omega;
      {
        const R c = (R) cos((double) omega), s = (R) -sin((double) omega);
        sr += f_hat[2 * kl] * c - f_hat[2 * kl + 1] * s;
        si += f_hat[2 * kl] * s + f_hat[2 * kl + 1] * c;
      }
    }
    f[2 * j] = sr;
    f[2 * j + 1] = si;
  }
}

/* exact adjoint NDFT, nfft.c:207-297:  f_hat_k = sum_j f_j exp(+2 pi i k x_j) */
void X(adjoint_direct)(int d, const INT *N, INT M, const R *x, const R *f, R *f_hat)
{
  INT N_total = 1, kl;
  int t;
  for (t = 0; t < d; t++) N_total *= N[t];
#ifdef _OPENMP
  #pragma omp parallel for schedule(static)
#endif
  for (kl = 0; kl < N_total; kl++)
  {
    INT rem = kl, ks[ORACLE_MAX_D], j;
    R sr = (R) 0.0, si = (R) 0.0;
    int tt;
    for (tt = d - 1; tt >= 0; tt--) { ks[tt] = rem % N[tt]; rem /= N[tt]; }
    for (j = 0; j < M; j++)
    {
      R omega = (R) 0.0;
      for (tt = 0; tt < d; tt++)
        omega += (R) (ks[tt] - N[tt] / 2) * K2PI * x[j * d + tt];
      {
        const R c = (R) cos((double) omega), s = (R) sin((double) omega);
        sr += f[2 * j] * c - f[2 * j + 1] * s;
        si += f[2 * j] * s + f[2 * j + 1] * c;
      }
    }
    f_hat[2 * kl] = sr;
    f_hat[2 * kl + 1] = si;
  }
}

static int X(needs_direct)(int d, const INT *N, const INT *n, INT m)
{
  int t;
  for (t = 0; t < d; t++)
    if (N[t] <= m || n[t] <= 2 * m + 2) return 1;   /* nfft.c:5658-5664 */
  return 0;
}

/* nfft_trafo, nfft.c:5655-5701: f = B F D f_hat.  `sorted` != 0 traverses nodes in the
 * NFFT_SORT_NODES order (no effect on the result of a gather). */
void X(trafo)(int d, const INT *N, const INT *n, INT m, INT M, const R *x, const R *f_hat, R *f)
{
  INT n_total = 1;
  R *g;
  int t;
  if (X(needs_direct)(d, N, n, m)) { X(trafo_direct)(d, N, M, x, f_hat, f); return; }
  for (t = 0; t < d; t++) n_total *= n[t];
  g = (R*) malloc(sizeof(R) * 2 * (size_t) n_total);
  X(stage_D)(d, N, n, m, f_hat, g);
  X(stage_F)(d, n, -1, g);
  X(stage_B)(d, N, n, m, M, x, g, f, NULL);
  free(g);
}

/* nfft_adjoint, nfft.c:5703-5749: f_hat = D^T F^H B^T f */
void X(adjoint)(int d, const INT *N, const INT *n, INT m, INT M, const R *x, const R *f,
    R *f_hat, int sorted)
{
  INT n_total = 1, *index_x = NULL;
  R *g;
  int t;
  if (X(needs_direct)(d, N, n, m)) { X(adjoint_direct)(d, N, M, x, f, f_hat); return; }
  for (t = 0; t < d; t++) n_total *= n[t];
  g = (R*) malloc(sizeof(R) * 2 * (size_t) n_total);
  if (sorted)
  {
    index_x = (INT*) malloc(sizeof(INT) * 2 * (size_t) (M > 0 ? M : 1));
    X(sort_nodes)(d, n, m, M, x, index_x);
  }
  X(stage_BT)(d, N, n, m, M, x, f, g, index_x);
  X(stage_F)(d, n, +1, g);
  X(stage_DT)(d, N, n, m, g, f_hat);
  free(index_x);
  free(g);
}
