/*
 * nfft3_host.c -- host side of the drop-in: the reference's plan API (include/nfft3.h:163-187)
 * implemented in C on top of the CUDA C ABI (include/nfftcu.h).  Compiled twice, like the
 * reference compiles kernel/nfft/nfft.c once per precision (include/infft.h:68-98):
 *     default            -> nfft_*   (R = double)
 *     -DNFFT_B200_SINGLE -> nfftf_*  (R = float)
 * and linked into libnfft3_b200.so.  There is no CPU compute path in this file: every
 * transform is a call into libnfftcu.so, and a failure there ends in nfft_die like every fatal
 * error of the reference (kernel/util/malloc.c).
 *
 * What stays on the host, as in init_help (kernel/nfft/nfft.c:5950-6046): the public plan
 * members N, n, sigma, b, N_total, n_total, M_total, m, K, flags, the MALLOC_X/F_HAT/F buffers,
 * c_phi_inv (PRE_PHI_HUT) and index_x (NFFT_SORT_NODES).  What moved to the device: the grid
 * g1/g2, the FFT plans, psi.  The context pointer sits in the my_fftw_plan1 slot.
 *
 * Host pointers are re-read on every call: the solver swaps f / f_hat around each transform
 * (kernel/solver/solver.c:240-242, 275-277) and fastsum assigns x, f, f_hat after init
 * (applications/fastsum/fastsum.c:919-921).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/nfft3_b200.h"
#include "../../include/nfftcu.h"

#ifdef NFFT_B200_SINGLE
typedef float R;
typedef nfft_b200_cfloat C;
#define X(name) nfftf_##name
#define PRECISION NFFTCU_FLOAT
#define DEFAULT_M 4 /* WINDOW_HELP_ESTIMATE_m, Kaiser-Bessel, include/infft.h:224-230 */
#else
typedef double R;
typedef nfft_b200_cdouble C;
#define X(name) nfft_##name
#define PRECISION NFFTCU_DOUBLE
#define DEFAULT_M 8
#endif

typedef NFFT_INT INT;

/* published FFTW3 flag values, kept only so that plan.fftw_flags carries what callers expect */
#define B200_FFTW_DESTROY_INPUT (1U << 0)
#define B200_FFTW_ESTIMATE (1U << 6)

/* ---- memory / abort API (include/nfft3.h:69-82, kernel/util/malloc.c) ------------------------ */
X(malloc_type_function) X(malloc_hook) = 0;
X(free_type_function) X(free_hook) = 0;
X(die_type_function) X(die_hook) = 0;

void X(die)(const char *s)
{
  if (X(die_hook)) X(die_hook)(s);
  fflush(stdout);
  fprintf(stderr, "nfft: %s\n", s);
  exit(EXIT_FAILURE);
}

void *X(malloc)(size_t n)
{
  void *p = NULL;
  if (X(malloc_hook)) return X(malloc_hook)(n);
  /* large buffers are page-locked (nfftcu_host_alloc) so that nfft_trafo / nfft_adjoint on MALLOC_X / MALLOC_F_HAT /
   * MALLOC_F members -- and on anything else callers take from nfft_malloc -- copy at the full host-link rate */
  p = nfftcu_host_alloc(n);
  if (!p) X(die)("nfft_malloc: out of memory");
  return p;
}

void X(free)(void *p)
{
  if (!p) return;
  if (X(free_hook)) { X(free_hook)(p); return; }
  nfftcu_host_free(p);
}

/* ---- small util entry points the initialisers need (kernel/util/int.c, window.c) -------------- */
INT X(next_power_of_2)(const INT x)
{
  INT v = 1;
  if (x < 0) return -1;
  if (x < 2) return x + 1;      /* documented special case: 1 -> 2 */
  while (v < x) v <<= 1;
  return v;
}

/* The reference fixes the window family at configure time (--with-window=..., include/infft.h:146-222); this build
 * carries Kaiser-Bessel (default) and Gaussian and selects per process with NFFT_B200_WINDOW=gaussian. */
static int gaussian_window(void)
{
  const char *w = getenv("NFFT_B200_WINDOW");
  return w && (strcmp(w, "gaussian") == 0 || strcmp(w, "GAUSSIAN") == 0);
}

#ifdef NFFT_B200_SINGLE
#define DEFAULT_M_GAUSSIAN 5   /* WINDOW_HELP_ESTIMATE_m, include/infft.h:166-172 */
#else
#define DEFAULT_M_GAUSSIAN 13
#endif

INT X(get_default_window_cut_off)(void) { return gaussian_window() ? DEFAULT_M_GAUSSIAN : DEFAULT_M; }
const char *X(get_window_name)(void) { return gaussian_window() ? "gaussian" : "kaiserbessel"; }

/* ---- plan <-> context ------------------------------------------------------------------------ */
static nfftcu_ctx *ctx_of(const X(plan) *ths) { return (nfftcu_ctx*) ths->my_fftw_plan1; }
/* multi-device plans (NFFT_B200_DEVICES=0,1,...): the group sits in the second FFTW-plan slot, ctx_of() is then the
 * plan of the group's first device (same geometry: window parameters, c_phi_inv, stage times) */
static nfftcu_group *group_of(const X(plan) *ths) { return (nfftcu_group*) ths->my_fftw_plan2; }

static void check_cu(int status)
{
  if (status != NFFTCU_OK) X(die)(nfftcu_last_error());
}

/* flags whose window data depends on the nodes: the caller must call precompute_* after
 * changing x (SURVEY 8b "Node updates"); without them the reference re-sorts and re-evaluates
 * on every transform (nfft.c:4889, 5351), so we refresh the device copy of x on every call. */
#define NODE_BOUND_FLAGS (PRE_PSI | PRE_FULL_PSI | PRE_FG_PSI)

static void refresh_index_x(X(plan) *ths)
{
  if ((ths->flags & NFFT_SORT_NODES) && ths->index_x && ths->M_total > 0)
  {
    int64_t *dst = (int64_t*) ths->index_x;   /* NFFT_INT is 64-bit on LP64 */
    int r = group_of(ths) ? nfftcu_group_get_index_x(group_of(ths), dst) : nfftcu_get_index_x(ctx_of(ths), dst);
    if (r != NFFTCU_OK && r != NFFTCU_ESTATE) check_cu(r);
  }
}

static int64_t nodes_version(const X(plan) *ths)
{
  return group_of(ths) ? nfftcu_group_nodes_version(group_of(ths)) : nfftcu_nodes_version(ctx_of(ths));
}

static void upload_nodes(X(plan) *ths)
{
  const int64_t before = nodes_version(ths);
  if (ths->M_total > 0 && !ths->x) X(die)("Member x not initialized.");
  if (group_of(ths)) check_cu(nfftcu_group_set_nodes(group_of(ths), ths->x));
  else check_cu(nfftcu_set_nodes(ctx_of(ths), ths->x));
  if (nodes_version(ths) != before) refresh_index_x(ths);
}

static void nodes_for_transform(X(plan) *ths)
{
  if (!(ths->flags & NODE_BOUND_FLAGS) || nodes_version(ths) == 0)
    upload_nodes(ths);
}

/* for solver_host.c: same freshness rule as nfft_trafo, once per solver step */
__attribute__((visibility("hidden"))) void X(b200_nodes_for_transform)(X(plan) *ths)
{
  if (!(ths->flags & NODE_BOUND_FLAGS)) upload_nodes(ths);
  else nodes_for_transform(ths);
}

static void store_times(X(plan) *ths)
{
  float ms[3];
  if (nfftcu_stage_times(ctx_of(ths), ms) == NFFTCU_OK)
  {
    ths->MEASURE_TIME_t[0] = (R) (ms[0] * 1e-3f);
    ths->MEASURE_TIME_t[1] = (R) (ms[1] * 1e-3f);
    ths->MEASURE_TIME_t[2] = (R) (ms[2] * 1e-3f);
  }
}

/* ---- initialisation: init_help, nfft.c:5950-6046 ---------------------------------------------- */
static void init_help(X(plan) *ths)
{
  nfftcu_ctx *ctx = NULL;
  nfftcu_group *grp = NULL;
  int64_t N64[NFFTCU_MAX_D], n64[NFFTCU_MAX_D];
  INT t;
  int device = 0, ndev = 0, devs[NFFTCU_MAX_PEERS], grid_plan = 1;
  const char *dev_env = getenv("NFFT_B200_DEVICE");
  const char *devs_env = getenv("NFFT_B200_DEVICES");   /* "0,1,2,3": node-sharded over these GPUs (nfftcu_group_*) */
  unsigned cu_flags;

  if (ths->d < 1 || ths->d > NFFTCU_MAX_D) X(die)("nfft_init: rank d out of range [1,8]");
  if (ths->flags & NFFT_OMP_BLOCKWISE_ADJOINT) ths->flags |= NFFT_SORT_NODES;   /* nfft.c:5955 */
  cu_flags = ths->flags;

  ths->N_total = 1;
  ths->n_total = 1;
  for (t = 0; t < ths->d; t++)
  {
    ths->N_total *= ths->N[t];
    ths->n_total *= ths->n[t];
    N64[t] = ths->N[t];
    n64[t] = ths->n[t];
  }
  if (dev_env) device = atoi(dev_env);
  if (gaussian_window()) cu_flags |= NFFTCU_FLAG_GAUSSIAN;
  for (t = 0; t < ths->d; t++)
    if (ths->N[t] <= ths->m || ths->n[t] <= 2 * ths->m + 2) grid_plan = 0;   /* NDFT fallback plans stay on one device */
  if (devs_env && grid_plan && ths->M_total >= 1024)
  {
    const char *s = devs_env;
    while (*s && ndev < NFFTCU_MAX_PEERS)
    {
      char *end = NULL;
      const long v = strtol(s, &end, 10);
      if (end == s) break;
      devs[ndev++] = (int) v;
      s = (*end == ',') ? end + 1 : end;
    }
  }
  if (ndev >= 2)
  {
    /* plans the group cannot take (an FFT axis long enough to be split) fall back to the first listed device */
    if (nfftcu_group_create(&grp, PRECISION, (int) ths->d, N64, n64, ths->m, ths->M_total, cu_flags, devs, ndev)
        == NFFTCU_OK)
      ctx = nfftcu_group_ctx(grp, 0);
    else
      device = devs[0];
  }
  if (!grp)
    check_cu(nfftcu_create(&ctx, PRECISION, (int) ths->d, N64, n64, ths->m, ths->M_total,
        cu_flags, device));
  ths->my_fftw_plan1 = ctx;
  ths->my_fftw_plan2 = grp;

  ths->sigma = (R*) X(malloc)((size_t) ths->d * sizeof(R));
  ths->b = (R*) X(malloc)((size_t) ths->d * sizeof(R));
  check_cu(nfftcu_get_window_params(ctx, ths->b, ths->sigma));

  ths->x = (ths->flags & MALLOC_X) ?
      (R*) X(malloc)((size_t) (ths->d * ths->M_total) * sizeof(R)) : NULL;
  ths->f_hat = (ths->flags & MALLOC_F_HAT) ?
      (C*) X(malloc)((size_t) ths->N_total * sizeof(C)) : NULL;
  ths->f = (ths->flags & MALLOC_F) ? (C*) X(malloc)((size_t) ths->M_total * sizeof(C)) : NULL;

  ths->c_phi_inv = NULL;
  if (ths->flags & PRE_PHI_HUT)
  {
    ths->c_phi_inv = (R**) X(malloc)((size_t) ths->d * sizeof(R*));
    for (t = 0; t < ths->d; t++)
    {
      ths->c_phi_inv[t] = (R*) X(malloc)((size_t) ths->N[t] * sizeof(R));
      check_cu(nfftcu_get_c_phi_inv(ctx, (int) t, ths->c_phi_inv[t]));
    }
  }

  if ((ths->flags & PRE_LIN_PSI) && ths->K == 0)   /* m2K, kernel/util/window.c */
  {
    static const int m2K_[] = {1, 3, 7, 9, 14, 17, 20, 23, 24};
    const int j = ths->m < 8 ? (int) ths->m : 8;
    ths->K = (INT) ((1U << m2K_[j]) * (unsigned) (ths->m + 2));
  }
  /* a per-node window table on the device stands in for psi of PRE_PSI / PRE_FULL_PSI */
  {
    int r, nctx = grp ? nfftcu_group_size(grp) : 1;
    for (r = 0; r < nctx; r++)
    {
      nfftcu_ctx *cr = grp ? nfftcu_group_ctx(grp, r) : ctx;
      if (ths->flags & (PRE_PSI | PRE_FULL_PSI))
        check_cu(nfftcu_set_option(cr, NFFTCU_OPT_PSI_TABLE, 1));
      check_cu(nfftcu_set_option(cr, NFFTCU_OPT_TIMING, getenv("NFFT_B200_MEASURE_TIME") ? 1 : 0));
    }
  }

  ths->psi = NULL;
  ths->psi_index_g = NULL;
  ths->psi_index_f = NULL;
  ths->g = ths->g_hat = ths->g1 = ths->g2 = NULL;
  ths->spline_coeffs = NULL;
  ths->MEASURE_TIME_t[0] = ths->MEASURE_TIME_t[1] = ths->MEASURE_TIME_t[2] = (R) 0;

  ths->index_x = (ths->flags & NFFT_SORT_NODES) ?
      (INT*) X(malloc)(sizeof(INT) * 2U * (size_t) ths->M_total) : NULL;

  ths->mv_trafo = (void (*)(void*)) X(trafo);
  ths->mv_adjoint = (void (*)(void*)) X(adjoint);
}

static void copy_dims(X(plan) *ths, int d, const int *N, const int *n)
{
  INT t;
  ths->d = (INT) d;
  ths->N = (INT*) X(malloc)((size_t) (d > 0 ? d : 1) * sizeof(INT));
  ths->n = (INT*) X(malloc)((size_t) (d > 0 ? d : 1) * sizeof(INT));
  for (t = 0; t < d; t++)
  {
    ths->N[t] = (INT) N[t];
    ths->n[t] = n ? (INT) n[t] : 2 * X(next_power_of_2)((INT) N[t]);   /* nfft.c:6064-6065 */
  }
}

void X(init)(X(plan) *ths, int d, int *N, int M_total)
{
  copy_dims(ths, d, N, NULL);
  ths->M_total = (INT) M_total;
  ths->m = X(get_default_window_cut_off)();
  /* defaults of the reference's OpenMP build (nfft.c:6068-6081) */
  if (d > 1)
    ths->flags = PRE_PHI_HUT | PRE_PSI | MALLOC_X | MALLOC_F_HAT | MALLOC_F | FFTW_INIT |
        NFFT_SORT_NODES | NFFT_OMP_BLOCKWISE_ADJOINT;
  else
    ths->flags = PRE_PHI_HUT | PRE_PSI | MALLOC_X | MALLOC_F_HAT | MALLOC_F | FFTW_INIT |
        FFT_OUT_OF_PLACE;
  ths->fftw_flags = B200_FFTW_ESTIMATE | B200_FFTW_DESTROY_INPUT;
  ths->K = 0;
  init_help(ths);
}

void X(init_guru)(X(plan) *ths, int d, int *N, int M_total, int *n, int m, unsigned flags,
    unsigned fftw_flags)
{
  copy_dims(ths, d, N, n);
  ths->M_total = (INT) M_total;
  ths->m = (INT) m;
  ths->flags = flags;
  ths->fftw_flags = fftw_flags;
  ths->K = 0;
  init_help(ths);
}

void X(init_lin)(X(plan) *ths, int d, int *N, int M_total, int *n, int m, int K, unsigned flags,
    unsigned fftw_flags)
{
  copy_dims(ths, d, N, n);
  ths->M_total = (INT) M_total;
  ths->m = (INT) m;
  ths->flags = flags;
  ths->fftw_flags = fftw_flags;
  ths->K = (INT) K;
  init_help(ths);
}

void X(init_1d)(X(plan) *ths, int N1, int M_total)
{
  int N[1];
  N[0] = N1;
  X(init)(ths, 1, N, M_total);
}

void X(init_2d)(X(plan) *ths, int N1, int N2, int M_total)
{
  int N[2];
  N[0] = N1;
  N[1] = N2;
  X(init)(ths, 2, N, M_total);
}

void X(init_3d)(X(plan) *ths, int N1, int N2, int N3, int M_total)
{
  int N[3];
  N[0] = N1;
  N[1] = N2;
  N[2] = N3;
  X(init)(ths, 3, N, M_total);
}

/* ---- psi precomputation: nfft.c:5776-5948 ------------------------------------------------------ */
void X(precompute_psi)(X(plan) *ths) { upload_nodes(ths); }       /* sort + table, nfft.c:5819-5844 */
void X(precompute_full_psi)(X(plan) *ths) { upload_nodes(ths); }  /* nfft.c:5891-5936 */
void X(precompute_fg_psi)(X(plan) *ths) { upload_nodes(ths); }    /* Gaussian-only maths; exact psi here */
void X(precompute_lin_psi)(X(plan) *ths) { (void) ths; }          /* node independent, nfft.c:5776-5790 */

void X(precompute_one_psi)(X(plan) *ths)
{
  if (ths->flags & PRE_LIN_PSI) X(precompute_lin_psi)(ths);
  if (ths->flags & PRE_FG_PSI) X(precompute_fg_psi)(ths);
  if (ths->flags & PRE_PSI) X(precompute_psi)(ths);
  if (ths->flags & PRE_FULL_PSI) X(precompute_full_psi)(ths);
}

/* ---- transforms: nfft.c:5655-5749 --------------------------------------------------------------- */
void X(trafo)(X(plan) *ths)
{
  if (!ths->f_hat || !ths->f) X(die)("nfft_trafo: f_hat or f is NULL");
  if (!(ths->flags & NODE_BOUND_FLAGS))
  {
    /* no psi flag: x may have changed without notice; refresh it overlapped with the transform */
    int changed = 0;
    if (ths->M_total > 0 && !ths->x) X(die)("Member x not initialized.");
    if (group_of(ths)) check_cu(nfftcu_group_trafo_refresh(group_of(ths), ths->x, ths->f_hat, ths->f, &changed));
    else check_cu(nfftcu_trafo_refresh(ctx_of(ths), ths->x, ths->f_hat, ths->f, &changed));
    if (changed) refresh_index_x(ths);
  }
  else
  {
    nodes_for_transform(ths);
    if (group_of(ths)) check_cu(nfftcu_group_trafo(group_of(ths), ths->f_hat, ths->f));
    else check_cu(nfftcu_trafo(ctx_of(ths), ths->f_hat, ths->f));
  }
  store_times(ths);
}

void X(adjoint)(X(plan) *ths)
{
  if (!ths->f_hat || !ths->f) X(die)("nfft_adjoint: f_hat or f is NULL");
  if (!(ths->flags & NODE_BOUND_FLAGS))
  {
    int changed = 0;
    if (ths->M_total > 0 && !ths->x) X(die)("Member x not initialized.");
    if (group_of(ths)) check_cu(nfftcu_group_adjoint_refresh(group_of(ths), ths->x, ths->f, ths->f_hat, &changed));
    else check_cu(nfftcu_adjoint_refresh(ctx_of(ths), ths->x, ths->f, ths->f_hat, &changed));
    if (changed) refresh_index_x(ths);
  }
  else
  {
    nodes_for_transform(ths);
    if (group_of(ths)) check_cu(nfftcu_group_adjoint(group_of(ths), ths->f, ths->f_hat));
    else check_cu(nfftcu_adjoint(ctx_of(ths), ths->f, ths->f_hat));
  }
  store_times(ths);
}

/* ---- split-phase extensions (not part of the reference API; include/nfft3_b200.h) ------------------------------
 * nfft_b200_trafo_begin / nfft_b200_adjoint_begin enqueue copy-in, transform and copy-out on the plan's stream and
 * return; nfft_b200_wait returns when they are done.  The nodes must be resident (nfft_precompute_* or an earlier
 * transform) and unchanged; f / f_hat must be page-locked (nfft_malloc) to overlap, and untouched until the wait. */
void X(b200_trafo_begin)(X(plan) *ths)
{
  if (!ths->f_hat || !ths->f) X(die)("nfft_b200_trafo_begin: f_hat or f is NULL");
  if (group_of(ths)) X(die)("nfft_b200_trafo_begin: not available on multi-device plans");
  if (nodes_version(ths) == 0) upload_nodes(ths);
  check_cu(nfftcu_trafo_begin(ctx_of(ths), ths->f_hat, ths->f));
}

void X(b200_adjoint_begin)(X(plan) *ths)
{
  if (!ths->f_hat || !ths->f) X(die)("nfft_b200_adjoint_begin: f_hat or f is NULL");
  if (group_of(ths)) X(die)("nfft_b200_adjoint_begin: not available on multi-device plans");
  if (nodes_version(ths) == 0) upload_nodes(ths);
  check_cu(nfftcu_adjoint_begin(ctx_of(ths), ths->f, ths->f_hat));
}

void X(b200_wait)(X(plan) *ths)
{
  if (!group_of(ths)) check_cu(nfftcu_end(ctx_of(ths)));
}

void X(trafo_1d)(X(plan) *ths) { X(trafo)(ths); }
void X(trafo_2d)(X(plan) *ths) { X(trafo)(ths); }
void X(trafo_3d)(X(plan) *ths) { X(trafo)(ths); }
void X(adjoint_1d)(X(plan) *ths) { X(adjoint)(ths); }
void X(adjoint_2d)(X(plan) *ths) { X(adjoint)(ths); }
void X(adjoint_3d)(X(plan) *ths) { X(adjoint)(ths); }

/* exact NDFT, nfft.c:145-297; reads ths->x like the reference does */
void X(trafo_direct)(const X(plan) *ths)
{
  if (!ths->f_hat || !ths->f) X(die)("nfft_trafo_direct: f_hat or f is NULL");
  upload_nodes((X(plan)*) ths);
  if (group_of(ths)) check_cu(nfftcu_group_direct(group_of(ths), 0, ths->f_hat, ths->f));
  else check_cu(nfftcu_trafo_direct(ctx_of(ths), ths->f_hat, ths->f));
}

void X(adjoint_direct)(const X(plan) *ths)
{
  if (!ths->f_hat || !ths->f) X(die)("nfft_adjoint_direct: f_hat or f is NULL");
  upload_nodes((X(plan)*) ths);
  if (group_of(ths)) check_cu(nfftcu_group_direct(group_of(ths), 1, ths->f, ths->f_hat));
  else check_cu(nfftcu_adjoint_direct(ctx_of(ths), ths->f, ths->f_hat));
}

/* ---- nfft_check, nfft.c:6169-6207 (same messages) ----------------------------------------------- */
const char *X(check)(X(plan) *ths)
{
  INT j;
  if (!ths->f) return "Member f not initialized.";
  if (!ths->x) return "Member x not initialized.";
  if (!ths->f_hat) return "Member f_hat not initialized.";
  if ((ths->flags & PRE_LIN_PSI) && ths->K < ths->M_total)
    return "Number of nodes too small to use PRE_LIN_PSI.";
  for (j = 0; j < ths->M_total * ths->d; j++)
    if ((ths->x[j] < (R) -0.5) || (ths->x[j] >= (R) 0.5))
      return "ths->x out of range [-0.5,0.5)";
  for (j = 0; j < ths->d; j++)
  {
    if (ths->sigma[j] <= 1) return "Oversampling factor too small";
    if (ths->N[j] % 2 == 1) return "polynomial degree N has to be even";
  }
  return 0;
}

/* ---- nfft_finalize, nfft.c:6209-6270 ------------------------------------------------------------ */
void X(finalize)(X(plan) *ths)
{
  INT t;
  if (ths->flags & NFFT_SORT_NODES) X(free)(ths->index_x);
  if (group_of(ths)) check_cu(nfftcu_group_destroy(group_of(ths)));
  else check_cu(nfftcu_destroy(ctx_of(ths)));
  ths->my_fftw_plan1 = NULL;
  ths->my_fftw_plan2 = NULL;
  if (ths->flags & PRE_PHI_HUT)
  {
    for (t = 0; t < ths->d; t++) X(free)(ths->c_phi_inv[t]);
    X(free)(ths->c_phi_inv);
  }
  if (ths->flags & MALLOC_F) X(free)(ths->f);
  if (ths->flags & MALLOC_F_HAT) X(free)(ths->f_hat);
  if (ths->flags & MALLOC_X) X(free)(ths->x);
  X(free)(ths->b);
  X(free)(ths->sigma);
  X(free)(ths->n);
  X(free)(ths->N);
}
