/*
 * solver_host.c -- the reference's inverse-NFFT solver API (include/nfft3.h:758-786: solver_init_advanced_complex,
 * solver_init_complex, solver_before_loop_complex, solver_loop_one_step_complex, solver_finalize_complex) on top of
 * the device-resident iteration of libnfftcu.so (include/nfftcu.h, nfftcu_solver_*).  Compiled twice like
 * nfft3_host.c: solver_* (double) and, with -DNFFT_B200_SINGLE, solverf_* (float).
 *
 * The host arrays of the plan are allocated as kernel/solver/solver.c:40-73 allocates them (which array exists
 * for which flag, z_hat_iter aliasing p_hat_iter unless CGNR); y, f_hat_iter and r_iter are page-locked so that
 * the per-step mirrors are asynchronous copies.  There is no host iteration in this file: an mv plan that is not
 * an NFFT plan of this library is refused.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/nfft3_b200.h"
#include "../../include/nfftcu.h"

#ifdef NFFT_B200_SINGLE
typedef float R;
typedef nfft_b200_cfloat C;
#define X(name) solverf_##name
#define Y(name) nfftf_##name
#else
typedef double R;
typedef nfft_b200_cdouble C;
#define X(name) solver_##name
#define Y(name) nfft_##name
#endif

/* nfft3_host.c: makes the device copy of the nodes current (same rule as nfft_trafo) */
__attribute__((visibility("hidden"))) void Y(b200_nodes_for_transform)(Y(plan) *ths);

/* solver_plan_complex has no spare member for the device object: keep a registry keyed by the plan address */
/* (distinct solver plans may be driven from different host threads, like distinct nfft plans) */
typedef struct reg_s { const void *key; nfftcu_solver *dev; struct reg_s *next; } reg_t;
static reg_t *registry = NULL;
static pthread_mutex_t registry_lock = PTHREAD_MUTEX_INITIALIZER;

static nfftcu_solver *reg_find(const void *key)
{
  reg_t *r;
  nfftcu_solver *dev = NULL;
  pthread_mutex_lock(&registry_lock);
  for (r = registry; r; r = r->next) if (r->key == key) { dev = r->dev; break; }
  pthread_mutex_unlock(&registry_lock);
  return dev;
}

static void reg_put(const void *key, nfftcu_solver *dev)
{
  reg_t *r;
  nfftcu_solver *stale = NULL;
  pthread_mutex_lock(&registry_lock);
  for (r = registry; r; r = r->next) if (r->key == key) break;
  if (r)
  {
    stale = r->dev;   /* plan re-initialised without finalize */
    r->dev = dev;
  }
  else
  {
    r = (reg_t*) malloc(sizeof(reg_t));
    if (r) { r->key = key; r->dev = dev; r->next = registry; registry = r; }
  }
  pthread_mutex_unlock(&registry_lock);
  if (!r) Y(die)("solver_init: out of memory");
  if (stale) nfftcu_solver_destroy(stale);
}

static nfftcu_solver *reg_take(const void *key)
{
  reg_t **pp, *r;
  nfftcu_solver *dev = NULL;
  pthread_mutex_lock(&registry_lock);
  for (pp = &registry; (r = *pp) != NULL; pp = &r->next)
    if (r->key == key)
    {
      dev = r->dev;
      *pp = r->next;
      free(r);
      break;
    }
  pthread_mutex_unlock(&registry_lock);
  return dev;
}

static void check_cu(int status)
{
  if (status != NFFTCU_OK) Y(die)(nfftcu_last_error());
}

static void *pinned(size_t bytes)
{
  void *p = NULL;
  check_cu(nfftcu_malloc_pinned(&p, bytes ? bytes : 1));
  return p;
}

static Y(plan) *nfft_plan_of(X(plan_complex) *ths)
{
  if (!ths->mv || ths->mv->mv_trafo != (void (*)(void*)) Y(trafo) || ths->mv->mv_adjoint != (void (*)(void*)) Y(adjoint))
    Y(die)("solver (B200): mv is not an NFFT plan of libnfft3_b200.so; the device-resident solver handles "
           "nfft plans only -- link the reference's kernel/solver/solver.c for other transforms");
  return (Y(plan)*) ths->mv;   /* MACRO_MV_PLAN is the head of nfft_plan, include/nfft3.h:54-60, 109 */
}

static void scalars_out(X(plan_complex) *ths, const double *sc)
{
  ths->alpha_iter = (R) sc[0];
  ths->beta_iter = (R) sc[1];
  ths->dot_r_iter = (R) sc[2];
  ths->dot_r_iter_old = (R) sc[3];
  ths->dot_z_hat_iter = (R) sc[4];
  ths->dot_z_hat_iter_old = (R) sc[5];
  ths->dot_p_hat_iter = (R) sc[6];
  ths->dot_v_iter = (R) sc[7];
}

static void mirror_all(X(plan_complex) *ths, nfftcu_solver *dev)
{
  if (!getenv("NFFT_B200_SOLVER_MIRROR_ALL")) return;
  check_cu(nfftcu_solver_download(dev, NFFTCU_SOLVER_P_HAT_ITER, ths->p_hat_iter));
  if (ths->flags & CGNR) check_cu(nfftcu_solver_download(dev, NFFTCU_SOLVER_Z_HAT_ITER, ths->z_hat_iter));
  if (ths->flags & (CGNR | STEEPEST_DESCENT)) check_cu(nfftcu_solver_download(dev, NFFTCU_SOLVER_V_ITER, ths->v_iter));
}

/* solver.c:40-73 */
void X(init_advanced_complex)(X(plan_complex) *ths, Y(mv_plan_complex) *mv, unsigned flags)
{
  Y(plan) *p;
  nfftcu_solver *dev = NULL;
  size_t M, N;
  ths->mv = mv;
  ths->flags = flags;
  p = nfft_plan_of(ths);
  if (p->my_fftw_plan2)
    Y(die)("solver (B200): the device-resident solver drives single-device plans; unset NFFT_B200_DEVICES for the "
           "plans a solver iterates on (coils are distributed plan-per-GPU, not node-sharded)");
  M = (size_t) mv->M_total;
  N = (size_t) mv->N_total;
  check_cu(nfftcu_solver_create(&dev, (nfftcu_ctx*) p->my_fftw_plan1, flags));
  reg_put(ths, dev);

  ths->y = (C*) pinned(M * sizeof(C));
  ths->r_iter = (C*) pinned(M * sizeof(C));
  ths->f_hat_iter = (C*) pinned(N * sizeof(C));
  ths->p_hat_iter = (C*) Y(malloc)(N * sizeof(C));
  ths->z_hat_iter = ths->p_hat_iter;
  ths->v_iter = NULL;
  ths->w = NULL;
  ths->w_hat = NULL;
  if (flags & STEEPEST_DESCENT) ths->v_iter = (C*) Y(malloc)(M * sizeof(C));
  if (flags & CGNR)
  {
    ths->z_hat_iter = (C*) Y(malloc)(N * sizeof(C));
    ths->v_iter = (C*) Y(malloc)(M * sizeof(C));
  }
  if (flags & PRECOMPUTE_WEIGHT) ths->w = (R*) Y(malloc)(M * sizeof(R));
  if (flags & PRECOMPUTE_DAMP) ths->w_hat = (R*) Y(malloc)(N * sizeof(R));
  ths->alpha_iter = ths->beta_iter = ths->dot_r_iter = ths->dot_r_iter_old = (R) 0;
  ths->dot_z_hat_iter = ths->dot_z_hat_iter_old = ths->dot_p_hat_iter = ths->dot_v_iter = (R) 0;
}

void X(init_complex)(X(plan_complex) *ths, Y(mv_plan_complex) *mv)
{
  X(init_advanced_complex)(ths, mv, CGNR);   /* solver.c:76-79 */
}

/* solver.c:81-125 */
void X(before_loop_complex)(X(plan_complex) *ths)
{
  double sc[8];
  nfftcu_solver *dev = reg_find(ths);
  Y(plan) *p = nfft_plan_of(ths);
  if (!dev) Y(die)("solver_before_loop: plan was not initialised by solver_init");
  Y(b200_nodes_for_transform)(p);
  check_cu(nfftcu_solver_upload(dev, NFFTCU_SOLVER_Y, ths->y));
  check_cu(nfftcu_solver_upload(dev, NFFTCU_SOLVER_F_HAT_ITER, ths->f_hat_iter));
  if (ths->flags & PRECOMPUTE_WEIGHT) check_cu(nfftcu_solver_upload(dev, NFFTCU_SOLVER_W, ths->w));
  if (ths->flags & PRECOMPUTE_DAMP) check_cu(nfftcu_solver_upload(dev, NFFTCU_SOLVER_W_HAT, ths->w_hat));
  check_cu(nfftcu_solver_before_loop(dev, ths->f_hat_iter, ths->r_iter, sc));
  sc[0] = (double) ths->alpha_iter;   /* caller-owned for LANDWEBER, untouched here as in the reference */
  sc[1] = (double) ths->beta_iter;
  scalars_out(ths, sc);
  mirror_all(ths, dev);
}

/* solver.c:347-360 */
void X(loop_one_step_complex)(X(plan_complex) *ths)
{
  double sc[8] = {0};
  nfftcu_solver *dev = reg_find(ths);
  Y(plan) *p = nfft_plan_of(ths);
  if (!dev) Y(die)("solver_loop_one_step: plan was not initialised by solver_init");
  Y(b200_nodes_for_transform)(p);
  sc[0] = (double) ths->alpha_iter;
  check_cu(nfftcu_solver_step(dev, ths->f_hat_iter, ths->r_iter, sc));
  scalars_out(ths, sc);
  mirror_all(ths, dev);
}

/* solver.c:363-389 */
void X(finalize_complex)(X(plan_complex) *ths)
{
  nfftcu_solver *dev = reg_take(ths);
  if (dev) check_cu(nfftcu_solver_destroy(dev));
  if (ths->flags & PRECOMPUTE_WEIGHT) Y(free)(ths->w);
  if (ths->flags & PRECOMPUTE_DAMP) Y(free)(ths->w_hat);
  if (ths->flags & CGNR)
  {
    Y(free)(ths->v_iter);
    Y(free)(ths->z_hat_iter);
  }
  if (ths->flags & STEEPEST_DESCENT) Y(free)(ths->v_iter);
  Y(free)(ths->p_hat_iter);
  check_cu(nfftcu_free_pinned(ths->f_hat_iter));
  check_cu(nfftcu_free_pinned(ths->r_iter));
  check_cu(nfftcu_free_pinned(ths->y));
}
