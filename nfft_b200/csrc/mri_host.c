/*
 * mri_host.c -- the reference's field-inhomogeneity API (include/nfft3.h:510-541: mri_inh_2d1d_trafo / _adjoint /
 * _init_guru / _finalize, mri_inh_3d_*; implemented by kernel/mri/mri.c in double precision only) on top of the
 * device-resident versions in libnfftcu.so (mri.cu: nfftcu_mri_inh_2d1d, nfftcu_mri_inh_3d).
 *
 * kernel/mri/mri.c loops N3 + 1 host-pointer NFFTs with host-side cexp / PHI_HUT / PHI scaling between them
 * (mri.c:76-91, 122-137); here the whole loop runs in HBM and only the arguments and the result cross PCIe.  The
 * host-visible side effects of the reference are kept, because callers rely on them:
 *   - 2d1d trafo: that->f is REPLACED by a fresh nfft_malloc buffer and the old one nfft_free'd (mri.c:94-96);
 *     2d1d adjoint does the same with that->f_hat (146-148); plan.f / plan.f_hat follow.
 *   - 3d adjoint scales that->f in place by 1/PHI_HUT (mri.c:233-236) before transforming.
 * The reference's mri.c still runs unmodified on top of nfft_trafo / nfft_adjoint of this library (tests:
 * test_reference_mri_inh_runs_on_the_engine); this file is the fused component SURVEY 8f rank 3 asks for.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/nfft3_b200.h"
#include "../../include/nfftcu.h"

typedef nfft_b200_cdouble C;

static void check_cu(int status)
{
  if (status != NFFTCU_OK) nfft_die(nfftcu_last_error());
}

static nfftcu_ctx *ctx_of(nfft_plan *p)
{
  if (p->my_fftw_plan2)
    nfft_die("mri_inh (B200): multi-device plans are not supported here; unset NFFT_B200_DEVICES");
  return (nfftcu_ctx*) p->my_fftw_plan1;
}

/* the nodes must be resident before the first transform: plans with PRE_PSI upload them in
 * nfft_precompute_psi, others on their first nfft_trafo -- which these wrappers never call */
static void nodes_resident(nfft_plan *p)
{
  if (!(p->flags & (PRE_PSI | PRE_FULL_PSI | PRE_FG_PSI)) || nfftcu_nodes_version(ctx_of(p)) == 0)
    check_cu(nfftcu_set_nodes(ctx_of(p), p->x));
}

/* ---- mri_inh_2d1d, mri.c:57-193 ---------------------------------------------------------------------------- */
void mri_inh_2d1d_trafo(mri_inh_2d1d_plan *that)
{
  C *f = (C*) nfft_malloc((size_t) that->M_total * sizeof(C));
  that->plan.f = that->f;           /* the solver may have swapped the pointers, mri.c:67-69 */
  that->plan.f_hat = that->f_hat;
  nodes_resident(&that->plan);
  check_cu(nfftcu_mri_inh_2d1d(ctx_of(&that->plan), 0, that->N3, that->sigma3, that->w, that->t, that->f_hat, f));
  nfft_free(that->plan.f);          /* mri.c:94-96 */
  that->f = f;
  that->plan.f = that->f;
}

void mri_inh_2d1d_adjoint(mri_inh_2d1d_plan *that)
{
  C *f_hat = (C*) nfft_malloc((size_t) that->N_total * sizeof(C));
  that->plan.f = that->f;
  that->plan.f_hat = that->f_hat;
  nodes_resident(&that->plan);
  check_cu(nfftcu_mri_inh_2d1d(ctx_of(&that->plan), 1, that->N3, that->sigma3, that->w, that->t, that->f, f_hat));
  nfft_free(that->plan.f_hat);      /* mri.c:146-148 */
  that->f_hat = f_hat;
  that->plan.f_hat = that->f_hat;
}

void mri_inh_2d1d_init_guru(mri_inh_2d1d_plan *ths, int *N, int M, int *n, int m, double sigma, unsigned nfft_flags,
    unsigned fftw_flags)
{
  nfft_init_guru(&ths->plan, 2, N, M, n, m, nfft_flags, fftw_flags);
  ths->N3 = N[2];
  ths->sigma3 = sigma;
  ths->N_total = ths->plan.N_total;
  ths->M_total = ths->plan.M_total;
  ths->f = ths->plan.f;
  ths->f_hat = ths->plan.f_hat;
  ths->t = (double*) nfft_malloc((size_t) ths->M_total * sizeof(double));
  ths->w = (double*) nfft_malloc((size_t) ths->N_total * sizeof(double));
  ths->mv_trafo = (void (*)(void*)) mri_inh_2d1d_trafo;
  ths->mv_adjoint = (void (*)(void*)) mri_inh_2d1d_adjoint;
}

void mri_inh_2d1d_finalize(mri_inh_2d1d_plan *ths)
{
  nfft_free(ths->t);
  nfft_free(ths->w);
  ths->plan.f = ths->f;             /* mri.c:186-190 */
  ths->plan.f_hat = ths->f_hat;
  nfft_finalize(&ths->plan);
}

/* ---- mri_inh_3d, mri.c:197-293 ------------------------------------------------------------------------------ */
void mri_inh_3d_trafo(mri_inh_3d_plan *that)
{
  that->plan.f = that->f;
  nodes_resident(&that->plan);
  check_cu(nfftcu_mri_inh_3d(ctx_of(&that->plan), 0, that->N3, that->sigma3, that->w, that->plan.x, that->f_hat,
      that->f, NULL));
}

void mri_inh_3d_adjoint(mri_inh_3d_plan *that)
{
  that->plan.f = that->f;
  nodes_resident(&that->plan);
  /* the reference divides that->f by PHI_HUT in place before the transform (mri.c:233-236): mirrored */
  check_cu(nfftcu_mri_inh_3d(ctx_of(&that->plan), 1, that->N3, that->sigma3, that->w, that->plan.x, that->f,
      that->f_hat, that->f));
}

void mri_inh_3d_init_guru(mri_inh_3d_plan *ths, int *N, int M, int *n, int m, double sigma, unsigned nfft_flags,
    unsigned fftw_flags)
{
  ths->N3 = N[2];
  ths->sigma3 = sigma;
  nfft_init_guru(&ths->plan, 3, N, M, n, m, nfft_flags, fftw_flags);
  ths->N_total = N[0] * N[1];
  ths->M_total = ths->plan.M_total;
  ths->f = ths->plan.f;
  ths->f_hat = (C*) nfft_malloc((size_t) ths->N_total * sizeof(C));
  ths->w = (double*) nfft_malloc((size_t) ths->N_total * sizeof(double));
  ths->mv_trafo = (void (*)(void*)) mri_inh_3d_trafo;
  ths->mv_adjoint = (void (*)(void*)) mri_inh_3d_adjoint;
}

void mri_inh_3d_finalize(mri_inh_3d_plan *ths)
{
  nfft_free(ths->w);
  nfft_free(ths->f_hat);
  nfft_finalize(&ths->plan);
}
