/* oracle/refbuild/apps_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Callers of the reference's kernel/mri/mri.c (field-inhomogeneity transforms mri_inh_2d1d_*, mri_inh_3d_*)
 * and applications/fastsum/fastsum.c (NFFT-based fast summation), written the way the reference's own
 * programs use them (applications/mri/mri2d/construct_data_inh_2d1d.c:40-140, construct_data_inh_3d.c,
 * applications/fastsum/fastsum_test.c:150-300), behind plain C entry points for ctypes.  Compiled twice by
 * oracle/refbuild/Makefile from the SAME sources:
 *   libapps_ref.so    mri.c + fastsum.c + kernels.c on the reference's own nfft.c (libnfft3_ref.so)
 *   libapps_b200.so   mri.c + fastsum.c + kernels.c + kernel/util/*.c compiled where they lie, unmodified,
 *                     on nfft_b200/lib/libnfft3_b200.so -- the reference applications running on the B200
 *                     engine through the unchanged plan API.
 * fastsum.c calls FFTW for its kernel coefficients b (fastsum.c:877, 771); FFTW is absent from this image, so
 * both builds link the same CPU shim (fftw_shim.c) for that call -- it is host-side set-up of the
 * application, not part of the NFFT path.
 * tests/test_gpu_parity.py::test_reference_mri_inh_runs_on_the_engine / test_reference_fastsum_runs_on_the_engine
 * compare the outputs of the two. */
#include <complex.h>
#include <string.h>

#include "config.h"
#include "nfft3.h"
#ifndef APPS_MRI_ONLY   /* libapps_dev_b200.so: only the mri drivers, on the product's own device-resident mri_inh_* */
#include "fastsum.h"
#include "kernels.h"
#endif

/* mri_inh_2d1d: f = trafo(f_hat) then f_hat_adj = adjoint(f_in) on the same plan.
 * N = {N0, N1, N3}, n = {n0, n1, N3}; x: M x 2, t: M, w: N0*N1. */
int apps_mri_inh_2d1d(const int *N, int M, const int *n, int m, double sigma, unsigned nfft_flags,
                      const double *x, const double *t, const double *w, const double *f_hat_in,
                      const double *f_in, double *f_out, double *f_hat_out)
{
  mri_inh_2d1d_plan p;
  int Nc[3] = {N[0], N[1], N[2]}, nc[3] = {n[0], n[1], n[2]};
  long long k;
  mri_inh_2d1d_init_guru(&p, Nc, M, nc, m, sigma, nfft_flags, FFTW_MEASURE | FFTW_DESTROY_INPUT);
  memcpy(p.plan.x, x, sizeof(double) * 2 * (size_t) M);
  memcpy(p.t, t, sizeof(double) * (size_t) M);
  memcpy(p.w, w, sizeof(double) * (size_t) p.N_total);
  if (p.plan.flags & PRE_PSI) nfft_precompute_psi(&p.plan);
  for (k = 0; k < p.N_total; k++) p.f_hat[k] = f_hat_in[2 * k] + _Complex_I * f_hat_in[2 * k + 1];
  mri_inh_2d1d_trafo(&p);
  memcpy(f_out, p.f, sizeof(double) * 2 * (size_t) M);
  for (k = 0; k < M; k++) p.f[k] = f_in[2 * k] + _Complex_I * f_in[2 * k + 1];
  mri_inh_2d1d_adjoint(&p);
  memcpy(f_hat_out, p.f_hat, sizeof(double) * 2 * (size_t) p.N_total);
  mri_inh_2d1d_finalize(&p);
  return 0;
}

/* mri_inh_3d: x: M x 3 (third coordinate = scaled read-out time), w: N0*N1. */
int apps_mri_inh_3d(const int *N, int M, const int *n, int m, double sigma, unsigned nfft_flags,
                    const double *x, const double *w, const double *f_hat_in, const double *f_in,
                    double *f_out, double *f_hat_out)
{
  mri_inh_3d_plan p;
  int Nc[3] = {N[0], N[1], N[2]}, nc[3] = {n[0], n[1], n[2]};
  long long k;
  mri_inh_3d_init_guru(&p, Nc, M, nc, m, sigma, nfft_flags, FFTW_MEASURE | FFTW_DESTROY_INPUT);
  memcpy(p.plan.x, x, sizeof(double) * 3 * (size_t) M);
  memcpy(p.w, w, sizeof(double) * (size_t) p.N_total);
  if (p.plan.flags & PRE_PSI) nfft_precompute_psi(&p.plan);
  for (k = 0; k < p.N_total; k++) p.f_hat[k] = f_hat_in[2 * k] + _Complex_I * f_hat_in[2 * k + 1];
  mri_inh_3d_trafo(&p);
  memcpy(f_out, p.f, sizeof(double) * 2 * (size_t) M);
  for (k = 0; k < M; k++) p.f[k] = f_in[2 * k] + _Complex_I * f_in[2 * k + 1];
  mri_inh_3d_adjoint(&p);
  memcpy(f_hat_out, p.f_hat, sizeof(double) * 2 * (size_t) p.N_total);
  mri_inh_3d_finalize(&p);
  return 0;
}

#ifndef APPS_MRI_ONLY
/* fastsum: f(y_j) = sum_k alpha_k K(|y_j - x_k|); kernel_id: 0 gaussian, 1 multiquadric, 2 one_over_x,
 * 3 inverse_multiquadric.  f_exact may be NULL (direct sum skipped). */
int apps_fastsum(int d, int N_total, int M_total, int nn, int m, int p, int kernel_id, double c,
                 double eps_I, double eps_B, unsigned fs_flags, const double *x, const double *alpha,
                 const double *y, double *f_out, double *f_exact)
{
  fastsum_plan fs;
  kernel kern = kernel_id == 0 ? gaussian : kernel_id == 1 ? multiquadric : kernel_id == 2 ? one_over_x
                                                                                            : inverse_multiquadric;
  long long k;
  double param = c;
  fastsum_init_guru(&fs, d, N_total, M_total, kern, &param, fs_flags, nn, m, p, eps_I, eps_B);
  memcpy(fs.x, x, sizeof(double) * (size_t) d * (size_t) N_total);
  memcpy(fs.y, y, sizeof(double) * (size_t) d * (size_t) M_total);
  for (k = 0; k < N_total; k++) fs.alpha[k] = alpha[2 * k] + _Complex_I * alpha[2 * k + 1];
  if (f_exact)
  {
    fastsum_exact(&fs);
    memcpy(f_exact, fs.f, sizeof(double) * 2 * (size_t) M_total);
  }
  fastsum_precompute(&fs);
  fastsum_trafo(&fs);
  memcpy(f_out, fs.f, sizeof(double) * 2 * (size_t) M_total);
  fastsum_finalize(&fs);
  return 0;
}
#endif /* APPS_MRI_ONLY */
