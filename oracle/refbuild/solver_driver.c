/* oracle/refbuild/solver_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A caller of the reference's iterative solver (kernel/solver/solver.c: solver_init_advanced_complex,
 * solver_before_loop_complex, solver_loop_one_step_complex; CGNR 232-296, CGNE 298-344) written the way
 * applications/mri/mri2d/reconstruct_data_2d.c:38-120 uses it, behind one plain C entry point so that the
 * tests can drive it with ctypes.  It is compiled twice by oracle/refbuild/Makefile, from the SAME sources:
 *   libsolver_ref.so    linked against the reference's own nfft.c (oracle/_ref/libnfft3_ref.so)
 *   libsolver_b200.so   the reference's solver.c + kernel/util/vector*.c compiled where they lie, unmodified,
 *                       linked against nfft_b200/lib/libnfft3_b200.so -- i.e. the reference solver running
 *                       on top of the B200 engine through the unchanged plan API (mv_trafo / mv_adjoint
 *                       function pointers, pointer swaps of f / f_hat around every call).
 *   libsolver_dev_b200.so   this driver alone, linked against libnfft3_b200.so, whose own solver_*_complex run
 *                           the iteration on the device (nfft_b200/csrc/solver_host.c, solver.cu)
 * With -DDRIVER_SINGLE the same driver is built for nfftf_ / solverf_ (entry point solver_driver_run_f).
 * tests/test_gpu_parity.py::test_reference_solver_runs_on_the_engine and ::test_device_solver_vs_reference
 * compare the iterates. */
#include <complex.h>
#include <string.h>

#include "nfft3.h"

#ifdef DRIVER_SINGLE
typedef float real_t;
#define nfft_plan nfftf_plan
#define solver_plan_complex solverf_plan_complex
#define nfft_mv_plan_complex nfftf_mv_plan_complex
#define nfft_init_guru nfftf_init_guru
#define nfft_precompute_one_psi nfftf_precompute_one_psi
#define nfft_finalize nfftf_finalize
#define solver_init_advanced_complex solverf_init_advanced_complex
#define solver_before_loop_complex solverf_before_loop_complex
#define solver_loop_one_step_complex solverf_loop_one_step_complex
#define solver_finalize_complex solverf_finalize_complex
#define solver_driver_run solver_driver_run_f
#else
typedef double real_t;
#endif

int solver_driver_run(int d, const int *N, int M, const int *n, int m, unsigned nfft_flags,
                      unsigned solver_flags, const real_t *x, const real_t *y, const real_t *w,
                      const real_t *w_hat, int iters, real_t *f_hat_out, real_t *dot_r_out, real_t landweber_alpha,
                      real_t *r_out)
{
  nfft_plan p;
  solver_plan_complex ip;
  int Nc[8], nc[8], l;
  long long k;
  for (l = 0; l < d; l++) { Nc[l] = N[l]; nc[l] = n[l]; }
  nfft_init_guru(&p, d, Nc, M, nc, m, nfft_flags, FFTW_MEASURE | FFTW_DESTROY_INPUT);
  memcpy(p.x, x, sizeof(real_t) * (size_t) d * (size_t) M);
  if (p.flags & PRE_ONE_PSI) nfft_precompute_one_psi(&p);
  solver_init_advanced_complex(&ip, (nfft_mv_plan_complex*) &p, solver_flags);
  memcpy(ip.y, y, sizeof(real_t) * 2 * (size_t) M);
  if (ip.flags & PRECOMPUTE_WEIGHT) memcpy(ip.w, w, sizeof(real_t) * (size_t) M);
  if (ip.flags & PRECOMPUTE_DAMP) memcpy(ip.w_hat, w_hat, sizeof(real_t) * (size_t) p.N_total);
  for (k = 0; k < p.N_total; k++) ip.f_hat_iter[k] = 0.0;
  if (ip.flags & LANDWEBER) ip.alpha_iter = landweber_alpha;
  solver_before_loop_complex(&ip);
  for (l = 0; l < iters; l++)
  {
    solver_loop_one_step_complex(&ip);
    dot_r_out[l] = ip.dot_r_iter;
  }
  memcpy(f_hat_out, ip.f_hat_iter, sizeof(real_t) * 2 * (size_t) p.N_total);
  if (r_out) memcpy(r_out, ip.r_iter, sizeof(real_t) * 2 * (size_t) M);
  solver_finalize_complex(&ip);
  nfft_finalize(&p);
  return 0;
}
