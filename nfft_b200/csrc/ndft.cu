// ndft.cu -- exact nonequispaced DFT, O(M * N_total).
//
// Replaces nfft_trafo_direct / nfft_adjoint_direct (kernel/nfft/nfft.c:145-205, 207-297 of the
// reference):   f_j     = sum_k f_hat_k exp(-2 pi i k.x_j)
//               f_hat_k = sum_j f_j     exp(+2 pi i k.x_j),   k_t in [-N_t/2, N_t/2).
// Public API in the reference, the oracle of its own tests (tests/nfft.c:458-541), and the
// fallback nfft_trafo/adjoint take when any N_t <= m or n_t <= 2m+2 (nfft.c:5658-5664).
// One CTA per output element; the phase k.x is formed in double and reduced exactly by
// sincospi, partial sums are kept in double for both precisions.
#include "common.cuh"

namespace nfftcu {

namespace {

constexpr int kThreads = 128;

struct BandGeom {
  long long N[NFFTCU_MAX_D];
  int d;
};

__device__ __forceinline__ void block_sum2(double &a, double &b) {
  __shared__ double sa[kThreads / 32], sb[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sa[warp] = a; sb[warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < kThreads / 32; w++) { ta += sa[w]; tb += sb[w]; }
    a = ta;
    b = tb;
  }
}

template <typename T>
__device__ __forceinline__ double phase_of(long long kl, const T *xj, const BandGeom &g) {
  double ph = 0.0;
  long long rem = kl;
  for (int t = g.d - 1; t >= 0; t--) {
    const long long ks = rem % g.N[t];
    rem /= g.N[t];
    ph += (double) (ks - g.N[t] / 2) * (double) xj[t];
  }
  return ph;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
ndft_trafo_kernel(const typename Cplx<T>::type *__restrict__ f_hat, const T *__restrict__ x,
                  typename Cplx<T>::type *__restrict__ f, long long N_total, BandGeom g) {
  const long long j = blockIdx.x;
  const T *xj = x + j * g.d;
  double sr = 0.0, si = 0.0;
  for (long long kl = threadIdx.x; kl < N_total; kl += kThreads) {
    double s, c;
    sincospi(-2.0 * phase_of<T>(kl, xj, g), &s, &c);
    const double vr = (double) f_hat[kl].x, vi = (double) f_hat[kl].y;
    sr += vr * c - vi * s;
    si += vr * s + vi * c;
  }
  block_sum2(sr, si);
  if (threadIdx.x == 0) f[j] = make_c<T>((T) sr, (T) si);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
ndft_adjoint_kernel(const typename Cplx<T>::type *__restrict__ f, const T *__restrict__ x,
                    typename Cplx<T>::type *__restrict__ f_hat, long long M, BandGeom g) {
  const long long kl = blockIdx.x;
  double sr = 0.0, si = 0.0;
  for (long long j = threadIdx.x; j < M; j += kThreads) {
    double s, c;
    sincospi(2.0 * phase_of<T>(kl, x + j * g.d, g), &s, &c);
    const double vr = (double) f[j].x, vi = (double) f[j].y;
    sr += vr * c - vi * s;
    si += vr * s + vi * c;
  }
  block_sum2(sr, si);
  if (threadIdx.x == 0) f_hat[kl] = make_c<T>((T) sr, (T) si);
}

BandGeom band_of(const nfftcu_ctx *c) {
  BandGeom g;
  g.d = c->d;
  for (int t = 0; t < c->d; t++) g.N[t] = c->N[t];
  return g;
}

}  // namespace

int ndft_trafo(nfftcu_ctx *c, const void *f_hat_dev, void *f_dev) {
  if (c->M == 0) return NFFTCU_OK;
  const BandGeom g = band_of(c);
  if (c->prec == NFFTCU_DOUBLE)
    ndft_trafo_kernel<double><<<(unsigned) c->M, kThreads, 0, c->stream>>>(
        (const double2 *) f_hat_dev, (const double *) c->x_dev, (double2 *) f_dev, c->N_total, g);
  else
    ndft_trafo_kernel<float><<<(unsigned) c->M, kThreads, 0, c->stream>>>(
        (const float2 *) f_hat_dev, (const float *) c->x_dev, (float2 *) f_dev, c->N_total, g);
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

int ndft_adjoint(nfftcu_ctx *c, const void *f_dev, void *f_hat_dev) {
  if (c->N_total == 0) return NFFTCU_OK;
  const BandGeom g = band_of(c);
  if (c->prec == NFFTCU_DOUBLE)
    ndft_adjoint_kernel<double><<<(unsigned) c->N_total, kThreads, 0, c->stream>>>(
        (const double2 *) f_dev, (const double *) c->x_dev, (double2 *) f_hat_dev, c->M, g);
  else
    ndft_adjoint_kernel<float><<<(unsigned) c->N_total, kThreads, 0, c->stream>>>(
        (const float2 *) f_dev, (const float *) c->x_dev, (float2 *) f_hat_dev, c->M, g);
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

}  // namespace nfftcu
