// deconv.cu -- D and D^T: phi_hat deconvolution fused with fftshift and zero padding.
//
// Replaces the D steps of the reference: trafo_1d/2d/3d (kernel/nfft/nfft.c:2793-2831,
// 3833-3897, 5415-5513), generic D_openmp_A (440-518) and the transposed adjoint versions
// (2874-2903, 3941-3997, 5560-5650, 535-613):
//   D  : g_hat[kappa(k)] = f_hat[k+N/2] * prod_t c_t[k_t+N_t/2],  0 elsewhere
//   D^T: f_hat[k+N/2]    = g_hat[kappa(k)] * prod_t c_t[k_t+N_t/2]
//   kappa_t = k_t (k_t >= 0) | n_t + k_t (k_t < 0)
// D is output-driven: one thread per oversampled-grid element decides whether it lies in the
// band and either gathers f_hat*c or writes zero, so the memset of the reference
// (nfft.c:5416-5430) and the scatter are one streaming pass (writes n_total, reads N_total).
// HBM-bound: algorithmic bytes C*(N_total + n_total) for D, C*2*N_total for D^T.
#include "common.cuh"

namespace nfftcu {

namespace {

struct DGeom {
  long long N[NFFTCU_MAX_D], n[NFFTCU_MAX_D];
  int d;
};

template <typename T> struct CPtrs { const T *c[NFFTCU_MAX_D]; };

template <typename T>
__global__ void deconv_pad_kernel(const typename Cplx<T>::type *__restrict__ f_hat,
                                  typename Cplx<T>::type *__restrict__ g, DGeom geo, CPtrs<T> cp,
                                  long long n_total, long long N_total) {
  typedef typename Cplx<T>::type C;
  f_hat += (size_t) blockIdx.y * N_total;   // right-hand side blockIdx.y of a batched transform
  g += (size_t) blockIdx.y * n_total;
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long gi = (long long) blockIdx.x * blockDim.x + threadIdx.x; gi < n_total; gi += stride) {
    long long rem = gi;
    long long ks[NFFTCU_MAX_D];
    bool in_band = true;
#pragma unroll 1
    for (int t = geo.d - 1; t >= 0; t--) {
      const long long l = rem % geo.n[t];
      rem /= geo.n[t];
      const long long half = geo.N[t] / 2;
      if (l <= geo.N[t] - 1 - half) ks[t] = l + half;          // k = l >= 0
      else if (l >= geo.n[t] - half) ks[t] = l - geo.n[t] + half;  // k = l - n < 0
      else { in_band = false; ks[t] = 0; }
    }
    C out = make_c<T>((T) 0, (T) 0);
    if (in_band) {
      long long kl = 0;
      T w = (T) 1;
      for (int t = 0; t < geo.d; t++) {
        kl = kl * geo.N[t] + ks[t];
        w = (t == 0) ? cp.c[0][ks[0]] : w * cp.c[t][ks[t]];
      }
      const C v = f_hat[kl];
      out = make_c<T>(v.x * w, v.y * w);
    }
    g[gi] = out;
  }
}

template <typename T>
__global__ void deconv_crop_kernel(const typename Cplx<T>::type *__restrict__ g,
                                   typename Cplx<T>::type *__restrict__ f_hat, DGeom geo,
                                   CPtrs<T> cp, long long N_total, long long n_total) {
  typedef typename Cplx<T>::type C;
  f_hat += (size_t) blockIdx.y * N_total;
  g += (size_t) blockIdx.y * n_total;
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long kl = (long long) blockIdx.x * blockDim.x + threadIdx.x; kl < N_total; kl += stride) {
    long long rem = kl;
    long long ks[NFFTCU_MAX_D];
#pragma unroll 1
    for (int t = geo.d - 1; t >= 0; t--) {
      ks[t] = rem % geo.N[t];
      rem /= geo.N[t];
    }
    long long gi = 0;
    T w = (T) 1;
    for (int t = 0; t < geo.d; t++) {
      const long long k = ks[t] - geo.N[t] / 2;
      gi = gi * geo.n[t] + (k >= 0 ? k : geo.n[t] + k);
      w = (t == 0) ? cp.c[0][ks[0]] : w * cp.c[t][ks[t]];
    }
    const C v = g[gi];
    f_hat[kl] = make_c<T>(v.x * w, v.y * w);
  }
}

// D without the zero padding: writes only the band of the grid (the pruned FFT passes of fft.cu never
// read anything else).  Input-driven, the mirror image of deconv_crop_kernel.
template <typename T>
__global__ void deconv_band_kernel(const typename Cplx<T>::type *__restrict__ f_hat,
                                   typename Cplx<T>::type *__restrict__ g, DGeom geo, CPtrs<T> cp,
                                   long long N_total, long long n_total) {
  typedef typename Cplx<T>::type C;
  f_hat += (size_t) blockIdx.y * N_total;
  g += (size_t) blockIdx.y * n_total;
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long kl = (long long) blockIdx.x * blockDim.x + threadIdx.x; kl < N_total; kl += stride) {
    long long rem = kl;
    long long ks[NFFTCU_MAX_D];
#pragma unroll 1
    for (int t = geo.d - 1; t >= 0; t--) {
      ks[t] = rem % geo.N[t];
      rem /= geo.N[t];
    }
    long long gi = 0;
    T w = (T) 1;
    for (int t = 0; t < geo.d; t++) {
      const long long k = ks[t] - geo.N[t] / 2;
      gi = gi * geo.n[t] + (k >= 0 ? k : geo.n[t] + k);
      w = (t == 0) ? cp.c[0][ks[0]] : w * cp.c[t][ks[t]];
    }
    const C v = f_hat[kl];
    g[gi] = make_c<T>(v.x * w, v.y * w);
  }
}

template <typename T>
int run(nfftcu_ctx *c, const void *f_hat_in, void *f_hat_out, bool transposed, bool band_only = false) {
  typedef typename Cplx<T>::type C;
  DGeom geo;
  CPtrs<T> cp;
  geo.d = c->d;
  for (int t = 0; t < c->d; t++) {
    geo.N[t] = c->N[t];
    geo.n[t] = c->n[t];
    cp.c[t] = (const T *) c->c_dev[t];
  }
  const int threads = 256;
  const long long work = (transposed || band_only) ? c->N_total : c->n_total;
  long long blocks = (work + threads - 1) / threads;
  const long long cap = (long long) c->sm_count * 16;   // grid-stride above 16 resident CTAs/SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const dim3 grid((unsigned) blocks, (unsigned) c->cur_batch);   // y: right-hand sides of a batched transform
  if (!transposed && band_only)
    deconv_band_kernel<T><<<grid, threads, 0, c->stream>>>((const C *) f_hat_in, (C *) c->grid, geo, cp, c->N_total,
                                                          c->n_total);
  else if (!transposed)
    deconv_pad_kernel<T><<<grid, threads, 0, c->stream>>>((const C *) f_hat_in, (C *) c->grid, geo, cp, c->n_total,
                                                         c->N_total);
  else
    deconv_crop_kernel<T><<<grid, threads, 0, c->stream>>>((const C *) c->grid, (C *) f_hat_out, geo, cp, c->N_total,
                                                          c->n_total);
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

}  // namespace

int stage_D(nfftcu_ctx *c, const void *f_hat_dev, bool band_only) {
  return c->prec == NFFTCU_DOUBLE ? run<double>(c, f_hat_dev, nullptr, false, band_only)
                                  : run<float>(c, f_hat_dev, nullptr, false, band_only);
}

int stage_DT(nfftcu_ctx *c, void *f_hat_dev) {
  return c->prec == NFFTCU_DOUBLE ? run<double>(c, nullptr, f_hat_dev, true)
                                  : run<float>(c, nullptr, f_hat_dev, true);
}

}  // namespace nfftcu
