// mri.cu -- field-inhomogeneity transforms of kernel/mri/mri.c kept in HBM.
//
// mri_inh_2d1d_trafo (mri.c:57-103) is a host loop over l = -N3/2 .. N3/2: scale f_hat by
// exp(-2 pi i w_j l) / PHI_HUT(N3, N3 w_j), run a 2-D nfft_trafo, accumulate f_j * PHI(N3, t_j - l/N3) on the
// window's support; the adjoint (105-150) mirrors it.  On top of a host-pointer NFFT every one of the N3 + 1
// iterations crosses PCIe twice.  Here f_hat (or f), w and t are uploaded once, the scaling and accumulation are
// element-wise kernels, and the N3 + 1 transforms run as batched transforms on one node set (nfftcu_*_batch_dev,
// up to kMriBatch values of l per launch sequence); only the result returns to the host.
// mri_inh_3d_* (mri.c:197-260): one 3-D transform with the window taken over the third (frequency) axis;
// the expansion f_hat[j] -> f_hat3[j][l] = f_hat[j] PHI(N3, w_j - l/N3) and the 1/PHI_HUT(N3, N3 x_j2) scaling of f
// run on the device around nfftcu_trafo_dev / nfftcu_adjoint_dev.
//
// The window here is the reference's window_funct_plan (mri.c:33-52): d = 1, n = N3, b = pi (2 - 1/sigma3)
// (Kaiser-Bessel; the engine's Gaussian option is not offered for these wrappers).  PHI_HUT needs I0, which is
// tabulated per call on the host (power series in long double, as for c_phi_inv) -- N_total resp. M values.
// Double precision only, like kernel/mri/mri.c itself.
#include "common.cuh"

#include <math.h>

namespace nfftcu {
long double bessel_i0_ld(long double x);   // api.cu
namespace {

constexpr int kMriBatch = 8;

// tmp[k][j] = f_hat[j] * s[j] * exp(-2 pi i w[j] (l0 + k))                                   mri.c:78-79
__global__ void mri_scale_fhat_kernel(const double2 *__restrict__ f_hat, const double *__restrict__ s,
                                      const double *__restrict__ w, double2 *__restrict__ tmp, long long N, int l0) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  const int l = l0 + (int) blockIdx.y;
  double2 *out = tmp + (size_t) blockIdx.y * (size_t) N;
  for (long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    double sn, cs;
    sincospi(-2.0 * w[j] * (double) l, &sn, &cs);
    const double2 v = f_hat[j];
    const double sc = s[j];
    out[j] = make_double2((v.x * cs - v.y * sn) * sc, (v.x * sn + v.y * cs) * sc);
  }
}

// acc[j] += sum_k fl[k][j] * PHI(N3, t[j] - (l0 + k)/N3) on the support |t - l/N3| < m/N3       mri.c:81-89
__global__ void mri_acc_f_kernel(double2 *__restrict__ acc, const double2 *__restrict__ fl, const double *__restrict__ t,
                                 long long M, int l0, int K, int N3, double m, double b) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x; j < M; j += stride) {
    double2 a = acc[j];
    const double tj = t[j];
    for (int k = 0; k < K; k++) {
      const double dx = tj - (double) (l0 + k) / (double) N3;
      if (fabs(dx) < m / (double) N3) {
        const double phi = kb_phi(dx * (double) N3, m * m, b);
        const double2 v = fl[(size_t) k * (size_t) M + j];
        a.x += v.x * phi;
        a.y += v.y * phi;
      }
    }
    acc[j] = a;
  }
}

// tmp[k][j] = f[j] * PHI(N3, t[j] - (l0 + k)/N3) on the support, else 0                        mri.c:126-132
__global__ void mri_scale_f_kernel(const double2 *__restrict__ f, const double *__restrict__ t, double2 *__restrict__ tmp,
                                   long long M, int l0, int N3, double m, double b) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  const int l = l0 + (int) blockIdx.y;
  double2 *out = tmp + (size_t) blockIdx.y * (size_t) M;
  for (long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x; j < M; j += stride) {
    const double dx = t[j] - (double) l / (double) N3;
    double2 v = make_double2(0.0, 0.0);
    if (fabs(dx) < m / (double) N3) {
      const double phi = kb_phi(dx * (double) N3, m * m, b);
      v = f[j];
      v.x *= phi;
      v.y *= phi;
    }
    out[j] = v;
  }
}

// acc[i] += sum_k fh[k][i] * exp(+2 pi i w[i] (l0 + k));  last: acc[i] *= s[i]                 mri.c:134-141
__global__ void mri_acc_fhat_kernel(double2 *__restrict__ acc, const double2 *__restrict__ fh, const double *__restrict__ w,
                                    const double *__restrict__ s, long long N, int l0, int K, int last) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    double2 a = acc[i];
    const double wi = w[i];
    for (int k = 0; k < K; k++) {
      double sn, cs;
      sincospi(2.0 * wi * (double) (l0 + k), &sn, &cs);
      const double2 v = fh[(size_t) k * (size_t) N + i];
      a.x += v.x * cs - v.y * sn;
      a.y += v.x * sn + v.y * cs;
    }
    if (last) { a.x *= s[i]; a.y *= s[i]; }
    acc[i] = a;
  }
}

// f_hat3[j][l + N3/2] = f_hat[j] * PHI(N3, w[j] - l/N3) on the support, else 0, l in [-N3/2, N3/2)   mri.c:207-216
__global__ void mri3_expand_kernel(const double2 *__restrict__ f_hat, const double *__restrict__ w,
                                   double2 *__restrict__ f_hat3, long long N2, int N3, double m, double b) {
  const long long total = N2 * N3, stride = (long long) gridDim.x * blockDim.x;
  for (long long q = (long long) blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
    const long long j = q / N3;
    const int l = (int) (q - j * N3) - N3 / 2;
    const double dx = w[j] - (double) l / (double) N3;
    double2 v = make_double2(0.0, 0.0);
    if (fabs(dx) < m / (double) N3) {
      const double phi = kb_phi(dx * (double) N3, m * m, b);
      v = f_hat[j];
      v.x *= phi;
      v.y *= phi;
    }
    f_hat3[q] = v;
  }
}

// f_hat[j] = sum_l f_hat3[j][l + N3/2] * PHI(N3, w[j] - l/N3) on the support                         mri.c:241-250
__global__ void mri3_collapse_kernel(const double2 *__restrict__ f_hat3, const double *__restrict__ w,
                                     double2 *__restrict__ f_hat, long long N2, int N3, double m, double b) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x; j < N2; j += stride) {
    double2 a = make_double2(0.0, 0.0);
    const double wj = w[j];
    for (int l = -N3 / 2; l < N3 / 2; l++) {
      const double dx = wj - (double) l / (double) N3;
      if (fabs(dx) < m / (double) N3) {
        const double phi = kb_phi(dx * (double) N3, m * m, b);
        const double2 v = f_hat3[j * N3 + (l + N3 / 2)];
        a.x += v.x * phi;
        a.y += v.y * phi;
      }
    }
    f_hat[j] = a;
  }
}

__global__ void mri_scale_real_kernel(double2 *__restrict__ f, const double *__restrict__ s, long long M) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x; j < M; j += stride) {
    f[j].x *= s[j];
    f[j].y *= s[j];
  }
}

unsigned nblocks(const nfftcu_ctx *c, long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long) c->sm_count * 16;
  if (b > cap) b = cap;
  return (unsigned) (b < 1 ? 1 : b);
}

// 1 / PHI_HUT(N3, N3 v) = 1 / I0(m sqrt(b^2 - (2 pi v)^2))                      include/infft.h:208, mri.c:79,139,221
void inv_phi_hut(const double *v, long long count, long long stride, int m, double b, std::vector<double> &out) {
  const long double two_pi = 6.283185307179586476925286766559005768394L;
  out.resize((size_t) count);
  for (long long j = 0; j < count; j++) {
    const long double a = two_pi * (long double) v[j * stride];
    const long double arg2 = (long double) b * (long double) b - a * a;
    out[(size_t) j] = (double) (1.0L / bessel_i0_ld((long double) m * sqrtl(arg2 > 0 ? arg2 : 0.0L)));
  }
}

struct DevBufs {
  std::vector<void *> p;
  ~DevBufs() { for (void *q : p) if (q) pool_free(q); }
  int get(void **out, size_t bytes) {
    NFFTCU_CUDA(pool_malloc(out, bytes ? bytes : 16));
    p.push_back(*out);
    return NFFTCU_OK;
  }
};

int check_plan(const nfftcu_ctx *c, int d, const char *who) {
  if (!c) { set_error("%s: null context", who); return NFFTCU_EINVAL; }
  if (c->prec != NFFTCU_DOUBLE || c->d != d || c->direct_only) {
    set_error("%s: needs a double-precision %d-D grid plan", who, d);
    return NFFTCU_EINVAL;
  }
  if (!c->have_nodes) { set_error("%s: called before the nodes were set", who); return NFFTCU_ESTATE; }
  return NFFTCU_OK;
}

}  // namespace
}  // namespace nfftcu

using namespace nfftcu;

extern "C" {

int nfftcu_mri_inh_2d1d(nfftcu_ctx *c, int adjoint, int N3, double sigma3, const double *w_host, const double *t_host,
                        const void *in_host, void *out_host) {
  NFFTCU_TRY(check_plan(c, 2, "nfftcu_mri_inh_2d1d"));
  NFFTCU_CUDA(cudaSetDevice(c->device));
  const long long N = c->N_total, M = c->M;
  const int m = (int) c->m;
  const double b = 3.1415926535897932384626433832795028841971693993751 * (2.0 - 1.0 / sigma3);
  const int nl = 2 * (N3 / 2) + 1;                       // l = -N3/2 .. N3/2 inclusive
  const int K = nl < kMriBatch ? nl : kMriBatch;
  NFFTCU_TRY(ensure_batch(c, K));
  std::vector<double> s;
  inv_phi_hut(w_host, N, 1, m, b, s);
  DevBufs bufs;
  double *w_d, *t_d, *s_d;
  double2 *in_d, *acc_d, *tmpN, *tmpM;
  NFFTCU_TRY(bufs.get((void **) &w_d, sizeof(double) * (size_t) N));
  NFFTCU_TRY(bufs.get((void **) &t_d, sizeof(double) * (size_t) M));
  NFFTCU_TRY(bufs.get((void **) &s_d, sizeof(double) * (size_t) N));
  NFFTCU_TRY(bufs.get((void **) &in_d, sizeof(double2) * (size_t) (adjoint ? M : N)));
  NFFTCU_TRY(bufs.get((void **) &acc_d, sizeof(double2) * (size_t) (adjoint ? N : M)));
  NFFTCU_TRY(bufs.get((void **) &tmpN, sizeof(double2) * (size_t) N * K));
  NFFTCU_TRY(bufs.get((void **) &tmpM, sizeof(double2) * (size_t) M * K));
  cudaStream_t st = c->stream;
  NFFTCU_CUDA(cudaMemcpyAsync(w_d, w_host, sizeof(double) * (size_t) N, cudaMemcpyHostToDevice, st));
  NFFTCU_CUDA(cudaMemcpyAsync(t_d, t_host, sizeof(double) * (size_t) M, cudaMemcpyHostToDevice, st));
  NFFTCU_CUDA(cudaMemcpyAsync(s_d, s.data(), sizeof(double) * (size_t) N, cudaMemcpyHostToDevice, st));
  NFFTCU_CUDA(cudaMemcpyAsync(in_d, in_host, sizeof(double2) * (size_t) (adjoint ? M : N), cudaMemcpyHostToDevice, st));
  NFFTCU_CUDA(cudaMemsetAsync(acc_d, 0, sizeof(double2) * (size_t) (adjoint ? N : M), st));
  for (int i0 = 0; i0 < nl; i0 += K) {
    const int kk = (nl - i0) < K ? (nl - i0) : K, l0 = -(N3 / 2) + i0;
    if (!adjoint) {
      mri_scale_fhat_kernel<<<dim3(nblocks(c, N), kk), 256, 0, st>>>(in_d, s_d, w_d, tmpN, N, l0);
      NFFTCU_TRY(nfftcu_trafo_batch_dev(c, kk, tmpN, tmpM));
      mri_acc_f_kernel<<<nblocks(c, M), 256, 0, st>>>(acc_d, tmpM, t_d, M, l0, kk, N3, (double) m, b);
    } else {
      mri_scale_f_kernel<<<dim3(nblocks(c, M), kk), 256, 0, st>>>(in_d, t_d, tmpM, M, l0, N3, (double) m, b);
      NFFTCU_TRY(nfftcu_adjoint_batch_dev(c, kk, tmpM, tmpN));
      mri_acc_fhat_kernel<<<nblocks(c, N), 256, 0, st>>>(acc_d, tmpN, w_d, s_d, N, l0, kk, i0 + kk >= nl);
    }
    c->launches += 2;
  }
  NFFTCU_CUDA(cudaGetLastError());
  NFFTCU_CUDA(cudaMemcpyAsync(out_host, acc_d, sizeof(double2) * (size_t) (adjoint ? N : M), cudaMemcpyDeviceToHost, st));
  NFFTCU_CUDA(cudaStreamSynchronize(st));
  return NFFTCU_OK;
}

int nfftcu_mri_inh_3d(nfftcu_ctx *c, int adjoint, int N3, double sigma3, const double *w_host, const double *x_host,
                      const void *in_host, void *out_host, void *f_scaled_host) {
  NFFTCU_TRY(check_plan(c, 3, "nfftcu_mri_inh_3d"));
  NFFTCU_CUDA(cudaSetDevice(c->device));
  if (c->N[2] != N3) { set_error("nfftcu_mri_inh_3d: plan N[2] = %lld, N3 = %d", (long long) c->N[2], N3); return NFFTCU_EINVAL; }
  const long long N2 = c->N[0] * c->N[1], M = c->M;
  const int m = (int) c->m;
  const double b = 3.1415926535897932384626433832795028841971693993751 * (2.0 - 1.0 / sigma3);
  std::vector<double> s;
  inv_phi_hut(x_host + 2, M, 3, m, b, s);                // 1 / PHI_HUT(N3, N3 x[3j+2])
  DevBufs bufs;
  double *w_d, *s_d;
  double2 *fh2, *fh3, *f_d;
  NFFTCU_TRY(bufs.get((void **) &w_d, sizeof(double) * (size_t) N2));
  NFFTCU_TRY(bufs.get((void **) &s_d, sizeof(double) * (size_t) M));
  NFFTCU_TRY(bufs.get((void **) &fh2, sizeof(double2) * (size_t) N2));
  NFFTCU_TRY(bufs.get((void **) &fh3, sizeof(double2) * (size_t) c->N_total));
  NFFTCU_TRY(bufs.get((void **) &f_d, sizeof(double2) * (size_t) M));
  cudaStream_t st = c->stream;
  NFFTCU_CUDA(cudaMemcpyAsync(w_d, w_host, sizeof(double) * (size_t) N2, cudaMemcpyHostToDevice, st));
  NFFTCU_CUDA(cudaMemcpyAsync(s_d, s.data(), sizeof(double) * (size_t) M, cudaMemcpyHostToDevice, st));
  if (!adjoint) {
    NFFTCU_CUDA(cudaMemcpyAsync(fh2, in_host, sizeof(double2) * (size_t) N2, cudaMemcpyHostToDevice, st));
    mri3_expand_kernel<<<nblocks(c, N2 * N3), 256, 0, st>>>(fh2, w_d, fh3, N2, N3, (double) m, b);
    NFFTCU_TRY(nfftcu_trafo_dev(c, fh3, f_d));
    mri_scale_real_kernel<<<nblocks(c, M), 256, 0, st>>>(f_d, s_d, M);
    NFFTCU_CUDA(cudaMemcpyAsync(out_host, f_d, sizeof(double2) * (size_t) M, cudaMemcpyDeviceToHost, st));
  } else {
    NFFTCU_CUDA(cudaMemcpyAsync(f_d, in_host, sizeof(double2) * (size_t) M, cudaMemcpyHostToDevice, st));
    mri_scale_real_kernel<<<nblocks(c, M), 256, 0, st>>>(f_d, s_d, M);
    if (f_scaled_host)
      NFFTCU_CUDA(cudaMemcpyAsync(f_scaled_host, f_d, sizeof(double2) * (size_t) M, cudaMemcpyDeviceToHost, st));
    NFFTCU_TRY(nfftcu_adjoint_dev(c, f_d, fh3));
    mri3_collapse_kernel<<<nblocks(c, N2), 256, 0, st>>>(fh3, w_d, fh2, N2, N3, (double) m, b);
    NFFTCU_CUDA(cudaMemcpyAsync(out_host, fh2, sizeof(double2) * (size_t) N2, cudaMemcpyDeviceToHost, st));
  }
  c->launches += 2;
  NFFTCU_CUDA(cudaGetLastError());
  NFFTCU_CUDA(cudaStreamSynchronize(st));
  return NFFTCU_OK;
}

}  // extern "C"
