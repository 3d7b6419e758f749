// solver.cu -- device-resident iterations of the reference's inverse-NFFT solver.
//
// Replaces the host loops of kernel/solver/solver.c (before_loop 81-125; one step of LANDWEBER 128-174,
// STEEPEST_DESCENT 177-229, CGNR 232-292, CGNE 295-344) and the vector kernels they call
// (kernel/util/vector1.c dot_complex / dot_w_complex, vector2.c cp_complex / cp_w_complex, vector3.c
// upd_axpy / upd_xpay / upd_xpawy) for the case that the matrix-vector plan is an NFFT plan of this library:
// y, w, w_hat, the iterate f_hat_iter and the work vectors r, z_hat, p_hat, v stay in HBM; the two transforms
// of a step are nfftcu_trafo_dev / nfftcu_adjoint_dev; the step sizes alpha and beta are computed on the
// device from device-side dot products, so that a step is ONE stream-ordered sequence of launches with a
// single synchronisation at its end (the scalars the reference publishes in the plan -- dot_r_iter etc. --
// come back in one 64-byte copy).  The host mirrors of f_hat_iter and r_iter are refreshed on a side stream
// while the rest of the step runs.
//
// Arithmetic: element-wise updates in the plan precision exactly as the reference's loops; dot products are
// accumulated in double (block partials, then one block) and rounded to the plan precision, alpha and beta
// are formed in the plan precision like the reference's R divisions.
#include "common.cuh"

#include <string.h>

struct nfftcu_solver_s {
  nfftcu_ctx *plan = nullptr;
  int device = 0;                // cached: nfftcu_solver_destroy works after the plan is gone
  int K = 1;                     // right-hand sides iterated in lock-step (nfftcu_solver_create_batch): vectors are
                                 // [K][len], scalars [K][8], weights w / w_hat shared
  unsigned flags = 0;
  void *vec[8] = {nullptr};      // device vectors, index NFFTCU_SOLVER_*; Z may alias P
  void *fhat_in = nullptr;       // N_total complex: argument of the transform
  void *f_in = nullptr;          // M complex: argument of the adjoint / result of CGNE's transform
  double *sc = nullptr;          // device scalars, index NFFTCU_SOLVER_SC_*
  double *partial = nullptr;     // block partial sums
  double *sc_host = nullptr;     // pinned copy
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fhat = nullptr, ev_r = nullptr;
};

namespace nfftcu {
namespace {

constexpr int kVecThreads = 256;
constexpr int kMaxBlocks = 1184;   // 8 x 148

enum { SC_ALPHA = 0, SC_BETA, SC_DOT_R, SC_DOT_R_OLD, SC_DOT_Z, SC_DOT_Z_OLD, SC_DOT_P, SC_DOT_V };
enum { LANDWEBER = 1u << 0, STEEPEST_DESCENT = 1u << 1, CGNR = 1u << 2, CGNE = 1u << 3, NORMS_FOR_LANDWEBER = 1u << 4,
       PRECOMPUTE_WEIGHT = 1u << 5, PRECOMPUTE_DAMP = 1u << 6 };

inline unsigned blocks_for(long long n) {
  long long b = (n + kVecThreads - 1) / kVecThreads;
  return (unsigned) (b < 1 ? 1 : (b > kMaxBlocks ? kMaxBlocks : b));
}

// x <- w .* y  (w == nullptr: x <- y)                                         vector2.c cp_complex / cp_w_complex
template <typename T>
__global__ void cp_w_kernel(typename Cplx<T>::type *__restrict__ x, const T *__restrict__ w,
                            const typename Cplx<T>::type *__restrict__ y, long long n) {
  x += (size_t) blockIdx.y * (size_t) n;   // right-hand side blockIdx.y
  y += (size_t) blockIdx.y * (size_t) n;
  for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long) gridDim.x * blockDim.x) {
    typename Cplx<T>::type v = y[k];
    if (w) { const T wk = w[k]; v.x = wk * v.x; v.y = wk * v.y; }
    x[k] = v;
  }
}

// x <- x + s * a * w .* y, a = sc[ia] rounded to T (ia < 0: a = 1)            vector3.c upd_xpay / upd_xpawy
template <typename T>
__global__ void xpawy_kernel(typename Cplx<T>::type *__restrict__ x, const double *__restrict__ sc, int ia, T s,
                             const T *__restrict__ w, const typename Cplx<T>::type *__restrict__ y, long long n) {
  x += (size_t) blockIdx.y * (size_t) n;
  y += (size_t) blockIdx.y * (size_t) n;
  sc += 8 * blockIdx.y;
  const T a = s * (ia >= 0 ? (T) sc[ia] : (T) 1);
  for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long) gridDim.x * blockDim.x) {
    typename Cplx<T>::type xv = x[k];
    const typename Cplx<T>::type yv = y[k];
    if (w) {
      const T aw = a * w[k];           // the reference evaluates a * w[k] * y[k] left to right
      xv.x += aw * yv.x; xv.y += aw * yv.y;
    } else {
      xv.x += a * yv.x; xv.y += a * yv.y;
    }
    x[k] = xv;
  }
}

// x <- a * x + y, a = sc[ia] rounded to T (ia < 0: a = aconst)                vector3.c upd_axpy_complex
template <typename T>
__global__ void axpy_kernel(typename Cplx<T>::type *__restrict__ x, const double *__restrict__ sc, int ia, T aconst,
                            const typename Cplx<T>::type *__restrict__ y, long long n) {
  x += (size_t) blockIdx.y * (size_t) n;
  y += (size_t) blockIdx.y * (size_t) n;
  sc += 8 * blockIdx.y;
  const T a = ia >= 0 ? (T) sc[ia] : aconst;
  for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long) gridDim.x * blockDim.x) {
    typename Cplx<T>::type xv = x[k];
    const typename Cplx<T>::type yv = y[k];
    xv.x = a * xv.x + yv.x;
    xv.y = a * xv.y + yv.y;
    x[k] = xv;
  }
}

// partial[b] = sum over the block's elements of w |x|^2 (double accumulation)  vector1.c dot_complex / dot_w_complex
// MODE 0: x as is.  MODE 1 (fused residual update): first x <- x + s * sc[ia] * v, and out <- wout .* x afterwards.
template <typename T, int MODE>
__global__ void dot_kernel(typename Cplx<T>::type *__restrict__ x, const T *__restrict__ w, long long n,
                           double *__restrict__ partial, const double *__restrict__ sc, int ia, T s,
                           const typename Cplx<T>::type *v, typename Cplx<T>::type *out) {   // out may alias v
  __shared__ double red[kVecThreads / 32];
  x += (size_t) blockIdx.y * (size_t) n;
  partial += (size_t) blockIdx.y * kMaxBlocks;
  if (MODE == 1) {
    v += (size_t) blockIdx.y * (size_t) n;
    out += (size_t) blockIdx.y * (size_t) n;
    sc += 8 * blockIdx.y;
  }
  double acc = 0.0;
  T a = 0;
  if (MODE == 1) a = s * (T) sc[ia];
  for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long) gridDim.x * blockDim.x) {
    typename Cplx<T>::type xv = x[k];
    if (MODE == 1) {
      const typename Cplx<T>::type vv = v[k];
      xv.x += a * vv.x; xv.y += a * vv.y;
      x[k] = xv;
    }
    const T wk = w ? w[k] : (T) 1;
    acc += (double) wk * ((double) xv.x * (double) xv.x + (double) xv.y * (double) xv.y);
    if (MODE == 1) {
      typename Cplx<T>::type o = xv;
      if (w) { o.x = wk * xv.x; o.y = wk * xv.y; }
      out[k] = o;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kVecThreads / 32; i++) t += red[i];
    partial[blockIdx.x] = t;
  }
}

// sc[dst] <- sum of the partials rounded to T, after sc[save] <- sc[dst] (save >= 0); then an optional quotient
// sc[q] <- sc[qn] / sc[qd] in T (alpha / beta of the reference).  One block.
template <typename T>
__global__ void dot_final_kernel(const double *__restrict__ partial, int nb, double *__restrict__ sc, int dst, int save,
                                 int q, int qn, int qd, int copy_to) {
  __shared__ double red[kVecThreads / 32];
  partial += (size_t) blockIdx.x * kMaxBlocks;   // one block per right-hand side
  sc += 8 * blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kVecThreads / 32; i++) t += red[i];
    if (save >= 0) sc[save] = sc[dst];
    sc[dst] = (double) (T) t;
    if (copy_to >= 0) sc[copy_to] = sc[dst];
    if (q >= 0) sc[q] = (double) ((T) sc[qn] / (T) sc[qd]);
  }
}

__global__ void quotient_kernel(double *sc, int q, int qn, int qd, int is_float) {
  sc += 8 * blockIdx.x;
  if (is_float) sc[q] = (double) ((float) sc[qn] / (float) sc[qd]);
  else sc[q] = sc[qn] / sc[qd];
}

template <typename T>
struct Ops {
  typedef typename Cplx<T>::type C;
  nfftcu_solver_s *s;
  cudaStream_t st;
  long long N, M;
  int K;
  explicit Ops(nfftcu_solver_s *s_) : s(s_), st(s_->plan->stream), N(s_->plan->N_total), M(s_->plan->M), K(s_->K) {}
  dim3 grid(long long n) const { return dim3(blocks_for(n), (unsigned) K); }
  int trafo(const void *fh, void *f) {
    return K > 1 ? nfftcu_trafo_batch_dev(s->plan, K, fh, f) : nfftcu_trafo_dev(s->plan, fh, f);
  }
  int adjoint(const void *f, void *fh) {
    return K > 1 ? nfftcu_adjoint_batch_dev(s->plan, K, f, fh) : nfftcu_adjoint_dev(s->plan, f, fh);
  }
  C *v(int i) const { return (C *) s->vec[i]; }
  const T *w() const { return (s->flags & PRECOMPUTE_WEIGHT) ? (const T *) s->vec[NFFTCU_SOLVER_W] : nullptr; }
  const T *wh() const { return (s->flags & PRECOMPUTE_DAMP) ? (const T *) s->vec[NFFTCU_SOLVER_W_HAT] : nullptr; }

  void cp_w(C *x, const T *wv, const C *y, long long n) {
    cp_w_kernel<T><<<grid(n), kVecThreads, 0, st>>>(x, wv, y, n);
    s->plan->launches++;
  }
  void xpawy(C *x, int ia, T sgn, const T *wv, const C *y, long long n) {
    xpawy_kernel<T><<<grid(n), kVecThreads, 0, st>>>(x, s->sc, ia, sgn, wv, y, n);
    s->plan->launches++;
  }
  void axpy(C *x, int ia, T aconst, const C *y, long long n) {
    axpy_kernel<T><<<grid(n), kVecThreads, 0, st>>>(x, s->sc, ia, aconst, y, n);
    s->plan->launches++;
  }
  // sc[dst] = sum w |x|^2, with the bookkeeping of dot_final_kernel
  void dot(C *x, const T *wv, long long n, int dst, int save = -1, int q = -1, int qn = -1, int qd = -1, int copy_to = -1) {
    const unsigned nb = blocks_for(n);
    dot_kernel<T, 0><<<dim3(nb, (unsigned) K), kVecThreads, 0, st>>>(x, wv, n, s->partial, nullptr, -1, (T) 0, nullptr, nullptr);
    dot_final_kernel<T><<<K, kVecThreads, 0, st>>>(s->partial, (int) nb, s->sc, dst, save, q, qn, qd, copy_to);
    s->plan->launches += 2;
  }
  // r <- r + sgn * sc[ia] * vv;  sc[dst] = sum w |r|^2;  out <- w .* r
  void upd_dot_cp(C *r, int ia, T sgn, const C *vv, const T *wv, long long n, C *out, int dst, int save = -1, int q = -1,
                  int qn = -1, int qd = -1) {
    const unsigned nb = blocks_for(n);
    dot_kernel<T, 1><<<dim3(nb, (unsigned) K), kVecThreads, 0, st>>>(r, wv, n, s->partial, s->sc, ia, sgn, vv, out);
    dot_final_kernel<T><<<K, kVecThreads, 0, st>>>(s->partial, (int) nb, s->sc, dst, save, q, qn, qd, -1);
    s->plan->launches += 2;
  }
  void quotient(int q, int qn, int qd) {
    quotient_kernel<<<K, 1, 0, st>>>(s->sc, q, qn, qd, sizeof(T) == 4);
    s->plan->launches++;
  }
  int mirror(cudaEvent_t ev, void *host, const void *dev, size_t bytes) {
    if (!host) return NFFTCU_OK;
    NFFTCU_CUDA(cudaEventRecord(ev, st));
    NFFTCU_CUDA(cudaStreamWaitEvent(s->side, ev, 0));
    NFFTCU_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s->side));
    return NFFTCU_OK;
  }
  int finish(double *scal) {
    NFFTCU_CUDA(cudaMemcpyAsync(s->sc_host, s->sc, 8 * sizeof(double) * (size_t) K, cudaMemcpyDeviceToHost, st));
    NFFTCU_CUDA(cudaStreamSynchronize(st));
    NFFTCU_CUDA(cudaStreamSynchronize(s->side));
    NFFTCU_CUDA(cudaGetLastError());
    if (scal) memcpy(scal, s->sc_host, 8 * sizeof(double) * (size_t) K);
    return NFFTCU_OK;
  }

  // r = y - A f_hat;  z = A^H (w r);  norms                                              solver.c:81-125
  int before_loop(void *fhat_host, void *r_host, double *scal) {
    const bool norms = !(s->flags & LANDWEBER) || (s->flags & NORMS_FOR_LANDWEBER);
    C *r = v(NFFTCU_SOLVER_R_ITER), *z = v(NFFTCU_SOLVER_Z_HAT_ITER), *p = v(NFFTCU_SOLVER_P_HAT_ITER);
    NFFTCU_TRY(trafo(v(NFFTCU_SOLVER_F_HAT_ITER), r));
    axpy(r, -1, (T) -1, v(NFFTCU_SOLVER_Y), M);
    if (norms) dot(r, w(), M, SC_DOT_R);
    NFFTCU_TRY(mirror(s->ev_r, r_host, r, sizeof(C) * (size_t) M * (size_t) K));
    cp_w((C *) s->f_in, w(), r, M);
    NFFTCU_TRY(adjoint(s->f_in, z));
    if (norms) dot(z, wh(), N, SC_DOT_Z, -1, -1, -1, -1, (s->flags & CGNE) ? SC_DOT_P : -1);
    if (s->flags & CGNR) cp_w(p, nullptr, z, N);
    (void) fhat_host;
    return finish(scal);
  }

  int step(void *fhat_host, void *r_host, double *scal) {
    C *fh = v(NFFTCU_SOLVER_F_HAT_ITER), *r = v(NFFTCU_SOLVER_R_ITER), *z = v(NFFTCU_SOLVER_Z_HAT_ITER);
    C *p = v(NFFTCU_SOLVER_P_HAT_ITER), *vv = v(NFFTCU_SOLVER_V_ITER), *y = v(NFFTCU_SOLVER_Y);
    C *fhat_in = (C *) s->fhat_in, *f_in = (C *) s->f_in;
    const size_t nb_fh = sizeof(C) * (size_t) N * (size_t) K, nb_r = sizeof(C) * (size_t) M * (size_t) K;
    if (s->flags & LANDWEBER) {                                                         // solver.c:128-174
      NFFTCU_CUDA(cudaMemcpy2DAsync(s->sc + SC_ALPHA, 8 * sizeof(double), scal + SC_ALPHA, 8 * sizeof(double),
                                    sizeof(double), (size_t) K, cudaMemcpyHostToDevice, st));
      xpawy(fh, SC_ALPHA, (T) 1, wh(), z, N);
      NFFTCU_TRY(mirror(s->ev_fhat, fhat_host, fh, nb_fh));
      NFFTCU_TRY(trafo(fh, r));
      axpy(r, -1, (T) -1, y, M);
      if (s->flags & NORMS_FOR_LANDWEBER) dot(r, w(), M, SC_DOT_R);
      NFFTCU_TRY(mirror(s->ev_r, r_host, r, nb_r));
      cp_w(f_in, w(), r, M);
      NFFTCU_TRY(adjoint(f_in, z));
      if (s->flags & NORMS_FOR_LANDWEBER) dot(z, wh(), N, SC_DOT_Z);
    }
    if (s->flags & (STEEPEST_DESCENT | CGNR)) {                                         // solver.c:177-229, 232-292
      C *dir = (s->flags & CGNR) ? p : z;     // search direction
      cp_w(fhat_in, wh(), dir, N);
      NFFTCU_TRY(trafo(fhat_in, vv));
      dot(vv, w(), M, SC_DOT_V, -1, SC_ALPHA, SC_DOT_Z, SC_DOT_V);                       // alpha = dot_z / dot_v
      xpawy(fh, SC_ALPHA, (T) 1, wh(), dir, N);
      NFFTCU_TRY(mirror(s->ev_fhat, fhat_host, fh, nb_fh));
      upd_dot_cp(r, SC_ALPHA, (T) -1, vv, w(), M, f_in, SC_DOT_R);                       // r -= alpha v; dot_r; f_in = w r
      NFFTCU_TRY(mirror(s->ev_r, r_host, r, nb_r));
      NFFTCU_TRY(adjoint(f_in, z));
      if (s->flags & CGNR) {
        dot(z, wh(), N, SC_DOT_Z, SC_DOT_Z_OLD, SC_BETA, SC_DOT_Z, SC_DOT_Z_OLD);        // beta = dot_z / dot_z_old
        axpy(p, SC_BETA, (T) 0, z, N);
      } else {
        dot(z, wh(), N, SC_DOT_Z);
      }
    }
    if (s->flags & CGNE) {                                                              // solver.c:295-344
      quotient(SC_ALPHA, SC_DOT_R, SC_DOT_P);
      xpawy(fh, SC_ALPHA, (T) 1, wh(), p, N);
      NFFTCU_TRY(mirror(s->ev_fhat, fhat_host, fh, nb_fh));
      cp_w(fhat_in, wh(), p, N);
      NFFTCU_TRY(trafo(fhat_in, f_in));
      // r -= alpha (A w_hat p); dot_r_old = dot_r; dot_r; beta = dot_r / dot_r_old; f_in = w r (in place: out == v)
      upd_dot_cp(r, SC_ALPHA, (T) -1, f_in, w(), M, f_in, SC_DOT_R, SC_DOT_R_OLD, SC_BETA, SC_DOT_R, SC_DOT_R_OLD);
      NFFTCU_TRY(mirror(s->ev_r, r_host, r, nb_r));
      NFFTCU_TRY(adjoint(f_in, fhat_in));
      axpy(p, SC_BETA, (T) 0, fhat_in, N);
      dot(p, wh(), N, SC_DOT_P);
    }
    return finish(scal);
  }
};

size_t vec_bytes(const nfftcu_solver_s *s, int which) {   // the weights are shared by all right-hand sides
  const size_t r = real_size(s->plan), K = (size_t) s->K;
  switch (which) {
    case NFFTCU_SOLVER_W: return r * (size_t) s->plan->M;
    case NFFTCU_SOLVER_W_HAT: return r * (size_t) s->plan->N_total;
    case NFFTCU_SOLVER_Y: case NFFTCU_SOLVER_R_ITER: case NFFTCU_SOLVER_V_ITER: return 2 * r * (size_t) s->plan->M * K;
    default: return 2 * r * (size_t) s->plan->N_total * K;
  }
}

}  // namespace
}  // namespace nfftcu

using namespace nfftcu;

extern "C" {

int nfftcu_solver_create(nfftcu_solver **out, nfftcu_ctx *plan, unsigned flags) {
  return nfftcu_solver_create_batch(out, plan, flags, 1);
}

int nfftcu_solver_create_batch(nfftcu_solver **out, nfftcu_ctx *plan, unsigned flags, int K) {
  if (!out || !plan || K < 1) { set_error("nfftcu_solver_create: null argument or K < 1"); return NFFTCU_EINVAL; }
  const unsigned methods = flags & (LANDWEBER | STEEPEST_DESCENT | CGNR | CGNE);
  if (methods == 0 || (methods & (methods - 1))) {
    set_error("nfftcu_solver_create: exactly one of LANDWEBER, STEEPEST_DESCENT, CGNR, CGNE must be set (flags 0x%x)", flags);
    return NFFTCU_EINVAL;
  }
  NFFTCU_CUDA(cudaSetDevice(plan->device));
  nfftcu_solver_s *s = new nfftcu_solver_s();
  s->plan = plan;
  s->device = plan->device;   // destroy must not touch the plan: nfft_finalize before solver_finalize is legal (solver.c:373-389)
  s->K = K;
  s->flags = flags;
  auto fail = [&](int rc) { nfftcu_solver_destroy(s); return rc; };
#define SOLVER_ALLOC(ptr, bytes)                                                                    \
  do {                                                                                              \
    if (pool_malloc(&(ptr), (bytes) ? (bytes) : 16) != cudaSuccess) {                                \
      set_error("nfftcu_solver_create: cudaMalloc of %zu bytes failed", (size_t) (bytes));          \
      return fail(NFFTCU_ENOMEM);                                                                   \
    }                                                                                               \
  } while (0)
  const int need[] = {NFFTCU_SOLVER_Y, NFFTCU_SOLVER_F_HAT_ITER, NFFTCU_SOLVER_R_ITER, NFFTCU_SOLVER_P_HAT_ITER};
  for (int which : need) SOLVER_ALLOC(s->vec[which], vec_bytes(s, which));
  if (flags & CGNR) SOLVER_ALLOC(s->vec[NFFTCU_SOLVER_Z_HAT_ITER], vec_bytes(s, NFFTCU_SOLVER_Z_HAT_ITER));
  else s->vec[NFFTCU_SOLVER_Z_HAT_ITER] = s->vec[NFFTCU_SOLVER_P_HAT_ITER];   // solver.c:52-69: z aliases p
  if (flags & (CGNR | STEEPEST_DESCENT)) SOLVER_ALLOC(s->vec[NFFTCU_SOLVER_V_ITER], vec_bytes(s, NFFTCU_SOLVER_V_ITER));
  if (flags & PRECOMPUTE_WEIGHT) SOLVER_ALLOC(s->vec[NFFTCU_SOLVER_W], vec_bytes(s, NFFTCU_SOLVER_W));
  if (flags & PRECOMPUTE_DAMP) SOLVER_ALLOC(s->vec[NFFTCU_SOLVER_W_HAT], vec_bytes(s, NFFTCU_SOLVER_W_HAT));
  SOLVER_ALLOC(s->fhat_in, vec_bytes(s, NFFTCU_SOLVER_F_HAT_ITER));
  SOLVER_ALLOC(s->f_in, vec_bytes(s, NFFTCU_SOLVER_Y));
  SOLVER_ALLOC(s->sc, 8 * sizeof(double) * (size_t) K);
  SOLVER_ALLOC(s->partial, kMaxBlocks * sizeof(double) * (size_t) K);
#undef SOLVER_ALLOC
  if (cudaMemset(s->sc, 0, 8 * sizeof(double) * (size_t) K) != cudaSuccess ||
      pool_malloc_host(&s->sc_host, 8 * sizeof(double) * (size_t) K) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_fhat, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_r, cudaEventDisableTiming) != cudaSuccess) {
    set_error("nfftcu_solver_create: stream / event / pinned allocation failed");
    return fail(NFFTCU_ECUDA);
  }
  *out = s;
  return NFFTCU_OK;
}

int nfftcu_solver_destroy(nfftcu_solver *s) {
  if (!s) return NFFTCU_OK;
  cudaSetDevice(s->device);
  if (s->side) { cudaStreamSynchronize(s->side); cudaStreamDestroy(s->side); }
  if (s->ev_fhat) cudaEventDestroy(s->ev_fhat);
  if (s->ev_r) cudaEventDestroy(s->ev_r);
  if (s->vec[NFFTCU_SOLVER_Z_HAT_ITER] == s->vec[NFFTCU_SOLVER_P_HAT_ITER]) s->vec[NFFTCU_SOLVER_Z_HAT_ITER] = nullptr;
  for (void *&p : s->vec) { if (p) pool_free(p); p = nullptr; }
  if (s->fhat_in) pool_free(s->fhat_in);
  if (s->f_in) pool_free(s->f_in);
  if (s->sc) pool_free(s->sc);
  if (s->partial) pool_free(s->partial);
  if (s->sc_host) pool_free_host(s->sc_host);
  delete s;
  return NFFTCU_OK;
}

int nfftcu_solver_upload(nfftcu_solver *s, int which, const void *host) {
  if (!s || which < 0 || which > 7 || !s->vec[which] || !host) {
    set_error("nfftcu_solver_upload: vector %d is not part of this solver (flags 0x%x)", which, s ? s->flags : 0);
    return NFFTCU_EINVAL;
  }
  NFFTCU_CUDA(cudaSetDevice(s->plan->device));
  NFFTCU_CUDA(cudaMemcpyAsync(s->vec[which], host, vec_bytes(s, which), cudaMemcpyHostToDevice, s->plan->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(s->plan->stream));
  return NFFTCU_OK;
}

int nfftcu_solver_download(nfftcu_solver *s, int which, void *host) {
  if (!s || which < 0 || which > 7 || !s->vec[which] || !host) {
    set_error("nfftcu_solver_download: vector %d is not part of this solver (flags 0x%x)", which, s ? s->flags : 0);
    return NFFTCU_EINVAL;
  }
  NFFTCU_CUDA(cudaSetDevice(s->plan->device));
  NFFTCU_CUDA(cudaMemcpyAsync(host, s->vec[which], vec_bytes(s, which), cudaMemcpyDeviceToHost, s->plan->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(s->plan->stream));
  return NFFTCU_OK;
}

void *nfftcu_solver_vector(nfftcu_solver *s, int which) { return (s && which >= 0 && which <= 7) ? s->vec[which] : nullptr; }

int nfftcu_solver_before_loop(nfftcu_solver *s, void *f_hat_iter_host, void *r_iter_host, double scal[8]) {
  if (!s) { set_error("nfftcu_solver_before_loop: null solver"); return NFFTCU_EINVAL; }
  NFFTCU_CUDA(cudaSetDevice(s->plan->device));
  if (s->plan->prec == NFFTCU_DOUBLE) return Ops<double>(s).before_loop(f_hat_iter_host, r_iter_host, scal);
  return Ops<float>(s).before_loop(f_hat_iter_host, r_iter_host, scal);
}

int nfftcu_solver_step(nfftcu_solver *s, void *f_hat_iter_host, void *r_iter_host, double scal[8]) {
  if (!s || !scal) { set_error("nfftcu_solver_step: null argument"); return NFFTCU_EINVAL; }
  NFFTCU_CUDA(cudaSetDevice(s->plan->device));
  if (s->plan->prec == NFFTCU_DOUBLE) return Ops<double>(s).step(f_hat_iter_host, r_iter_host, scal);
  return Ops<float>(s).step(f_hat_iter_host, r_iter_host, scal);
}

}  // extern "C"
