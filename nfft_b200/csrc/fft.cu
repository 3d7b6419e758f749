// fft.cu -- F step: the unnormalised d-dimensional complex DFT over the oversampled grid.
//
// Replaces the FFTW plan pair of the reference (kernel/nfft/nfft.c:6030-6031: FORWARD g1->g2,
// BACKWARD g2->g1) and their execution (5516 trafo, 5557 adjoint):
//     out[k] = sum_l in[l] exp(sign 2 pi i <k,l>/n),  sign=-1 trafo, +1 adjoint, in place.
// Hand-written: one pass per axis; a CTA stages a bundle of lines of that axis in shared memory,
// runs a Stockham autosort radix-4 (+ one radix-2 when log2 is odd) between two shared buffers
// and writes the lines back.  For strided axes the bundle is a run of neighbouring lines so that
// every global access is a contiguous segment of bundle*sizeof(complex) bytes.
// Twiddles come from a per-axis table exp(-2 pi i q/len) computed on the host in long double.
// Lengths that are not a power of two use an O(len^2) table DFT in shared memory (correct for
// any length that fits; mixed radix / Bluestein are the planned replacement).
// HBM-bound: algorithmic bytes 2 * C * n_total per axis pass.
//
// Pruned passes.  Inside a transform the oversampled spectrum is non-zero (trafo: after D) or needed
// (adjoint: before D^T) only in the band k_t in [-N_t/2, N_t/2), i.e. at indices [0, N_t - N_t/2) and
// [n_t - N_t/2, n_t) of every axis.  The forward transform therefore runs the axes last to first, and the
// pass over axis t (a) only visits lines whose coordinates in the axes before t lie in the band -- all
// other lines are still zero -- and (b) loads only the band elements of a line, taking the rest as zero
// WITHOUT reading them, so D never has to write the zero padding.  The backward transform runs first to
// last, visits the same lines and stores only the band elements.  At sigma = 2 in 3-D this is
// (1/4 + 1/2 + 1)/3 of the lines and 2.5x less HBM traffic for D + F.
#include "common.cuh"

#include <math.h>

namespace nfftcu {

namespace {

constexpr int kFftThreads = 256;
constexpr size_t kSmemBudget = 200 * 1024;

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}

struct LineGeom {
  long long len;      // transform length L
  long long inner;    // element stride of the axis (product of faster dims)
  long long lines;    // number of lines = total / len
  int bundle;         // lines per CTA
  long long bundles_inner;  // ceil(inner / bundle) when inner > 1
  int pitch;          // shared-memory pitch of one line (complex elements)
  // pruning (see the header): 0 none | 1 forward: band lines, band loads | 2 backward: band lines, band stores
  int prune;
  int nouter;                       // axes before this one
  long long oN[NFFTCU_MAX_D];       // their band sizes N_s
  long long olow[NFFTCU_MAX_D];     // ... of which the first olow[s] map to l = c, the rest to l = c + n_s - N_s
  long long on[NFFTCU_MAX_D];       // ... and full lengths n_s
  long long elow, ehigh;            // band of this axis: e < elow || e >= ehigh
};

// compact (band) index over the axes before t -> row-major index over their full lengths
__device__ __forceinline__ long long map_outer(const LineGeom &g, long long oc) {
  if (g.prune == 0) return oc;
  long long o = 0, mul = 1;
  for (int s = g.nouter - 1; s >= 0; s--) {
    const long long dgt = oc % g.oN[s];
    oc /= g.oN[s];
    o += (dgt < g.olow[s] ? dgt : dgt + (g.on[s] - g.oN[s])) * mul;
    mul *= g.on[s];
  }
  return o;
}

// global <-> shared staging shared by both kernels.  Line c of bundle b:
//   inner == 1 : line index q = b*bundle + c,              element e at q*len + e
//   inner  > 1 : (o, i) = (b / bundles_inner, (b % bundles_inner)*bundle + c),
//                element e at (o*len + e)*inner + i
template <typename C, bool STORE>
__device__ __forceinline__ void stage_lines(C *__restrict__ data, C *__restrict__ sm,
                                            const LineGeom &g, long long b) {
  const int L = (int) g.len;
  const bool band_only = STORE ? g.prune == 2 : g.prune == 1;
  __shared__ long long line_base[64];
  if (g.inner == 1) {
    const long long q0 = b * g.bundle;
    const int cnt = (int) min((long long) g.bundle, g.lines - q0);
    if (threadIdx.x < cnt) line_base[threadIdx.x] = map_outer(g, q0 + threadIdx.x) * g.len;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * L; i += blockDim.x) {
      const int cl = i / L, e = i - cl * L;
      const bool in_band = !band_only || e < g.elow || e >= g.ehigh;
      if (STORE) { if (in_band) data[line_base[cl] + e] = sm[cl * g.pitch + e]; }
      else sm[cl * g.pitch + e] = in_band ? data[line_base[cl] + e] : C{0, 0};
    }
  } else {
    const long long oc = b / g.bundles_inner;
    const long long i0 = (b - oc * g.bundles_inner) * g.bundle;
    const int cnt = (int) min((long long) g.bundle, g.inner - i0);
    C *base = data + map_outer(g, oc) * g.len * g.inner + i0;
    for (int i = threadIdx.x; i < cnt * L; i += blockDim.x) {
      const int e = i / cnt, cl = i - e * cnt;
      const bool in_band = !band_only || e < g.elow || e >= g.ehigh;
      if (STORE) { if (in_band) base[(long long) e * g.inner + cl] = sm[cl * g.pitch + e]; }
      else sm[cl * g.pitch + e] = in_band ? base[(long long) e * g.inner + cl] : C{0, 0};
    }
  }
}

template <typename C>
__device__ __forceinline__ int bundle_count(const LineGeom &g, long long b) {
  if (g.inner == 1) return (int) min((long long) g.bundle, g.lines - b * g.bundle);
  const long long o = b / g.bundles_inner;
  return (int) min((long long) g.bundle, g.inner - (b - o * g.bundles_inner) * g.bundle);
}

// ---- power-of-two lengths: Stockham autosort in shared memory -----------------------------------
template <typename T>
__global__ void __launch_bounds__(kFftThreads)
fft_stockham_kernel(typename Cplx<T>::type *__restrict__ data,
                    const typename Cplx<T>::type *__restrict__ tw, LineGeom g, int sign,
                    int log2len) {
  typedef typename Cplx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *buf0 = reinterpret_cast<C *>(smem_raw);
  C *buf1 = buf0 + (size_t) g.bundle * g.pitch;
  const int L = (int) g.len;
  const long long b = blockIdx.x;
  const int cnt = bundle_count<C>(g, b);

  stage_lines<C, false>(data, buf0, g, b);
  __syncthreads();

  C *src = buf0, *dst = buf1;
  int Ns = 1;
  if (log2len & 1) {   // one radix-2 stage first
    const int half = L >> 1;
    for (int w = threadIdx.x; w < cnt * half; w += blockDim.x) {
      const int cl = w / half, j = w - cl * half;
      const C a = src[cl * g.pitch + j], bb = src[cl * g.pitch + j + half];
      dst[cl * g.pitch + 2 * j] = cadd(a, bb);
      dst[cl * g.pitch + 2 * j + 1] = csub(a, bb);
    }
    __syncthreads();
    C *t = src; src = dst; dst = t;
    Ns = 2;
  }
  const int quarter = L >> 2;
  for (; Ns < L; Ns <<= 2) {
    const int tstep = L / (Ns * 4);   // table stride: W_{4Ns} = W_L^tstep
    for (int w = threadIdx.x; w < cnt * quarter; w += blockDim.x) {
      const int cl = w / quarter, j = w - cl * quarter;
      const int k = j & (Ns - 1);
      const C *in = src + cl * g.pitch + j;
      C v0 = in[0], v1 = in[quarter], v2 = in[2 * quarter], v3 = in[3 * quarter];
      if (k) {
        C w1 = tw[k * tstep], w2 = tw[2 * k * tstep], w3 = tw[3 * k * tstep];
        if (sign > 0) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        v1 = cmul(v1, w1);
        v2 = cmul(v2, w2);
        v3 = cmul(v3, w3);
      }
      const C t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3);
      C t3 = csub(v1, v3);
      // multiply by sign*i: forward (-i): (x,y)->(y,-x); backward (+i): (x,y)->(-y,x)
      { const T x = t3.x, y = t3.y; if (sign < 0) { t3.x = y; t3.y = -x; } else { t3.x = -y; t3.y = x; } }
      C *out = dst + cl * g.pitch + ((j - k) << 2) + k;
      out[0] = cadd(t0, t2);
      out[Ns] = cadd(t1, t3);
      out[2 * Ns] = csub(t0, t2);
      out[3 * Ns] = csub(t1, t3);
    }
    __syncthreads();
    C *t = src; src = dst; dst = t;
  }
  stage_lines<C, true>(data, src, g, b);
}

// ---- any length that fits: O(len^2) DFT from the twiddle table ---------------------------------
template <typename T>
__global__ void __launch_bounds__(kFftThreads)
dft_table_kernel(typename Cplx<T>::type *__restrict__ data,
                 const typename Cplx<T>::type *__restrict__ tw, LineGeom g, int sign) {
  typedef typename Cplx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *buf0 = reinterpret_cast<C *>(smem_raw);
  C *buf1 = buf0 + (size_t) g.bundle * g.pitch;
  const int L = (int) g.len;
  const long long b = blockIdx.x;
  const int cnt = bundle_count<C>(g, b);
  stage_lines<C, false>(data, buf0, g, b);
  __syncthreads();
  for (int w = threadIdx.x; w < cnt * L; w += blockDim.x) {
    const int cl = w / L, k = w - cl * L;
    const C *in = buf0 + cl * g.pitch;
    // accumulate in double for both precisions: len products of O(1) terms
    double sr = 0.0, si = 0.0;
    int q = 0;
    for (int l = 0; l < L; l++) {
      C wv = tw[q];
      if (sign > 0) wv.y = -wv.y;
      sr += (double) in[l].x * (double) wv.x - (double) in[l].y * (double) wv.y;
      si += (double) in[l].x * (double) wv.y + (double) in[l].y * (double) wv.x;
      q += k;
      if (q >= L) q -= L;
    }
    C o;
    o.x = (T) sr;
    o.y = (T) si;
    buf1[cl * g.pitch + k] = o;
  }
  __syncthreads();
  stage_lines<C, true>(data, buf1, g, b);
}

int ilog2_exact(long long v) {
  int l = 0;
  while ((1ll << l) < v) l++;
  return ((1ll << l) == v) ? l : -1;
}

template <typename T>
int run_axis(nfftcu_ctx *c, int t, int sign, bool pruned) {
  typedef typename Cplx<T>::type C;
  const FftAxis &ax = c->fft[t];
  if (ax.kind == 0) return NFFTCU_OK;
  LineGeom g;
  g.len = ax.len;
  g.inner = 1;
  for (int t2 = t + 1; t2 < c->d; t2++) g.inner *= c->n[t2];
  g.lines = c->n_total / ax.len;
  g.prune = pruned ? (sign < 0 ? 1 : 2) : 0;
  g.nouter = t;
  long long outer = 1;
  for (int s2 = 0; s2 < t; s2++) {
    g.oN[s2] = pruned ? c->N[s2] : c->n[s2];
    g.olow[s2] = pruned ? c->N[s2] - c->N[s2] / 2 : c->n[s2];
    g.on[s2] = c->n[s2];
    outer *= g.oN[s2];
  }
  g.elow = c->N[t] - c->N[t] / 2;
  g.ehigh = c->n[t] - c->N[t] / 2;
  g.lines = outer * g.inner;
  const int pad = 1;
  g.pitch = (int) ax.len + pad;
  const size_t line_bytes = 2 * (size_t) g.pitch * sizeof(C);   // two buffers
  int pref = (int) (128 / sizeof(C));                           // 128-byte global segments
  if (g.inner == 1) pref = (int) max(1ll, min(8ll, 2048ll / ax.len));
  int fit = (int) (kSmemBudget / line_bytes);
  if (fit < 1) {
    set_error("FFT axis %d: length %lld does not fit the shared-memory kernel", t, (long long) ax.len);
    return NFFTCU_EINVAL;
  }
  g.bundle = pref < fit ? pref : fit;
  if (g.inner > 1 && (long long) g.bundle > g.inner) g.bundle = (int) g.inner;
  g.bundles_inner = g.inner == 1 ? 1 : (g.inner + g.bundle - 1) / g.bundle;
  if (g.bundle > 64) g.bundle = 64;   // line_base[] of stage_lines
  const long long nb = g.inner == 1 ? (g.lines + g.bundle - 1) / g.bundle
                                    : (g.lines / g.inner) * g.bundles_inner;
  const size_t smem = line_bytes * g.bundle;
  if (ax.kind == 1) {
    NFFTCU_CUDA(cudaFuncSetAttribute(fft_stockham_kernel<T>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    fft_stockham_kernel<T><<<(unsigned) nb, kFftThreads, smem, c->stream>>>(
        (C *) c->grid, (const C *) ax.tw, g, sign, ilog2_exact(ax.len));
  } else {
    NFFTCU_CUDA(cudaFuncSetAttribute(dft_table_kernel<T>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    dft_table_kernel<T><<<(unsigned) nb, kFftThreads, smem, c->stream>>>((C *) c->grid,
                                                                         (const C *) ax.tw, g, sign);
  }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

}  // namespace

int fft_plan_axes(nfftcu_ctx *c) {
  const long double two_pi = 6.283185307179586476925286766559005768394L;
  for (int t = 0; t < c->d; t++) {
    FftAxis &ax = c->fft[t];
    ax.len = c->n[t];
    if (ax.len == 1) { ax.kind = 0; continue; }
    ax.kind = ilog2_exact(ax.len) >= 0 ? 1 : 2;
    const size_t esz = 2 * real_size(c);
    std::vector<unsigned char> host(esz * (size_t) ax.len);
    for (long long q = 0; q < ax.len; q++) {
      const long double ang = two_pi * (long double) q / (long double) ax.len;
      const long double cr = cosl(ang), ci = -sinl(ang);
      if (c->prec == NFFTCU_DOUBLE) {
        ((double *) host.data())[2 * q] = (double) cr;
        ((double *) host.data())[2 * q + 1] = (double) ci;
      } else {
        ((float *) host.data())[2 * q] = (float) cr;
        ((float *) host.data())[2 * q + 1] = (float) ci;
      }
    }
    NFFTCU_CUDA(cudaMalloc(&ax.tw, host.size()));
    NFFTCU_CUDA(cudaMemcpy(ax.tw, host.data(), host.size(), cudaMemcpyHostToDevice));
  }
  return NFFTCU_OK;
}

void fft_free_axes(nfftcu_ctx *c) {
  for (int t = 0; t < c->d; t++) {
    if (c->fft[t].tw) cudaFree(c->fft[t].tw);
    c->fft[t].tw = nullptr;
  }
}

// pruned: the caller guarantees the band structure described in the header (trafo: grid written by the
// sparse D; adjoint: only the band of the result is read by D^T)
int stage_F(nfftcu_ctx *c, int sign, bool pruned) {
  // forward (and every unpruned transform): last axis first; pruned backward: first axis first
  const bool last_first = !(pruned && sign > 0);
  for (int i = 0; i < c->d; i++) {
    const int t = last_first ? c->d - 1 - i : i;
    if (c->prec == NFFTCU_DOUBLE) NFFTCU_TRY(run_axis<double>(c, t, sign, pruned));
    else NFFTCU_TRY(run_axis<float>(c, t, sign, pruned));
  }
  return NFFTCU_OK;
}

}  // namespace nfftcu
