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
// Lengths that are not a power of two run a mixed-radix Stockham (radix 4, 2, 3, 5, 7 in registers, any
// other prime factor <= 61 as a direct r-point stage); a length with a larger prime factor runs Bluestein's
// chirp-z through a power-of-two convolution inside the CTA (fft_bluestein_kernel) when the padded length fits
// (len <= 2048 fp64 / 4096 fp32), else the O(len^2) table DFT in shared memory.  An axis too long for one CTA's shared memory is split len = L1 * L2
// (four-step): L1-point transforms down the columns with the twiddle W_len^(l2 k1) fused into their store
// (two-level table, one extra complex multiply), then L2-point transforms along the rows whose store
// transposes into a second grid buffer; the two buffers are swapped afterwards.
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
#include <string.h>

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
  long long icount;   // lines per outer index: inner, or the number of band lines when the inner axes are pruned
  long long lines;    // number of lines visited = outer count * icount
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
  // slab mode (multi-GPU node slabs, see run_axis_slab): a plane window [w0, w0 + wc) mod n_0 of the first axis.
  //   ooff[s]: offset added (mod on[s]) to the mapped outer coordinate -- the window start when axis s is the first axis
  //   wload / wstore: on the transform axis itself, elements outside the window are zero (not read) / not needed
  //   iprune: the axes BEHIND this one are still band-limited: only the lines whose inner coordinates lie in the band
  //           are visited (compact inner index -> element offset through iN / ilow / in, last inner axis fastest)
  long long ooff[NFFTCU_MAX_D];
  int wload, wstore;
  long long w0, wc;
  int iprune, niax;
  long long iN[NFFTCU_MAX_D], ilow[NFFTCU_MAX_D], in_[NFFTCU_MAX_D];
  long long ialign;   // iprune: a bundle must divide this (both band halves of the last inner axis), host side only
  // four-step split of a long axis (see the header).  twist: this pass is the column pass, element e (= k1) of
  // the line with inner index i is multiplied by W_Ltot^((i / tw_I) * e) = twA[q >> tw_shift] * twB[q & mask]
  // at the store.  tstore: this pass is the row pass over [o][k1][e = k2][i], stored as [o][k2][k1][i].
  int twist, tstore;
  int tw_shift;
  long long tw_I;
  const void *twA, *twB;
  long long tL1, tL2, tI;
};

struct RadixPlan {
  int nst;
  int radix[24];
};

// compact (band) index over the axes before t -> row-major index over their full lengths
__device__ __forceinline__ long long map_outer(const LineGeom &g, long long oc) {
  if (g.prune == 0) return oc;
  long long o = 0, mul = 1;
  for (int s = g.nouter - 1; s >= 0; s--) {
    const long long dgt = oc % g.oN[s];
    oc /= g.oN[s];
    long long dr = (dgt < g.olow[s] ? dgt : dgt + (g.on[s] - g.oN[s])) + g.ooff[s];
    if (dr >= g.on[s]) dr -= g.on[s];
    o += dr * mul;
    mul *= g.on[s];
  }
  return o;
}

// compact index over the band lines of the inner axes -> element offset (stride 1 on the last axis)
__device__ __forceinline__ long long map_inner(const LineGeom &g, long long ic) {
  if (!g.iprune) return ic;
  long long i = 0, mul = 1;
  for (int s = g.niax - 1; s >= 0; s--) {
    const long long dgt = ic % g.iN[s];
    ic /= g.iN[s];
    i += (dgt < g.ilow[s] ? dgt : dgt + (g.in_[s] - g.iN[s])) * mul;
    mul *= g.in_[s];
  }
  return i;
}

__device__ __forceinline__ bool in_window(const LineGeom &g, int e) {
  long long ee = (long long) e - g.w0;
  if (ee < 0) ee += g.len;
  return ee < g.wc;
}
// element e of a line: does it have to be read (else it is zero) / written (else nobody reads it)?
__device__ __forceinline__ bool keep_load(const LineGeom &g, int e) {
  if (g.prune == 1 && !(e < g.elow || e >= g.ehigh)) return false;
  if (g.wload && !in_window(g, e)) return false;
  return true;
}
__device__ __forceinline__ bool keep_store(const LineGeom &g, int e) {
  if (g.prune == 2 && !(e < g.elow || e >= g.ehigh)) return false;
  if (g.wstore && !in_window(g, e)) return false;
  return true;
}

// global <-> shared staging shared by both kernels.  Line c of bundle b:
//   inner == 1 : line index q = b*bundle + c,              element e at q*len + e
//   inner  > 1 : (o, i) = (b / bundles_inner, (b % bundles_inner)*bundle + c),
//                element e at (o*len + e)*inner + i
template <typename C>
__device__ __forceinline__ C twist_factor(const LineGeom &g, long long l2, int e, int sign) {
  const long long q = l2 * e;
  const C a = reinterpret_cast<const C *>(g.twA)[q >> g.tw_shift];
  const C bq = reinterpret_cast<const C *>(g.twB)[q & ((1ll << g.tw_shift) - 1)];
  C w = cmul(a, bq);
  if (sign > 0) w.y = -w.y;
  return w;
}

template <typename C, bool STORE>
__device__ __forceinline__ void stage_lines(C *__restrict__ data, C *__restrict__ sm,
                                            const LineGeom &g, long long b, int sign = 0) {
  const int L = (int) g.len;
  __shared__ long long line_base[64];
  if (STORE && g.tstore) {
    // row pass of a split axis: line (oc = o * L1 + k1, i), element e = k2 -> [o][k2][k1][i]; the bundle runs
    // along k1 (inner == 1) or i (inner > 1), so consecutive cl are consecutive addresses either way
    long long oc, i0;
    int cnt;
    if (g.inner == 1) {
      oc = b * g.bundle; i0 = 0;
      cnt = (int) min((long long) g.bundle, g.lines - oc);
    } else {
      oc = b / g.bundles_inner;
      i0 = (b - oc * g.bundles_inner) * g.bundle;
      cnt = (int) min((long long) g.bundle, g.icount - i0);
    }
    for (int i = threadIdx.x; i < cnt * L; i += blockDim.x) {
      const int e = i / cnt, cl = i - e * cnt;
      const long long ocl = g.inner == 1 ? oc + cl : oc;
      const long long o = ocl / g.tL1, k1 = ocl - o * g.tL1;
      const long long ii = g.inner == 1 ? 0 : i0 + cl;
      data[((o * g.tL2 + e) * g.tL1 + k1) * g.tI + ii] = sm[cl * g.pitch + e];
    }
    return;
  }
  if (g.inner == 1) {
    const long long q0 = b * g.bundle;
    const int cnt = (int) min((long long) g.bundle, g.lines - q0);
    if (threadIdx.x < cnt) line_base[threadIdx.x] = map_outer(g, q0 + threadIdx.x) * g.len;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * L; i += blockDim.x) {
      const int cl = i / L, e = i - cl * L;
      const bool in_band = STORE ? keep_store(g, e) : keep_load(g, e);
      if (STORE) { if (in_band) data[line_base[cl] + e] = sm[cl * g.pitch + e]; }
      else sm[cl * g.pitch + e] = in_band ? data[line_base[cl] + e] : C{0, 0};
    }
  } else {
    const long long oc = b / g.bundles_inner;
    const long long i0 = (b - oc * g.bundles_inner) * g.bundle;
    const int cnt = (int) min((long long) g.bundle, g.icount - i0);
    C *base = data + map_outer(g, oc) * g.len * g.inner + map_inner(g, i0);   // a bundle never straddles a band gap
    for (int i = threadIdx.x; i < cnt * L; i += blockDim.x) {
      const int e = i / cnt, cl = i - e * cnt;
      const bool in_band = STORE ? keep_store(g, e) : keep_load(g, e);
      if (STORE) {
        C v = sm[cl * g.pitch + e];
        if (g.twist && e) v = cmul(v, twist_factor<C>(g, (i0 + cl) / g.tw_I, e, sign));
        if (in_band) base[(long long) e * g.inner + cl] = v;
      }
      else sm[cl * g.pitch + e] = in_band ? base[(long long) e * g.inner + cl] : C{0, 0};
    }
  }
}

template <typename C>
__device__ __forceinline__ int bundle_count(const LineGeom &g, long long b) {
  if (g.inner == 1) return (int) min((long long) g.bundle, g.lines - b * g.bundle);
  const long long o = b / g.bundles_inner;
  return (int) min((long long) g.bundle, g.icount - (b - o * g.bundles_inner) * g.bundle);
}

// ---- power-of-two lengths: Stockham autosort in shared memory -----------------------------------
// radix-4 (+ one radix-2) Stockham autosort of `cnt` lines of length L = 2^log2len between two shared buffers;
// returns the buffer that holds the result.  tw = exp(-2 pi i q / L).
template <typename T>
__device__ __forceinline__ typename Cplx<T>::type *pow2_stockham_smem(typename Cplx<T>::type *buf0,
                                                                      typename Cplx<T>::type *buf1,
                                                                      const typename Cplx<T>::type *__restrict__ tw,
                                                                      int L, int log2len, int cnt, int pitch, int sign) {
  typedef typename Cplx<T>::type C;
  C *src = buf0, *dst = buf1;
  int Ns = 1;
  if (log2len & 1) {   // one radix-2 stage first
    const int half = L >> 1;
    for (int w = threadIdx.x; w < cnt * half; w += blockDim.x) {
      const int cl = w / half, j = w - cl * half;
      const C a = src[cl * pitch + j], bb = src[cl * pitch + j + half];
      dst[cl * pitch + 2 * j] = cadd(a, bb);
      dst[cl * pitch + 2 * j + 1] = csub(a, bb);
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
      const C *in = src + cl * pitch + j;
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
      C *out = dst + cl * pitch + ((j - k) << 2) + k;
      out[0] = cadd(t0, t2);
      out[Ns] = cadd(t1, t3);
      out[2 * Ns] = csub(t0, t2);
      out[3 * Ns] = csub(t1, t3);
    }
    __syncthreads();
    C *t = src; src = dst; dst = t;
  }
  return src;
}

template <typename T>
__global__ void __launch_bounds__(kFftThreads)
fft_stockham_kernel(typename Cplx<T>::type *__restrict__ data, typename Cplx<T>::type *__restrict__ out_data,
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

  C *src = pow2_stockham_smem<T>(buf0, buf1, tw, L, log2len, cnt, g.pitch, sign);
  stage_lines<C, true>(out_data, src, g, b, sign);
}

// ---- power-of-two lengths 64 ... 2048: register-resident Stockham ---------------------------------------
// A line of L = 16 * TPL points is owned by TPL threads, 16 points per thread in registers for the whole
// transform.  L = R1 * R2 (* R3) with radices 4, 8, 16: a step with radix R does 16 / R butterflies per thread
// entirely in registers (DFT8 = 2 x 4, DFT16 = 4 x 4 with constant twiddles), the inter-step twiddles come from the
// per-axis table, and steps exchange data through ONE shared-memory buffer (one write + one read per element and
// exchange, against two per radix-4 stage in fft_stockham_kernel).  The Stockham index maps make every thread's
// global loads (first step) and stores (last step) the elements t + TPL * e, e < 16: for strided axes the lanes run
// along the bundle of neighbouring lines (128-byte segments), for the contiguous axis along t.
// Element i of a line sits at i + i / 16 of its shared-memory row (conflict-free 16-byte quarter-warp accesses), rows
// are pitched to an odd multiple of 16 bytes.
template <typename C, int SIGN> __device__ __forceinline__ C mul_i(C a) {   // a * (SIGN * i)
  C r;
  if (SIGN < 0) { r.x = a.y; r.y = -a.x; } else { r.x = -a.y; r.y = a.x; }
  return r;
}

// exp(SIGN * 2 pi i e / 16), e compile-time after unrolling
template <typename T, int SIGN> __device__ __forceinline__ typename Cplx<T>::type w16(int e) {
  const T c1 = (T) 0.92387953251128675613, c2 = (T) 0.70710678118654752440, c3 = (T) 0.38268343236508977173;
  T cr, ci;
  switch (e & 15) {
    case 0: cr = 1; ci = 0; break;
    case 1: cr = c1; ci = c3; break;
    case 2: cr = c2; ci = c2; break;
    case 3: cr = c3; ci = c1; break;
    case 4: cr = 0; ci = 1; break;
    case 5: cr = -c3; ci = c1; break;
    case 6: cr = -c2; ci = c2; break;
    case 7: cr = -c1; ci = c3; break;
    case 8: cr = -1; ci = 0; break;
    case 9: cr = -c1; ci = -c3; break;
    case 10: cr = -c2; ci = -c2; break;
    case 11: cr = -c3; ci = -c1; break;
    case 12: cr = 0; ci = -1; break;
    case 13: cr = c3; ci = -c1; break;
    case 14: cr = c2; ci = -c2; break;
    default: cr = c1; ci = -c3; break;
  }
  typename Cplx<T>::type r;
  r.x = cr;
  r.y = SIGN < 0 ? -ci : ci;
  return r;
}

template <typename T, int SIGN> __device__ __forceinline__ void dft4(typename Cplx<T>::type &a, typename Cplx<T>::type &b,
                                                                     typename Cplx<T>::type &c, typename Cplx<T>::type &d) {
  typedef typename Cplx<T>::type C;
  const C t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = mul_i<C, SIGN>(csub(b, d));
  a = cadd(t0, t2);
  b = cadd(t1, t3);
  c = csub(t0, t2);
  d = csub(t1, t3);
}

// in-place DFT of x[0..R), natural order in and out
template <typename T, int SIGN, int R> __device__ __forceinline__ void dft_reg(typename Cplx<T>::type *x) {
  typedef typename Cplx<T>::type C;
  if constexpr (R == 4) {
    dft4<T, SIGN>(x[0], x[1], x[2], x[3]);
  } else if constexpr (R == 8) {   // n = 4 n1 + n2: DFT2 over n1, twiddle W8^(n2 k1), DFT4 over n2 -> X[k1 + 2 k2]
    C y[8];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
      y[n2] = cadd(x[n2], x[4 + n2]);
      C o = csub(x[n2], x[4 + n2]);
      if (n2 == 1 || n2 == 3) o = cmul(o, w16<T, SIGN>(2 * n2));
      if (n2 == 2) o = mul_i<C, SIGN>(o);
      y[4 + n2] = o;
    }
    dft4<T, SIGN>(y[0], y[1], y[2], y[3]);
    dft4<T, SIGN>(y[4], y[5], y[6], y[7]);
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) { x[2 * k2] = y[k2]; x[2 * k2 + 1] = y[4 + k2]; }
  } else {               // R == 16, n = 4 n1 + n2: DFT4 over n1, twiddle W16^(n2 k1), DFT4 over n2 -> X[k1 + 4 k2]
    C y[16];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
      C a = x[n2], b = x[4 + n2], c = x[8 + n2], d = x[12 + n2];
      dft4<T, SIGN>(a, b, c, d);
      y[n2] = a;   // k1 = 0
      if (n2 == 0) { y[4] = b; y[8] = c; y[12] = d; }
      else {
        y[4 + n2] = cmul(b, w16<T, SIGN>(n2));
        y[8 + n2] = (2 * n2 == 4) ? mul_i<C, SIGN>(c) : cmul(c, w16<T, SIGN>(2 * n2));
        y[12 + n2] = cmul(d, w16<T, SIGN>(3 * n2));
      }
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
      dft4<T, SIGN>(y[4 * k1], y[4 * k1 + 1], y[4 * k1 + 2], y[4 * k1 + 3]);
#pragma unroll
      for (int k2 = 0; k2 < 4; k2++) x[k1 + 4 * k2] = y[4 * k1 + k2];
    }
  }
}

__device__ __forceinline__ int reg_pad(int i) { return i + (i >> 4); }

// one Stockham step with radix R after NS = product of the earlier radices on the 16 register values of thread t;
// results go to the shared-memory row (LAST = false) or stay in v in global order t + TPL * e (LAST = true)
template <typename T, int SIGN, int L, int R, int NS, bool LAST>
__device__ __forceinline__ void reg_step(typename Cplx<T>::type *v, int t, const typename Cplx<T>::type *__restrict__ tw,
                                         typename Cplx<T>::type *row) {
  typedef typename Cplx<T>::type C;
  constexpr int TPL = L / 16, U = 16 / R;
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int jj = t + TPL * u;
    const int k = jj & (NS - 1);
    C x[R];
#pragma unroll
    for (int q = 0; q < R; q++) {
      x[q] = v[u + q * U];
      if (NS > 1 && q > 0) {
        C wq = tw[q * k * (L / (NS * R))];
        if (SIGN > 0) wq.y = -wq.y;
        x[q] = cmul(x[q], wq);
      }
    }
    dft_reg<T, SIGN, R>(x);
#pragma unroll
    for (int pq = 0; pq < R; pq++) {
      if (LAST) v[u + pq * U] = x[pq];   // (jj - k) R + k + pq NS == jj + pq L / R == t + TPL (u + pq U)
      else row[reg_pad((jj - k) * R + k + pq * NS)] = x[pq];
    }
  }
}

template <typename T, int SIGN, int R1, int R2, int R3>
__global__ void __launch_bounds__(256)
fft_reg_kernel(typename Cplx<T>::type *__restrict__ data, const typename Cplx<T>::type *__restrict__ tw, LineGeom g) {
  typedef typename Cplx<T>::type C;
  constexpr int L = R1 * R2 * R3, TPL = L / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long b = blockIdx.x;
  const int cnt = bundle_count<C>(g, b);
  int cl, t;
  if (g.inner == 1) { t = threadIdx.x % TPL; cl = threadIdx.x / TPL; }
  else { cl = threadIdx.x % g.bundle; t = threadIdx.x / g.bundle; }
  const bool active = cl < cnt;
  C *row = reinterpret_cast<C *>(smem_raw) + (size_t) cl * g.pitch;
  C *base = data;
  long long stride = 1;
  if (active) {
    if (g.inner == 1) base += map_outer(g, b * g.bundle + cl) * L;
    else {
      const long long oc = b / g.bundles_inner;
      base += map_outer(g, oc) * L * g.inner + map_inner(g, (b - oc * g.bundles_inner) * g.bundle) + cl;
      stride = g.inner;
    }
  }
  C v[16];
#pragma unroll
  for (int e = 0; e < 16; e++) {
    const int i = t + TPL * e;
    v[e] = (active && keep_load(g, i)) ? base[(long long) i * stride] : C{0, 0};
  }
  reg_step<T, SIGN, L, R1, 1, false>(v, t, tw, row);
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 16; e++) v[e] = row[reg_pad(t + TPL * e)];
  if (R3 > 1) {
    __syncthreads();
    reg_step<T, SIGN, L, R2, R1, false>(v, t, tw, row);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = row[reg_pad(t + TPL * e)];
    reg_step<T, SIGN, L, (R3 > 1 ? R3 : 4), R1 * R2, true>(v, t, tw, row);
  } else {
    reg_step<T, SIGN, L, R2, R1, true>(v, t, tw, row);
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const int i = t + TPL * e;
      if (keep_store(g, i)) base[(long long) i * stride] = v[e];
    }
  }
}

// ---- any length with prime factors <= 61: mixed-radix Stockham in shared memory ------------------
// Stage with radix R after Ns = product of the earlier radices: butterfly j in [0, L/R), k = j mod Ns,
//   v[q] = src[j + q L/R] * W_L^(q k L/(Ns R)),  dst[(j - k) R + k + p Ns] = sum_q v[q] W_R^(p q).
template <typename T, int R>
__device__ __forceinline__ void radix_stage(const typename Cplx<T>::type *__restrict__ src,
                                            typename Cplx<T>::type *__restrict__ dst,
                                            const typename Cplx<T>::type *__restrict__ tw, int L, int Ns, int cnt,
                                            int pitch, int sign) {
  typedef typename Cplx<T>::type C;
  const int sub = L / R, tstep = L / (Ns * R), rstep = L / R;
  for (int w = threadIdx.x; w < cnt * sub; w += blockDim.x) {
    const int cl = w / sub, j = w - cl * sub;
    const int k = j % Ns;
    C v[R];
#pragma unroll
    for (int q = 0; q < R; q++) {
      v[q] = src[cl * pitch + j + q * sub];
      if (q && k) {
        C wq = tw[q * k * tstep];
        if (sign > 0) wq.y = -wq.y;
        v[q] = cmul(v[q], wq);
      }
    }
    C *out = dst + cl * pitch + (j - k) * R + k;
#pragma unroll
    for (int pq = 0; pq < R; pq++) {
      C acc = v[0];
#pragma unroll
      for (int q = 1; q < R; q++) {
        C wr = tw[((pq * q) % R) * rstep];
        if (sign > 0) wr.y = -wr.y;
        acc = cadd(acc, cmul(v[q], wr));
      }
      out[pq * Ns] = acc;
    }
  }
}

// any other (odd prime) radix: one output per inner iteration, inputs re-read from shared memory
template <typename T>
__device__ __forceinline__ void radix_stage_any(const typename Cplx<T>::type *__restrict__ src,
                                                typename Cplx<T>::type *__restrict__ dst,
                                                const typename Cplx<T>::type *__restrict__ tw, int L, int R, int Ns,
                                                int cnt, int pitch, int sign) {
  typedef typename Cplx<T>::type C;
  const int sub = L / R, tstep = L / (Ns * R), rstep = L / R;
  for (int w = threadIdx.x; w < cnt * sub * R; w += blockDim.x) {
    const int pq = w % R, rest = w / R;
    const int cl = rest / sub, j = rest - cl * sub;
    const int k = j % Ns;
    C acc = src[cl * pitch + j];
    for (int q = 1; q < R; q++) {
      // W_L^(q k tstep) * W_R^(pq q): both exponents are < L, their sum is reduced once
      int e = q * k * tstep + ((pq * q) % R) * rstep;
      if (e >= L) e -= L;
      C wv = tw[e];
      if (sign > 0) wv.y = -wv.y;
      acc = cadd(acc, cmul(src[cl * pitch + j + q * sub], wv));
    }
    dst[cl * pitch + (j - k) * R + k + pq * Ns] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(kFftThreads)
fft_mixed_kernel(typename Cplx<T>::type *__restrict__ data, typename Cplx<T>::type *__restrict__ out_data,
                 const typename Cplx<T>::type *__restrict__ tw, LineGeom g, int sign, RadixPlan rp) {
  typedef typename Cplx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *src = reinterpret_cast<C *>(smem_raw);
  C *dst = src + (size_t) g.bundle * g.pitch;
  const int L = (int) g.len;
  const long long b = blockIdx.x;
  const int cnt = bundle_count<C>(g, b);
  stage_lines<C, false>(data, src, g, b);
  __syncthreads();
  int Ns = 1;
  for (int st = 0; st < rp.nst; st++) {
    const int R = rp.radix[st];
    switch (R) {
      case 2: radix_stage<T, 2>(src, dst, tw, L, Ns, cnt, g.pitch, sign); break;
      case 3: radix_stage<T, 3>(src, dst, tw, L, Ns, cnt, g.pitch, sign); break;
      case 4: radix_stage<T, 4>(src, dst, tw, L, Ns, cnt, g.pitch, sign); break;
      case 5: radix_stage<T, 5>(src, dst, tw, L, Ns, cnt, g.pitch, sign); break;
      case 7: radix_stage<T, 7>(src, dst, tw, L, Ns, cnt, g.pitch, sign); break;
      default: radix_stage_any<T>(src, dst, tw, L, R, Ns, cnt, g.pitch, sign); break;
    }
    __syncthreads();
    C *t = src; src = dst; dst = t;
    Ns *= R;
  }
  stage_lines<C, true>(out_data, src, g, b, sign);
}

// ---- any length that fits: O(len^2) DFT from the twiddle table ---------------------------------
template <typename T>
__global__ void __launch_bounds__(kFftThreads)
dft_table_kernel(typename Cplx<T>::type *__restrict__ data, typename Cplx<T>::type *__restrict__ out_data,
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
  stage_lines<C, true>(out_data, buf1, g, b, sign);
}

// ---- lengths with a prime factor > 61: Bluestein (chirp-z) inside one CTA -----------------------------------------
// X[k] = w[k] * sum_l (x[l] w[l]) conj(w)[k - l],  w[k] = exp(-i pi k^2 / L): a cyclic convolution of length
// P = 2^p >= 2L - 1, done with the power-of-two Stockham above: a = x.w zero-padded, A = DFT_P(a), A *= Bhat
// (= DFT_P of the wrapped conj(w), tabulated on the host in long double, 1/P folded in), y = IDFT_P(A),
// X[k] = w[k] y[k].  The backward transform is conj(DFT(conj x)).
template <typename T>
__global__ void __launch_bounds__(kFftThreads)
fft_bluestein_kernel(typename Cplx<T>::type *__restrict__ data, typename Cplx<T>::type *__restrict__ out_data,
                     const typename Cplx<T>::type *__restrict__ chirp, const typename Cplx<T>::type *__restrict__ bhat,
                     const typename Cplx<T>::type *__restrict__ twP, LineGeom g, int sign, int P, int log2P) {
  typedef typename Cplx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *buf0 = reinterpret_cast<C *>(smem_raw);
  C *buf1 = buf0 + (size_t) g.bundle * g.pitch;
  const int L = (int) g.len;
  const long long b = blockIdx.x;
  const int cnt = bundle_count<C>(g, b);
  stage_lines<C, false>(data, buf0, g, b);
  __syncthreads();
  for (int w = threadIdx.x; w < cnt * P; w += blockDim.x) {
    const int cl = w / P, k = w - cl * P;
    C v = C{0, 0};
    if (k < L) {
      v = buf0[cl * g.pitch + k];
      if (sign > 0) v.y = -v.y;
      v = cmul(v, chirp[k]);
    }
    buf0[cl * g.pitch + k] = v;
  }
  __syncthreads();
  C *A = pow2_stockham_smem<T>(buf0, buf1, twP, P, log2P, cnt, g.pitch, -1);
  for (int w = threadIdx.x; w < cnt * P; w += blockDim.x) {
    const int cl = w / P, k = w - cl * P;
    A[cl * g.pitch + k] = cmul(A[cl * g.pitch + k], bhat[k]);
  }
  __syncthreads();
  C *other = (A == buf0) ? buf1 : buf0;
  C *y = pow2_stockham_smem<T>(A, other, twP, P, log2P, cnt, g.pitch, +1);
  C *res = (y == buf0) ? buf1 : buf0;
  for (int w = threadIdx.x; w < cnt * L; w += blockDim.x) {
    const int cl = w / L, k = w - cl * L;
    C v = cmul(y[cl * g.pitch + k], chirp[k]);
    if (sign > 0) v.y = -v.y;
    res[cl * g.pitch + k] = v;
  }
  __syncthreads();
  stage_lines<C, true>(out_data, res, g, b, sign);
}

int ilog2_exact(long long v) {
  int l = 0;
  while ((1ll << l) < v) l++;
  return ((1ll << l) == v) ? l : -1;
}

constexpr int kMaxPrimeRadix = 61;

// len = 4^a * 2^b * odd primes <= kMaxPrimeRadix ?
bool factorize(long long len, FftLine &ln) {
  ln.nst = 0;
  while (len % 4 == 0 && ln.nst < 24) { ln.radix[ln.nst++] = 4; len /= 4; }
  if (len % 2 == 0 && ln.nst < 24) { ln.radix[ln.nst++] = 2; len /= 2; }
  for (int p = 3; p <= kMaxPrimeRadix; p += 2)
    while (len % p == 0) {
      if (ln.nst >= 24) return false;
      ln.radix[ln.nst++] = p;
      len /= p;
    }
  return len == 1;
}

// longest line whose two shared-memory buffers (pitch len + 1) fit one CTA
long long max_line_len(const nfftcu_ctx *c) { return (long long) (kSmemBudget / (2 * 2 * real_size(c))) - 1; }

int upload_twiddles(const nfftcu_ctx *c, long long count, long long num_stride, long long den, void **out) {
  // table[q] = exp(-2 pi i q * num_stride / den), q < count, in long double -> plan precision
  const long double two_pi = 6.283185307179586476925286766559005768394L;
  const size_t esz = 2 * real_size(c);
  std::vector<unsigned char> host(esz * (size_t) count);
  for (long long q = 0; q < count; q++) {
    const long long r = (long long) (((__int128) q * num_stride) % den);
    const long double ang = two_pi * (long double) r / (long double) den;
    const long double cr = cosl(ang), ci = -sinl(ang);
    if (c->prec == NFFTCU_DOUBLE) {
      ((double *) host.data())[2 * q] = (double) cr;
      ((double *) host.data())[2 * q + 1] = (double) ci;
    } else {
      ((float *) host.data())[2 * q] = (float) cr;
      ((float *) host.data())[2 * q + 1] = (float) ci;
    }
  }
  NFFTCU_CUDA(pool_malloc(out, host.size()));
  NFFTCU_CUDA(cudaMemcpy(*out, host.data(), host.size(), cudaMemcpyHostToDevice));
  return NFFTCU_OK;
}

int upload_complex(const nfftcu_ctx *c, const std::vector<long double> &re, const std::vector<long double> &im, void **out) {
  const size_t esz = 2 * real_size(c), count = re.size();
  std::vector<unsigned char> host(esz * count);
  for (size_t q = 0; q < count; q++) {
    if (c->prec == NFFTCU_DOUBLE) {
      ((double *) host.data())[2 * q] = (double) re[q];
      ((double *) host.data())[2 * q + 1] = (double) im[q];
    } else {
      ((float *) host.data())[2 * q] = (float) re[q];
      ((float *) host.data())[2 * q + 1] = (float) im[q];
    }
  }
  NFFTCU_CUDA(pool_malloc(out, host.size()));
  NFFTCU_CUDA(cudaMemcpy(*out, host.data(), host.size(), cudaMemcpyHostToDevice));
  return NFFTCU_OK;
}

// Bluestein tables of a line of length len (host, long double): chirp w[k] = exp(-i pi k^2/len) with k^2 reduced
// mod 2 len, and Bhat = DFT_P(b)/P for b[j] = conj(w[|j|]), |j| < len, wrapped into P points.
int plan_bluestein(const nfftcu_ctx *c, FftLine &ln) {
  const long long L = ln.len, P = ln.blu_P;
  const long double pi = 3.141592653589793238462643383279502884L;
  std::vector<long double> wr((size_t) L), wi((size_t) L), br((size_t) P, 0.0L), bi((size_t) P, 0.0L);
  for (long long k = 0; k < L; k++) {
    const long long r = (long long) (((__int128) k * k) % (2 * L));
    const long double ang = pi * (long double) r / (long double) L;
    wr[(size_t) k] = cosl(ang);
    wi[(size_t) k] = -sinl(ang);
  }
  for (long long j = 0; j < L; j++) {
    br[(size_t) j] = wr[(size_t) j];
    bi[(size_t) j] = -wi[(size_t) j];
    if (j) { br[(size_t) (P - j)] = wr[(size_t) j]; bi[(size_t) (P - j)] = -wi[(size_t) j]; }
  }
  // in-place iterative radix-2 DFT (sign -1) of b in long double
  for (long long i = 1, j = 0; i < P; i++) {
    long long bit = P >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(br[(size_t) i], br[(size_t) j]); std::swap(bi[(size_t) i], bi[(size_t) j]); }
  }
  for (long long len2 = 2; len2 <= P; len2 <<= 1) {
    for (long long i = 0; i < P; i += len2)
      for (long long k = 0; k < len2 / 2; k++) {
        const long double ang = -2.0L * pi * (long double) k / (long double) len2;
        const long double cr = cosl(ang), ci = sinl(ang);
        const size_t u = (size_t) (i + k), v = (size_t) (i + k + len2 / 2);
        const long double tr = br[v] * cr - bi[v] * ci, ti = br[v] * ci + bi[v] * cr;
        br[v] = br[u] - tr; bi[v] = bi[u] - ti;
        br[u] += tr; bi[u] += ti;
      }
  }
  for (long long k = 0; k < P; k++) { br[(size_t) k] /= (long double) P; bi[(size_t) k] /= (long double) P; }
  NFFTCU_TRY(upload_complex(c, wr, wi, &ln.blu_chirp));
  NFFTCU_TRY(upload_complex(c, br, bi, &ln.blu_bhat));
  return upload_twiddles(c, P, 1, P, &ln.blu_tw);
}

int plan_line(const nfftcu_ctx *c, long long len, FftLine &ln) {
  ln.len = len;
  if (len == 1) { ln.kind = 0; return NFFTCU_OK; }
  if (ilog2_exact(len) >= 0) ln.kind = 1;
  else if (factorize(len, ln)) ln.kind = 3;
  else {
    // a prime factor > 61: Bluestein when the padded convolution (two buffers of P + 1 points) fits one CTA,
    // else the O(len^2) table DFT
    long long P = 1;
    while (P < 2 * len - 1) P <<= 1;
    if (2 * (size_t) (P + 1) * 2 * real_size(c) <= kSmemBudget) {
      ln.kind = 4;
      ln.blu_P = P;
      NFFTCU_TRY(plan_bluestein(c, ln));
    } else {
      ln.kind = 2;
    }
  }
  return upload_twiddles(c, len, 1, len, &ln.tw);
}

template <typename T, int R1, int R2, int R3>
int launch_reg(nfftcu_ctx *c, const FftLine &ln, LineGeom g, int sign, void *data) {
  typedef typename Cplx<T>::type C;
  constexpr int L = R1 * R2 * R3, TPL = L / 16;
  int bundle = 256 / TPL;                                  // 256 threads per CTA at most
  const int seg = (int) (128 / sizeof(C));                 // lines per 128-byte segment of a strided axis
  if (g.inner > 1 && bundle > seg) bundle = seg;
  if (g.inner == 1 && bundle > 8) bundle = 8;
  if (g.icount == 0) g.icount = g.inner;
  if (g.inner > 1 && (long long) bundle > g.icount) bundle = (int) g.icount;
  if (bundle < 1) bundle = 1;
  while (g.iprune && bundle > 1 && g.ialign % bundle) bundle--;   // bundles must not straddle a band gap
  g.len = L;
  g.bundle = bundle;
  g.bundles_inner = g.inner == 1 ? 1 : (g.icount + bundle - 1) / bundle;
  g.pitch = L + L / 16;
  while (g.pitch % 8 != 1) g.pitch++;                      // odd multiple of 16 bytes (fp64) between rows
  const long long nb = g.inner == 1 ? (g.lines + bundle - 1) / bundle : (g.lines / g.icount) * g.bundles_inner;
  if (nb > 0x7fffffffll) { set_error("FFT: too many line bundles (%lld)", nb); return NFFTCU_EINVAL; }
  const size_t smem = sizeof(C) * (size_t) g.pitch * bundle;
  const unsigned threads = (unsigned) (bundle * TPL);
  if (sign < 0) {
    NFFTCU_CUDA(cudaFuncSetAttribute(fft_reg_kernel<T, -1, R1, R2, R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    fft_reg_kernel<T, -1, R1, R2, R3><<<(unsigned) nb, threads, smem, c->stream>>>((C *) data, (const C *) ln.tw, g);
  } else {
    NFFTCU_CUDA(cudaFuncSetAttribute(fft_reg_kernel<T, +1, R1, R2, R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    fft_reg_kernel<T, +1, R1, R2, R3><<<(unsigned) nb, threads, smem, c->stream>>>((C *) data, (const C *) ln.tw, g);
  }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

// register-resident kernel for this length?  (in place, no twist / transposed store)
template <typename T>
int try_reg(nfftcu_ctx *c, const FftLine &ln, const LineGeom &g, int sign, void *data, bool *done) {
  *done = true;
  switch (ln.len) {
    case 64: return launch_reg<T, 8, 8, 1>(c, ln, g, sign, data);
    case 128: return launch_reg<T, 16, 8, 1>(c, ln, g, sign, data);
    case 256: return launch_reg<T, 16, 16, 1>(c, ln, g, sign, data);
    case 512: return launch_reg<T, 8, 8, 8>(c, ln, g, sign, data);
    case 1024: return launch_reg<T, 16, 16, 4>(c, ln, g, sign, data);
    case 2048: return launch_reg<T, 16, 16, 8>(c, ln, g, sign, data);
    default: break;
  }
  *done = false;
  return NFFTCU_OK;
}

// one pass of shared-memory transforms of length ln.len over lines described by g (bundle etc. filled in here)
template <typename T>
int run_pass(nfftcu_ctx *c, const FftLine &ln, LineGeom g, int sign, void *src, void *dst) {
  typedef typename Cplx<T>::type C;
  if (ln.kind == 1 && !g.twist && !g.tstore && src == dst && c->opt_fft_kernel != 1) {
    bool done = false;
    NFFTCU_TRY(try_reg<T>(c, ln, g, sign, src, &done));
    if (done) return NFFTCU_OK;
  }
  const int pad = 1;
  g.len = ln.len;
  g.pitch = (int) (ln.kind == 4 ? ln.blu_P : ln.len) + pad;
  const size_t line_bytes = 2 * (size_t) g.pitch * sizeof(C);   // two buffers
  int pref = (int) (128 / sizeof(C));                           // 128-byte global segments
  if (g.inner == 1) pref = (int) max(1ll, min(8ll, 2048ll / ln.len));
  if (g.inner == 1 && g.tstore) pref = (int) (128 / sizeof(C));  // the transposed store runs along the bundle
  int fit = (int) (kSmemBudget / line_bytes);
  if (fit < 1) {
    set_error("FFT: length %lld does not fit the shared-memory kernel", (long long) ln.len);
    return NFFTCU_EINVAL;
  }
  g.bundle = pref < fit ? pref : fit;
  if (g.icount == 0) g.icount = g.inner;
  if (g.inner > 1 && (long long) g.bundle > g.icount) g.bundle = (int) g.icount;
  if (g.bundle > 64) g.bundle = 64;   // line_base[] of stage_lines
  while (g.iprune && g.bundle > 1 && g.ialign % g.bundle) g.bundle--;   // bundles must not straddle a band gap
  g.bundles_inner = g.inner == 1 ? 1 : (g.icount + g.bundle - 1) / g.bundle;
  const long long nb = g.inner == 1 ? (g.lines + g.bundle - 1) / g.bundle
                                    : (g.lines / g.icount) * g.bundles_inner;
  if (nb > 0x7fffffffll) {
    set_error("FFT: too many line bundles (%lld)", nb);
    return NFFTCU_EINVAL;
  }
  const size_t smem = line_bytes * g.bundle;
  if (ln.kind == 1) {
    NFFTCU_CUDA(cudaFuncSetAttribute(fft_stockham_kernel<T>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    fft_stockham_kernel<T><<<(unsigned) nb, kFftThreads, smem, c->stream>>>(
        (C *) src, (C *) dst, (const C *) ln.tw, g, sign, ilog2_exact(ln.len));
  } else if (ln.kind == 3) {
    RadixPlan rp;
    rp.nst = ln.nst;
    for (int i = 0; i < 24; i++) rp.radix[i] = ln.radix[i];
    NFFTCU_CUDA(cudaFuncSetAttribute(fft_mixed_kernel<T>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    fft_mixed_kernel<T><<<(unsigned) nb, kFftThreads, smem, c->stream>>>((C *) src, (C *) dst, (const C *) ln.tw, g,
                                                                         sign, rp);
  } else if (ln.kind == 4) {
    NFFTCU_CUDA(cudaFuncSetAttribute(fft_bluestein_kernel<T>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    fft_bluestein_kernel<T><<<(unsigned) nb, kFftThreads, smem, c->stream>>>(
        (C *) src, (C *) dst, (const C *) ln.blu_chirp, (const C *) ln.blu_bhat, (const C *) ln.blu_tw, g, sign,
        (int) ln.blu_P, ilog2_exact(ln.blu_P));
  } else {
    NFFTCU_CUDA(cudaFuncSetAttribute(dft_table_kernel<T>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    dft_table_kernel<T><<<(unsigned) nb, kFftThreads, smem, c->stream>>>((C *) src, (C *) dst, (const C *) ln.tw, g,
                                                                         sign);
  }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

template <typename T>
int run_axis(nfftcu_ctx *c, int t, int sign, bool pruned) {
  const FftAxis &ax = c->fft[t];
  if (ax.len == 1) return NFFTCU_OK;
  LineGeom g;
  memset(&g, 0, sizeof(g));
  long long inner = 1, outer_full = 1;
  for (int t2 = t + 1; t2 < c->d; t2++) inner *= c->n[t2];
  for (int s2 = 0; s2 < t; s2++) outer_full *= c->n[s2];
  // a batched transform (nfftcu_*_batch): the grid is [K][n_0]...[n_{d-1}], the right-hand side is one more
  // (unpruned) axis in front of all others
  const long long K = c->cur_batch;
  outer_full *= K;
  if (ax.split) {
    // column pass: [outer][L1][L2 * inner], transform over L1, twiddle at the store, in place
    const long long L1 = ax.sub1.len, L2 = ax.sub2.len;
    g.inner = L2 * inner;
    g.lines = outer_full * g.inner;
    g.twist = 1;
    g.tw_I = inner;
    g.twA = ax.twA;
    g.twB = ax.twB;
    g.tw_shift = ax.tw_shift;
    NFFTCU_TRY(run_pass<T>(c, ax.sub1, g, sign, c->grid, c->grid));
    // row pass: [outer * L1][L2][inner], transform over L2, transposed store into the second buffer
    memset(&g, 0, sizeof(g));
    g.inner = inner;
    g.lines = outer_full * L1 * inner;
    g.tstore = 1;
    g.tL1 = L1;
    g.tL2 = L2;
    g.tI = inner;
    NFFTCU_TRY(run_pass<T>(c, ax.sub2, g, sign, c->grid, c->grid2));
    void *tmp = c->grid; c->grid = c->grid2; c->grid2 = tmp;
    return NFFTCU_OK;
  }
  g.inner = inner;
  g.prune = pruned ? (sign < 0 ? 1 : 2) : 0;
  long long outer = 1;
  int o0 = 0;
  if (K > 1) {
    if (t + 1 > NFFTCU_MAX_D) { set_error("batched FFT: d = %d leaves no room for the batch axis", c->d); return NFFTCU_EINVAL; }
    g.oN[0] = g.olow[0] = g.on[0] = K;
    outer = K;
    o0 = 1;
  }
  g.nouter = t + o0;
  for (int s2 = 0; s2 < t; s2++) {
    g.oN[o0 + s2] = pruned ? c->N[s2] : c->n[s2];
    g.olow[o0 + s2] = pruned ? c->N[s2] - c->N[s2] / 2 : c->n[s2];
    g.on[o0 + s2] = c->n[s2];
    outer *= g.oN[o0 + s2];
  }
  g.elow = c->N[t] - c->N[t] / 2;
  g.ehigh = c->n[t] - c->N[t] / 2;
  g.lines = outer * g.inner;
  return run_pass<T>(c, ax.whole, g, sign, c->grid, c->grid);
}

// ---- slab mode: the F step of a GPU that holds a node SLAB (multi-GPU node sharding) ------------------------------
// The nodes of the plan touch only the planes [w0, w0 + wc) (mod n_0) of the first axis (slab_detect in api.cu).
// Forward (after D, before B) only those planes of the result are read, backward (after B^T, before D^T) only
// those planes of the input are non-zero.  The passes therefore run FIRST AXIS FIRST in the forward direction:
//   axis 0: the N_1 x N_2 band lines (inner axes still band-limited: iprune), band loads, window stores;
//   axis 1: window planes x N_2 band lines, band loads;      axis 2: window planes x n_1 lines, band loads;
// and in the opposite order with loads and stores exchanged in the backward direction.  At sigma = 2 with a window
// of 78 of 512 planes (cfg4 on 8 GPUs) this visits 27 % of the lines of the full pruned transform.
template <typename T>
int run_axis_slab(nfftcu_ctx *c, int t, int sign) {
  const FftAxis &ax = c->fft[t];
  LineGeom g;
  memset(&g, 0, sizeof(g));
  const long long K = c->cur_batch;
  long long inner = 1;
  for (int t2 = t + 1; t2 < c->d; t2++) inner *= c->n[t2];
  g.inner = inner;
  g.prune = sign < 0 ? 1 : 2;               // band loads (forward) / band stores (backward) on the transform axis
  g.elow = c->N[t] - c->N[t] / 2;
  g.ehigh = c->n[t] - c->N[t] / 2;
  long long outer = 1;
  int o0 = 0;
  if (K > 1) { g.oN[0] = g.olow[0] = g.on[0] = K; outer = K; o0 = 1; }
  if (t == 0) {
    // window on the transform axis itself; the inner axes are band-limited
    if (sign < 0) g.wstore = 1; else g.wload = 1;
    g.w0 = c->slab_w0;
    g.wc = c->slab_wc;
  } else {
    // the first axis is an outer axis restricted to the window planes
    g.oN[o0] = g.olow[o0] = c->slab_wc;
    g.on[o0] = c->n[0];
    g.ooff[o0] = c->slab_w0;
    outer *= c->slab_wc;
    for (int s2 = 1; s2 < t; s2++) {        // axes between the first and this one: already / still in grid space
      g.oN[o0 + s2] = g.olow[o0 + s2] = g.on[o0 + s2] = c->n[s2];
      outer *= c->n[s2];
    }
  }
  g.nouter = t + o0;
  g.icount = inner;
  if (t + 1 < c->d) {                       // axes behind this one are band-limited: visit the band lines only
    g.iprune = 1;
    g.niax = c->d - 1 - t;
    long long ic = 1;
    for (int s2 = t + 1; s2 < c->d; s2++) {
      g.iN[s2 - t - 1] = c->N[s2];
      g.ilow[s2 - t - 1] = c->N[s2] - c->N[s2] / 2;
      g.in_[s2 - t - 1] = c->n[s2];
      ic *= c->N[s2];
    }
    g.icount = ic;
    const long long lo = c->N[c->d - 1] - c->N[c->d - 1] / 2, hi = c->N[c->d - 1] / 2;
    long long a = lo, b2 = hi;               // gcd
    while (b2) { const long long r = a % b2; a = b2; b2 = r; }
    g.ialign = a;
  }
  g.lines = outer * g.icount;
  return run_pass<T>(c, ax.whole, g, sign, c->grid, c->grid);
}

}  // namespace

int fft_plan_axes(nfftcu_ctx *c) {
  const long long fit = max_line_len(c);
  for (int t = 0; t < c->d; t++) {
    FftAxis &ax = c->fft[t];
    ax.len = c->n[t];
    ax.split = false;
    if (ax.len <= fit) {
      NFFTCU_TRY(plan_line(c, ax.len, ax.whole));
      continue;
    }
    // four-step: the divisor pair (L1, L2) closest to sqrt(len) with both factors fitting one CTA
    long long best = 0;
    for (long long a = 2; a * a <= ax.len; a++)
      if (ax.len % a == 0 && ax.len / a <= fit) best = a;   // a <= sqrt(len) <= len / a
    if (!best) {
      set_error("FFT axis %d: length %lld neither fits shared memory (max %lld) nor splits into two such factors",
                t, (long long) ax.len, fit);
      return NFFTCU_EINVAL;
    }
    ax.split = true;
    NFFTCU_TRY(plan_line(c, best, ax.sub1));
    NFFTCU_TRY(plan_line(c, ax.len / best, ax.sub2));
    // W_len^q, q < len, as twA[q >> s] * twB[q & (2^s - 1)]
    int s = 0;
    while ((1ll << (2 * s)) < ax.len) s++;
    ax.tw_shift = s;
    NFFTCU_TRY(upload_twiddles(c, (ax.len >> s) + 1, 1ll << s, ax.len, &ax.twA));
    NFFTCU_TRY(upload_twiddles(c, 1ll << s, 1, ax.len, &ax.twB));
    c->fft_no_prune = true;
    if (!c->grid2) NFFTCU_CUDA(pool_malloc(&c->grid2, 2 * real_size(c) * (size_t) c->n_total));
  }
  return NFFTCU_OK;
}

void fft_free_axes(nfftcu_ctx *c) {
  for (int t = 0; t < c->d; t++) {
    FftAxis &ax = c->fft[t];
    void **ptrs[] = {&ax.whole.tw, &ax.sub1.tw, &ax.sub2.tw, &ax.twA, &ax.twB,
                     &ax.whole.blu_chirp, &ax.whole.blu_bhat, &ax.whole.blu_tw, &ax.sub1.blu_chirp, &ax.sub1.blu_bhat,
                     &ax.sub1.blu_tw, &ax.sub2.blu_chirp, &ax.sub2.blu_bhat, &ax.sub2.blu_tw};
    for (void **p : ptrs) {
      if (*p) pool_free(*p);
      *p = nullptr;
    }
  }
  if (c->grid2) pool_free(c->grid2);
  c->grid2 = nullptr;
}

// pruned: the caller guarantees the band structure described in the header (trafo: grid written by the
// sparse D; adjoint: only the band of the result is read by D^T)
int stage_F(nfftcu_ctx *c, int sign, bool pruned) {
  if (c->fft_no_prune) pruned = false;
  if (pruned && c->slab_on && c->d >= 2) {
    // forward: first axis first; backward: last axis first
    for (int i = 0; i < c->d; i++) {
      const int t = sign < 0 ? i : c->d - 1 - i;
      if (c->prec == NFFTCU_DOUBLE) NFFTCU_TRY(run_axis_slab<double>(c, t, sign));
      else NFFTCU_TRY(run_axis_slab<float>(c, t, sign));
    }
    return NFFTCU_OK;
  }
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
