// tile3d.cu -- 3-D interpolation (B) and spreading (B^T) as a "pencil sweep" with register-resident
// grid windows: the fast path for d = 3.
//
// Reference being replaced: nfft_trafo_3d_B / nfft_trafo_3d_compute (kernel/nfft/nfft.c:4687-4914,
// 4020-4265) and nfft_adjoint_3d_B with its atomic and blockwise compute variants (5126-5384,
// 4393-4436, 4289-4388, slab assignment 1345-1420).  The reference's blockwise adjoint gives every
// thread a slab of the grid and lets it walk the sorted nodes touching that slab; this kernel is
// the GPU form of the same owner-computes idea, taken down to the register level.
//
// Geometry.  Nodes are binned by the corner u = floor(x n) - m of their (2m+2)^3 tap box:
//   tile (a,b)  = (u0 / T0, u1 / T1), T0 = T1 = 3        slab s = u2 / SZ, SZ = 2.
// A CTA owns one tile and sweeps a range of slabs along the contiguous axis z.  Every tap box of
// the tile lies inside the tile's FOOTPRINT of F0 x F1 = (T0+W-1) x (T1+W-1) grid rows (W = 2m+2;
// 16 x 16 rows for m = 6) and, for the current slab, inside a z-window of WZ = W+SZ-1 cells.
// Each thread owns two footprint rows and keeps their z-windows -- 2 x WZ complex values -- in
// REGISTERS:
//   spreading      the windows are accumulators; a node adds (psi0 psi1 f_j) * psi2[k] to all WZ cells
//                  of the thread's rows (psi vectors are zero-padded to the footprint / window, so the
//                  register indices are static); when the sweep leaves a slab, the SZ cells that
//                  fall out of the window are retired to the grid with RED.ADD and the window shifts.
//                  No shared-memory accumulation, no intra-CTA conflicts: a row has one owner.
//   interpolation  the windows hold grid values, refilled SZ cells per slab straight from L2;
//                  a node reduces them against psi2, weights by psi0 psi1, and the per-thread partial
//                  sums of a batch of NB nodes are reduced across the CTA through shared memory.
// Per tap this costs 2 FP64 FMAs and, per node and thread, WZ broadcast shared-memory loads of psi2
// -- instead of one 16-byte shared/L1 load per tap -- which moves the kernel from the LSU roof
// (128 B/clk/SM) to the FP64 roof (64 FMA/clk/SM); see DESIGN.md for the arithmetic and
// profiles/ for the measurements.  Zero padding costs (W/F0)(W/F1)(W/WZ) = 71% lane efficiency at m = 6.
#include "common.cuh"

#include <type_traits>

namespace nfftcu {

namespace {

constexpr int kT0 = 3, kT1 = 3, kSZ = 2, kNB = 8;

template <int W_>
struct Cfg {
  static constexpr int W = W_, T0 = kT0, T1 = kT1, SZ = kSZ, NB = kNB;
  static constexpr int F0 = T0 + W - 1, F1 = T1 + W - 1, ROWS = F0 * F1;
  static constexpr int WZ = W + SZ - 1;
  static constexpr int WZP = (WZ + 1) & ~1;
  static constexpr int THREADS = ((((ROWS + 1) / 2) + 31) / 32) * 32;
  static constexpr int PADLEN = F0 + F1 + WZP;
  static constexpr int MINB = THREADS <= 128 ? 2 : 1;
};

struct TileParams {
  int n0, n1, n2;
  int NT0, NT1, NS;
  int zseg;
  int m;
  double m2, b0, b1, b2;
};

__device__ __forceinline__ int wrap_idx(long long v, int n) {
  long long r = v % n;
  if (r < 0) r += n;
  return (int) r;
}

__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }

template <typename T>
__global__ void tile_keys_kernel(const T *__restrict__ x, uint64_t *__restrict__ keys,
                                 uint32_t *__restrict__ vals, long long M, TileParams P) {
  const long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int u0 = wrap_idx(cell_of(x[3 * j], P.n0) - P.m, P.n0);
  const int u1 = wrap_idx(cell_of(x[3 * j + 1], P.n1) - P.m, P.n1);
  const int u2 = wrap_idx(cell_of(x[3 * j + 2], P.n2) - P.m, P.n2);
  const unsigned long long tile = (unsigned long long) (u0 / kT0) * P.NT1 + (u1 / kT1);
  keys[j] = tile * P.NS + (u2 / kSZ);
  vals[j] = (uint32_t) j;
}

// bin_start[b] = first position whose key >= b, b = 0..nbins
__global__ void bin_bounds_kernel(const uint64_t *__restrict__ keys, uint32_t *__restrict__ bin_start,
                                  long long nbins, long long M) {
  const long long b = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nbins) return;
  long long lo = 0, hi = M;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < (uint64_t) b) lo = mid + 1;
    else hi = mid;
  }
  bin_start[b] = (uint32_t) lo;
}

// ---- shared by both kernels: stage a batch of nodes (zero-padded window vectors) ------------------
// pads[i] = [ psi0 padded to F0 | psi1 padded to F1 | psi2 padded to WZP ], slab[i] = u2 / SZ
template <typename T, int W>
__device__ __forceinline__ void stage_batch(T (*pads)[Cfg<W>::PADLEN], int *slab,
                                            const T *__restrict__ xt, const T *__restrict__ table,
                                            long long k, int nb, int a, int b, const TileParams &P) {
  typedef Cfg<W> CF;
  T *flat = &pads[0][0];
  for (int i = threadIdx.x; i < CF::NB * CF::PADLEN; i += CF::THREADS) flat[i] = (T) 0;
  __syncthreads();
  for (int it = threadIdx.x; it < nb * 3 * W; it += CF::THREADS) {
    const int i = it / (3 * W), rem = it - i * (3 * W);
    const int t = rem / W, l = rem - t * W;
    const T x = xt[(k + i) * 3 + t];
    const int n = (t == 0) ? P.n0 : (t == 1) ? P.n1 : P.n2;
    const long long uu = cell_of(x, n) - P.m;       // unwrapped corner
    const int u = wrap_idx(uu, n);
    T psi;
    if (table) psi = table[((k + i) * 3 + t) * W + l];
    else {
      const double bb = (t == 0) ? P.b0 : (t == 1) ? P.b1 : P.b2;
      psi = (T) kb_phi((double) x * (double) n - (double) (uu + l), P.m2, bb);
    }
    int pos;
    if (t == 0) pos = (u - a * CF::T0) + l;
    else if (t == 1) pos = CF::F0 + (u - b * CF::T1) + l;
    else {
      pos = CF::F0 + CF::F1 + (u % CF::SZ) + l;
      if (l == 0) slab[i] = u / CF::SZ;
    }
    pads[i][pos] = psi;
  }
}

template <typename T, int W>
struct RowSetup {
  long long off[2];
  int l0[2], l1[2];
  bool valid[2];
  __device__ __forceinline__ RowSetup(int a, int b, const TileParams &P) {
    typedef Cfg<W> CF;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int r = threadIdx.x + j * CF::THREADS;
      valid[j] = r < CF::ROWS;
      const int rr = valid[j] ? r : 0;
      l0[j] = rr / CF::F1;
      l1[j] = rr - l0[j] * CF::F1;
      const int g0 = wrap_idx((long long) a * CF::T0 + l0[j], P.n0);
      const int g1 = wrap_idx((long long) b * CF::T1 + l1[j], P.n1);
      off[j] = ((long long) g0 * P.n1 + g1) * P.n2;
    }
  }
};

// ---- spreading ---------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(Cfg<W>::THREADS, Cfg<W>::MINB)
spread_tile_kernel(typename Cplx<T>::type *__restrict__ G, const T *__restrict__ xt,
                   const uint32_t *__restrict__ perm, const typename Cplx<T>::type *__restrict__ f,
                   const uint32_t *__restrict__ bin_start, const T *__restrict__ table, TileParams P) {
  typedef Cfg<W> CF;
  typedef typename Cplx<T>::type C;
  __shared__ __align__(16) T pads[CF::NB][CF::PADLEN];
  __shared__ C fv[CF::NB];
  __shared__ int slab[CF::NB];

  const int tile = blockIdx.x / P.zseg, seg = blockIdx.x - tile * P.zseg;
  const int a = tile / P.NT1, b = tile - a * P.NT1;
  const int s_begin = (int) ((long long) P.NS * seg / P.zseg);
  const int s_end = (int) ((long long) P.NS * (seg + 1) / P.zseg);
  const long long bin0 = (long long) tile * P.NS;
  const long long k0 = bin_start[bin0 + s_begin], k1 = bin_start[bin0 + s_end];
  if (k0 == k1) return;

  const RowSetup<T, W> rows(a, b, P);
  T *Gr = reinterpret_cast<T *>(G);
  T accr[2][CF::WZ], acci[2][CF::WZ];
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int kz = 0; kz < CF::WZ; kz++) { accr[j][kz] = (T) 0; acci[j][kz] = (T) 0; }
  int cur = -1;   // slab the window is aligned to; window covers z = cur*SZ .. cur*SZ+WZ-1 (mod n2)

  // retire cells [0, cnt) of the window to the grid and shift the window down by cnt (static cnt)
  auto retire = [&](auto cnt_tag) {
    constexpr int CNT = decltype(cnt_tag)::value;
    const int zb = cur * CF::SZ;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      if (rows.valid[j]) {
#pragma unroll
        for (int kz = 0; kz < CNT; kz++) {
          int z = zb + kz;
          if (z >= P.n2) z -= P.n2;
          if (z >= P.n2) z %= P.n2;
          T *p = Gr + 2 * (rows.off[j] + z);
          red_add(p, accr[j][kz]);
          red_add(p + 1, acci[j][kz]);
        }
      }
#pragma unroll
      for (int kz = 0; kz < CF::WZ; kz++) {
        if (kz + CNT < CF::WZ) { accr[j][kz] = accr[j][kz + CNT]; acci[j][kz] = acci[j][kz + CNT]; }
        else { accr[j][kz] = (T) 0; acci[j][kz] = (T) 0; }
      }
    }
  };

  for (long long k = k0; k < k1; k += CF::NB) {
    const int nb = (int) min((long long) CF::NB, k1 - k);
    __syncthreads();   // previous batch fully consumed before the pads are rewritten
    stage_batch<T, W>(pads, slab, xt, table, k, nb, a, b, P);
    if (threadIdx.x < nb) fv[threadIdx.x] = f[perm[k + threadIdx.x]];
    __syncthreads();
    for (int i = 0; i < nb; i++) {
      const int s = slab[i];
      if (cur < 0) cur = s;
      while (cur < s) {
        if (s - cur >= (CF::WZ + CF::SZ - 1) / CF::SZ) {   // the whole window leaves: flush it all
          retire(std::integral_constant<int, CF::WZ>());
          cur = s;
        } else {
          retire(std::integral_constant<int, CF::SZ>());
          cur++;
        }
      }
      const T *pd = pads[i];
      const C fj = fv[i];
      const T w0 = rows.valid[0] ? pd[rows.l0[0]] * pd[CF::F0 + rows.l1[0]] : (T) 0;
      const T w1 = rows.valid[1] ? pd[rows.l0[1]] * pd[CF::F0 + rows.l1[1]] : (T) 0;
      const T ar0 = w0 * fj.x, ai0 = w0 * fj.y, ar1 = w1 * fj.x, ai1 = w1 * fj.y;
      const T *p2 = pd + CF::F0 + CF::F1;
#pragma unroll
      for (int kz = 0; kz < CF::WZ; kz++) {
        const T p = p2[kz];
        accr[0][kz] += ar0 * p;
        acci[0][kz] += ai0 * p;
        accr[1][kz] += ar1 * p;
        acci[1][kz] += ai1 * p;
      }
    }
  }
  if (cur >= 0) retire(std::integral_constant<int, CF::WZ>());
}

// ---- interpolation -----------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(Cfg<W>::THREADS, Cfg<W>::MINB)
interp_tile_kernel(const typename Cplx<T>::type *__restrict__ G, const T *__restrict__ xt,
                   const uint32_t *__restrict__ perm, typename Cplx<T>::type *__restrict__ f,
                   const uint32_t *__restrict__ bin_start, const T *__restrict__ table, TileParams P) {
  typedef Cfg<W> CF;
  typedef typename Cplx<T>::type C;
  __shared__ __align__(16) T pads[CF::NB][CF::PADLEN];
  __shared__ int slab[CF::NB];
  __shared__ __align__(16) C red[CF::NB][CF::THREADS];

  const int tile = blockIdx.x / P.zseg, seg = blockIdx.x - tile * P.zseg;
  const int a = tile / P.NT1, b = tile - a * P.NT1;
  const int s_begin = (int) ((long long) P.NS * seg / P.zseg);
  const int s_end = (int) ((long long) P.NS * (seg + 1) / P.zseg);
  const long long bin0 = (long long) tile * P.NS;
  const long long k0 = bin_start[bin0 + s_begin], k1 = bin_start[bin0 + s_end];
  if (k0 == k1) return;

  const RowSetup<T, W> rows(a, b, P);
  T winr[2][CF::WZ], wini[2][CF::WZ];
  int cur = -1;

  // (re)load window cells [FROM, WZ) for the current alignment
  auto fill = [&](auto from_tag) {
    constexpr int FROM = decltype(from_tag)::value;
    const int zb = cur * CF::SZ;
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
      for (int kz = FROM; kz < CF::WZ; kz++) {
        int z = zb + kz;
        if (z >= P.n2) z -= P.n2;
        if (z >= P.n2) z %= P.n2;
        const C v = rows.valid[j] ? G[rows.off[j] + z] : make_c<T>((T) 0, (T) 0);
        winr[j][kz] = v.x;
        wini[j][kz] = v.y;
      }
    }
  };
  auto shift = [&]() {
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int kz = 0; kz + CF::SZ < CF::WZ; kz++) {
        winr[j][kz] = winr[j][kz + CF::SZ];
        wini[j][kz] = wini[j][kz + CF::SZ];
      }
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NWARPS = CF::THREADS / 32;

  for (long long k = k0; k < k1; k += CF::NB) {
    const int nb = (int) min((long long) CF::NB, k1 - k);
    __syncthreads();
    stage_batch<T, W>(pads, slab, xt, table, k, nb, a, b, P);
    __syncthreads();
    T pr[CF::NB], pi[CF::NB];
#pragma unroll
    for (int i = 0; i < CF::NB; i++) {
      pr[i] = (T) 0;
      pi[i] = (T) 0;
      if (i < nb) {
        const int s = slab[i];
        if (cur < 0) { cur = s; fill(std::integral_constant<int, 0>()); }
        while (cur < s) {
          if (s - cur >= (CF::WZ + CF::SZ - 1) / CF::SZ) {
            cur = s;
            fill(std::integral_constant<int, 0>());
          } else {
            shift();
            cur++;
            fill(std::integral_constant<int, CF::WZ - CF::SZ>());
          }
        }
        const T *pd = pads[i];
        const T w0 = rows.valid[0] ? pd[rows.l0[0]] * pd[CF::F0 + rows.l1[0]] : (T) 0;
        const T w1 = rows.valid[1] ? pd[rows.l0[1]] * pd[CF::F0 + rows.l1[1]] : (T) 0;
        const T *p2 = pd + CF::F0 + CF::F1;
        T t0r = (T) 0, t0i = (T) 0, t1r = (T) 0, t1i = (T) 0;
#pragma unroll
        for (int kz = 0; kz < CF::WZ; kz++) {
          const T p = p2[kz];
          t0r += p * winr[0][kz];
          t0i += p * wini[0][kz];
          t1r += p * winr[1][kz];
          t1i += p * wini[1][kz];
        }
        pr[i] = w0 * t0r + w1 * t1r;
        pi[i] = w0 * t0i + w1 * t1i;
      }
    }
    // CTA-wide reduction of the NB partial sums: shared-memory transpose, one warp per node
#pragma unroll
    for (int i = 0; i < CF::NB; i++) red[i][threadIdx.x] = make_c<T>(pr[i], pi[i]);
    __syncthreads();
    for (int i = warp; i < nb; i += NWARPS) {
      T sr = (T) 0, si = (T) 0;
#pragma unroll
      for (int q = 0; q < NWARPS; q++) {
        const C v = red[i][lane + 32 * q];
        sr += v.x;
        si += v.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
      }
      if (lane == 0) f[perm[k + i]] = make_c<T>(sr, si);
    }
  }
}

TileParams make_params(const nfftcu_ctx *c) {
  TileParams P;
  P.n0 = (int) c->n[0];
  P.n1 = (int) c->n[1];
  P.n2 = (int) c->n[2];
  P.NT0 = (P.n0 + kT0 - 1) / kT0;
  P.NT1 = (P.n1 + kT1 - 1) / kT1;
  P.NS = (P.n2 + kSZ - 1) / kSZ;
  const long long tiles = (long long) P.NT0 * P.NT1;
  long long zseg = (8ll * c->sm_count + tiles - 1) / tiles;
  if (zseg < 1) zseg = 1;
  if (zseg > P.NS) zseg = P.NS;
  P.zseg = (int) zseg;
  P.m = (int) c->m;
  P.m2 = (double) c->m * (double) c->m;
  P.b0 = c->b[0];
  P.b1 = c->b[1];
  P.b2 = c->b[2];
  return P;
}

template <typename T, int W>
int launch_spread(nfftcu_ctx *c, const void *f_dev, const TileParams &P) {
  typedef typename Cplx<T>::type C;
  const unsigned grid = (unsigned) ((long long) P.NT0 * P.NT1 * P.zseg);
  spread_tile_kernel<T, W><<<grid, Cfg<W>::THREADS, 0, c->stream>>>(
      (C *) c->grid, (const T *) c->tile_x, c->tile_perm, (const C *) f_dev, c->bin_start,
      (const T *) c->tile_psi, P);
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

template <typename T, int W>
int launch_interp(nfftcu_ctx *c, void *f_dev, const TileParams &P) {
  typedef typename Cplx<T>::type C;
  const unsigned grid = (unsigned) ((long long) P.NT0 * P.NT1 * P.zseg);
  interp_tile_kernel<T, W><<<grid, Cfg<W>::THREADS, 0, c->stream>>>(
      (const C *) c->grid, (const T *) c->tile_x, c->tile_perm, (C *) f_dev, c->bin_start,
      (const T *) c->tile_psi, P);
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

template <typename T>
int dispatch(nfftcu_ctx *c, const void *f_in, void *f_out, bool spread) {
  const TileParams P = make_params(c);
#define NFFTCU_TILE_CASE(Wv)                                                        \
  case Wv:                                                                          \
    return spread ? launch_spread<T, Wv>(c, f_in, P) : launch_interp<T, Wv>(c, f_out, P);
  switch (2 * (int) c->m + 2) {
    NFFTCU_TILE_CASE(6)
    NFFTCU_TILE_CASE(8)
    NFFTCU_TILE_CASE(10)
    NFFTCU_TILE_CASE(12)
    NFFTCU_TILE_CASE(14)
    NFFTCU_TILE_CASE(16)
    NFFTCU_TILE_CASE(18)
    default: break;
  }
#undef NFFTCU_TILE_CASE
  set_error("tile3d: unsupported window cut-off m=%lld", (long long) c->m);
  return NFFTCU_EINVAL;
}

template <typename T>
__global__ void tile_psi_kernel(const T *__restrict__ xt, T *__restrict__ table, long long M,
                                TileParams P) {
  const int W = 2 * P.m + 2;
  const long long total = M * 3 * W;
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long k = i / (3 * W);
    const int r = (int) (i - k * 3 * W);
    const int t = r / W, l = r - t * W;
    const T x = xt[k * 3 + t];
    const int n = (t == 0) ? P.n0 : (t == 1) ? P.n1 : P.n2;
    const double bb = (t == 0) ? P.b0 : (t == 1) ? P.b1 : P.b2;
    const long long uu = cell_of(x, n) - P.m;
    table[i] = (T) kb_phi((double) x * (double) n - (double) (uu + l), P.m2, bb);
  }
}

}  // namespace

bool tile3d_supported(const nfftcu_ctx *c) {
  if (c->d != 3 || c->direct_only) return false;
  if (c->m < 2 || c->m > 8) return false;
  for (int t = 0; t < 3; t++)
    if (c->n[t] > 0x3fffffff) return false;
  return true;
}

// tile-binned processing order: keys, stable sort, node gather, bin offsets, optional psi table
int tile3d_bin_nodes(nfftcu_ctx *c) {
  const long long M = c->M;
  c->tile_ready = false;
  if (M == 0) return NFFTCU_OK;
  const TileParams P = make_params(c);
  const long long nbins = (long long) P.NT0 * P.NT1 * P.NS;
  if (!c->tile_keys) NFFTCU_CUDA(cudaMalloc(&c->tile_keys, sizeof(uint64_t) * (size_t) M));
  if (!c->tile_perm) NFFTCU_CUDA(cudaMalloc((void **) &c->tile_perm, sizeof(uint32_t) * (size_t) M));
  if (!c->tile_x) NFFTCU_CUDA(cudaMalloc(&c->tile_x, real_size(c) * (size_t) M * 3));
  if (!c->bin_start || c->tile_nbins != nbins) {
    if (c->bin_start) cudaFree(c->bin_start);
    NFFTCU_CUDA(cudaMalloc((void **) &c->bin_start, sizeof(uint32_t) * (size_t) (nbins + 1)));
    c->tile_nbins = nbins;
  }
  const int kb = 256;
  const unsigned kgrid = (unsigned) ((M + kb - 1) / kb);
  if (c->prec == NFFTCU_DOUBLE)
    tile_keys_kernel<double><<<kgrid, kb, 0, c->stream>>>((const double *) c->x_dev,
                                                         (uint64_t *) c->tile_keys, c->tile_perm, M, P);
  else
    tile_keys_kernel<float><<<kgrid, kb, 0, c->stream>>>((const float *) c->x_dev,
                                                        (uint64_t *) c->tile_keys, c->tile_perm, M, P);
  c->launches++;
  int bits = 0;
  while ((1ll << bits) < nbins && bits < 62) bits++;
  NFFTCU_TRY(radix_sort_pairs(c, (uint64_t *) c->tile_keys, c->tile_perm, M, bits));
  NFFTCU_TRY(gather_nodes(c, c->tile_perm, c->tile_x));
  bin_bounds_kernel<<<(unsigned) ((nbins + 1 + kb - 1) / kb), kb, 0, c->stream>>>(
      (const uint64_t *) c->tile_keys, c->bin_start, nbins, M);
  c->launches++;
  if (c->opt_psi_table) {
    const size_t bytes = real_size(c) * (size_t) M * 3 * (2 * (size_t) c->m + 2);
    if (!c->tile_psi) NFFTCU_CUDA(cudaMalloc(&c->tile_psi, bytes));
    long long blocks = (M * 3 * (2 * c->m + 2) + kb - 1) / kb;
    if (blocks > (long long) c->sm_count * 16) blocks = (long long) c->sm_count * 16;
    if (c->prec == NFFTCU_DOUBLE)
      tile_psi_kernel<double><<<(unsigned) blocks, kb, 0, c->stream>>>((const double *) c->tile_x,
                                                                      (double *) c->tile_psi, M, P);
    else
      tile_psi_kernel<float><<<(unsigned) blocks, kb, 0, c->stream>>>((const float *) c->tile_x,
                                                                     (float *) c->tile_psi, M, P);
    c->launches++;
  }
  NFFTCU_CUDA(cudaGetLastError());
  c->tile_ready = true;
  return NFFTCU_OK;
}

int tile3d_interp(nfftcu_ctx *c, void *f_dev) {
  return c->prec == NFFTCU_DOUBLE ? dispatch<double>(c, nullptr, f_dev, false)
                                  : dispatch<float>(c, nullptr, f_dev, false);
}

int tile3d_spread(nfftcu_ctx *c, const void *f_dev) {
  return c->prec == NFFTCU_DOUBLE ? dispatch<double>(c, f_dev, nullptr, true)
                                  : dispatch<float>(c, f_dev, nullptr, true);
}

}  // namespace nfftcu
