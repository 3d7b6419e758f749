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
//   interpolation  the windows hold grid values, refilled SZ cells per slab straight from L2 (the
//                  cells of the next slab are prefetched one slab ahead); a node reduces them against
//                  psi2, weights by psi0 psi1, and the per-thread partial sums of a batch are reduced
//                  across the CTA through shared memory one batch later.
// Per tap this costs 2 FP64 FMAs and, per node and thread, WZ broadcast shared-memory loads of psi2
// -- instead of one 16-byte shared/L1 load per tap -- which moves the kernel from the LSU roof
// (128 B/clk/SM) to the FP64 roof (64 FMA/clk/SM); see DESIGN.md for the arithmetic and
// profiles/ for the measurements.  Zero padding costs (W/F0)(W/F1)(W/WZ) = 71% lane efficiency at m = 6.
//
// Pipeline inside a CTA.  Nodes are consumed in batches of NB.  While batch b is consumed, the raw
// data (x, f) of batch b+2 is in flight from HBM into registers and the padded window vectors of
// batch b+1 are produced into the other half of a double buffer (window values from the
// piecewise polynomial of kbpoly.cu, the optional per-node table, or the closed form); there is
// ONE __syncthreads per batch.
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
  static constexpr int NWARPS = THREADS / 32;
  static constexpr int PADLEN = F0 + F1 + WZP;
  static constexpr int MINB = THREADS <= 128 ? 2 : 1;
  static constexpr int RETIRE_ALL = (WZ + SZ - 1) / SZ;   // slabs after which the whole window has left
};

struct TileParams {
  int n0, n1, n2;
  int NT0, NT1, NS;
  int zseg;
  int m;
  int deg;          // polynomial degree, -1: closed form
  double m2, b0, b1, b2;
};

__device__ __forceinline__ int wrap_fast(long long v, int n) {
  if (v >= 0 && v < n) return (int) v;
  if (v < 0 && v >= -(long long) n) return (int) (v + n);
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
  const int u0 = wrap_fast(cell_of(x[3 * j], P.n0) - P.m, P.n0);
  const int u1 = wrap_fast(cell_of(x[3 * j + 1], P.n1) - P.m, P.n1);
  const int u2 = wrap_fast(cell_of(x[3 * j + 2], P.n2) - P.m, P.n2);
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

template <typename C>
__global__ void gather_f_kernel(const C *__restrict__ f, const uint32_t *__restrict__ perm,
                                C *__restrict__ ft, long long M) {
  const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (k < M) ft[k] = f[perm[k]];
}

template <typename C>
__global__ void scatter_f_kernel(const C *__restrict__ ft, const uint32_t *__restrict__ perm,
                                 C *__restrict__ f, long long M) {
  const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (k < M) f[perm[k]] = ft[k];
}

// ---- shared-memory carve-up ------------------------------------------------------------------------
template <typename T, int W, bool SPREAD>
struct Smem {
  typedef Cfg<W> CF;
  typedef typename Cplx<T>::type C;
  C *red;       // [2][NB][THREADS]   interpolation partial sums
  C *padf;      // [2][NB]            f_j of the batch (spreading)
  C *rawf;      // [2][NB]
  double *poly; // [polyN]
  T *pads;      // [2][NB][PADLEN]
  T *rawx;      // [2][NB*3]
  int *slab;    // [2][NB]

  __host__ __device__ static size_t bytes(int polyN) {
    size_t b = 0;
    if (!SPREAD) b += sizeof(C) * 2 * CF::NB * CF::THREADS;
    b += sizeof(C) * 2 * CF::NB * 2;
    b += sizeof(double) * (size_t) polyN;
    b += sizeof(T) * 2 * CF::NB * CF::PADLEN;
    b += sizeof(T) * 2 * CF::NB * 3;
    b = (b + 7) & ~(size_t) 7;
    b += sizeof(int) * 2 * CF::NB;
    return b;
  }
  __device__ Smem(unsigned char *base, int polyN) {
    size_t o = 0;
    red = reinterpret_cast<C *>(base);
    if (!SPREAD) o += sizeof(C) * 2 * CF::NB * CF::THREADS;
    padf = reinterpret_cast<C *>(base + o);
    o += sizeof(C) * 2 * CF::NB;
    rawf = reinterpret_cast<C *>(base + o);
    o += sizeof(C) * 2 * CF::NB;
    poly = reinterpret_cast<double *>(base + o);
    o += sizeof(double) * (size_t) polyN;
    pads = reinterpret_cast<T *>(base + o);
    o += sizeof(T) * 2 * CF::NB * CF::PADLEN;
    rawx = reinterpret_cast<T *>(base + o);
    o += sizeof(T) * 2 * CF::NB * 3;
    o = (o + 7) & ~(size_t) 7;
    slab = reinterpret_cast<int *>(base + o);
  }
};

// raw node data of one batch: issue the global loads (into registers) / park them in shared memory
template <typename T, int W, bool SPREAD>
struct RawRegs {
  typedef Cfg<W> CF;
  typedef typename Cplx<T>::type C;
  T px;
  C pf;
  __device__ __forceinline__ void load(const T *__restrict__ xt, const C *__restrict__ ft,
                                       long long kb, long long k1) {
    const int tid = threadIdx.x;
    px = (T) 0;
    pf = make_c<T>((T) 0, (T) 0);
    if (tid < CF::NB * 3) {
      const long long idx = kb * 3 + tid;
      if (idx < k1 * 3) px = xt[idx];
    } else if (SPREAD && tid < CF::NB * 4) {
      const long long idx = kb + (tid - CF::NB * 3);
      if (idx < k1) pf = ft[idx];
    }
  }
  __device__ __forceinline__ void park(const Smem<T, W, SPREAD> &S, int rb) const {
    const int tid = threadIdx.x;
    if (tid < CF::NB * 3) S.rawx[rb * CF::NB * 3 + tid] = px;
    else if (SPREAD && tid < CF::NB * 4) S.rawf[rb * CF::NB + (tid - CF::NB * 3)] = pf;
  }
};

// padded window vectors of one batch: every element of pads[pb] is written exactly once
//   pads[i] = [ psi0 padded to F0 | psi1 padded to F1 | psi2 padded to WZP ],  slab[i] = u2 / SZ
template <typename T, int W, bool SPREAD>
__device__ __forceinline__ void produce(const Smem<T, W, SPREAD> &S, int pb, int rb, long long kb,
                                        int nb, int a, int b, const TileParams &P,
                                        const T *__restrict__ table) {
  typedef Cfg<W> CF;
  for (int it = threadIdx.x; it < CF::NB * CF::PADLEN; it += CF::THREADS) {
    const int i = it / CF::PADLEN, q = it - i * CF::PADLEN;
    int t, pos;
    if (q < CF::F0) { t = 0; pos = q; }
    else if (q < CF::F0 + CF::F1) { t = 1; pos = q - CF::F0; }
    else { t = 2; pos = q - CF::F0 - CF::F1; }
    T val = (T) 0;
    if (i < nb) {
      const T x = S.rawx[(rb * CF::NB + i) * 3 + t];
      const int n = (t == 0) ? P.n0 : (t == 1) ? P.n1 : P.n2;
      const long long cc = cell_of(x, n);
      const int u = wrap_fast(cc - P.m, n);
      const int delta = (t == 0) ? u - a * CF::T0 : (t == 1) ? u - b * CF::T1 : u % CF::SZ;
      const int l = pos - delta;
      if (t == 2 && pos == 0) S.slab[pb * CF::NB + i] = u / CF::SZ;
      if (l >= 0 && l < W) {
        if (table) val = table[((kb + i) * 3 + t) * W + l];
        else if (P.deg >= 0) {
          const double y = 2.0 * ((double) x * (double) n - (double) cc) - 1.0;
          const double *cf = S.poly + (size_t) t * (P.deg + 1) * W + l;
          double acc = cf[P.deg * W];
          for (int k = P.deg - 1; k >= 0; k--) acc = fma(acc, y, cf[k * W]);
          val = (T) acc;
        } else {
          const double bb = (t == 0) ? P.b0 : (t == 1) ? P.b1 : P.b2;
          val = (T) kb_phi((double) x * (double) n - (double) (cc - P.m + l), P.m2, bb);
        }
      }
    }
    S.pads[(pb * CF::NB + i) * CF::PADLEN + q] = val;
  }
  if (SPREAD && threadIdx.x < CF::NB)
    S.padf[pb * CF::NB + threadIdx.x] = S.rawf[rb * CF::NB + threadIdx.x];
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
      const int g0 = wrap_fast((long long) a * CF::T0 + l0[j], P.n0);
      const int g1 = wrap_fast((long long) b * CF::T1 + l1[j], P.n1);
      off[j] = ((long long) g0 * P.n1 + g1) * P.n2;
    }
  }
};

struct TileRange {
  int a, b;
  long long k0, k1;
  __device__ __forceinline__ TileRange(const uint32_t *__restrict__ bin_start, const TileParams &P) {
    const int tile = blockIdx.x / P.zseg, seg = blockIdx.x - tile * P.zseg;
    a = tile / P.NT1;
    b = tile - a * P.NT1;
    const int s_begin = (int) ((long long) P.NS * seg / P.zseg);
    const int s_end = (int) ((long long) P.NS * (seg + 1) / P.zseg);
    const long long bin0 = (long long) tile * P.NS;
    k0 = bin_start[bin0 + s_begin];
    k1 = bin_start[bin0 + s_end];
  }
};

__device__ __forceinline__ int wrap_z(int z, int n2) {
  if (z >= n2) z -= n2;
  if (z >= n2) z %= n2;
  return z;
}

// ---- spreading ---------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(Cfg<W>::THREADS, Cfg<W>::MINB)
spread_tile_kernel(typename Cplx<T>::type *__restrict__ G, const T *__restrict__ xt,
                   const typename Cplx<T>::type *__restrict__ ft,
                   const uint32_t *__restrict__ bin_start, const T *__restrict__ table,
                   const double *__restrict__ poly, int polyN, TileParams P) {
  typedef Cfg<W> CF;
  typedef typename Cplx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TileRange R(bin_start, P);
  if (R.k0 == R.k1) return;
  const Smem<T, W, true> S(smem_raw, polyN);
  for (int i = threadIdx.x; i < polyN; i += CF::THREADS) S.poly[i] = poly[i];

  const RowSetup<T, W> rows(R.a, R.b, P);
  T *Gr = reinterpret_cast<T *>(G);
  T accr[2][CF::WZ], acci[2][CF::WZ];
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int kz = 0; kz < CF::WZ; kz++) { accr[j][kz] = (T) 0; acci[j][kz] = (T) 0; }
  int cur = -1;   // slab the window is aligned to: it covers z = cur*SZ .. cur*SZ+WZ-1 (mod n2)

  // retire cells [0, CNT) of the window to the grid and shift the window down by CNT
  auto retire = [&](auto cnt_tag) {
    constexpr int CNT = decltype(cnt_tag)::value;
    const int zb = cur * CF::SZ;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      if (rows.valid[j]) {
#pragma unroll
        for (int kz = 0; kz < CNT; kz++) {
          T *p = Gr + 2 * (rows.off[j] + wrap_z(zb + kz, P.n2));
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

  const int nbatch = (int) ((R.k1 - R.k0 + CF::NB - 1) / CF::NB);
  RawRegs<T, W, true> raw;
  raw.load(xt, ft, R.k0, R.k1);
  raw.park(S, 0);
  raw.load(xt, ft, R.k0 + CF::NB, R.k1);
  raw.park(S, 1);
  __syncthreads();
  produce<T, W, true>(S, 0, 0, R.k0, (int) min((long long) CF::NB, R.k1 - R.k0), R.a, R.b, P, table);
  __syncthreads();

  for (int bb = 0; bb < nbatch; bb++) {
    const long long kb = R.k0 + (long long) bb * CF::NB;
    const int nb = (int) min((long long) CF::NB, R.k1 - kb);
    const int pb = bb & 1;
    raw.load(xt, ft, kb + 2 * CF::NB, R.k1);            // batch bb+2: in flight during the consume

    for (int i = 0; i < nb; i++) {
      const int s = S.slab[pb * CF::NB + i];
      if (cur < 0) cur = s;
      while (cur < s) {
        if (s - cur >= CF::RETIRE_ALL) {
          retire(std::integral_constant<int, CF::WZ>());
          cur = s;
        } else {
          retire(std::integral_constant<int, CF::SZ>());
          cur++;
        }
      }
      const T *pd = S.pads + (pb * CF::NB + i) * CF::PADLEN;
      const C fj = S.padf[pb * CF::NB + i];
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

    if (bb + 1 < nbatch)
      produce<T, W, true>(S, pb ^ 1, pb ^ 1, kb + CF::NB, (int) min((long long) CF::NB, R.k1 - kb - CF::NB),
                          R.a, R.b, P, table);
    raw.park(S, pb);   // raw[pb] was last read by the produce of the previous iteration
    __syncthreads();
  }
  if (cur >= 0) retire(std::integral_constant<int, CF::WZ>());
}

// ---- interpolation -----------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(Cfg<W>::THREADS, Cfg<W>::MINB)
interp_tile_kernel(const typename Cplx<T>::type *__restrict__ G, const T *__restrict__ xt,
                   typename Cplx<T>::type *__restrict__ ft, const uint32_t *__restrict__ bin_start,
                   const T *__restrict__ table, const double *__restrict__ poly, int polyN,
                   TileParams P) {
  typedef Cfg<W> CF;
  typedef typename Cplx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TileRange R(bin_start, P);
  if (R.k0 == R.k1) return;
  const Smem<T, W, false> S(smem_raw, polyN);
  for (int i = threadIdx.x; i < polyN; i += CF::THREADS) S.poly[i] = poly[i];

  const RowSetup<T, W> rows(R.a, R.b, P);
  T winr[2][CF::WZ], wini[2][CF::WZ];
  T nxr[2][CF::SZ], nxi[2][CF::SZ];   // the SZ cells that enter the window at the next slab
  int cur = -1;

  auto fill_all = [&]() {
    const int zb = cur * CF::SZ;
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int kz = 0; kz < CF::WZ; kz++) {
        const C v = rows.valid[j] ? G[rows.off[j] + wrap_z(zb + kz, P.n2)] : make_c<T>((T) 0, (T) 0);
        winr[j][kz] = v.x;
        wini[j][kz] = v.y;
      }
  };
  auto prefetch_next = [&]() {
    const int zb = cur * CF::SZ + CF::WZ;
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int q = 0; q < CF::SZ; q++) {
        const C v = rows.valid[j] ? G[rows.off[j] + wrap_z(zb + q, P.n2)] : make_c<T>((T) 0, (T) 0);
        nxr[j][q] = v.x;
        nxi[j][q] = v.y;
      }
  };
  auto step_one = [&]() {   // move the window up by one slab using the prefetched cells
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
      for (int kz = 0; kz + CF::SZ < CF::WZ; kz++) {
        winr[j][kz] = winr[j][kz + CF::SZ];
        wini[j][kz] = wini[j][kz + CF::SZ];
      }
#pragma unroll
      for (int q = 0; q < CF::SZ; q++) {
        winr[j][CF::WZ - CF::SZ + q] = nxr[j][q];
        wini[j][CF::WZ - CF::SZ + q] = nxi[j][q];
      }
    }
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto reduce_batch = [&](int bb) {   // partial sums of batch bb -> ft (tile order), one warp per node
    const long long kb = R.k0 + (long long) bb * CF::NB;
    const int nb = (int) min((long long) CF::NB, R.k1 - kb);
    const C *red = S.red + (size_t) (bb & 1) * CF::NB * CF::THREADS;
    for (int i = warp; i < nb; i += CF::NWARPS) {
      T sr = (T) 0, si = (T) 0;
#pragma unroll
      for (int q = 0; q < CF::NWARPS; q++) {
        const C v = red[i * CF::THREADS + lane + 32 * q];
        sr += v.x;
        si += v.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
      }
      if (lane == 0) ft[kb + i] = make_c<T>(sr, si);
    }
  };

  const int nbatch = (int) ((R.k1 - R.k0 + CF::NB - 1) / CF::NB);
  RawRegs<T, W, false> raw;
  raw.load(xt, nullptr, R.k0, R.k1);
  raw.park(S, 0);
  raw.load(xt, nullptr, R.k0 + CF::NB, R.k1);
  raw.park(S, 1);
  __syncthreads();
  produce<T, W, false>(S, 0, 0, R.k0, (int) min((long long) CF::NB, R.k1 - R.k0), R.a, R.b, P, table);
  __syncthreads();

  for (int bb = 0; bb < nbatch; bb++) {
    const long long kb = R.k0 + (long long) bb * CF::NB;
    const int nb = (int) min((long long) CF::NB, R.k1 - kb);
    const int pb = bb & 1;
    raw.load(xt, nullptr, kb + 2 * CF::NB, R.k1);
    if (bb > 0) reduce_batch(bb - 1);

    C *red = S.red + (size_t) pb * CF::NB * CF::THREADS;
    for (int i = 0; i < nb; i++) {
      const int s = S.slab[pb * CF::NB + i];
      if (cur < 0) { cur = s; fill_all(); prefetch_next(); }
      while (cur < s) {
        if (s - cur >= CF::RETIRE_ALL) {
          cur = s;
          fill_all();
        } else {
          step_one();
          cur++;
        }
        prefetch_next();
      }
      const T *pd = S.pads + (pb * CF::NB + i) * CF::PADLEN;
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
      red[i * CF::THREADS + threadIdx.x] = make_c<T>(w0 * t0r + w1 * t1r, w0 * t0i + w1 * t1i);
    }

    if (bb + 1 < nbatch)
      produce<T, W, false>(S, pb ^ 1, pb ^ 1, kb + CF::NB, (int) min((long long) CF::NB, R.k1 - kb - CF::NB),
                           R.a, R.b, P, table);
    raw.park(S, pb);
    __syncthreads();
  }
  reduce_batch(nbatch - 1);
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
  P.deg = c->kbpoly_deg;
  P.m2 = (double) c->m * (double) c->m;
  P.b0 = c->b[0];
  P.b1 = c->b[1];
  P.b2 = c->b[2];
  return P;
}

template <typename T, int W>
int launch_spread(nfftcu_ctx *c, const void *f_dev, const TileParams &P) {
  typedef typename Cplx<T>::type C;
  const int kb = 256;
  gather_f_kernel<C><<<(unsigned) ((c->M + kb - 1) / kb), kb, 0, c->stream>>>(
      (const C *) f_dev, c->tile_perm, (C *) c->f_tile, c->M);
  const int polyN = (c->tile_psi || P.deg < 0) ? 0 : 3 * (P.deg + 1) * W;
  const size_t smem = Smem<T, W, true>::bytes(polyN);
  NFFTCU_CUDA(cudaFuncSetAttribute(spread_tile_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  const unsigned grid = (unsigned) ((long long) P.NT0 * P.NT1 * P.zseg);
  spread_tile_kernel<T, W><<<grid, Cfg<W>::THREADS, smem, c->stream>>>(
      (C *) c->grid, (const T *) c->tile_x, (const C *) c->f_tile, c->bin_start,
      (const T *) c->tile_psi, (const double *) c->kbpoly_dev, polyN, P);
  c->launches += 2;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

template <typename T, int W>
int launch_interp(nfftcu_ctx *c, void *f_dev, const TileParams &P) {
  typedef typename Cplx<T>::type C;
  const int polyN = (c->tile_psi || P.deg < 0) ? 0 : 3 * (P.deg + 1) * W;
  const size_t smem = Smem<T, W, false>::bytes(polyN);
  NFFTCU_CUDA(cudaFuncSetAttribute(interp_tile_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  const unsigned grid = (unsigned) ((long long) P.NT0 * P.NT1 * P.zseg);
  interp_tile_kernel<T, W><<<grid, Cfg<W>::THREADS, smem, c->stream>>>(
      (const C *) c->grid, (const T *) c->tile_x, (C *) c->f_tile, c->bin_start,
      (const T *) c->tile_psi, (const double *) c->kbpoly_dev, polyN, P);
  const int kb = 256;
  scatter_f_kernel<C><<<(unsigned) ((c->M + kb - 1) / kb), kb, 0, c->stream>>>(
      (const C *) c->f_tile, c->tile_perm, (C *) f_dev, c->M);
  c->launches += 2;
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
  if (!c->f_tile) NFFTCU_CUDA(cudaMalloc(&c->f_tile, 2 * real_size(c) * (size_t) M));
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
