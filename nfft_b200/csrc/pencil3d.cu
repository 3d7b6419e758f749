// pencil3d.cu -- 3-D interpolation (B) and spreading (B^T) as a "pencil sweep" with register-resident
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
// Warp specialisation inside a CTA.  The consumer warps (two footprint rows per thread) do nothing but
// the node loop: read the node's padded window vectors from shared memory, FMA into / out of their
// register windows, advance the window when the slab changes.  One producer warp runs up to
// STAGES-1 batches (of NB nodes) ahead: it streams the raw node data (x, f) from HBM, evaluates the
// window (piecewise polynomial of kbpoly.cu, the optional per-node table, or the closed form), writes
// the zero-padded vectors into a ring of STAGES shared-memory stages and, for interpolation,
// reduces the consumers' per-thread partial sums of finished batches.  Stages are handed over with
// named barriers (bar.sync / bar.arrive, one full/empty pair per stage); the consumers never wait
// on global memory and there is no CTA-wide __syncthreads in the steady state.
#include "common.cuh"

namespace nfftcu {

namespace {

constexpr int kT0 = 3, kT1 = 3, kSZ = 2, kNB = 8, kStages = 4;

template <int W_>
struct Cfg {
  static constexpr int W = W_, T0 = kT0, T1 = kT1, SZ = kSZ, NB = kNB, STAGES = kStages;
  static constexpr int F0 = T0 + W - 1, F1 = T1 + W - 1, ROWS = F0 * F1;
  static constexpr int WZ = W + SZ - 1;
  static constexpr int WZP = (WZ + 1) & ~1;
  static constexpr int CT = ((((ROWS + 1) / 2) + 31) / 32) * 32;   // consumer threads
  static constexpr int NWARPS = CT / 32;                           // consumer warps
  static constexpr int THREADS = CT + 32;                          // + one producer warp
  static constexpr int PADLEN = F0 + F1 + WZP;
  static constexpr int MINB = THREADS <= 160 ? 2 : 1;
  static constexpr int RETIRE_ALL = (WZ + SZ - 1) / SZ;   // slabs after which the whole window has left
  static_assert(NB * 4 == 32, "the producer warp loads 3 coordinates + 1 sample per node with one lane each");
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

__device__ __forceinline__ int wrap_z(int z, int n2) {
  if (z >= n2) z -= n2;
  if (z >= n2) z %= n2;
  return z;
}

__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }

__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
#define BAR_FULL(s) (1 + (s))
#define BAR_EMPTY(s) (1 + kStages + (s))

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
  C *red;       // [STAGES][NB][CT]     interpolation partial sums
  C *padf;      // [STAGES][NB]         f_j of the batch (spreading)
  double *poly; // [polyN]
  T *pads;      // [STAGES][NB][PADLEN]
  T *rawx;      // [NB*3]               producer staging
  int *slab;    // [STAGES][NB]
  int *nbv;     // [STAGES]             nodes in the stage

  __host__ __device__ static size_t bytes(int polyN) {
    size_t b = 0;
    if (!SPREAD) b += sizeof(C) * CF::STAGES * CF::NB * CF::CT;
    b += sizeof(C) * CF::STAGES * CF::NB;
    b += sizeof(double) * (size_t) polyN;
    b += sizeof(T) * CF::STAGES * CF::NB * CF::PADLEN;
    b = (b + 15) & ~(size_t) 15;
    b += sizeof(T) * CF::NB * 3;
    b = (b + 7) & ~(size_t) 7;
    b += sizeof(int) * (CF::STAGES * CF::NB + CF::STAGES);
    return b;
  }
  __device__ __forceinline__ Smem(unsigned char *base, int polyN) {
    size_t o = 0;
    red = reinterpret_cast<C *>(base);
    if (!SPREAD) o += sizeof(C) * CF::STAGES * CF::NB * CF::CT;
    padf = reinterpret_cast<C *>(base + o);
    o += sizeof(C) * CF::STAGES * CF::NB;
    poly = reinterpret_cast<double *>(base + o);
    o += sizeof(double) * (size_t) polyN;
    pads = reinterpret_cast<T *>(base + o);
    o += sizeof(T) * CF::STAGES * CF::NB * CF::PADLEN;
    o = (o + 15) & ~(size_t) 15;
    rawx = reinterpret_cast<T *>(base + o);
    o += sizeof(T) * CF::NB * 3;
    o = (o + 7) & ~(size_t) 7;
    slab = reinterpret_cast<int *>(base + o);
    nbv = slab + CF::STAGES * CF::NB;
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

// ---- producer warp -----------------------------------------------------------------------------------
// Fills ring stage after ring stage: pads[s][i] = [ psi0 padded to F0 | psi1 padded to F1 | psi2 padded to
// WZP ], slab[s][i] = u2 / SZ, padf[s][i] = f_j (spreading).  Every pad element is written exactly once.
// For interpolation it also turns the consumers' partial sums of a finished stage into ft[k].
template <typename T, int W, bool SPREAD>
__device__ __forceinline__ void producer_warp(const Smem<T, W, SPREAD> &S, const TileRange &R,
                                              const TileParams &P, const T *__restrict__ xt,
                                              typename Cplx<T>::type *__restrict__ ft,
                                              const T *__restrict__ table) {
  typedef Cfg<W> CF;
  typedef typename Cplx<T>::type C;
  const int lane = threadIdx.x & 31;
  const int nbatch = (int) ((R.k1 - R.k0 + CF::NB - 1) / CF::NB);

  auto reduce_stage = [&](int bb) {   // interpolation: sum the CT partials of every node of batch bb
    const int s = bb % CF::STAGES;
    const long long kb = R.k0 + (long long) bb * CF::NB;
    const int nb = (int) min((long long) CF::NB, R.k1 - kb);
    const C *red = S.red + (size_t) s * CF::NB * CF::CT;
    for (int i = 0; i < nb; i++) {
      T sr = (T) 0, si = (T) 0;
#pragma unroll
      for (int q = 0; q < CF::NWARPS; q++) {
        const C v = red[i * CF::CT + lane + 32 * q];
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

  // raw data of batch 0 (lanes 0..23: coordinates, lanes 24..31: samples)
  T px = (T) 0;
  C pf = make_c<T>((T) 0, (T) 0);
  auto load_raw = [&](long long kb) {
    px = (T) 0;
    pf = make_c<T>((T) 0, (T) 0);
    if (lane < CF::NB * 3) {
      const long long idx = kb * 3 + lane;
      if (idx < R.k1 * 3) px = xt[idx];
    } else if (SPREAD) {
      const long long idx = kb + (lane - CF::NB * 3);
      if (idx < R.k1) pf = ft[idx];
    }
  };
  load_raw(R.k0);

  for (int bb = 0; bb < nbatch; bb++) {
    const int s = bb % CF::STAGES;
    const long long kb = R.k0 + (long long) bb * CF::NB;
    const int nb = (int) min((long long) CF::NB, R.k1 - kb);
    if (bb >= CF::STAGES) {
      bar_sync(BAR_EMPTY(s), CF::THREADS);      // consumers are done with what was in this stage
      if (!SPREAD) reduce_stage(bb - CF::STAGES);
    }
    if (lane < CF::NB * 3) S.rawx[lane] = px;
    else if (SPREAD) S.padf[s * CF::NB + (lane - CF::NB * 3)] = pf;
    __syncwarp();
    load_raw(kb + CF::NB);                       // next batch: in flight while this one is evaluated

    T *pads = S.pads + (size_t) s * CF::NB * CF::PADLEN;
#pragma unroll 4
    for (int it = lane; it < CF::NB * CF::PADLEN; it += 32) {
      const int i = it / CF::PADLEN, q = it - i * CF::PADLEN;
      int t, pos;
      if (q < CF::F0) { t = 0; pos = q; }
      else if (q < CF::F0 + CF::F1) { t = 1; pos = q - CF::F0; }
      else { t = 2; pos = q - CF::F0 - CF::F1; }
      T val = (T) 0;
      if (i < nb) {
        const T x = S.rawx[i * 3 + t];
        const int n = (t == 0) ? P.n0 : (t == 1) ? P.n1 : P.n2;
        const long long cc = cell_of(x, n);
        const int u = wrap_fast(cc - P.m, n);
        const int delta = (t == 0) ? u - R.a * CF::T0 : (t == 1) ? u - R.b * CF::T1 : u % CF::SZ;
        const int l = pos - delta;
        if (t == 2 && pos == 0) S.slab[s * CF::NB + i] = u / CF::SZ;
        if (l >= 0 && l < W) {
          if (table) val = table[((kb + i) * 3 + t) * W + l];
          else if (P.deg >= 0) {
            const double y = 2.0 * ((double) x * (double) n - (double) cc) - 1.0;
            const double *cf = S.poly + (size_t) t * (P.deg + 1) * W + l;
            double acc = cf[P.deg * W];
            for (int k = P.deg - 1; k >= 0; k--) acc = fma(acc, y, cf[k * W]);
            val = (T) acc;
          } else {
            const double bb2 = (t == 0) ? P.b0 : (t == 1) ? P.b1 : P.b2;
            val = (T) kb_phi((double) x * (double) n - (double) (cc - P.m + l), P.m2, bb2);
          }
        }
      }
      pads[it] = val;
    }
    if (lane == 0) S.nbv[s] = nb;
    __syncwarp();
    __threadfence_block();
    bar_arrive(BAR_FULL(s), CF::THREADS);
  }
  if (!SPREAD) {
    for (int bb = max(0, nbatch - CF::STAGES); bb < nbatch; bb++) {
      bar_sync(BAR_EMPTY(bb % CF::STAGES), CF::THREADS);
      reduce_stage(bb);
    }
  }
}

template <typename T> struct PsiLoad;
template <> struct PsiLoad<double> {
  template <int N> static __device__ __forceinline__ void load(const double *p, double (&out)[N]) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
      const double2 v = *reinterpret_cast<const double2 *>(p + k);
      out[k] = v.x;
      out[k + 1] = v.y;
    }
    if (N & 1) out[N - 1] = p[N - 1];
  }
};
template <> struct PsiLoad<float> {
  template <int N> static __device__ __forceinline__ void load(const float *p, float (&out)[N]) {
#pragma unroll
    for (int k = 0; k + 3 < N; k += 4) {
      const float4 v = *reinterpret_cast<const float4 *>(p + k);
      out[k] = v.x; out[k + 1] = v.y; out[k + 2] = v.z; out[k + 3] = v.w;
    }
#pragma unroll
    for (int k = N & ~3; k < N; k++) out[k] = p[k];
  }
};

// ---- spreading ---------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(Cfg<W>::THREADS, Cfg<W>::MINB)
spread_tile_kernel(typename Cplx<T>::type *__restrict__ G, const T *__restrict__ xt,
                   typename Cplx<T>::type *__restrict__ ft,
                   const uint32_t *__restrict__ bin_start, const T *__restrict__ table,
                   const double *__restrict__ poly, int polyN, TileParams P) {
  typedef Cfg<W> CF;
  typedef typename Cplx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TileRange R(bin_start, P);
  if (R.k0 == R.k1) return;
  const Smem<T, W, true> S(smem_raw, polyN);
  for (int i = threadIdx.x; i < polyN; i += CF::THREADS) S.poly[i] = poly[i];
  __syncthreads();

  if (threadIdx.x >= CF::CT) {
    producer_warp<T, W, true>(S, R, P, xt, ft, table);
    return;
  }

  // ---- consumers ----
  const int r0 = threadIdx.x, r1 = threadIdx.x + CF::CT;
  const bool v1 = r1 < CF::ROWS;
  const int l0a = r0 / CF::F1, l1a = r0 - l0a * CF::F1;
  const int l0b = v1 ? r1 / CF::F1 : 0, l1b = v1 ? r1 - l0b * CF::F1 : 0;
  T *const Ga = reinterpret_cast<T *>(G) +
                2 * (((long long) wrap_fast((long long) R.a * CF::T0 + l0a, P.n0) * P.n1 +
                      wrap_fast((long long) R.b * CF::T1 + l1a, P.n1)) * P.n2);
  T *const Gb = reinterpret_cast<T *>(G) +
                2 * (((long long) wrap_fast((long long) R.a * CF::T0 + l0b, P.n0) * P.n1 +
                      wrap_fast((long long) R.b * CF::T1 + l1b, P.n1)) * P.n2);
  const int n2 = P.n2;

  T ar[2][CF::WZ], ai[2][CF::WZ];
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int kz = 0; kz < CF::WZ; kz++) { ar[j][kz] = (T) 0; ai[j][kz] = (T) 0; }
  int cur = -1;   // slab the window is aligned to: it covers z = cur*SZ .. cur*SZ+WZ-1 (mod n2)

#define NFFTCU_RETIRE(CNT)                                                                   \
  {                                                                                          \
    const int zb = cur * CF::SZ;                                                             \
    _Pragma("unroll") for (int kz = 0; kz < (CNT); kz++) {                                   \
      const int z = wrap_z(zb + kz, n2);                                                     \
      red_add(Ga + 2 * z, ar[0][kz]);                                                        \
      red_add(Ga + 2 * z + 1, ai[0][kz]);                                                    \
      if (v1) {                                                                              \
        red_add(Gb + 2 * z, ar[1][kz]);                                                      \
        red_add(Gb + 2 * z + 1, ai[1][kz]);                                                  \
      }                                                                                      \
    }                                                                                        \
    _Pragma("unroll") for (int j = 0; j < 2; j++)                                            \
    _Pragma("unroll") for (int kz = 0; kz < CF::WZ; kz++) {                                  \
      if (kz + (CNT) < CF::WZ) { ar[j][kz] = ar[j][kz + (CNT)]; ai[j][kz] = ai[j][kz + (CNT)]; } \
      else { ar[j][kz] = (T) 0; ai[j][kz] = (T) 0; }                                         \
    }                                                                                        \
  }

  const int nbatch = (int) ((R.k1 - R.k0 + CF::NB - 1) / CF::NB);
  for (int bb = 0; bb < nbatch; bb++) {
    const int s = bb % CF::STAGES;
    bar_sync(BAR_FULL(s), CF::THREADS);
    const int nb = S.nbv[s];
    const T *pd = S.pads + (size_t) s * CF::NB * CF::PADLEN;
    for (int i = 0; i < nb; i++, pd += CF::PADLEN) {
      const int sl = S.slab[s * CF::NB + i];
      if (sl != cur) {
        if (cur < 0) cur = sl;
        while (cur < sl) {
          if (sl - cur >= CF::RETIRE_ALL) {
            NFFTCU_RETIRE(CF::WZ)
            cur = sl;
          } else {
            NFFTCU_RETIRE(CF::SZ)
            cur++;
          }
        }
      }
      const C fj = S.padf[s * CF::NB + i];
      const T w0 = pd[l0a] * pd[CF::F0 + l1a];
      const T w1 = v1 ? pd[l0b] * pd[CF::F0 + l1b] : (T) 0;
      const T a0 = w0 * fj.x, b0 = w0 * fj.y, a1 = w1 * fj.x, b1 = w1 * fj.y;
      T p2[CF::WZ];
      PsiLoad<T>::load(pd + CF::F0 + CF::F1, p2);
#pragma unroll
      for (int kz = 0; kz < CF::WZ; kz++) {
        ar[0][kz] += a0 * p2[kz];
        ai[0][kz] += b0 * p2[kz];
        ar[1][kz] += a1 * p2[kz];
        ai[1][kz] += b1 * p2[kz];
      }
    }
    bar_arrive(BAR_EMPTY(s), CF::THREADS);
  }
  if (cur >= 0) NFFTCU_RETIRE(CF::WZ)
#undef NFFTCU_RETIRE
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
  __syncthreads();

  if (threadIdx.x >= CF::CT) {
    producer_warp<T, W, false>(S, R, P, xt, ft, table);
    return;
  }

  // ---- consumers ----
  const int r0 = threadIdx.x, r1 = threadIdx.x + CF::CT;
  const bool v1 = r1 < CF::ROWS;
  const int l0a = r0 / CF::F1, l1a = r0 - l0a * CF::F1;
  const int l0b = v1 ? r1 / CF::F1 : 0, l1b = v1 ? r1 - l0b * CF::F1 : 0;
  const C *const Ga = G + ((long long) wrap_fast((long long) R.a * CF::T0 + l0a, P.n0) * P.n1 +
                           wrap_fast((long long) R.b * CF::T1 + l1a, P.n1)) * P.n2;
  const C *const Gb = G + ((long long) wrap_fast((long long) R.a * CF::T0 + l0b, P.n0) * P.n1 +
                           wrap_fast((long long) R.b * CF::T1 + l1b, P.n1)) * P.n2;
  const int n2 = P.n2;

  T wr[2][CF::WZ], wi[2][CF::WZ];
  T nr[2][CF::SZ], ni[2][CF::SZ];   // the SZ cells that enter the window at the next slab
  int cur = -1;

#define NFFTCU_FILL_ALL()                                                       \
  {                                                                             \
    const int zb = cur * CF::SZ;                                                \
    _Pragma("unroll") for (int kz = 0; kz < CF::WZ; kz++) {                     \
      const int z = wrap_z(zb + kz, n2);                                        \
      const C va = Ga[z], vb = Gb[z];                                           \
      wr[0][kz] = va.x; wi[0][kz] = va.y; wr[1][kz] = vb.x; wi[1][kz] = vb.y;   \
    }                                                                           \
  }
#define NFFTCU_PREFETCH()                                                       \
  {                                                                             \
    const int zb = cur * CF::SZ + CF::WZ;                                       \
    _Pragma("unroll") for (int q = 0; q < CF::SZ; q++) {                        \
      const int z = wrap_z(zb + q, n2);                                         \
      const C va = Ga[z], vb = Gb[z];                                           \
      nr[0][q] = va.x; ni[0][q] = va.y; nr[1][q] = vb.x; ni[1][q] = vb.y;       \
    }                                                                           \
  }
#define NFFTCU_STEP()                                                           \
  {                                                                             \
    _Pragma("unroll") for (int j = 0; j < 2; j++) {                             \
      _Pragma("unroll") for (int kz = 0; kz + CF::SZ < CF::WZ; kz++) {          \
        wr[j][kz] = wr[j][kz + CF::SZ]; wi[j][kz] = wi[j][kz + CF::SZ];         \
      }                                                                         \
      _Pragma("unroll") for (int q = 0; q < CF::SZ; q++) {                      \
        wr[j][CF::WZ - CF::SZ + q] = nr[j][q]; wi[j][CF::WZ - CF::SZ + q] = ni[j][q]; \
      }                                                                         \
    }                                                                           \
  }

  const int nbatch = (int) ((R.k1 - R.k0 + CF::NB - 1) / CF::NB);
  for (int bb = 0; bb < nbatch; bb++) {
    const int s = bb % CF::STAGES;
    bar_sync(BAR_FULL(s), CF::THREADS);
    const int nb = S.nbv[s];
    const T *pd = S.pads + (size_t) s * CF::NB * CF::PADLEN;
    C *red = S.red + (size_t) s * CF::NB * CF::CT + threadIdx.x;
    for (int i = 0; i < nb; i++, pd += CF::PADLEN, red += CF::CT) {
      const int sl = S.slab[s * CF::NB + i];
      if (sl != cur) {
        if (cur < 0) { cur = sl; NFFTCU_FILL_ALL() NFFTCU_PREFETCH() }
        while (cur < sl) {
          if (sl - cur >= CF::RETIRE_ALL) {
            cur = sl;
            NFFTCU_FILL_ALL()
          } else {
            NFFTCU_STEP()
            cur++;
          }
          NFFTCU_PREFETCH()
        }
      }
      const T w0 = pd[l0a] * pd[CF::F0 + l1a];
      const T w1 = v1 ? pd[l0b] * pd[CF::F0 + l1b] : (T) 0;
      T p2[CF::WZ];
      PsiLoad<T>::load(pd + CF::F0 + CF::F1, p2);
      T t0r = (T) 0, t0i = (T) 0, t1r = (T) 0, t1i = (T) 0;
#pragma unroll
      for (int kz = 0; kz < CF::WZ; kz++) {
        t0r += p2[kz] * wr[0][kz];
        t0i += p2[kz] * wi[0][kz];
        t1r += p2[kz] * wr[1][kz];
        t1i += p2[kz] * wi[1][kz];
      }
      *red = make_c<T>(w0 * t0r + w1 * t1r, w0 * t0i + w1 * t1i);
    }
    __threadfence_block();
    bar_arrive(BAR_EMPTY(s), CF::THREADS);
  }
#undef NFFTCU_FILL_ALL
#undef NFFTCU_PREFETCH
#undef NFFTCU_STEP
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
      (C *) c->grid, (const T *) c->tile_x, (C *) c->f_tile, c->bin_start,
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
