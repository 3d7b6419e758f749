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
// A group of consumer warps owns one tile and sweeps a range of slabs along the contiguous axis z.
// Every tap box of the tile lies inside the tile's FOOTPRINT of F0 x F1 = (T0+W-1) x (T1+W-1) grid rows
// (W = 2m+2; 16 x 16 rows for m = 6) and, for the current slab, inside a z-window of WZ = W+SZ-1
// cells.  Each consumer thread owns two footprint rows and keeps their z-windows -- 2 x WZ complex
// values -- in REGISTERS:
//   spreading      the windows are accumulators; a node adds (psi0 psi1 f_j) * psi2[k] to all WZ cells
//                  of the thread's rows (psi vectors are zero-padded to the footprint / window, so the
//                  register indices are static); when the sweep leaves a slab, the SZ cells that
//                  fall out of the window are retired to the grid with RED.ADD and the window shifts.
//                  No shared-memory accumulation, no intra-CTA conflicts: a row has one owner.
//   interpolation  the windows hold grid values, refilled SZ cells per slab straight from L2 (the
//                  cells of the next slab are prefetched one slab ahead); a node reduces them against
//                  psi2, weights by psi0 psi1, and the per-thread partial sums are reduced across the
//                  group through shared memory by the (otherwise idle) producer warp.
// Per tap this costs 2 FP64 FMAs and, per node and thread, WZ broadcast shared-memory loads of psi2
// -- instead of one 16-byte shared/L1 load per tap -- which moves the kernel from the LSU roof
// (128 B/clk/SM) to the FP64 roof (64 FMA/clk/SM); see DESIGN.md for the arithmetic and
// profiles/ for the measurements.  Zero padding costs (W/F0)(W/F1)(W/WZ) = 71% lane efficiency at m = 6.
//
// Two kernels per transform.
//  (1) expand_nodes_kernel: a streaming pass over the nodes (tile order) that evaluates the window --
//      piecewise polynomial of kbpoly.cu (Horner chains with the coefficients of a tap in registers),
//      or the closed form -- and writes one fixed-size RECORD per node:
//          [ psi0 padded to F0 | psi1 padded to F1 | psi2 padded to WZP | slab | f.re | f.im ]
//      HBM-bound (REC*sizeof(T) = 416 B per node in fp64) and trivially parallel.  With the PRE_PSI
//      analogue (NFFTCU_OPT_PSI_TABLE) the records are kept across transforms and only f is refreshed.
//  (2) the pencil kernel.  A CTA is one group of consumer warps (128 threads for m <= 6: two CTAs per
//      SM at 255 registers per thread -- the 2 x WZ complex windows alone are 120; the register file is
//      handed out in units of 4 warps, so adding a producer warp to the CTA would cap every thread at
//      168 registers and spill the windows, profiles/r01h..r01q).  Thread 0 doubles as the TMA issuer:
//      it keeps a ring of STAGES shared-memory stages filled with bulk copies (cp.async.bulk + mbarrier
//      complete_tx) of NB records at a time, up to STAGES-1 batches ahead.  The consumers wait on the stage's
//      "full" mbarrier, run the node loop out of shared memory and registers only, and release the
//      stage through its "empty" mbarrier; they never wait on global memory except for the window
//      refill of interpolation, which is prefetched one slab ahead.
#include "common.cuh"

namespace nfftcu {

namespace {

constexpr int kT0 = 3, kT1 = 3, kSZ = 2, kNB = 8, kStages = 4;

template <typename T, int W_>
struct Cfg {
  static constexpr int W = W_, T0 = kT0, T1 = kT1, SZ = kSZ, NB = kNB, STAGES = kStages;
  static constexpr int F0 = T0 + W - 1, F1 = T1 + W - 1, ROWS = F0 * F1;
  static constexpr int WZ = W + SZ - 1;
  static constexpr int WZP = (WZ + 3) & ~3;
  static constexpr int PADLEN = F0 + F1 + WZP;
  static constexpr int RALIGN = 16 / (int) sizeof(T);
  static constexpr int REC = ((PADLEN + 3 + RALIGN - 1) / RALIGN) * RALIGN;   // record length in T
  static constexpr int CT = ((((ROWS + 1) / 2) + 31) / 32) * 32;   // consumer threads of a group
  static constexpr int NWARPS = CT / 32;
  static constexpr int MINB = CT <= 128 ? 2 : 1;                   // CTAs per SM the register budget is sized for
  static constexpr int RETIRE_ALL = (WZ + SZ - 1) / SZ;   // slabs after which the whole window has left
  static constexpr int VMAX = F0 > WZP ? (F0 > F1 ? F0 : F1) : (WZP > F1 ? WZP : F1);
  static constexpr int LPI = VMAX <= 16 ? 16 : 32, IPP = 32 / LPI;   // lanes per item, items per pass
  static_assert(NB * 4 == 32, "expand: 3 coordinate lanes + 1 sample lane per node");
  static_assert(VMAX <= 32, "padded vectors longer than a warp");
};

struct TileParams {
  int n0, n1, n2;
  int NT0, NT1, NS;
  int zseg;
  int m;
  int deg;          // kKbPolyDeg when the polynomial window is available, -1: closed form
  double m2, b0, b1, b2;
  double ws0, ws1, ws2;   // power-of-two window scale per dimension
  int window;
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

// ---- mbarrier / TMA bulk copy ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// producer-side wait: the producer runs STAGES batches ahead and mostly waits; sleep between polls so
// that the polling loop does not take issue slots from the consumer warps of its scheduler
// (profiles/r01p: the bare try_wait loop was 25% of all executed instructions)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, int parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(200);
  }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

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
__global__ void scatter_f_kernel(const C *__restrict__ ft, const uint32_t *__restrict__ perm,
                                 C *__restrict__ f, long long M) {
  const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (k < M) f[perm[k]] = ft[k];
}

// only the samples of the records (spreading with cached records)
template <typename T, int W>
__global__ void refresh_f_kernel(T *__restrict__ rec, const typename Cplx<T>::type *__restrict__ f,
                                 const uint32_t *__restrict__ perm, long long M) {
  typedef Cfg<T, W> CF;
  const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M) return;
  const typename Cplx<T>::type v = f[perm[k]];
  rec[k * CF::REC + CF::PADLEN + 1] = v.x;
  rec[k * CF::REC + CF::PADLEN + 2] = v.y;
}

// ---- (1) window expansion ------------------------------------------------------------------------------
// One warp per chunk of NB nodes (tile order).  Lanes 0..23 fetch the 3*NB coordinates, lanes 24..31
// the samples.  Polynomial window, lanes <-> taps: a lane keeps the kKbPolyDeg+1 coefficients of ITS
// tap in registers (reloaded per dimension: the shape parameter b may differ), LPI lanes serve one
// (node, dimension) item, spare lanes write the zero padding; IPP items per pass and 4 passes are
// independent Horner chains.
template <typename T, int W, bool SPREAD>
__global__ void __launch_bounds__(256)
expand_nodes_kernel(T *__restrict__ rec, const T *__restrict__ xt,
                    const typename Cplx<T>::type *__restrict__ f, const uint32_t *__restrict__ perm,
                    const double *__restrict__ poly, long long M, TileParams P) {
  typedef Cfg<T, W> CF;
  typedef typename Cplx<T>::type C;
  __shared__ double spoly[3 * (kKbPolyDeg + 1) * W];
  if (P.deg >= 0)
    for (int i = threadIdx.x; i < 3 * (kKbPolyDeg + 1) * W; i += blockDim.x) spoly[i] = poly[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
  const long long nchunks = (M + CF::NB - 1) / CF::NB;
  const bool xlane = lane < CF::NB * 3;
  const int half = lane / CF::LPI, l = lane - half * CF::LPI;
  for (long long ch = warp; ch < nchunks; ch += nwarps) {
    const long long kb = ch * CF::NB;
    const int nb = (int) min((long long) CF::NB, M - kb);
    T x = (T) 0;
    if (xlane) {
      if (kb * 3 + lane < M * 3) x = xt[kb * 3 + lane];
    } else if (SPREAD) {
      const long long k = kb + (lane - CF::NB * 3);
      if (k < M) {
        const C v = f[perm[k]];
        rec[k * CF::REC + CF::PADLEN + 1] = v.x;
        rec[k * CF::REC + CF::PADLEN + 2] = v.y;
      }
    }
#pragma unroll 1
    for (int t = 0; t < 3; t++) {
      const int nt = (t == 0) ? P.n0 : (t == 1) ? P.n1 : P.n2;
      const int vb = (t == 0) ? 0 : (t == 1) ? CF::F0 : CF::F0 + CF::F1;
      const int vl = (t == 0) ? CF::F0 : (t == 1) ? CF::F1 : CF::WZP;
      const int tmod = (t == 0) ? CF::T0 : (t == 1) ? CF::T1 : CF::SZ;
      double cf[kKbPolyDeg + 1];
      if (P.deg >= 0) {
#pragma unroll
        for (int k = 0; k <= kKbPolyDeg; k++)
          cf[k] = (l < W) ? spoly[((size_t) t * (kKbPolyDeg + 1) + k) * W + l] : 0.0;
      }
      const double bt = (t == 0) ? P.b0 : (t == 1) ? P.b1 : P.b2;
      const double wst = (t == 0) ? P.ws0 : (t == 1) ? P.ws1 : P.ws2;
#pragma unroll 4
      for (int i0 = 0; i0 < CF::NB; i0 += CF::IPP) {
        const int i = i0 + half;
        const T xi = __shfl_sync(0xffffffffu, x, 3 * i + t);
        const long long cc = cell_of(xi, nt);
        const int u = wrap_fast(cc - P.m, nt);
        const int delta = u % tmod;
        double acc;
        if (P.deg >= 0) {
          const double y = 2.0 * ((double) xi * (double) nt - (double) cc) - 1.0;
          acc = cf[kKbPolyDeg];
#pragma unroll
          for (int k = kKbPolyDeg - 1; k >= 0; k--) acc = fma(acc, y, cf[k]);
        } else {
          acc = (l < W) ? window_phi((double) xi * (double) nt - (double) (cc - P.m + l), P.m2, bt, P.window, wst) : 0.0;
        }
        if (i < nb) {
          T *dst = rec + (kb + i) * CF::REC + vb;
          if (t == 2 && l == 0) rec[(kb + i) * CF::REC + CF::PADLEN] = (T) (u / CF::SZ);
          if (l < W) dst[delta + l] = (T) acc;
          else if (l - W < vl - W) dst[(l - W < delta) ? l - W : l] = (T) 0;
        }
      }
    }
  }
}

// ---- (2) pencil kernels ----------------------------------------------------------------------------------
template <typename T, int W, bool SPREAD>
struct Smem {
  typedef Cfg<T, W> CF;
  typedef typename Cplx<T>::type C;
  C *red;          // [2][NB][CT]        interpolation partial sums (double-buffered across batches)
  C *stg;          // [2*CT][8]          spreading: retired cells staged per row, flushed as 128-byte runs
  long long *rowoff; // [2*CT]           spreading: grid offset of every footprint row
  T *stage;        // [STAGES][NB*REC]   node records
  uint64_t *full;  // [STAGES]
  uint64_t *empty; // [STAGES]

  __host__ __device__ static size_t bytes() {
    size_t b = 0;
    if (!SPREAD) b += sizeof(C) * 2 * CF::NB * CF::CT;
    else b += sizeof(C) * 2 * CF::CT * 8 + sizeof(long long) * 2 * CF::CT;
    b += sizeof(T) * CF::STAGES * CF::NB * CF::REC;
    b += sizeof(uint64_t) * 2 * CF::STAGES;
    return (b + 127) & ~(size_t) 127;
  }
  __device__ __forceinline__ Smem(unsigned char *base) {
    size_t o = 0;
    red = reinterpret_cast<C *>(base);
    stg = reinterpret_cast<C *>(base);
    rowoff = reinterpret_cast<long long *>(base + sizeof(C) * 2 * CF::CT * 8);
    if (!SPREAD) o += sizeof(C) * 2 * CF::NB * CF::CT;
    else o += sizeof(C) * 2 * CF::CT * 8 + sizeof(long long) * 2 * CF::CT;
    stage = reinterpret_cast<T *>(base + o);
    o += sizeof(T) * CF::STAGES * CF::NB * CF::REC;
    full = reinterpret_cast<uint64_t *>(base + o);
    empty = full + CF::STAGES;
  }
};

struct TileRange {
  int a, b;
  long long k0, k1;
  __device__ __forceinline__ TileRange(const uint32_t *__restrict__ bin_start, const TileParams &P,
                                       long long unit, long long units) {
    if (unit >= units) { a = b = 0; k0 = k1 = 0; return; }
    const int tile = (int) (unit / P.zseg), seg = (int) (unit - (long long) tile * P.zseg);
    a = tile / P.NT1;
    b = tile - a * P.NT1;
    const int s_begin = (int) ((long long) P.NS * seg / P.zseg);
    const int s_end = (int) ((long long) P.NS * (seg + 1) / P.zseg);
    const long long bin0 = (long long) tile * P.NS;
    k0 = bin_start[bin0 + s_begin];
    k1 = bin_start[bin0 + s_end];
  }
};

// thread 0: bulk-copy the records of batch bb into its ring stage (arms the stage's full barrier)
template <typename T, int W, bool SPREAD>
__device__ __forceinline__ void tma_fill(const Smem<T, W, SPREAD> &S, const TileRange &R,
                                         const T *__restrict__ rec, int bb) {
  typedef Cfg<T, W> CF;
  const int s = bb % CF::STAGES;
  const long long kb = R.k0 + (long long) bb * CF::NB;
  const int nb = (int) min((long long) CF::NB, R.k1 - kb);
  const uint32_t bytes = (uint32_t) (nb * CF::REC * sizeof(T));
  mbar_expect_tx(&S.full[s], bytes);
  bulk_g2s(S.stage + (size_t) s * CF::NB * CF::REC, rec + kb * CF::REC, bytes, &S.full[s]);
}

// four consecutive psi2 values with vector loads (the psi2 segment of a record starts 16-byte aligned
// and is padded so that a 4-wide read never leaves it: see Cfg::WZP)
template <typename T> struct PsiLoad;
template <> struct PsiLoad<double> {
  static __device__ __forceinline__ void load4(const double *p, double (&q)[4]) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y;
  }
};
template <> struct PsiLoad<float> {
  static __device__ __forceinline__ void load4(const float *p, float (&q)[4]) {
    const float4 a = *reinterpret_cast<const float4 *>(p);
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w;
  }
};

// common prologue: work unit, barriers, first fills, row ownership
#define NFFTCU_PENCIL_PROLOGUE(SPREADV)                                                              \
  typedef Cfg<T, W> CF;                                                                              \
  typedef typename Cplx<T>::type C;                                                                  \
  extern __shared__ __align__(128) unsigned char smem_raw[];                                         \
  const int tid = threadIdx.x, lane = threadIdx.x & 31;                                              \
  const long long units = (long long) P.NT0 * P.NT1 * P.zseg;                                        \
  const TileRange R(bin_start, P, (long long) blockIdx.x, units);                                    \
  if (R.k0 == R.k1) return;                                                                          \
  const Smem<T, W, SPREADV> S(smem_raw);                                                             \
  const int nbatch = (int) ((R.k1 - R.k0 + CF::NB - 1) / CF::NB);                                    \
  if (tid == 0) {                                                                                    \
    for (int s = 0; s < CF::STAGES; s++) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], CF::NWARPS); } \
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                               \
    for (int bb = 0; bb < CF::STAGES && bb < nbatch; bb++) tma_fill<T, W, SPREADV>(S, R, rec, bb);   \
  }                                                                                                  \
  __syncthreads();                                                                                   \
  const int r0 = tid, r1 = tid + CF::CT;                                                             \
  const bool v1 = r1 < CF::ROWS;                                                                     \
  const int l0a = r0 / CF::F1, l1a = r0 - l0a * CF::F1;                                              \
  const int l0b = v1 ? r1 / CF::F1 : 0, l1b = v1 ? r1 - l0b * CF::F1 : 0;                            \
  const long long offa = ((long long) wrap_fast((long long) R.a * CF::T0 + l0a, P.n0) * P.n1 +       \
                          wrap_fast((long long) R.b * CF::T1 + l1a, P.n1)) * P.n2;                   \
  const long long offb = ((long long) wrap_fast((long long) R.a * CF::T0 + l0b, P.n0) * P.n1 +       \
                          wrap_fast((long long) R.b * CF::T1 + l1b, P.n1)) * P.n2;                   \
  const int n2 = P.n2;

// thread 0, once per batch: refill the stage that batch bb-1 occupied (all warps have released it, or
// are about to) with batch bb-1+STAGES
#define NFFTCU_REFILL(SPREADV)                                                                       \
  if (tid == 0 && bb >= 1 && bb - 1 + CF::STAGES < nbatch) {                                         \
    mbar_wait(&S.empty[(bb - 1) % CF::STAGES], ((bb - 1) / CF::STAGES) & 1);                         \
    tma_fill<T, W, SPREADV>(S, R, rec, bb - 1 + CF::STAGES);                                         \
  }

// ---- spreading ---------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(Cfg<T, W>::CT, Cfg<T, W>::MINB)
spread_tile_kernel(typename Cplx<T>::type *__restrict__ G, const T *__restrict__ rec,
                   typename Cplx<T>::type *__restrict__ ft, const uint32_t *__restrict__ bin_start,
                   TileParams P) {
  NFFTCU_PENCIL_PROLOGUE(true)
  (void) offa; (void) offb;
  T ar[2][CF::WZ], ai[2][CF::WZ];
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int kz = 0; kz < CF::WZ; kz++) { ar[j][kz] = (T) 0; ai[j][kz] = (T) 0; }
  int cur = -1;   // slab the window is aligned to: it covers z = cur*SZ .. cur*SZ+WZ-1 (mod n2)

  // Retired cells are not sent to the grid one by one (a warp's rows are 4 KB apart: 32 different
  // lines per RED instruction, and L2 resolves scattered reductions at ~60-100 G/s, which bounded
  // the kernel at ~10 ms, profiles/r01u).  They are staged per row in shared memory, 8 cells = 128 bytes
  // per row, and every warp flushes the finished 8-cell blocks of ITS OWN rows as contiguous runs:
  // 16 lanes cover one 128-byte line.  Untouched cells of a block are zero, so partial blocks
  // (start, end, jumps, wrap-around) are flushed the same way.
  const int warp = tid >> 5;
  for (int r = tid; r < 2 * CF::CT; r += CF::CT) {
    const int l0 = r / CF::F1, l1 = r - l0 * CF::F1;
    S.rowoff[r] = r < CF::ROWS ? ((long long) wrap_fast((long long) R.a * CF::T0 + l0, P.n0) * P.n1 +
                                  wrap_fast((long long) R.b * CF::T1 + l1, P.n1)) * P.n2 : -1;
#pragma unroll
    for (int q = 0; q < 8; q++) S.stg[r * 8 + q] = make_c<T>((T) 0, (T) 0);
  }
  __syncwarp();   // a warp only ever touches the staging rows it owns: rows tid and tid+CT of its lanes
  int pend = -1;  // 8-cell block (index z/8) with staged, unflushed cells
  T *const Gr = reinterpret_cast<T *>(G);
  auto flush_block = [&](int blk) {
    __syncwarp();
    const int zb8 = blk * 8;
    const int cell = (lane & 15) >> 1, comp = lane & 1;
#pragma unroll 4
    for (int it = 0; it < 32; it++) {
      const int rr = 2 * it + (lane >> 4);
      const int row = (rr < 32) ? warp * 32 + rr : CF::CT + warp * 32 + (rr - 32);
      const long long off = S.rowoff[row];
      T *slot = reinterpret_cast<T *>(S.stg + row * 8 + cell) + comp;
      const T v = *slot;
      if (off >= 0 && zb8 + cell < n2) red_add(Gr + 2 * (off + zb8 + cell) + comp, v);
      *slot = (T) 0;
    }
    __syncwarp();
  };

#define NFFTCU_RETIRE(CNT)                                                                   \
  {                                                                                          \
    const int zb = cur * CF::SZ;                                                             \
    _Pragma("unroll") for (int kz = 0; kz < (CNT); kz++) {                                   \
      const int z = wrap_z(zb + kz, n2);                                                     \
      if ((z >> 3) != pend) {                                                                \
        if (pend >= 0) flush_block(pend);                                                    \
        pend = z >> 3;                                                                       \
      }                                                                                      \
      S.stg[r0 * 8 + (z & 7)] = make_c<T>(ar[0][kz], ai[0][kz]);                             \
      S.stg[r1 * 8 + (z & 7)] = make_c<T>(ar[1][kz], ai[1][kz]);                             \
    }                                                                                        \
    _Pragma("unroll") for (int j = 0; j < 2; j++)                                            \
    _Pragma("unroll") for (int kz = 0; kz < CF::WZ; kz++) {                                  \
      if (kz + (CNT) < CF::WZ) { ar[j][kz] = ar[j][kz + (CNT)]; ai[j][kz] = ai[j][kz + (CNT)]; } \
      else { ar[j][kz] = (T) 0; ai[j][kz] = (T) 0; }                                         \
    }                                                                                        \
  }

  for (int bb = 0; bb < nbatch; bb++) {
    const int s = bb % CF::STAGES;
    const int nb = (int) min((long long) CF::NB, R.k1 - R.k0 - (long long) bb * CF::NB);
    NFFTCU_REFILL(true)
    mbar_wait(&S.full[s], (bb / CF::STAGES) & 1);
    const T *pd = S.stage + (size_t) s * CF::NB * CF::REC;
    for (int i = 0; i < nb; i++, pd += CF::REC) {
      const int sl = (int) pd[CF::PADLEN];
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
      const T fx = pd[CF::PADLEN + 1], fy = pd[CF::PADLEN + 2];
      const T w0 = pd[l0a] * pd[CF::F0 + l1a];
      const T w1 = v1 ? pd[l0b] * pd[CF::F0 + l1b] : (T) 0;
      const T a0 = w0 * fx, b0 = w0 * fy, a1 = w1 * fx, b1 = w1 * fy;
      const T *p2 = pd + CF::F0 + CF::F1;
#pragma unroll
      for (int k0 = 0; k0 < CF::WZ; k0 += 4) {   // psi2 in chunks of 4: few live registers
        T q[4];
        PsiLoad<T>::load4(p2 + k0, q);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          if (k0 + kk < CF::WZ) {
            ar[0][k0 + kk] += a0 * q[kk];
            ai[0][k0 + kk] += b0 * q[kk];
            ar[1][k0 + kk] += a1 * q[kk];
            ai[1][k0 + kk] += b1 * q[kk];
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[s]);
  }
  if (cur >= 0) {
    NFFTCU_RETIRE(CF::WZ)
    if (pend >= 0) flush_block(pend);
  }
#undef NFFTCU_RETIRE
}

// ---- interpolation -----------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(Cfg<T, W>::CT, Cfg<T, W>::MINB)
interp_tile_kernel(const typename Cplx<T>::type *__restrict__ G, const T *__restrict__ rec,
                   typename Cplx<T>::type *__restrict__ ft, const uint32_t *__restrict__ bin_start,
                   TileParams P) {
  NFFTCU_PENCIL_PROLOGUE(false)
  const C *const Ga = G + offa;
  const C *const Gb = G + offb;

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

  for (int bb = 0; bb < nbatch; bb++) {
    const int s = bb % CF::STAGES;
    const int nb = (int) min((long long) CF::NB, R.k1 - R.k0 - (long long) bb * CF::NB);
    NFFTCU_REFILL(false)
    mbar_wait(&S.full[s], (bb / CF::STAGES) & 1);
    const T *pd = S.stage + (size_t) s * CF::NB * CF::REC;
    C *const redb = S.red + (size_t) (bb & 1) * CF::NB * CF::CT;
    C *red = redb + tid;
    for (int i = 0; i < nb; i++, pd += CF::REC, red += CF::CT) {
      const int sl = (int) pd[CF::PADLEN];
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
      const T *p2 = pd + CF::F0 + CF::F1;
      // two partial sums per component: 8 independent FMA chains of WZ/2 instead of 4 of WZ
      // (DFMA latency 8.4 cycles, issue 2.1: profiles/r01u)
      T t0r[2] = {(T) 0, (T) 0}, t0i[2] = {(T) 0, (T) 0}, t1r[2] = {(T) 0, (T) 0}, t1i[2] = {(T) 0, (T) 0};
#pragma unroll
      for (int k0 = 0; k0 < CF::WZ; k0 += 4) {
        T q[4];
        PsiLoad<T>::load4(p2 + k0, q);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          if (k0 + kk < CF::WZ) {
            t0r[kk & 1] += q[kk] * wr[0][k0 + kk];
            t0i[kk & 1] += q[kk] * wi[0][k0 + kk];
            t1r[kk & 1] += q[kk] * wr[1][k0 + kk];
            t1i[kk & 1] += q[kk] * wi[1][k0 + kk];
          }
        }
      }
      *red = make_c<T>(w0 * (t0r[0] + t0r[1]) + w1 * (t1r[0] + t1r[1]),
                       w0 * (t0i[0] + t0i[1]) + w1 * (t1i[0] + t1i[1]));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[s]);          // the stage's records are no longer needed
    // CTA-wide reduction of the batch's partial sums: after the barrier every warp sums the CT partials
    // of its share of the nodes.  red is double-buffered over batches; a warp can be at most one batch
    // ahead of the slowest one because of this barrier.
    __syncthreads();
    for (int i = tid >> 5; i < nb; i += CF::NWARPS) {
      T sr = (T) 0, si = (T) 0;
#pragma unroll
      for (int q = 0; q < CF::NWARPS; q++) {
        const C v = redb[i * CF::CT + lane + 32 * q];
        sr += v.x;
        si += v.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
      }
      if (lane == 0) ft[R.k0 + (long long) bb * CF::NB + i] = make_c<T>(sr, si);
    }
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
  P.ws0 = c->wscale[0];
  P.ws1 = c->wscale[1];
  P.ws2 = c->wscale[2];
  P.window = c->window;
  return P;
}

template <typename T, int W>
size_t record_bytes(long long M) { return sizeof(T) * (size_t) Cfg<T, W>::REC * (size_t) M; }

// make sure the node records exist and are current: window part once per node set when cached
// (NFFTCU_OPT_PSI_TABLE), else on every transform; samples on every spreading call
template <typename T, int W>
int prepare_records(nfftcu_ctx *c, const void *f_dev, const TileParams &P) {
  typedef typename Cplx<T>::type C;
  const size_t bytes = record_bytes<T, W>(c->M);
  if (!c->tile_psi) {
    NFFTCU_CUDA(pool_malloc(&c->tile_psi, bytes));
    NFFTCU_CUDA(cudaMemsetAsync(c->tile_psi, 0, bytes, c->stream));
    c->tile_psi_valid = false;
  }
  const int kb = 256;
  const bool cached = c->opt_psi_table && c->tile_psi_valid;
  if (!cached) {
    long long blocks = ((c->M + kNB - 1) / kNB + 7) / 8;
    const long long cap = (long long) c->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (f_dev)
      expand_nodes_kernel<T, W, true><<<(unsigned) blocks, 256, 0, c->stream>>>(
          (T *) c->tile_psi, (const T *) c->tile_x, (const C *) f_dev, c->tile_perm,
          (const double *) c->kbpoly_dev, c->M, P);
    else
      expand_nodes_kernel<T, W, false><<<(unsigned) blocks, 256, 0, c->stream>>>(
          (T *) c->tile_psi, (const T *) c->tile_x, nullptr, c->tile_perm,
          (const double *) c->kbpoly_dev, c->M, P);
    c->tile_psi_valid = true;
    c->launches++;
  } else if (f_dev) {
    refresh_f_kernel<T, W><<<(unsigned) ((c->M + kb - 1) / kb), kb, 0, c->stream>>>(
        (T *) c->tile_psi, (const C *) f_dev, c->tile_perm, c->M);
    c->launches++;
  }
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

template <typename T, int W>
int launch_spread(nfftcu_ctx *c, const void *f_dev, const TileParams &P) {
  typedef typename Cplx<T>::type C;
  typedef Cfg<T, W> CF;
  NFFTCU_TRY((prepare_records<T, W>(c, f_dev, P)));
  const size_t smem = Smem<T, W, true>::bytes();
  NFFTCU_CUDA(cudaFuncSetAttribute(spread_tile_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const unsigned grid = (unsigned) ((long long) P.NT0 * P.NT1 * P.zseg);
  if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
  spread_tile_kernel<T, W><<<grid, CF::CT, smem, c->stream>>>(
      (C *) c->grid, (const T *) c->tile_psi, (C *) c->f_tile, c->bin_start, P);
  if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

template <typename T, int W>
int launch_interp(nfftcu_ctx *c, void *f_dev, const TileParams &P) {
  typedef typename Cplx<T>::type C;
  typedef Cfg<T, W> CF;
  NFFTCU_TRY((prepare_records<T, W>(c, nullptr, P)));
  const size_t smem = Smem<T, W, false>::bytes();
  NFFTCU_CUDA(cudaFuncSetAttribute(interp_tile_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const unsigned grid = (unsigned) ((long long) P.NT0 * P.NT1 * P.zseg);
  if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
  interp_tile_kernel<T, W><<<grid, CF::CT, smem, c->stream>>>(
      (const C *) c->grid, (const T *) c->tile_psi, (C *) c->f_tile, c->bin_start, P);
  if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
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
  set_error("pencil3d: unsupported window cut-off m=%lld", (long long) c->m);
  return NFFTCU_EINVAL;
}

}  // namespace

bool tile3d_supported(const nfftcu_ctx *c) {
  if (c->d != 3 || c->direct_only) return false;
  if (c->m < 2 || c->m > 8) return false;
  for (int t = 0; t < 3; t++)
    if (c->n[t] > 0x3fffff) return false;   // slab index is carried as a float in fp32 records
  return true;
}

// tile-binned processing order: keys, stable sort, node gather, bin offsets
int tile3d_bin_nodes(nfftcu_ctx *c) {
  const long long M = c->M;
  c->tile_ready = false;
  c->tile_psi_valid = false;
  if (M == 0) return NFFTCU_OK;
  const TileParams P = make_params(c);
  const long long nbins = (long long) P.NT0 * P.NT1 * P.NS;
  if (!c->tile_keys) NFFTCU_CUDA(pool_malloc(&c->tile_keys, sizeof(uint64_t) * (size_t) M));
  if (!c->tile_perm) NFFTCU_CUDA(pool_malloc((void **) &c->tile_perm, sizeof(uint32_t) * (size_t) M));
  if (!c->tile_x) NFFTCU_CUDA(pool_malloc(&c->tile_x, real_size(c) * (size_t) M * 3));
  if (!c->f_tile) NFFTCU_CUDA(pool_malloc(&c->f_tile, 2 * real_size(c) * (size_t) M));
  if (!c->bin_start || c->tile_nbins != nbins) {
    if (c->bin_start) pool_free(c->bin_start);
    NFFTCU_CUDA(pool_malloc((void **) &c->bin_start, sizeof(uint32_t) * (size_t) (nbins + 1)));
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
