// tile2d.cu -- B and B^T for d = 2: grid tiles staged in shared memory.
//
// Replaces nfft_trafo_2d_B / nfft_adjoint_2d_B and their compute loops (kernel/nfft/nfft.c:3221-3410, 2927-3004;
// 3583-3805, 3010-3216) for 2-D plans.  The nodes are binned by the 32 x 32-cell tile of their first tap
// (u0 / 32, u1 / 32), and a tile's node list is cut into chunks of at most 128 nodes, one CTA each (MRI trajectories
// put thousands of nodes into the tiles around the origin and a handful into the outer ones).  A CTA
//   * stages the tile's footprint -- (32 + 2m + 1)^2 grid cells, wrapped -- in shared memory as double complex
//     (interpolation: coalesced row loads; spreading: zeroed, flushed at the end with coalesced RED.ADD rows, zero
//     cells skipped),
//   * evaluates the 2 (2m+2) window values of its nodes once into shared memory (piecewise polynomials of
//     kbpoly.cu, or sinh / sqrt when no polynomial fit exists for this m), in double for both precisions,
//   * interpolation: one half-warp per node, lane l owns column u1 + l, runs down the 2m+2 rows with one 16-byte
//     shared load and two FMAs per tap, then a 16-lane shuffle reduction;
//   * spreading: race-free without atomics inside the tile: thread (h, l) of the 16 x 16 CTA owns the footprint cells
//     with row = h, column = l (mod 16); a node's 2m+2 <= 16 consecutive rows / columns meet every residue at most
//     once, so per node every thread adds at most one cell with a plain read-modify-write and no two threads share a
//     cell; all threads walk all nodes of the chunk.
// Both kernels are bound by shared-memory wavefronts (one 16-byte access per tap), not by HBM: the footprint is
// read / flushed once per chunk.
#include "common.cuh"

namespace nfftcu {

namespace {

constexpr int kT2 = 32;            // tile edge in cells
constexpr int kChunkNodes = 128;   // nodes per round: their window values sit in shared memory
constexpr int kCtaNodes = 1024;    // nodes per CTA (8 rounds): one footprint load / flush per CTA
constexpr int kThreads2 = 256;
constexpr int kMaxW2 = 16;         // 2m+2 <= 16 (half-warp per node)

struct Tile2Params {
  int n0, n1, m, W, F, pitch;      // F = kT2 + W - 1 footprint edge, pitch = F + 1 (complex elements)
  int NT0, NT1;
  int use_poly, deg;
  double b0, b1, m2;
  double ws0, ws1;   // power-of-two window scale per dimension
  int window;
};

__device__ __forceinline__ int wrap2(int v, int n) {
  v %= n;
  return v < 0 ? v + n : v;
}

template <typename TS>
__global__ void tile2_keys_kernel(const TS *__restrict__ x, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                  long long M, Tile2Params P) {
  const long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int u0 = wrap2((int) (cell_of(x[2 * j], (long long) P.n0) - P.m), P.n0);
  const int u1 = wrap2((int) (cell_of(x[2 * j + 1], (long long) P.n1) - P.m), P.n1);
  keys[j] = (uint64_t) (u0 / kT2) * P.NT1 + (u1 / kT2);
  vals[j] = (uint32_t) j;
}

// tile_start[t] = first position whose key >= t, t = 0..tiles
__global__ void tile2_bounds_kernel(const uint64_t *__restrict__ keys, uint32_t *__restrict__ tile_start, long long tiles,
                                    long long M) {
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t > tiles) return;
  long long lo = 0, hi = M;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < (uint64_t) t) lo = mid + 1;
    else hi = mid;
  }
  tile_start[t] = (uint32_t) lo;
}

__global__ void tile2_chunk_count_kernel(const uint32_t *__restrict__ tile_start, uint32_t *__restrict__ counts, long long tiles) {
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= tiles) return;
  counts[t] = (tile_start[t + 1] - tile_start[t] + kCtaNodes - 1) / kCtaNodes;
}

// exclusive scan, single CTA (tile counts are small: n_total / 1024 entries)
__global__ void tile2_scan_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ out, long long n) {
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  const long long chunk = (n + 1023) / 1024;
  const long long lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
  uint32_t s = 0;
  for (long long i = lo; i < hi; i++) s += counts[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = t > 0 ? part[t - 1] : 0;
  for (long long i = lo; i < hi; i++) { out[i] = run; run += counts[i]; }
  if (t == 1023) out[n] = part[1023];
}

__global__ void tile2_chunk_fill_kernel(const uint32_t *__restrict__ tile_start, const uint32_t *__restrict__ chunk_start,
                                        uint4 *__restrict__ chunks, long long tiles) {
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= tiles) return;
  const uint32_t k0 = tile_start[t], k1 = tile_start[t + 1];
  uint32_t c = chunk_start[t];
  for (uint32_t k = k0; k < k1; k += kCtaNodes, c++)
    chunks[c] = make_uint4((uint32_t) t, k, k + kCtaNodes < k1 ? k + kCtaNodes : k1, 0u);
}

template <typename C2>
__global__ void tile2_gather_f_kernel(const C2 *__restrict__ f, const uint32_t *__restrict__ perm, C2 *__restrict__ ft,
                                      long long M) {
  const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const size_t off = (size_t) blockIdx.y * (size_t) M;   // right-hand side of a batched transform
  if (k < M) ft[off + k] = f[off + perm[k]];
}

// shared memory: tile [F][pitch] double2 | psi0 [kChunkNodes][kMaxW2] double | psi1 (interp: double, spread: double2)
// | node origin (r0, c0) relative to the footprint
struct NodeOrg { short r0, c0; };

// Node data of a round (coordinates, and the samples for spreading), staged in shared memory ONE ROUND AHEAD: the
// global loads of round r + 1 are issued before the node loop of round r and stored after it, so that the window
// evaluation never waits for L2 (the per-value loads it replaces were the top stall of both kernels:
// profiles/r2_full_new_kernels.md, 30 % of the samples on the first use of x / f).
template <typename TS>
struct NodeStage {
  TS x[kChunkNodes * 2];
  typename Cplx<TS>::type f[kChunkNodes];
};
// thread t carries coordinate t (t < 2 cnt) and sample t (t < cnt) of the next round in registers
template <typename TS, bool SPREAD>
__device__ __forceinline__ void node_load(const TS *__restrict__ xt, const typename Cplx<TS>::type *__restrict__ ft,
                                          int k0, int cnt, TS &rx, TS &rfx, TS &rfy) {
  const int t = threadIdx.x;
  rx = (t < 2 * cnt) ? xt[2 * (size_t) k0 + t] : (TS) 0;
  rfx = rfy = (TS) 0;
  if (SPREAD && t < cnt) {
    const typename Cplx<TS>::type v = ft[(size_t) k0 + t];
    rfx = v.x;
    rfy = v.y;
  }
}
template <typename TS, bool SPREAD>
__device__ __forceinline__ void node_store(NodeStage<TS> &st, int cnt, TS rx, TS rfx, TS rfy) {
  const int t = threadIdx.x;
  if (t < 2 * cnt) st.x[t] = rx;
  if (SPREAD && t < cnt) {
    st.f[t].x = rfx;
    st.f[t].y = rfy;
  }
}

// window values of the chunk's nodes: thread per (node, dim, tap); node data from the staged copy
template <typename TS, bool SPREAD>
__device__ __forceinline__ void chunk_windows(const NodeStage<TS> &ns, const Tile2Params &P, int cnt,
                                              int ta, int tb, double *psi0, double *psi1, NodeOrg *org,
                                              const double *coef) {
  for (int i = threadIdx.x; i < cnt * 2 * kMaxW2; i += blockDim.x) {
    const int l = i % kMaxW2, t = (i / kMaxW2) & 1, j = i / (2 * kMaxW2);
    const TS x = ns.x[2 * j + t];
    const int n = t == 0 ? P.n0 : P.n1;
    const long long c = cell_of(x, (long long) n);
    double v = 0.0;
    if (l < P.W) {
      if (P.use_poly) {
        const double y = 2.0 * ((double) x * (double) n - (double) c) - 1.0;
        const double *cf = coef + (size_t) t * (kKbPolyDeg + 1) * P.W + l;
        v = cf[(size_t) P.deg * P.W];
        for (int k = P.deg - 1; k >= 0; k--) v = fma(v, y, cf[(size_t) k * P.W]);
      } else {
        const double dist = (double) x * (double) n - (double) (c - P.m + l);
        v = window_phi(dist, P.m2, t == 0 ? P.b0 : P.b1, P.window, t == 0 ? P.ws0 : P.ws1);
      }
    }
    if (t == 0) psi0[j * kMaxW2 + l] = v;
    else if (!SPREAD) psi1[j * kMaxW2 + l] = v;
    else {
      const typename Cplx<TS>::type fv = ns.f[j];
      psi1[2 * (j * kMaxW2 + l)] = v * (double) fv.x;
      psi1[2 * (j * kMaxW2 + l) + 1] = v * (double) fv.y;
    }
    if (l == 0) {
      const int u = wrap2((int) (c - P.m), n);
      if (t == 0) org[j].r0 = (short) (u - kT2 * ta);
      else org[j].c0 = (short) (u - kT2 * tb);
    }
  }
}

template <typename TS>
__global__ void __launch_bounds__(kThreads2)
interp_tile2_kernel(const typename Cplx<TS>::type *__restrict__ G, const TS *__restrict__ xt,
                    const uint32_t *__restrict__ perm, typename Cplx<TS>::type *__restrict__ f,
                    const uint4 *__restrict__ chunks, const double *__restrict__ poly, Tile2Params P,
                    long long gstride, long long fstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  G += (size_t) blockIdx.y * gstride;   // right-hand side blockIdx.y of a batched transform
  f += (size_t) blockIdx.y * fstride;
  double2 *tile = reinterpret_cast<double2 *>(smem_raw);
  double *psi0 = reinterpret_cast<double *>(tile + (size_t) P.F * P.pitch);
  double *psi1 = psi0 + kChunkNodes * kMaxW2;
  NodeOrg *org = reinterpret_cast<NodeOrg *>(psi1 + kChunkNodes * kMaxW2);
  double *coef = reinterpret_cast<double *>(org + kChunkNodes);
  NodeStage<TS> &ns = *reinterpret_cast<NodeStage<TS> *>(coef + 2 * (kKbPolyDeg + 1) * kMaxW2);
  const uint4 ch = chunks[blockIdx.x];
  const int ta = (int) ch.x / P.NT1, tb = (int) ch.x - ta * P.NT1;
  const int kbeg = (int) ch.y, kend = (int) ch.z;
  TS nx, nfx, nfy;   // next round's node data of this thread
  node_load<TS, false>(xt, nullptr, kbeg, min(kChunkNodes, kend - kbeg), nx, nfx, nfy);
  if (P.use_poly)   // polynomial coefficients of both dimensions, once per CTA
    for (int i = threadIdx.x; i < 2 * (kKbPolyDeg + 1) * P.W; i += blockDim.x) coef[i] = poly[i];

  for (int i = threadIdx.x; i < P.F * P.F; i += blockDim.x) {
    const int r = i / P.F, cc = i - r * P.F;
    const typename Cplx<TS>::type v = G[(size_t) wrap2(kT2 * ta + r, P.n0) * P.n1 + wrap2(kT2 * tb + cc, P.n1)];
    tile[r * P.pitch + cc] = make_double2((double) v.x, (double) v.y);
  }
  const int hw = threadIdx.x >> 4, l = threadIdx.x & 15;
  for (int k0 = kbeg; k0 < kend; k0 += kChunkNodes) {
    const int cnt = min(kChunkNodes, kend - k0);
    node_store<TS, false>(ns, cnt, nx, nfx, nfy);
    __syncthreads();   // staged nodes (and, in the first round, the footprint and the coefficients) are visible
    if (k0 + kChunkNodes < kend)   // next round's nodes: in flight during this round
      node_load<TS, false>(xt, nullptr, k0 + kChunkNodes, min(kChunkNodes, kend - k0 - kChunkNodes), nx, nfx, nfy);
    chunk_windows<TS, false>(ns, P, cnt, ta, tb, psi0, psi1, org, coef);
    __syncthreads();
    for (int j0 = 0; j0 < cnt; j0 += kThreads2 / 16) {   // warp-uniform trip count: the shuffles below need all 32 lanes
      const int j = j0 + hw;
      const bool live = j < cnt;
      const NodeOrg o = org[live ? j : 0];
      double ar = 0.0, ai = 0.0;
      if (live && l < P.W) {
        const double2 *col = tile + (size_t) o.r0 * P.pitch + o.c0 + l;
        const double *p0 = psi0 + j * kMaxW2;
#pragma unroll 2
        for (int l0 = 0; l0 < P.W; l0++) {
          const double2 v = col[(size_t) l0 * P.pitch];
          const double w = p0[l0];
          ar = fma(w, v.x, ar);
          ai = fma(w, v.y, ai);
        }
        const double w1 = psi1[j * kMaxW2 + l];
        ar *= w1;
        ai *= w1;
      }
#pragma unroll
      for (int s = 8; s > 0; s >>= 1) {
        ar += __shfl_xor_sync(0xffffffffu, ar, s);
        ai += __shfl_xor_sync(0xffffffffu, ai, s);
      }
      if (live && l == 0) {
        typename Cplx<TS>::type out;
        out.x = (TS) ar;
        out.y = (TS) ai;
        f[perm[k0 + j]] = out;
      }
    }
    __syncthreads();   // the next round overwrites the window values
  }
}

__device__ __forceinline__ void red_add2(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add2(float *p, double v) { atomicAdd(p, (float) v); }

template <typename TS>
__global__ void __launch_bounds__(kThreads2)
spread_tile2_kernel(typename Cplx<TS>::type *__restrict__ G, const TS *__restrict__ xt,
                    const typename Cplx<TS>::type *__restrict__ ft, const uint4 *__restrict__ chunks,
                    const double *__restrict__ poly, Tile2Params P, long long gstride, long long fstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  G += (size_t) blockIdx.y * gstride;
  ft += (size_t) blockIdx.y * fstride;
  double2 *tile = reinterpret_cast<double2 *>(smem_raw);
  double *psi0 = reinterpret_cast<double *>(tile + (size_t) P.F * P.pitch);
  double *psi1 = psi0 + kChunkNodes * kMaxW2;                       // (psi1 * f) complex
  NodeOrg *org = reinterpret_cast<NodeOrg *>(psi1 + 2 * kChunkNodes * kMaxW2);
  double *coef = reinterpret_cast<double *>(org + kChunkNodes);
  NodeStage<TS> &ns = *reinterpret_cast<NodeStage<TS> *>(coef + 2 * (kKbPolyDeg + 1) * kMaxW2);
  const uint4 ch = chunks[blockIdx.x];
  const int ta = (int) ch.x / P.NT1, tb = (int) ch.x - ta * P.NT1;
  const int kbeg = (int) ch.y, kend = (int) ch.z;
  TS nx, nfx, nfy;   // next round's node data of this thread
  node_load<TS, true>(xt, ft, kbeg, min(kChunkNodes, kend - kbeg), nx, nfx, nfy);
  if (P.use_poly)
    for (int i = threadIdx.x; i < 2 * (kKbPolyDeg + 1) * P.W; i += blockDim.x) coef[i] = poly[i];

  for (int i = threadIdx.x; i < P.F * P.pitch; i += blockDim.x) tile[i] = make_double2(0.0, 0.0);
  // thread (h, l) owns the footprint cells with row = h and column = l (mod 16): a node's <= 16 consecutive rows and
  // columns meet every residue at most once, so per node a thread updates at most one cell, no two threads ever touch
  // the same cell, and a thread's own read-modify-writes are ordered by the hardware -- no atomics, no barriers
  // inside the node loop.
  const int h = threadIdx.x >> 4, l = threadIdx.x & 15;
  const double2 *pf = reinterpret_cast<const double2 *>(psi1);
  for (int k0 = kbeg; k0 < kend; k0 += kChunkNodes) {
    const int cnt = min(kChunkNodes, kend - k0);
    node_store<TS, true>(ns, cnt, nx, nfx, nfy);
    __syncthreads();
    if (k0 + kChunkNodes < kend)
      node_load<TS, true>(xt, ft, k0 + kChunkNodes, min(kChunkNodes, kend - k0 - kChunkNodes), nx, nfx, nfy);
    chunk_windows<TS, true>(ns, P, cnt, ta, tb, psi0, psi1, org, coef);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; j++) {
      const NodeOrg o = org[j];
      const int l0 = (h - o.r0) & 15, l1 = (l - o.c0) & 15;
      if (l0 < P.W && l1 < P.W) {
        double2 *cell = tile + (size_t) (o.r0 + l0) * P.pitch + o.c0 + l1;
        const double w = psi0[j * kMaxW2 + l0];
        const double2 a = pf[j * kMaxW2 + l1];
        double2 v = *cell;
        v.x = fma(w, a.x, v.x);
        v.y = fma(w, a.y, v.y);
        *cell = v;
      }
    }
    __syncthreads();   // the next round overwrites the window values; the flush reads the tile
  }

  TS *Gs = reinterpret_cast<TS *>(G);
  for (int i = threadIdx.x; i < P.F * P.F; i += blockDim.x) {
    const int r = i / P.F, cc = i - r * P.F;
    const double2 v = tile[r * P.pitch + cc];
    if (v.x == 0.0 && v.y == 0.0) continue;
    TS *dst = Gs + 2 * ((size_t) wrap2(kT2 * ta + r, P.n0) * P.n1 + wrap2(kT2 * tb + cc, P.n1));
    red_add2(dst, v.x);
    red_add2(dst + 1, v.y);
  }
}

Tile2Params make_params2(const nfftcu_ctx *c) {
  Tile2Params P;
  P.n0 = (int) c->n[0];
  P.n1 = (int) c->n[1];
  P.m = (int) c->m;
  P.W = 2 * P.m + 2;
  P.F = kT2 + P.W - 1;
  P.pitch = P.F + 1;
  P.NT0 = (P.n0 + kT2 - 1) / kT2;
  P.NT1 = (P.n1 + kT2 - 1) / kT2;
  P.use_poly = c->kbpoly_fit >= 0 && c->kbpoly_dev != nullptr;
  P.deg = c->kbpoly_fit;
  P.b0 = c->b[0];
  P.b1 = c->b[1];
  P.ws0 = c->wscale[0];
  P.ws1 = c->wscale[1];
  P.window = c->window;
  P.m2 = (double) c->m * (double) c->m;
  return P;
}

size_t smem2(const Tile2Params &P, bool spread) {
  return sizeof(double2) * (size_t) P.F * P.pitch + sizeof(double) * kChunkNodes * kMaxW2 * (spread ? 3 : 2) +
         sizeof(NodeOrg) * kChunkNodes + sizeof(double) * 2 * (kKbPolyDeg + 1) * kMaxW2 +
         sizeof(double) * 2 * kChunkNodes + sizeof(double2) * kChunkNodes + 16;   // NodeStage (sized for double)
}

template <typename TS>
int run2(nfftcu_ctx *c, const void *f_in, void *f_out, bool spread) {
  typedef typename Cplx<TS>::type C2;
  const Tile2Params P = make_params2(c);
  const dim3 grid((unsigned) c->mma_nchunks, (unsigned) c->cur_batch);
  if (grid.x == 0) return NFFTCU_OK;
  const size_t smem = smem2(P, spread);
  const uint4 *chunks = (const uint4 *) c->mma_chunks;
  const double *poly = (const double *) c->kbpoly_dev;
  if (!spread) {
    NFFTCU_CUDA(cudaFuncSetAttribute(interp_tile2_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
    interp_tile2_kernel<TS><<<grid, kThreads2, smem, c->stream>>>((const C2 *) c->grid, (const TS *) c->tile_x, c->tile_perm,
                                                                 (C2 *) f_out, chunks, poly, P, c->n_total, c->M);
    if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
    c->launches++;
  } else {
    const int kb = 256;
    tile2_gather_f_kernel<C2><<<dim3((unsigned) ((c->M + kb - 1) / kb), grid.y), kb, 0, c->stream>>>(
        (const C2 *) f_in, c->tile_perm, (C2 *) c->f_tile, c->M);
    NFFTCU_CUDA(cudaFuncSetAttribute(spread_tile2_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
    spread_tile2_kernel<TS><<<grid, kThreads2, smem, c->stream>>>((C2 *) c->grid, (const TS *) c->tile_x,
                                                                 (const C2 *) c->f_tile, chunks, poly, P, c->n_total, c->M);
    if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
    c->launches += 2;
  }
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

}  // namespace

bool tile2d_supported(const nfftcu_ctx *c) {
  if (c->d != 2 || c->direct_only) return false;
  if (2 * c->m + 2 > kMaxW2) return false;
  for (int t = 0; t < 2; t++)
    if (c->n[t] < 2 * c->m + 2 || c->n[t] > 0x3fffffff) return false;
  if (c->M >= (1ll << 31)) return false;
  return true;
}

// plan-time binning: tile order (stable within a tile), node copies in that order, chunk list
int tile2d_bin_nodes(nfftcu_ctx *c) {
  const long long M = c->M;
  c->tile2_ready = false;
  if (M == 0) return NFFTCU_OK;
  const Tile2Params P = make_params2(c);
  const long long tiles = (long long) P.NT0 * P.NT1;
  if (!c->tile_keys) NFFTCU_CUDA(pool_malloc(&c->tile_keys, sizeof(uint64_t) * (size_t) M));
  if (!c->tile_perm) NFFTCU_CUDA(pool_malloc((void **) &c->tile_perm, sizeof(uint32_t) * (size_t) M));
  if (!c->tile_x) NFFTCU_CUDA(pool_malloc(&c->tile_x, real_size(c) * (size_t) M * 2));
  if (!c->f_tile) NFFTCU_CUDA(pool_malloc(&c->f_tile, 2 * real_size(c) * (size_t) M * (size_t) c->batch_cap));
  if (!c->bin_start || c->tile_nbins != tiles) {
    if (c->bin_start) pool_free(c->bin_start);
    if (c->mma_counts) pool_free(c->mma_counts);
    if (c->mma_chunk_start) pool_free(c->mma_chunk_start);
    c->bin_start = c->mma_counts = c->mma_chunk_start = nullptr;
    NFFTCU_CUDA(pool_malloc((void **) &c->bin_start, sizeof(uint32_t) * (size_t) (tiles + 1)));
    NFFTCU_CUDA(pool_malloc((void **) &c->mma_counts, sizeof(uint32_t) * (size_t) tiles));
    NFFTCU_CUDA(pool_malloc((void **) &c->mma_chunk_start, sizeof(uint32_t) * (size_t) (tiles + 1)));
    c->tile_nbins = tiles;
  }
  const int kb = 256;
  const unsigned ngrid = (unsigned) ((M + kb - 1) / kb);
  if (c->prec == NFFTCU_DOUBLE)
    tile2_keys_kernel<double><<<ngrid, kb, 0, c->stream>>>((const double *) c->x_dev, (uint64_t *) c->tile_keys, c->tile_perm, M, P);
  else
    tile2_keys_kernel<float><<<ngrid, kb, 0, c->stream>>>((const float *) c->x_dev, (uint64_t *) c->tile_keys, c->tile_perm, M, P);
  c->launches++;
  int bits = 0;
  while ((1ll << bits) < tiles && bits < 62) bits++;
  NFFTCU_TRY(radix_sort_pairs(c, (uint64_t *) c->tile_keys, c->tile_perm, M, bits));
  NFFTCU_TRY(gather_nodes(c, c->tile_perm, c->tile_x));
  const unsigned tgrid = (unsigned) ((tiles + 1 + kb - 1) / kb);
  tile2_bounds_kernel<<<tgrid, kb, 0, c->stream>>>((const uint64_t *) c->tile_keys, c->bin_start, tiles, M);
  tile2_chunk_count_kernel<<<tgrid, kb, 0, c->stream>>>(c->bin_start, c->mma_counts, tiles);
  tile2_scan_kernel<<<1, 1024, 0, c->stream>>>(c->mma_counts, c->mma_chunk_start, tiles);
  uint32_t nchunks = 0;
  NFFTCU_CUDA(cudaMemcpyAsync(&nchunks, c->mma_chunk_start + tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  if ((long long) nchunks > c->mma_chunk_cap) {
    if (c->mma_chunks) pool_free(c->mma_chunks);
    c->mma_chunks = nullptr;
    c->mma_chunk_cap = (long long) nchunks + nchunks / 8 + 1024;
    NFFTCU_CUDA(pool_malloc(&c->mma_chunks, sizeof(uint4) * (size_t) c->mma_chunk_cap));
  }
  tile2_chunk_fill_kernel<<<tgrid, kb, 0, c->stream>>>(c->bin_start, c->mma_chunk_start, (uint4 *) c->mma_chunks, tiles);
  c->mma_nchunks = nchunks;
  c->launches += 4;
  NFFTCU_CUDA(cudaGetLastError());
  c->tile2_ready = true;
  return NFFTCU_OK;
}

int tile2d_interp(nfftcu_ctx *c, void *f_dev) {
  return c->prec == NFFTCU_DOUBLE ? run2<double>(c, nullptr, f_dev, false) : run2<float>(c, nullptr, f_dev, false);
}

int tile2d_spread(nfftcu_ctx *c, const void *f_dev) {
  return c->prec == NFFTCU_DOUBLE ? run2<double>(c, f_dev, nullptr, true) : run2<float>(c, f_dev, nullptr, true);
}

}  // namespace nfftcu
