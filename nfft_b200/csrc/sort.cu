// sort.cu -- node keys and the stable LSD radix sort of (key, node index) pairs.
//
// Replaces sort0 (kernel/nfft/nfft.c:75-109 of the reference) and
// nfft_sort_node_indices_radix_lsdf (kernel/util/sort.c:91-167).  Contract to reproduce
// bit-exactly: index_x = stable ascending sort of
//     key_j = row-major linearisation of ((floor(n_t*x_jt - m) mod n_t) + n_t) mod n_t
// with ties kept in ascending j.  An LSD radix sort with a stable scatter per digit is exactly
// that, independent of the digit width (the reference uses 9-bit digits, we use 8).
//
// Pass structure (all on the plan's stream, no host round trip):
//   radix_hist    per-tile digit histogram             -> counts[digit][tile]
//   radix_scan    exclusive scan over counts (digit-major), one CTA
//   radix_scatter per-tile stable ranking (warp match + per-warp running counters) and scatter
// Tiles are 4096 pairs (8 warps x 512 consecutive pairs) so that every warp walks a contiguous
// run in order, which is what makes the scatter stable.
#include "common.cuh"

namespace nfftcu {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kWarpChunk = 512;
constexpr int kTile = kSortWarps * kWarpChunk;

struct KeyGeom {
  long long n[NFFTCU_MAX_D];
  int d;
  long long m;
};

// floor(n*x - m): one rounded multiply, one rounded subtract, in the plan's precision -- the
// reference's expression (nfft.c:88) without FMA contraction.
__device__ __forceinline__ long long key_floor(double x, long long n, long long m) {
  return (long long) floor(__dsub_rn(__dmul_rn((double) n, x), (double) m));
}
__device__ __forceinline__ long long key_floor(float x, long long n, long long m) {
  return (long long) floorf(__fsub_rn(__fmul_rn((float) n, x), (float) m));
}

template <typename T>
__global__ void make_keys_kernel(const T *__restrict__ x, uint64_t *__restrict__ keys,
                                 uint32_t *__restrict__ vals, long long M, KeyGeom g) {
  const long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  unsigned long long key = 0;
  for (int t = 0; t < g.d; t++) {
    const long long help = key_floor(x[j * g.d + t], g.n[t], g.m);
    const long long u = (help % g.n[t] + g.n[t]) % g.n[t];
    key += (unsigned long long) u;
    if (t + 1 < g.d) key *= (unsigned long long) g.n[t + 1];
  }
  keys[j] = key;
  vals[j] = (uint32_t) j;
}

__global__ void radix_hist_kernel(const uint64_t *__restrict__ keys, uint32_t *__restrict__ counts,
                                  long long M, int shift, int ntiles) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const long long beg = (long long) blockIdx.x * kTile;
  const long long end = min(beg + (long long) kTile, M);
  for (long long i = beg + threadIdx.x; i < end; i += kSortThreads)
    atomicAdd(&hist[(unsigned) (keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  counts[(size_t) threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

// exclusive scan of `total` counters in place, one CTA of 1024 threads
__global__ void radix_scan_kernel(uint32_t *__restrict__ counts, long long total) {
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  const long long chunk = (total + 1023) / 1024;
  const long long beg = min((long long) t * chunk, total), end = min(beg + chunk, total);
  uint32_t s = 0;
  for (long long i = beg; i < end; i++) s += counts[i];
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    uint32_t v = (t >= off) ? part[t - off] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = part[t] - s;
  for (long long i = beg; i < end; i++) {
    const uint32_t c = counts[i];
    counts[i] = run;
    run += c;
  }
}

__global__ void radix_scatter_kernel(const uint64_t *__restrict__ keys_in,
                                     const uint32_t *__restrict__ vals_in,
                                     uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                     const uint32_t *__restrict__ offsets, long long M, int shift,
                                     int ntiles) {
  __shared__ uint32_t wc[kSortWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wc[0][0])[i] = 0;
  __syncthreads();
  const long long wbeg = min((long long) blockIdx.x * kTile + (long long) warp * kWarpChunk, M);
  const long long wend = min(wbeg + (long long) kWarpChunk, M);

  // phase A: per-warp digit counts
  for (long long i = wbeg; i < wend; i += 32) {
    const long long idx = i + lane;
    const bool valid = idx < wend;
    const unsigned digit = valid ? ((unsigned) (keys_in[idx] >> shift) & 255u) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    if (valid && (int) (__ffs(peers) - 1) == lane) wc[warp][digit] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // phase B: turn counts into start positions: global digit offset of this tile, then warps in order
  {
    const int digit = threadIdx.x;
    uint32_t run = offsets[(size_t) digit * ntiles + blockIdx.x];
    for (int w = 0; w < kSortWarps; w++) {
      const uint32_t c = wc[w][digit];
      wc[w][digit] = run;
      run += c;
    }
  }
  __syncthreads();
  // phase C: stable scatter
  for (long long i = wbeg; i < wend; i += 32) {
    const long long idx = i + lane;
    const bool valid = idx < wend;
    uint64_t key = 0;
    unsigned digit = 0xffffffffu;
    if (valid) {
      key = keys_in[idx];
      digit = (unsigned) (key >> shift) & 255u;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    uint32_t pos = 0;
    if (valid) pos = wc[warp][digit] + __popc(peers & lt_mask);
    __syncwarp();
    if (valid) {
      keys_out[pos] = key;
      vals_out[pos] = vals_in[idx];
      if ((int) (__ffs(peers) - 1) == lane) wc[warp][digit] += __popc(peers);
    }
    __syncwarp();
  }
}

template <typename T>
__global__ void gather_nodes_kernel(const T *__restrict__ x, const uint32_t *__restrict__ perm,
                                    T *__restrict__ xs, long long M, int d) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * d) return;
  const long long k = i / d;
  const int t = (int) (i % d);
  xs[i] = x[(long long) perm[k] * d + t];
}

int ilog2_ceil(unsigned long long v) {
  int b = 0;
  while ((1ull << b) < v && b < 63) b++;
  return b;
}

}  // namespace

// Stable LSD radix sort of (keys, vals) pairs over the low `bits` key bits; the result is left in
// the arrays passed in.  Scratch (second pair of buffers + counters) lives in c->sort_tmp.
int radix_sort_pairs(nfftcu_ctx *c, uint64_t *keys, uint32_t *vals, long long M, int bits) {
  if (M == 0) return NFFTCU_OK;
  const int ntiles = (int) ((M + kTile - 1) / kTile);
  const size_t kbytes = sizeof(uint64_t) * (size_t) M, vbytes = sizeof(uint32_t) * (size_t) M;
  const size_t vpad = ((vbytes + 255) / 256) * 256;
  const size_t cbytes = sizeof(uint32_t) * 256 * (size_t) ntiles;
  const size_t need = kbytes + vpad + cbytes;
  if (c->sort_tmp_bytes < need) {
    if (c->sort_tmp) pool_free(c->sort_tmp);
    c->sort_tmp = nullptr;
    c->sort_tmp_bytes = 0;
    NFFTCU_CUDA(pool_malloc(&c->sort_tmp, need));
    c->sort_tmp_bytes = need;
  }
  uint64_t *kA = keys, *kB = (uint64_t *) c->sort_tmp;
  uint32_t *vA = vals, *vB = (uint32_t *) ((char *) c->sort_tmp + kbytes);
  uint32_t *counts = (uint32_t *) ((char *) c->sort_tmp + kbytes + vpad);
  const int passes = bits <= 0 ? 1 : (bits + 7) / 8;
  for (int p = 0; p < passes; p++) {
    const int shift = 8 * p;
    radix_hist_kernel<<<ntiles, kSortThreads, 0, c->stream>>>(kA, counts, M, shift, ntiles);
    radix_scan_kernel<<<1, 1024, 0, c->stream>>>(counts, 256ll * ntiles);
    radix_scatter_kernel<<<ntiles, kSortThreads, 0, c->stream>>>(kA, vA, kB, vB, counts, M, shift,
                                                               ntiles);
    c->launches += 3;
    uint64_t *tk = kA; kA = kB; kB = tk;
    uint32_t *tv = vA; vA = vB; vB = tv;
  }
  if (kA != keys) {   // odd number of passes: result sits in the scratch
    NFFTCU_CUDA(cudaMemcpyAsync(keys, kA, kbytes, cudaMemcpyDeviceToDevice, c->stream));
    NFFTCU_CUDA(cudaMemcpyAsync(vals, vA, vbytes, cudaMemcpyDeviceToDevice, c->stream));
  }
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

// dst[k*d+t] = c->x_dev[perm[k]*d+t]
int gather_nodes(nfftcu_ctx *c, const uint32_t *perm, void *dst) {
  if (c->M == 0) return NFFTCU_OK;
  const int kb = 256;
  const unsigned ggrid = (unsigned) ((c->M * c->d + kb - 1) / kb);
  if (c->prec == NFFTCU_DOUBLE)
    gather_nodes_kernel<double><<<ggrid, kb, 0, c->stream>>>((const double *) c->x_dev, perm,
                                                            (double *) dst, c->M, c->d);
  else
    gather_nodes_kernel<float><<<ggrid, kb, 0, c->stream>>>((const float *) c->x_dev, perm,
                                                           (float *) dst, c->M, c->d);
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

// Reference order: keys from c->x_dev, stable sort, leaves c->keys_ref (sorted keys),
// c->perm (= c->perm_ref) and c->x_sorted.
int sort_nodes(nfftcu_ctx *c) {
  const long long M = c->M;
  if (M == 0) return NFFTCU_OK;
  const size_t kbytes = sizeof(uint64_t) * (size_t) M, vbytes = sizeof(uint32_t) * (size_t) M;
  if (!c->keys_ref) NFFTCU_CUDA(pool_malloc(&c->keys_ref, kbytes));
  if (!c->perm) NFFTCU_CUDA(pool_malloc((void **) &c->perm, vbytes));
  if (!c->x_sorted) NFFTCU_CUDA(pool_malloc(&c->x_sorted, real_size(c) * (size_t) M * c->d));
  c->perm_ref = c->perm;
  KeyGeom g;
  g.d = c->d;
  g.m = c->m;
  for (int t = 0; t < c->d; t++) g.n[t] = c->n[t];
  const int kb = 256;
  const unsigned kgrid = (unsigned) ((M + kb - 1) / kb);
  if (c->prec == NFFTCU_DOUBLE)
    make_keys_kernel<double><<<kgrid, kb, 0, c->stream>>>((const double *) c->x_dev,
                                                         (uint64_t *) c->keys_ref, c->perm, M, g);
  else
    make_keys_kernel<float><<<kgrid, kb, 0, c->stream>>>((const float *) c->x_dev,
                                                        (uint64_t *) c->keys_ref, c->perm, M, g);
  c->launches++;
  NFFTCU_TRY(radix_sort_pairs(c, (uint64_t *) c->keys_ref, c->perm, M,
                              ilog2_ceil((unsigned long long) c->n_total)));
  return gather_nodes(c, c->perm, c->x_sorted);
}

}  // namespace nfftcu
