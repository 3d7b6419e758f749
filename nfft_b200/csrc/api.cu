// api.cu -- the extern "C" boundary of libnfftcu.so (include/nfftcu.h): plan context life cycle,
// node upload/sort, the trafo/adjoint drivers and the plumbing entry points.
//
// Driver structure mirrors the reference's nfft_trafo / nfft_adjoint (kernel/nfft/nfft.c:5655-5749):
//   trafo   = D (deconv.cu) -> F forward (fft.cu)  -> B   (interp.cu)
//   adjoint = B^T (spread.cu) -> F backward (fft.cu) -> D^T (deconv.cu)
// with the exact NDFT (ndft.cu) when any N_t <= m or n_t <= 2m+2.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <thread>

namespace nfftcu {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

// I0(x) by its power series sum_k ((x/2)^2)^k/(k!)^2 in long double: all terms positive, so the
// double result is good to ~1 ulp over the window's argument range.  Takes the role of
// nfft_bessel_i0 (kernel/util/bessel_i0.c:300-338) inside PHI_HUT (include/infft.h:208).
long double bessel_i0_series(long double x) {
  const long double q = 0.25L * x * x;
  long double term = 1.0L, sum = 1.0L;
  for (int k = 1; k < 4000; k++) {
    term *= q / ((long double) k * (long double) k);
    sum += term;
    if (term < sum * 0x1p-70L) break;
  }
  return sum;
}

// ---- 64-bit fingerprint of a host array (node change detection without moving the nodes) ----------------------
// Plans without a node-bound psi flag get no notice when the caller rewrites x; the reference simply re-reads x on
// every transform (nfft.c:4889, 5351).  Re-uploading x to compare it on the device costs 24 B/node of PCIe and host
// memory traffic per transform; instead the host array is hashed by a few host threads WHILE the device already runs
// the transform with the resident nodes, and only a mismatch triggers upload + re-sort + a repeated transform.
// Chunks of 1 MiB are hashed independently (four multiply-rotate lanes over 8-byte words, xxhash-style constants)
// and combined in order, so the result does not depend on the thread count.
inline uint64_t rotl64(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }
inline uint64_t mix64(uint64_t h) {
  h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
  return h;
}
uint64_t hash_chunk(const unsigned char *p, size_t bytes, uint64_t seed) {
  const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL;
  uint64_t a0 = seed + P1, a1 = seed ^ P2, a2 = seed * P1 + 1, a3 = seed - P2;
  size_t i = 0;
  for (; i + 32 <= bytes; i += 32) {
    uint64_t w[4];
    memcpy(w, p + i, 32);
    a0 = rotl64(a0 + w[0] * P2, 31) * P1;
    a1 = rotl64(a1 + w[1] * P2, 31) * P1;
    a2 = rotl64(a2 + w[2] * P2, 31) * P1;
    a3 = rotl64(a3 + w[3] * P2, 31) * P1;
  }
  uint64_t tail = 0;
  for (int sh = 0; i < bytes; i++, sh += 8) {
    if (sh == 64) { a0 = rotl64(a0 + tail * P2, 31) * P1; tail = 0; sh = 0; }
    tail |= (uint64_t) p[i] << sh;
  }
  a1 = rotl64(a1 + tail * P2, 31) * P1;
  return mix64(rotl64(a0, 1) + rotl64(a1, 7) + rotl64(a2, 12) + rotl64(a3, 18) + (uint64_t) bytes);
}
}  // namespace
uint64_t fingerprint(const void *data, size_t bytes) {
  const size_t kChunk = (size_t) 1 << 20;
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  const unsigned char *p = (const unsigned char *) data;
  if (nchunks <= 4) {
    uint64_t h = 0x1234567ULL;
    for (size_t k = 0; k < nchunks; k++)
      h = mix64(h ^ hash_chunk(p + k * kChunk, k + 1 < nchunks ? kChunk : bytes - k * kChunk, k));
    return h ^ bytes;
  }
  static const unsigned max_threads = [] {
    const char *e = getenv("NFFT_B200_HASH_THREADS");
    unsigned t = e ? (unsigned) atoi(e) : std::thread::hardware_concurrency();
    if (const char *o = getenv("OMP_NUM_THREADS")) { const unsigned ot = (unsigned) atoi(o); if (!e && ot > 0 && ot < t) t = ot; }
    return t < 1 ? 1u : (t > 32 ? 32u : t);
  }();
  unsigned nt = max_threads;
  if (nt > nchunks / 2) nt = (unsigned) (nchunks / 2);
  if (nt < 1) nt = 1;
  std::vector<uint64_t> part(nchunks);
  auto work = [&](unsigned tid) {
    for (size_t k = tid; k < nchunks; k += nt)
      part[k] = hash_chunk(p + k * kChunk, k + 1 < nchunks ? kChunk : bytes - k * kChunk, k);
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
  work(0);
  for (auto &t : th) t.join();
  uint64_t h = 0x1234567ULL;
  for (size_t k = 0; k < nchunks; k++) h = mix64(h ^ part[k]);
  return h ^ bytes;
}
namespace {
bool exact_node_check() {
  static const bool v = [] { const char *e = getenv("NFFT_B200_EXACT_NODE_CHECK"); return e && atoi(e) != 0; }();
  return v;
}

}  // namespace
long double bessel_i0_ld(long double x) { return bessel_i0_series(x); }   // for mri.cu
namespace {

int check_ctx(const nfftcu_ctx *c) {
  if (!c) {
    set_error("null context");
    return NFFTCU_EINVAL;
  }
  return NFFTCU_OK;
}

int bind_device(const nfftcu_ctx *c) {
  NFFTCU_CUDA(cudaSetDevice(c->device));
  return NFFTCU_OK;
}

size_t cbytes(const nfftcu_ctx *c, long long count) { return 2 * real_size(c) * (size_t) count; }

int ensure_staging(nfftcu_ctx *c) {
  if (!c->fhat_dev && c->N_total > 0) NFFTCU_CUDA(pool_malloc(&c->fhat_dev, cbytes(c, c->N_total) * (size_t) c->batch_cap));
  if (!c->f_dev && c->M > 0) NFFTCU_CUDA(pool_malloc(&c->f_dev, cbytes(c, c->M) * (size_t) c->batch_cap));
  return NFFTCU_OK;
}

struct StageTimer {
  nfftcu_ctx *c;
  explicit StageTimer(nfftcu_ctx *ctx) : c(ctx) {}
  void mark(int i) { if (c->opt_timing) cudaEventRecord(c->ev[i], c->stream); }
  void finish(int s0, int s1, int s2) {
    if (!c->opt_timing) return;
    cudaEventSynchronize(c->ev[3]);
    float a = 0, b = 0, d = 0;
    cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&d, c->ev[2], c->ev[3]);
    c->stage_ms[s0] = a;
    c->stage_ms[s1] = b;
    c->stage_ms[s2] = d;
    c->bkernel_ms = 0.f;
    if (c->evk_recorded) cudaEventElapsedTime(&c->bkernel_ms, c->evk[0], c->evk[1]);
    c->evk_recorded = false;
  }
};

int need_nodes(const nfftcu_ctx *c) {
  if (!c->have_nodes) {
    set_error("transform called before nfftcu_set_nodes");
    return NFFTCU_ESTATE;
  }
  return NFFTCU_OK;
}

int trafo_dev_impl(nfftcu_ctx *c, const void *f_hat_dev, void *f_dev) {
  NFFTCU_TRY(need_nodes(c));
  if (c->direct_only) return ndft_trafo(c, f_hat_dev, f_dev);
  StageTimer tm(c);
  tm.mark(0);
  NFFTCU_TRY(stage_D(c, f_hat_dev, c->opt_fft_prune != 0 && !c->fft_no_prune));
  tm.mark(1);
  NFFTCU_TRY(stage_F(c, -1, c->opt_fft_prune != 0 && !c->fft_no_prune));
  tm.mark(2);
  NFFTCU_TRY(stage_B(c, f_dev));
  tm.mark(3);
  tm.finish(0, 1, 2);   // MEASURE_TIME_t[0..2] = D, F, B (nfft.c:5415,5513,5515-5521)
  return NFFTCU_OK;
}

int adjoint_dev_impl(nfftcu_ctx *c, const void *f_dev, void *f_hat_dev, bool peer_reduce = false) {
  NFFTCU_TRY(need_nodes(c));
  if (c->direct_only) return ndft_adjoint(c, f_dev, f_hat_dev);
  StageTimer tm(c);
  tm.mark(0);
  const bool pruned = c->opt_fft_prune != 0 && !c->fft_no_prune;
  NFFTCU_TRY(stage_BT(c, f_dev, pruned));
  tm.mark(1);
  NFFTCU_TRY(stage_F(c, +1, c->opt_fft_prune != 0 && !c->fft_no_prune));
  tm.mark(2);
  // peer_reduce: D^T with the cross-GPU sum taken inside the kernel through peer pointers (peer.cu)
  NFFTCU_TRY(peer_reduce ? peer_reduce_DT(c, f_hat_dev) : stage_DT(c, f_hat_dev));
  tm.mark(3);
  tm.finish(2, 1, 0);
  return NFFTCU_OK;
}

}  // namespace

// make room for K right-hand sides: the grid(s), the gathered-sample buffer of the spreading kernels and the
// host-pointer staging buffers grow to K slices (contents are scratch between transforms)
int ensure_batch(nfftcu_ctx *c, int K) {
  if (K < 1) { set_error("batched transform: K = %d", K); return NFFTCU_EINVAL; }
  if (K <= c->batch_cap) return NFFTCU_OK;
  if (c->direct_only || !c->grid) { set_error("batched transform: plan has no grid (NDFT fallback plan)"); return NFFTCU_ESTATE; }
  if (c->peer) { set_error("batched transform on a peer-attached (multi-GPU) plan is not supported"); return NFFTCU_ESTATE; }
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  const size_t gb = cbytes(c, c->n_total) * (size_t) K;
  void *g = nullptr;
  NFFTCU_CUDA(pool_malloc(&g, gb));
  pool_free(c->grid);
  c->grid = g;
  NFFTCU_CUDA(cudaMemsetAsync(c->grid, 0, gb, c->stream));   // see nfftcu_create
  if (c->grid2) {
    void *g2 = nullptr;
    NFFTCU_CUDA(pool_malloc(&g2, gb));
    pool_free(c->grid2);
    c->grid2 = g2;
  }
  if (c->f_tile) { pool_free(c->f_tile); c->f_tile = nullptr; NFFTCU_CUDA(pool_malloc(&c->f_tile, cbytes(c, c->M) * (size_t) K)); }
  if (c->fhat_dev) { pool_free(c->fhat_dev); c->fhat_dev = nullptr; }
  if (c->f_dev) { pool_free(c->f_dev); c->f_dev = nullptr; }
  c->batch_cap = K;
  return NFFTCU_OK;
}

// min / max over the nodes of u_0 = (floor(n_0 x_0) - m) mod n_0, the first plane a node's taps touch
template <typename T>
__global__ void slab_minmax_kernel(const T *__restrict__ x, long long M, int d, long long n0, long long m, int *mm) {
  int lo = 0x7fffffff, hi = -1;
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x; j < M; j += stride) {
    long long u = (cell_of(x[j * d], n0) - m) % n0;
    if (u < 0) u += n0;
    lo = min(lo, (int) u);
    hi = max(hi, (int) u);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && hi >= 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}

// Slab mode: when the nodes of the plan occupy a slab of the first axis -- a rank of a node-sharded multi-GPU run holds
// 1/P of the globally sorted nodes -- only the planes [u_min, u_max + 2m + 2) of the grid are ever touched by B / B^T.
// The pruned F passes and the B^T memset then skip the rest (fft.cu run_axis_slab).  Off when the window is wider than
// 3/4 of the axis (single-GPU runs), for plans with a split FFT axis and with NFFTCU_OPT_SLAB_FFT = 1.
static int slab_detect(nfftcu_ctx *c) {
  c->slab_on = false;
  if (c->direct_only || c->nodes_only || c->opt_slab == 1 || c->d < 2 || c->M == 0 || c->fft_no_prune ||
      c->n[0] > 0x3fffffff)
    return NFFTCU_OK;
  int *mm = nullptr;
  NFFTCU_CUDA(pool_malloc((void **) &mm, 2 * sizeof(int)));
  const int init[2] = {0x7fffffff, -1};
  NFFTCU_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  long long blocks = (c->M + 255) / 256;
  if (blocks > (long long) c->sm_count * 8) blocks = (long long) c->sm_count * 8;
  if (c->prec == NFFTCU_DOUBLE)
    slab_minmax_kernel<double><<<(unsigned) blocks, 256, 0, c->stream>>>((const double *) c->x_dev, c->M, c->d, c->n[0], c->m, mm);
  else
    slab_minmax_kernel<float><<<(unsigned) blocks, 256, 0, c->stream>>>((const float *) c->x_dev, c->M, c->d, c->n[0], c->m, mm);
  c->launches++;
  int h[2] = {0, 0};
  NFFTCU_CUDA(cudaMemcpyAsync(h, mm, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  pool_free(mm);
  if (h[1] < h[0]) return NFFTCU_OK;
  const long long wc = (long long) h[1] - h[0] + 2 * c->m + 2;
  if (4 * wc <= 3 * c->n[0]) {
    c->slab_on = true;
    c->slab_w0 = h[0];
    c->slab_wc = wc;
  }
  return NFFTCU_OK;
}

// the options the node-dependent state (orders, bins, tables, images) was built under
static unsigned node_opts_signature(const nfftcu_ctx *c) {
  return (unsigned) (c->opt_b_kernel & 15) | ((unsigned) (c->opt_node_order & 15) << 4) | ((unsigned) (c->opt_psi_table & 1) << 8) |
         ((unsigned) (c->opt_window_images & 3) << 9) | ((c->flags & (1u << 11)) ? 1u << 11 : 0u) |
         ((unsigned) (c->opt_slab & 1) << 12) | ((unsigned) (c->opt_tc5 & 3) << 13);
}

int nodes_ready(nfftcu_ctx *c) {
  c->built_sig = node_opts_signature(c);
  if (c->nodes_only) {   // sorter of a multi-GPU group: the reference order is all that is needed
    NFFTCU_TRY(sort_nodes(c));
    c->ref_sorted = true;
    c->have_nodes = true;
    c->nodes_version++;
    return NFFTCU_OK;
  }
  c->ref_sorted = false;
  c->tile_ready = false;
  c->mma_ready = false;
  c->tile2_ready = false;
  c->psi_table_valid = false;   // the generic kernels use the table only when it was rebuilt for these nodes
  if (!c->direct_only) {
    // B / B^T kernel family: DMMA (mma3d.cu) > register pencils (pencil3d.cu) > generic warp-per-node
    const bool use_mma = mma3d_supported(c) && (c->opt_b_kernel == 0 || c->opt_b_kernel == 3);
    const bool use_tile = !use_mma && tile3d_supported(c) && c->opt_b_kernel != 1;
    const bool use_tile2 = tile2d_supported(c) && c->opt_b_kernel != 1;   // d = 2: shared-memory tiles (tile2d.cu)
    // the reference-order sort is the index_x witness and the order the generic kernels walk
    if (!(use_tile || use_mma || use_tile2) || (c->flags & (1u << 11))) {   // NFFT_SORT_NODES
      NFFTCU_TRY(sort_nodes(c));
      c->ref_sorted = true;
    }
    if (use_mma) NFFTCU_TRY(mma3d_bin_nodes(c));
    else if (use_tile) NFFTCU_TRY(tile3d_bin_nodes(c));
    else if (use_tile2) NFFTCU_TRY(tile2d_bin_nodes(c));
    else if (c->opt_psi_table) NFFTCU_TRY(build_psi_table(c));
  }
  NFFTCU_TRY(slab_detect(c));
  c->have_nodes = true;
  c->nodes_version++;
  return NFFTCU_OK;
}

namespace {
template <typename C>
__global__ void cmul_diag_kernel(C *__restrict__ x, const C *__restrict__ b, long long n) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const C a = x[k], bb = b[k];
    C r;
    r.x = bb.x * a.x - bb.y * a.y;     // b[k] * f_hat[k], operand order of fastsum.c:1206
    r.y = bb.x * a.y + bb.y * a.x;
    x[k] = r;
  }
}

__global__ void differs_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                               long long words, int *flag) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  bool diff = false;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride)
    diff |= (a[i] != b[i]);
  if (__syncthreads_or(diff) && threadIdx.x == 0) *flag = 1;
}

// upload x into the staging buffer; returns changed=false when it equals the resident nodes
int stage_and_compare(nfftcu_ctx *c, const void *x, cudaMemcpyKind kind, bool *changed, uint64_t *fp_out) {
  const size_t bytes = real_size(c) * (size_t) c->M * c->d;
  *changed = true;
  if (fp_out && bytes == 0) *fp_out = fingerprint(x, 0);
  if (bytes == 0) { *changed = !c->have_nodes; return NFFTCU_OK; }
  if (!c->x_stage) NFFTCU_CUDA(pool_malloc(&c->x_stage, bytes));
  if (!c->x_dev) NFFTCU_CUDA(pool_malloc(&c->x_dev, bytes));
  if (!c->diff_flag) NFFTCU_CUDA(pool_malloc((void **) &c->diff_flag, sizeof(int)));
  NFFTCU_CUDA(cudaMemcpyAsync(c->x_stage, x, bytes, kind, c->stream));
  if (fp_out) *fp_out = fingerprint(x, bytes);   // host threads, while the copy is in flight
  if (c->have_nodes) {
    int h = 0;
    NFFTCU_CUDA(cudaMemsetAsync(c->diff_flag, 0, sizeof(int), c->stream));
    const long long words = (long long) (bytes / 4);
    long long blocks = (words + 255) / 256;
    if (blocks > (long long) c->sm_count * 16) blocks = (long long) c->sm_count * 16;
    differs_kernel<<<(unsigned) blocks, 256, 0, c->stream>>>((const uint32_t *) c->x_stage,
                                                           (const uint32_t *) c->x_dev, words,
                                                           c->diff_flag);
    c->launches++;
    NFFTCU_CUDA(cudaMemcpyAsync(&h, c->diff_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
    *changed = (h != 0);
  }
  if (*changed) { void *t = c->x_dev; c->x_dev = c->x_stage; c->x_stage = t; }
  return NFFTCU_OK;
}

}  // namespace
}  // namespace nfftcu

using namespace nfftcu;

extern "C" {

const char *nfftcu_last_error(void) { return g_err; }

int nfftcu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int nfftcu_create(nfftcu_ctx **out, int precision, int d, const int64_t *N, const int64_t *n,
                  int64_t m, int64_t M, unsigned flags, int device) {
  return create_ctx(out, precision, d, N, n, m, M, flags, device, false);
}

}  // extern "C"

// ---- plan cache: finalized plans parked for the next identical nfft_init ------------------------------------------
// Plan-per-coil callers (applications/mri/mri2d/reconstruct_data_2d.c:38-139; BASELINE configs[4]) run
// nfft_init_guru -> x -> nfft_precompute_one_psi -> solver -> nfft_finalize once per coil, every time with the same
// geometry and the SAME nodes.  A finalized plan is therefore parked (at most kPlanCacheSlots, small plans only)
// and handed out again by the next create with an identical key; its resident nodes keep their fingerprint, so the
// following set_nodes on an unchanged x is a no-op and the sort / binning / window tables are reused -- the
// "many vectors, one node set" sharing of SURVEY 8f rank 2 for callers that cannot use the batch API.
// NFFT_B200_PLAN_CACHE=0 disables it; nfftcu_pool_trim() empties it.
namespace nfftcu {
namespace {
constexpr int kPlanCacheSlots = 2;
std::mutex g_plan_cache_mu;
std::vector<nfftcu_ctx *> g_plan_cache;

bool plan_cache_enabled() {
  static const bool v = [] { const char *e = getenv("NFFT_B200_PLAN_CACHE"); return !(e && atoi(e) == 0); }();
  return v;
}
bool same_plan(const nfftcu_ctx *c, int precision, int d, const int64_t *N, const int64_t *n, int64_t m, int64_t M,
               unsigned flags, int device, bool nodes_only) {
  const int window = (flags & NFFTCU_FLAG_GAUSSIAN) ? NFFTCU_WINDOW_GAUSSIAN : NFFTCU_WINDOW_KAISER_BESSEL;
  if (c->prec != precision || c->d != d || c->m != m || c->M != M || c->flags != (flags & ~NFFTCU_FLAG_GAUSSIAN) ||
      c->device != device || c->nodes_only != nodes_only || c->window != window)
    return false;
  for (int t = 0; t < d; t++)
    if (c->N[t] != N[t] || c->n[t] != n[t]) return false;
  return true;
}
size_t plan_device_bytes(const nfftcu_ctx *c) {   // rough: what parking this plan keeps allocated
  const size_t r = real_size(c);
  return 2 * r * (size_t) c->n_total * (size_t) c->batch_cap * (c->grid2 ? 2 : 1) + c->mma_images_bytes + c->tc5i.images_bytes + c->tc5s.images_bytes +
         (size_t) c->M * (size_t) (r * c->d * 4 + 48);
}
}  // namespace
void plan_cache_clear() {
  std::vector<nfftcu_ctx *> victims;
  {
    std::lock_guard<std::mutex> lock(g_plan_cache_mu);
    victims.swap(g_plan_cache);
  }
  for (nfftcu_ctx *c : victims) { c->no_cache = true; nfftcu_destroy(c); }
}
}  // namespace nfftcu

int nfftcu::create_ctx(nfftcu_ctx **out, int precision, int d, const int64_t *N, const int64_t *n,
                       int64_t m, int64_t M, unsigned flags, int device, bool nodes_only) {
  if (out && N && n && d >= 1 && d <= NFFTCU_MAX_D && plan_cache_enabled()) {
    std::lock_guard<std::mutex> lock(g_plan_cache_mu);
    for (size_t i = 0; i < g_plan_cache.size(); i++)
      if (same_plan(g_plan_cache[i], precision, d, N, n, m, M, flags, device, nodes_only)) {
        nfftcu_ctx *c = g_plan_cache[i];
        g_plan_cache.erase(g_plan_cache.begin() + (long) i);
        // a revived plan starts with the defaults a new one has; its node state (and fingerprint) is kept
        c->opt_timing = 0;
        c->opt_psi_table = 0;
        c->opt_b_flush = 0;
        c->opt_fft_prune = 1;
        c->opt_slab = 0;
        c->cur_batch = 1;
        // to its new owner the plan has no nodes yet; the parked ones are adopted by the first nfftcu_set_nodes
        // whose host array has the same fingerprint
        c->parked_nodes = c->have_nodes && c->x_fp_valid;
        c->have_nodes = false;
        c->nodes_version = 0;
        c->launches = 0;
        *out = c;
        return NFFTCU_OK;
      }
  }
  if (!out || !N || !n || d < 1 || d > NFFTCU_MAX_D || m < 0 || M < 0 ||
      (precision != NFFTCU_DOUBLE && precision != NFFTCU_FLOAT)) {
    set_error("nfftcu_create: invalid argument (d=%d, m=%lld, M=%lld, precision=%d)", d,
              (long long) m, (long long) M, precision);
    return NFFTCU_EINVAL;
  }
  if (M > 0xffffffffll) {
    set_error("nfftcu_create: M=%lld exceeds the 32-bit node index of the sort", (long long) M);
    return NFFTCU_EINVAL;
  }
  int ndev = nfftcu_device_count();
  if (ndev <= 0 || device < 0 || device >= ndev) {
    set_error("nfftcu_create: no usable CUDA device (count=%d, requested %d); there is no CPU path",
              ndev, device);
    return NFFTCU_ENODEV;
  }
  NFFTCU_CUDA(cudaSetDevice(device));
  nfftcu_ctx *c = new nfftcu_ctx_s();
  // every failure below releases what the context already holds (nfftcu_destroy copes with a half-built context)
  const int status = [&]() -> int {
  c->prec = precision;
  c->d = d;
  c->device = device;
  c->m = m;
  c->M = M;
  c->flags = flags & ~NFFTCU_FLAG_GAUSSIAN;
  c->window = (flags & NFFTCU_FLAG_GAUSSIAN) ? NFFTCU_WINDOW_GAUSSIAN : NFFTCU_WINDOW_KAISER_BESSEL;
  c->nodes_only = nodes_only;
  c->N_total = 1;
  c->n_total = 1;
  for (int t = 0; t < d; t++) {
    if (N[t] < 1 || n[t] < 1) {
      set_error("nfftcu_create: N[%d]=%lld, n[%d]=%lld", t, (long long) N[t], t, (long long) n[t]);
      return NFFTCU_EINVAL;
    }
    c->N[t] = N[t];
    c->n[t] = n[t];
    c->N_total *= N[t];
    c->n_total *= n[t];
    if (N[t] <= m || n[t] <= 2 * m + 2) c->direct_only = true;
    // sigma and b in the plan precision like init_help (nfft.c:5961-5964, infft.h:216-222)
    // Gaussian: b = 2 sigma / (2 sigma - 1) * m / pi (WINDOW_HELP_INIT, infft.h:157-165)
    const double kPi = 3.1415926535897932384626433832795028841971693993751;
    if (precision == NFFTCU_DOUBLE) {
      c->sigma[t] = (double) n[t] / (double) N[t];
      c->b[t] = c->window == NFFTCU_WINDOW_GAUSSIAN
                    ? (2.0 * c->sigma[t]) / (2.0 * c->sigma[t] - 1.0) * ((double) m / kPi)
                    : kPi * (2.0 - 1.0 / c->sigma[t]);
    } else {
      const float sg = (float) n[t] / (float) N[t];
      c->sigma[t] = sg;
      c->b[t] = c->window == NFFTCU_WINDOW_GAUSSIAN
                    ? (double) ((2.0f * sg) / (2.0f * sg - 1.0f) * ((float) m / (float) kPi))
                    : (double) ((float) kPi * (2.0f - 1.0f / sg));
    }
  }
  cudaDeviceProp prop;
  NFFTCU_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  NFFTCU_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  for (int i = 0; i < 4; i++) NFFTCU_CUDA(cudaEventCreate(&c->ev[i]));
  for (int i = 0; i < 2; i++) NFFTCU_CUDA(cudaEventCreate(&c->evk[i]));

  // c_t[k+N/2] = 1/phi_hat_t(k), k = -N/2..  (precompute_phi_hut, nfft.c:5754-5770):
  //   Kaiser-Bessel 1/I0(m sqrt(b^2 - (2 pi k/n)^2)) (infft.h:208), Gaussian exp((pi k/n)^2 b) (infft.h:154)
  const long double two_pi = 6.283185307179586476925286766559005768394L;
  for (int t = 0; t < d; t++) {
    c->c_host[t].resize((size_t) N[t]);
    for (long long ks = 0; ks < N[t]; ks++) {
      const long double w = two_pi * (long double) (ks - N[t] / 2) / (long double) n[t];
      const long double bb = (long double) c->b[t];
      if (c->window == NFFTCU_WINDOW_GAUSSIAN) {
        c->c_host[t][(size_t) ks] = (double) expl(0.25L * w * w * bb);
      } else {
        const long double arg2 = bb * bb - w * w;
        const long double arg = (long double) m * sqrtl(arg2 > 0 ? arg2 : 0.0L);
        c->c_host[t][(size_t) ks] = (double) (1.0L / bessel_i0_series(arg));
      }
    }
    // fp32 plans: move phi_hat_t(0) ~ 1e11 (Kaiser-Bessel, m = 6) out of the window and into c as an exact power of two
    c->wscale[t] = 1.0;
    if (precision == NFFTCU_FLOAT) {
      int e = 0;
      frexp(c->c_host[t][(size_t) (N[t] / 2)], &e);   // c(0) = f * 2^e, f in [0.5, 1)
      c->wscale[t] = ldexp(1.0, e);
    }
    const size_t bytes = real_size(c) * (size_t) N[t];
    NFFTCU_CUDA(pool_malloc(&c->c_dev[t], bytes));
    if (precision == NFFTCU_DOUBLE) {
      NFFTCU_CUDA(cudaMemcpy(c->c_dev[t], c->c_host[t].data(), bytes, cudaMemcpyHostToDevice));
    } else {
      std::vector<float> tmp((size_t) N[t]);
      // round to float FIRST (the value the reference multiplies with), then scale exactly
      for (size_t i = 0; i < tmp.size(); i++) tmp[i] = (float) ((double) (float) c->c_host[t][i] / c->wscale[t]);
      NFFTCU_CUDA(cudaMemcpy(c->c_dev[t], tmp.data(), bytes, cudaMemcpyHostToDevice));
    }
  }
  if (!c->direct_only && !nodes_only) {
    NFFTCU_CUDA(pool_malloc(&c->grid, cbytes(c, c->n_total)));
    // never leave pool garbage in the grid: the slab mode of the F step writes only the planes the nodes touch, and the
    // tiled B kernels read their whole footprint (zero weights times a NaN bit pattern would still be NaN)
    NFFTCU_CUDA(cudaMemsetAsync(c->grid, 0, cbytes(c, c->n_total), c->stream));
    NFFTCU_TRY(fft_plan_axes(c));
    if (tile3d_supported(c) || tile2d_supported(c)) NFFTCU_TRY(build_kb_poly(c));
  }
  return NFFTCU_OK;
  }();
  if (status != NFFTCU_OK) {
    const std::string msg = g_err;   // nfftcu_destroy must not clobber the reason
    nfftcu_destroy(c);
    set_error("%s", msg.c_str());
    return status;
  }
  *out = c;
  return NFFTCU_OK;
}

extern "C" {

int nfftcu_destroy(nfftcu_ctx *c) {
  if (!c) return NFFTCU_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  // park small, plain plans for the next identical create (see "plan cache" above)
  if (!c->no_cache && plan_cache_enabled() && !c->peer && c->own_stream && !c->direct_only &&
      c->opt_b_kernel == 0 && c->opt_node_order == 0 && c->opt_fft_kernel == 0 && c->opt_window_images == 0 &&
      plan_device_bytes(c) <= ((size_t) 1 << 30)) {
    nfftcu_ctx *evict = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_plan_cache_mu);
      g_plan_cache.push_back(c);
      if (g_plan_cache.size() > (size_t) kPlanCacheSlots) {
        evict = g_plan_cache.front();
        g_plan_cache.erase(g_plan_cache.begin());
      }
    }
    if (!evict) return NFFTCU_OK;
    c = evict;
    c->no_cache = true;
    cudaSetDevice(c->device);
  }
  peer_detach(c);
  fft_free_axes(c);
  for (int t = 0; t < NFFTCU_MAX_D; t++)
    if (c->c_dev[t]) pool_free(c->c_dev[t]);
  void *bufs[] = {c->tc5i.images, c->tc5i.batches, (void *) c->tc5i.batch_start, (void *) c->tc5i.chunk_start, c->tc5i.chunks,
                  c->tc5s.images, c->tc5s.batches, (void *) c->tc5s.batch_start, (void *) c->tc5s.chunk_start, c->tc5s.chunks, (void *) c->tc5_counts, c->tc5_ft,
                  c->mma_images, c->mma_batches, (void *) c->mma_batch_start, (void *) c->mma_counts, (void *) c->mma_chunk_start, c->mma_chunks, c->f_tile, c->kbpoly_dev, c->tile_keys, (void *) c->tile_perm, c->tile_x, (void *) c->bin_start, c->tile_psi, c->grid, c->x_dev, c->x_stage, (void *) c->diff_flag, c->x_sorted, (void *) c->perm, c->keys_ref, c->psi_table,
                  c->sort_tmp, c->fhat_dev, c->f_dev};
  for (void *p : bufs)
    if (p) pool_free(p);
  for (int i = 0; i < 4; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 2; i++)
    if (c->evk[i]) cudaEventDestroy(c->evk[i]);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->ev_side) cudaEventDestroy(c->ev_side);
  if (c->h_flag) pool_free_host(c->h_flag);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return NFFTCU_OK;
}

int nfftcu_get_c_phi_inv(nfftcu_ctx *c, int t, void *out_host) {
  NFFTCU_TRY(check_ctx(c));
  if (t < 0 || t >= c->d || !out_host) {
    set_error("nfftcu_get_c_phi_inv: bad dimension %d", t);
    return NFFTCU_EINVAL;
  }
  for (size_t i = 0; i < c->c_host[t].size(); i++) {
    if (c->prec == NFFTCU_DOUBLE) ((double *) out_host)[i] = c->c_host[t][i];
    else ((float *) out_host)[i] = (float) c->c_host[t][i];
  }
  return NFFTCU_OK;
}

int nfftcu_get_window_scale(nfftcu_ctx *c, double *scale_host) {
  NFFTCU_TRY(check_ctx(c));
  for (int t = 0; t < c->d; t++) scale_host[t] = c->wscale[t];
  return NFFTCU_OK;
}

int nfftcu_get_window_params(nfftcu_ctx *c, void *b_host, void *sigma_host) {
  NFFTCU_TRY(check_ctx(c));
  for (int t = 0; t < c->d; t++) {
    if (c->prec == NFFTCU_DOUBLE) {
      if (b_host) ((double *) b_host)[t] = c->b[t];
      if (sigma_host) ((double *) sigma_host)[t] = c->sigma[t];
    } else {
      if (b_host) ((float *) b_host)[t] = (float) c->b[t];
      if (sigma_host) ((float *) sigma_host)[t] = (float) c->sigma[t];
    }
  }
  return NFFTCU_OK;
}

int nfftcu_set_nodes(nfftcu_ctx *c, const void *x_host) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  if (c->M > 0 && !x_host) {
    set_error("nfftcu_set_nodes: x is NULL");
    return NFFTCU_EINVAL;
  }
  const size_t xbytes = real_size(c) * (size_t) c->M * c->d;
  uint64_t fp = 0;
  bool have_fp = false;
  if ((c->have_nodes || c->parked_nodes) && c->x_fp_valid && !exact_node_check()) {
    fp = fingerprint(x_host, xbytes);
    have_fp = true;
    if (fp == c->x_fp) {   // same nodes as the resident ones: nothing to upload, nothing to sort
      if (c->parked_nodes) {   // a revived plan (plan cache) adopts the node state its predecessor left
        c->parked_nodes = false;
        c->have_nodes = true;
        c->nodes_version++;
      }
      // ... unless that state was built under other kernel / order options than the ones in force now
      if (c->built_sig != node_opts_signature(c)) {
        NFFTCU_TRY(nodes_ready(c));
        NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
      }
      return NFFTCU_OK;
    }
    have_fp = c->have_nodes;   // parked nodes that do not match: a fresh upload, nothing to swap with
  }
  c->parked_nodes = false;
  bool changed = true;
  if (have_fp) {   // known to differ: upload straight into the resident buffer's partner and swap
    if (xbytes) {
      if (!c->x_stage) NFFTCU_CUDA(pool_malloc(&c->x_stage, xbytes));
      NFFTCU_CUDA(cudaMemcpyAsync(c->x_stage, x_host, xbytes, cudaMemcpyHostToDevice, c->stream));
      void *t = c->x_dev; c->x_dev = c->x_stage; c->x_stage = t;
    }
  } else {
    NFFTCU_TRY(stage_and_compare(c, x_host, cudaMemcpyHostToDevice, &changed, &fp));
  }
  c->x_fp = fp;
  c->x_fp_valid = true;
  if (changed) NFFTCU_TRY(nodes_ready(c));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  return NFFTCU_OK;
}

int nfftcu_set_nodes_dev(nfftcu_ctx *c, const void *x_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  bool changed = true;
  c->parked_nodes = false;
  NFFTCU_TRY(stage_and_compare(c, x_dev, cudaMemcpyDefault, &changed, nullptr));   // x_dev may live on a peer device
  c->x_fp_valid = false;   // no host array to fingerprint: transforms with a node refresh compare on the device
  if (changed) NFFTCU_TRY(nodes_ready(c));
  return NFFTCU_OK;
}

int64_t nfftcu_nodes_version(nfftcu_ctx *c) { return c ? c->nodes_version : 0; }

int nfftcu_get_index_x(nfftcu_ctx *c, int64_t *index_x_host) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(need_nodes(c));
  if (c->M == 0) return NFFTCU_OK;
  if (c->direct_only) {
    set_error("nfftcu_get_index_x: plan has no sorted nodes (direct-only plan)");
    return NFFTCU_ESTATE;
  }
  if (!c->ref_sorted) {   // plans without NFFT_SORT_NODES on the tile path: sort on demand
    NFFTCU_TRY(sort_nodes(c));
    c->ref_sorted = true;
  }
  std::vector<uint64_t> keys((size_t) c->M);
  std::vector<uint32_t> perm((size_t) c->M);
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  NFFTCU_CUDA(cudaMemcpy(keys.data(), c->keys_ref, sizeof(uint64_t) * keys.size(), cudaMemcpyDeviceToHost));
  NFFTCU_CUDA(cudaMemcpy(perm.data(), c->perm_ref, sizeof(uint32_t) * perm.size(), cudaMemcpyDeviceToHost));
  for (size_t k = 0; k < keys.size(); k++) {
    index_x_host[2 * k] = (int64_t) keys[k];
    index_x_host[2 * k + 1] = (int64_t) perm[k];
  }
  return NFFTCU_OK;
}

int nfftcu_get_sorted_slab(nfftcu_ctx *c, int64_t begin, int64_t end, void *x_out_dev, uint32_t *perm_out_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(need_nodes(c));
  if (begin < 0 || end < begin || end > c->M) {
    set_error("nfftcu_get_sorted_slab: range [%lld, %lld) outside [0, %lld)", (long long) begin, (long long) end, (long long) c->M);
    return NFFTCU_EINVAL;
  }
  if (c->direct_only && !c->nodes_only) {
    set_error("nfftcu_get_sorted_slab: plan has no sorted nodes (direct-only plan)");
    return NFFTCU_ESTATE;
  }
  if (!c->ref_sorted) {
    NFFTCU_TRY(sort_nodes(c));
    c->ref_sorted = true;
  }
  const size_t rs = real_size(c) * (size_t) c->d;
  if (x_out_dev && end > begin)
    NFFTCU_CUDA(cudaMemcpyAsync(x_out_dev, (const char *) c->x_sorted + rs * (size_t) begin, rs * (size_t) (end - begin),
                                cudaMemcpyDeviceToDevice, c->stream));
  if (perm_out_dev && end > begin)
    NFFTCU_CUDA(cudaMemcpyAsync(perm_out_dev, c->perm_ref + begin, sizeof(uint32_t) * (size_t) (end - begin),
                                cudaMemcpyDeviceToDevice, c->stream));
  return NFFTCU_OK;
}

static int host_transform(nfftcu_ctx *c, const void *in_host, void *out_host, int which) {
  // which: 0 trafo, 1 adjoint, 2 trafo_direct, 3 adjoint_direct
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(need_nodes(c));
  NFFTCU_TRY(ensure_staging(c));
  const bool forward = (which == 0 || which == 2);
  const size_t in_bytes = forward ? cbytes(c, c->N_total) : cbytes(c, c->M);
  const size_t out_bytes = forward ? cbytes(c, c->M) : cbytes(c, c->N_total);
  void *in_dev = forward ? c->fhat_dev : c->f_dev;
  void *out_dev = forward ? c->f_dev : c->fhat_dev;
  if (in_bytes) NFFTCU_CUDA(cudaMemcpyAsync(in_dev, in_host, in_bytes, cudaMemcpyHostToDevice, c->stream));
  int r;
  switch (which) {
    case 0: r = trafo_dev_impl(c, in_dev, out_dev); break;
    case 1: r = adjoint_dev_impl(c, in_dev, out_dev); break;
    case 2: r = ndft_trafo(c, in_dev, out_dev); break;
    default: r = ndft_adjoint(c, in_dev, out_dev); break;
  }
  if (r != NFFTCU_OK) return r;
  if (out_bytes) NFFTCU_CUDA(cudaMemcpyAsync(out_host, out_dev, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  return NFFTCU_OK;
}

// Transform with an unannounced node refresh (plans without a psi flag: the reference re-reads x on every call,
// nfft.c:4889, 5351).  The transform is launched at once with the resident nodes; meanwhile host threads fingerprint
// the caller's x (no PCIe traffic, no device work).  Only if the fingerprint differs from that of the array the
// resident nodes came from are the nodes uploaded and re-sorted and the transform repeated.  Unchanged nodes -- every
// call of a solver loop -- cost nothing on the device or the host link.  Resident nodes without a host fingerprint
// (set through nfftcu_set_nodes_dev) or NFFT_B200_EXACT_NODE_CHECK=1 use the exact variant instead: x is uploaded on a
// side stream and compared word for word on the device while the transform runs.
static int host_transform_refresh(nfftcu_ctx *c, const void *x_host, const void *in_host, void *out_host,
                                  int which, int *changed_out) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  if (changed_out) *changed_out = 0;
  const size_t xbytes = real_size(c) * (size_t) c->M * c->d;
  if (!c->have_nodes || xbytes == 0 || c->direct_only) {
    const int64_t before = c->nodes_version;
    NFFTCU_TRY(nfftcu_set_nodes(c, x_host));
    if (changed_out) *changed_out = c->nodes_version != before;
    return host_transform(c, in_host, out_host, which);
  }
  if (!x_host) {
    set_error("transform: x is NULL");
    return NFFTCU_EINVAL;
  }
  NFFTCU_TRY(ensure_staging(c));
  const bool forward = which == 0;
  const size_t in_bytes = forward ? cbytes(c, c->N_total) : cbytes(c, c->M);
  const size_t out_bytes = forward ? cbytes(c, c->M) : cbytes(c, c->N_total);
  void *in_dev = forward ? c->fhat_dev : c->f_dev;
  void *out_dev = forward ? c->f_dev : c->fhat_dev;
  const bool by_fingerprint = c->x_fp_valid && !exact_node_check();
  if (in_bytes) NFFTCU_CUDA(cudaMemcpyAsync(in_dev, in_host, in_bytes, cudaMemcpyHostToDevice, c->stream));
  if (!by_fingerprint) {
    if (!c->side_stream) NFFTCU_CUDA(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    if (!c->h_flag) NFFTCU_CUDA(pool_malloc_host((void **) &c->h_flag, sizeof(int)));
    if (!c->x_stage) NFFTCU_CUDA(pool_malloc(&c->x_stage, xbytes));
    if (!c->diff_flag) NFFTCU_CUDA(pool_malloc((void **) &c->diff_flag, sizeof(int)));
    // the side stream must not start before earlier work on the plan's stream that may still read x_stage, and not
    // before the transform's own input is on the device: the node upload would share the host link with it
    if (!c->ev_side) NFFTCU_CUDA(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
    NFFTCU_CUDA(cudaEventRecord(c->ev_side, c->stream));
    NFFTCU_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_side, 0));
    NFFTCU_CUDA(cudaMemcpyAsync(c->x_stage, x_host, xbytes, cudaMemcpyHostToDevice, c->side_stream));
    NFFTCU_CUDA(cudaMemsetAsync(c->diff_flag, 0, sizeof(int), c->side_stream));
    const long long words = (long long) (xbytes / 4);
    long long blocks = (words + 255) / 256;
    if (blocks > (long long) c->sm_count * 16) blocks = (long long) c->sm_count * 16;
    differs_kernel<<<(unsigned) blocks, 256, 0, c->side_stream>>>((const uint32_t *) c->x_stage,
                                                                (const uint32_t *) c->x_dev, words, c->diff_flag);
    c->launches++;
    NFFTCU_CUDA(cudaMemcpyAsync(c->h_flag, c->diff_flag, sizeof(int), cudaMemcpyDeviceToHost, c->side_stream));
  }
  int r = forward ? trafo_dev_impl(c, in_dev, out_dev) : adjoint_dev_impl(c, in_dev, out_dev);
  if (r != NFFTCU_OK) return r;
  bool changed;
  uint64_t fp = 0;
  if (by_fingerprint) {
    fp = fingerprint(x_host, xbytes);   // host threads; the device is busy with the transform meanwhile
    changed = fp != c->x_fp;
  } else {
    NFFTCU_CUDA(cudaStreamSynchronize(c->side_stream));
    changed = *c->h_flag != 0;
  }
  if (changed) {   // the nodes did change: adopt them and redo the transform (its input is still on the device)
    NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
    if (by_fingerprint) {
      if (!c->x_stage) NFFTCU_CUDA(pool_malloc(&c->x_stage, xbytes));
      NFFTCU_CUDA(cudaMemcpyAsync(c->x_stage, x_host, xbytes, cudaMemcpyHostToDevice, c->stream));
      c->x_fp = fp;
    } else {
      c->x_fp_valid = false;
    }
    void *t = c->x_dev; c->x_dev = c->x_stage; c->x_stage = t;
    NFFTCU_TRY(nodes_ready(c));
    if (changed_out) *changed_out = 1;
    r = forward ? trafo_dev_impl(c, in_dev, out_dev) : adjoint_dev_impl(c, in_dev, out_dev);
    if (r != NFFTCU_OK) return r;
  }
  if (out_bytes) NFFTCU_CUDA(cudaMemcpyAsync(out_host, out_dev, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  return NFFTCU_OK;
}

int nfftcu_trafo_refresh(nfftcu_ctx *c, const void *x_host, const void *f_hat_host, void *f_host, int *changed) {
  return host_transform_refresh(c, x_host, f_hat_host, f_host, 0, changed);
}
int nfftcu_adjoint_refresh(nfftcu_ctx *c, const void *x_host, const void *f_host, void *f_hat_host, int *changed) {
  return host_transform_refresh(c, x_host, f_host, f_hat_host, 1, changed);
}

// Split-phase host-pointer transforms: *_begin enqueues H2D, the transform and D2H on the plan's stream and returns;
// nfftcu_end waits.  Two plans (or one plan and the caller's own work) can then overlap their copies with each other's
// kernels -- the synchronous nfft_trafo / nfft_adjoint cannot, the reference API has no such notion.  The host buffers
// must be page-locked (nfft_malloc'ed) for the copies to be asynchronous and must stay untouched until nfftcu_end;
// the resident nodes are used as they are (no refresh, like a plan with PRE_PSI).
static int host_begin(nfftcu_ctx *c, const void *in_host, void *out_host, bool forward) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(need_nodes(c));
  NFFTCU_TRY(ensure_staging(c));
  const size_t in_bytes = forward ? cbytes(c, c->N_total) : cbytes(c, c->M);
  const size_t out_bytes = forward ? cbytes(c, c->M) : cbytes(c, c->N_total);
  void *in_dev = forward ? c->fhat_dev : c->f_dev, *out_dev = forward ? c->f_dev : c->fhat_dev;
  if (in_bytes) NFFTCU_CUDA(cudaMemcpyAsync(in_dev, in_host, in_bytes, cudaMemcpyHostToDevice, c->stream));
  const int timing = c->opt_timing;
  c->opt_timing = 0;   // the stage timers synchronise
  const int r = forward ? trafo_dev_impl(c, in_dev, out_dev) : adjoint_dev_impl(c, in_dev, out_dev);
  c->opt_timing = timing;
  if (r != NFFTCU_OK) return r;
  if (out_bytes) NFFTCU_CUDA(cudaMemcpyAsync(out_host, out_dev, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  return NFFTCU_OK;
}
int nfftcu_trafo_begin(nfftcu_ctx *c, const void *f_hat_host, void *f_host) { return host_begin(c, f_hat_host, f_host, true); }
int nfftcu_adjoint_begin(nfftcu_ctx *c, const void *f_host, void *f_hat_host) { return host_begin(c, f_host, f_hat_host, false); }
int nfftcu_end(nfftcu_ctx *c) { return nfftcu_sync(c); }

int nfftcu_trafo(nfftcu_ctx *c, const void *f_hat_host, void *f_host) {
  return host_transform(c, f_hat_host, f_host, 0);
}
int nfftcu_adjoint(nfftcu_ctx *c, const void *f_host, void *f_hat_host) {
  return host_transform(c, f_host, f_hat_host, 1);
}
int nfftcu_trafo_direct(nfftcu_ctx *c, const void *f_hat_host, void *f_host) {
  return host_transform(c, f_hat_host, f_host, 2);
}
int nfftcu_adjoint_direct(nfftcu_ctx *c, const void *f_host, void *f_hat_host) {
  return host_transform(c, f_host, f_hat_host, 3);
}

// ---- K right-hand sides on ONE node set -----------------------------------------------------------
// The node-dependent state -- sort, tile binning, chunk lists, window images / psi tables -- is built once and
// serves every right-hand side; D, the FFT passes, the 2-D tile kernels and D^T take the right-hand side as one more
// grid dimension (one launch for all K), the 3-D kernel families walk the batch launch by launch.
int nfftcu_trafo_batch_dev(nfftcu_ctx *c, int K, const void *f_hat_dev, void *f_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(ensure_batch(c, K));
  c->cur_batch = K;
  const int r = trafo_dev_impl(c, f_hat_dev, f_dev);
  c->cur_batch = 1;
  return r;
}
int nfftcu_adjoint_batch_dev(nfftcu_ctx *c, int K, const void *f_dev, void *f_hat_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(ensure_batch(c, K));
  c->cur_batch = K;
  const int r = adjoint_dev_impl(c, f_dev, f_hat_dev);
  c->cur_batch = 1;
  return r;
}
static int host_batch(nfftcu_ctx *c, int K, const void *in_host, void *out_host, bool forward) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(need_nodes(c));
  NFFTCU_TRY(ensure_batch(c, K));
  NFFTCU_TRY(ensure_staging(c));
  const size_t in_bytes = (forward ? cbytes(c, c->N_total) : cbytes(c, c->M)) * (size_t) K;
  const size_t out_bytes = (forward ? cbytes(c, c->M) : cbytes(c, c->N_total)) * (size_t) K;
  void *in_dev = forward ? c->fhat_dev : c->f_dev, *out_dev = forward ? c->f_dev : c->fhat_dev;
  if (in_bytes) NFFTCU_CUDA(cudaMemcpyAsync(in_dev, in_host, in_bytes, cudaMemcpyHostToDevice, c->stream));
  c->cur_batch = K;
  const int r = forward ? trafo_dev_impl(c, in_dev, out_dev) : adjoint_dev_impl(c, in_dev, out_dev);
  c->cur_batch = 1;
  if (r != NFFTCU_OK) return r;
  if (out_bytes) NFFTCU_CUDA(cudaMemcpyAsync(out_host, out_dev, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  return NFFTCU_OK;
}
int nfftcu_trafo_batch(nfftcu_ctx *c, int K, const void *f_hat_host, void *f_host) {
  return host_batch(c, K, f_hat_host, f_host, true);
}
int nfftcu_adjoint_batch(nfftcu_ctx *c, int K, const void *f_host, void *f_hat_host) {
  return host_batch(c, K, f_host, f_hat_host, false);
}

// ---- adjoint -> diagonal multiply -> trafo without leaving the device (fastsum far field) ----------
int nfftcu_adjoint_mul_trafo(nfftcu_ctx *src, nfftcu_ctx *dst, const void *f_src_host, const void *b_host,
                             void *f_dst_host) {
  NFFTCU_TRY(check_ctx(src));
  NFFTCU_TRY(check_ctx(dst));
  if (src->prec != dst->prec || src->N_total != dst->N_total || src->device != dst->device) {
    set_error("nfftcu_adjoint_mul_trafo: the two plans must share precision, bandwidths and device");
    return NFFTCU_EINVAL;
  }
  NFFTCU_TRY(bind_device(src));
  NFFTCU_TRY(need_nodes(src));
  NFFTCU_TRY(need_nodes(dst));
  NFFTCU_TRY(ensure_staging(src));
  NFFTCU_TRY(ensure_staging(dst));
  const size_t nb = cbytes(src, src->N_total);
  if (src->M) NFFTCU_CUDA(cudaMemcpyAsync(src->f_dev, f_src_host, cbytes(src, src->M), cudaMemcpyHostToDevice, src->stream));
  if (b_host) NFFTCU_CUDA(cudaMemcpyAsync(dst->fhat_dev, b_host, nb, cudaMemcpyHostToDevice, src->stream));   // b parks in dst's staging
  NFFTCU_TRY(adjoint_dev_impl(src, src->f_dev, src->fhat_dev));
  if (b_host) {
    long long blocks = (src->N_total + 255) / 256;
    if (blocks > (long long) src->sm_count * 16) blocks = (long long) src->sm_count * 16;
    if (src->prec == NFFTCU_DOUBLE)
      cmul_diag_kernel<double2><<<(unsigned) blocks, 256, 0, src->stream>>>((double2 *) src->fhat_dev, (const double2 *) dst->fhat_dev, src->N_total);
    else
      cmul_diag_kernel<float2><<<(unsigned) blocks, 256, 0, src->stream>>>((float2 *) src->fhat_dev, (const float2 *) dst->fhat_dev, src->N_total);
    src->launches++;
    NFFTCU_CUDA(cudaGetLastError());
  }
  // hand over from the source plan's stream to the target plan's
  if (!src->ev_side) NFFTCU_CUDA(cudaEventCreateWithFlags(&src->ev_side, cudaEventDisableTiming));
  NFFTCU_CUDA(cudaEventRecord(src->ev_side, src->stream));
  NFFTCU_CUDA(cudaStreamWaitEvent(dst->stream, src->ev_side, 0));
  NFFTCU_TRY(trafo_dev_impl(dst, src->fhat_dev, dst->f_dev));
  if (dst->M) NFFTCU_CUDA(cudaMemcpyAsync(f_dst_host, dst->f_dev, cbytes(dst, dst->M), cudaMemcpyDeviceToHost, dst->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(dst->stream));
  return NFFTCU_OK;
}

int nfftcu_trafo_dev(nfftcu_ctx *c, const void *f_hat_dev, void *f_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  return trafo_dev_impl(c, f_hat_dev, f_dev);
}
int nfftcu_adjoint_dev(nfftcu_ctx *c, const void *f_dev, void *f_hat_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  return adjoint_dev_impl(c, f_dev, f_hat_dev);
}
// f_hat := sum over ranks of D^T F^H B_r^T f_r, every rank gets the full result (the collective is part of D^T)
int nfftcu_adjoint_dev_peer(nfftcu_ctx *c, const void *f_dev, void *f_hat_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  if (!c->peer) { set_error("nfftcu_adjoint_dev_peer: plan is not attached to its peers"); return NFFTCU_ESTATE; }
  if (c->direct_only) { set_error("nfftcu_adjoint_dev_peer: NDFT-fallback plan"); return NFFTCU_ESTATE; }
  return adjoint_dev_impl(c, f_dev, f_hat_dev, true);
}
int nfftcu_trafo_direct_dev(nfftcu_ctx *c, const void *f_hat_dev, void *f_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(need_nodes(c));
  return ndft_trafo(c, f_hat_dev, f_dev);
}
int nfftcu_adjoint_direct_dev(nfftcu_ctx *c, const void *f_dev, void *f_hat_dev) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_TRY(need_nodes(c));
  return ndft_adjoint(c, f_dev, f_hat_dev);
}

static int need_grid(nfftcu_ctx *c) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  if (c->direct_only) {
    set_error("stage call on a direct-only plan (some N_t <= m or n_t <= 2m+2): no grid");
    return NFFTCU_ESTATE;
  }
  return NFFTCU_OK;
}

int nfftcu_stage_D(nfftcu_ctx *c, const void *f_hat_dev) {
  NFFTCU_TRY(need_grid(c));
  return stage_D(c, f_hat_dev);
}
int nfftcu_stage_F(nfftcu_ctx *c, int sign) {
  NFFTCU_TRY(need_grid(c));
  return stage_F(c, sign < 0 ? -1 : +1);
}
int nfftcu_stage_B(nfftcu_ctx *c, void *f_dev) {
  NFFTCU_TRY(need_grid(c));
  NFFTCU_TRY(need_nodes(c));
  return stage_B(c, f_dev);
}
int nfftcu_stage_BT(nfftcu_ctx *c, const void *f_dev) {
  NFFTCU_TRY(need_grid(c));
  NFFTCU_TRY(need_nodes(c));
  return stage_BT(c, f_dev);
}
int nfftcu_stage_DT(nfftcu_ctx *c, void *f_hat_dev) {
  NFFTCU_TRY(need_grid(c));
  return stage_DT(c, f_hat_dev);
}
void *nfftcu_grid_ptr(nfftcu_ctx *c) { return c ? c->grid : nullptr; }

int nfftcu_set_option(nfftcu_ctx *c, int option, int64_t value) {
  NFFTCU_TRY(check_ctx(c));
  switch (option) {
    case NFFTCU_OPT_TIMING: c->opt_timing = (int) value; break;
    case NFFTCU_OPT_PSI_TABLE: c->opt_psi_table = (int) value; break;
    case NFFTCU_OPT_B_KERNEL: c->opt_b_kernel = (int) value; break;
    case NFFTCU_OPT_NODE_ORDER: c->opt_node_order = (int) value; break;
    case NFFTCU_OPT_B_FLUSH: c->opt_b_flush = (int) value; break;
    case NFFTCU_OPT_FFT_PRUNE: c->opt_fft_prune = (int) value; break;
    case NFFTCU_OPT_FFT_KERNEL: c->opt_fft_kernel = (int) value; break;
    case NFFTCU_OPT_WINDOW_IMAGES: c->opt_window_images = (int) value; break;
    case NFFTCU_OPT_SLAB_FFT: c->opt_slab = (int) value; break;
    case NFFTCU_OPT_TC5: c->opt_tc5 = (int) value; break;
    default:
      set_error("nfftcu_set_option: unknown option %d", option);
      return NFFTCU_EINVAL;
  }
  return NFFTCU_OK;
}

int nfftcu_set_stream(nfftcu_ctx *c, void *cuda_stream) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  if (c->stream) NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (cuda_stream) {
    c->stream = (cudaStream_t) cuda_stream;
    c->own_stream = false;
  } else {
    NFFTCU_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  return NFFTCU_OK;
}

void *nfftcu_get_stream(nfftcu_ctx *c) { return c ? (void *) c->stream : nullptr; }

int nfftcu_sync(nfftcu_ctx *c) {
  NFFTCU_TRY(check_ctx(c));
  NFFTCU_TRY(bind_device(c));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  return NFFTCU_OK;
}

int nfftcu_stage_times(nfftcu_ctx *c, float ms[3]) {
  NFFTCU_TRY(check_ctx(c));
  for (int i = 0; i < 3; i++) ms[i] = c->stage_ms[i];
  return NFFTCU_OK;
}

int nfftcu_b_kernel_time(nfftcu_ctx *c, float *ms) {
  NFFTCU_TRY(check_ctx(c));
  *ms = c->bkernel_ms;
  return NFFTCU_OK;
}

int64_t nfftcu_launch_count(nfftcu_ctx *c) { return c ? c->launches : 0; }

int nfftcu_malloc_device(void **ptr, size_t bytes, int device) {
  NFFTCU_CUDA(cudaSetDevice(device));
  NFFTCU_CUDA(pool_malloc(ptr, bytes ? bytes : 1));
  return NFFTCU_OK;
}
int nfftcu_free_device(void *ptr) {
  if (ptr) NFFTCU_CUDA(pool_free(ptr));
  return NFFTCU_OK;
}
int nfftcu_malloc_pinned(void **ptr, size_t bytes) {
  NFFTCU_CUDA(pool_malloc_host(ptr, bytes ? bytes : 1));
  return NFFTCU_OK;
}
int nfftcu_free_pinned(void *ptr) {
  if (ptr) NFFTCU_CUDA(pool_free_host(ptr));
  return NFFTCU_OK;
}
// nfft_malloc / nfft_free backing store (nfft3_host.c): large host buffers of the plan API (MALLOC_X / MALLOC_F_HAT /
// MALLOC_F, index_x, and whatever callers allocate through nfft_malloc, e.g. kernel/mri/mri.c:94-96) are page-locked
// so that the host-pointer transforms copy at the full link rate; small ones and the no-device case use the C heap.
void *nfftcu_host_alloc(size_t bytes) {
  static const size_t threshold = [] {
    const char *e = getenv("NFFT_B200_PINNED_MALLOC_MIN");   // bytes; 0 disables page-locking
    return e ? (size_t) atoll(e) : ((size_t) 1 << 18);
  }();
  void *p = nullptr;
  if (bytes == 0) bytes = 1;
  if (threshold > 0 && bytes >= threshold && nfftcu_device_count() > 0) {
    if (pool_malloc_host(&p, bytes) == cudaSuccess && p) return p;
    cudaGetLastError();
    p = nullptr;
  }
  if (posix_memalign(&p, 64, bytes) != 0) return nullptr;
  return p;
}
void nfftcu_host_free(void *p) {
  if (!p) return;
  if (pool_owns_host(p)) pool_free_host(p);
  else free(p);
}
void nfftcu_pool_trim(void) { plan_cache_clear(); pool_trim(); }
// the 64-bit fingerprint nfftcu_set_nodes / nfftcu_*_refresh use to detect a changed node array (host only)
uint64_t nfftcu_fingerprint(const void *data, size_t bytes) { return fingerprint(data, bytes); }

int nfftcu_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes) {
  NFFTCU_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
  return NFFTCU_OK;
}
int nfftcu_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes) {
  NFFTCU_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return NFFTCU_OK;
}

}  // extern "C"
