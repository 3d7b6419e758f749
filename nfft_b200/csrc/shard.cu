// shard.cu -- nfftcu_group_*: ONE process, P GPUs of one NVSwitch domain behind the host-pointer interface of
// nfft_trafo / nfft_adjoint (north_star configs[3]: N=256^3, M=1e8 node-sharded over 1/2/4/8 B200).
//
//   set_nodes   x -> device 0, reference sort (sort.cu: the same key and stable order as nfft.c:75-123, so index_x is
//               still the reference's), sorted order cut into P equal-count slabs; slab r and its slice of the
//               permutation go to device r, whose plan bins them for the tensor kernels as usual.  Every device
//               then works on a compact slab of the grid at the FULL node density (the tensor kernels' batches
//               stay full, which a split by caller index would not give).
//   trafo       f_hat -> every device (each over its own host link, concurrently); D + F replicated; B on the slab;
//               results leave slab order through an all-to-all over peer memory: device r stores f[perm[k]] into
//               the caller-order chunk of the device that owns index perm[k]; every device copies its chunk
//               (M/P samples) to the host over its own link.
//   adjoint     chunk q of f -> device q; device r pulls f[perm[k]] for its slab from the owners (peer loads);
//               B^T + F on the slab; fused D^T + reduce-scatter over peer memory (peer.cu): device r produces slice
//               r of f_hat = c * sum_p grid_p and copies it to the host.  No NCCL call, no partial f_hat in HBM.
//
// Cross-device ordering inside one transform: CUDA events recorded / waited on the per-device plan streams by the
// single host thread, plus the flag barriers of peer.cu around the fused reduce.  Everything that can block the host
// (copies from / to pageable memory) is issued only after all kernels of the phase are enqueued on all devices.
#include "common.cuh"

#include <string.h>

#include <thread>

struct nfftcu_group_s {
  int P = 0, prec = 0, d = 0;
  int64_t M = 0, N_total = 0, L = 0;      // L: caller-order chunk length, chunk q = [qL, min(M, (q+1)L))
  std::vector<int> dev;
  nfftcu_ctx *sorter = nullptr;           // device 0, nodes only
  std::vector<nfftcu_ctx *> shard;
  std::vector<int64_t> begin;             // P+1 offsets into the sorted order
  std::vector<uint32_t *> perm;           // slab r: sorted position -> original index
  std::vector<void *> f_slab, f_chunk, fhat;
  std::vector<cudaEvent_t> ev, ev_t;      // ordering; timing (4 per device)
  int64_t nodes_version = 0;
  float ms[3] = {0.f, 0.f, 0.f};
};

namespace nfftcu {
namespace {

struct ChunkPtrs { void *p[NFFTCU_MAX_PEERS]; };

// slab order -> caller order, all-to-all over peer memory (trafo) ...
template <typename C>
__global__ void slab_scatter_kernel(const C *__restrict__ f_slab, const uint32_t *__restrict__ perm, long long count,
                                    ChunkPtrs chunk, long long L) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const long long j = perm[k];
    const long long q = j / L;
    ((C *) chunk.p[q])[j - q * L] = f_slab[k];
  }
  __threadfence_system();
}
// ... and caller order -> slab order (adjoint)
template <typename C>
__global__ void slab_gather_kernel(C *__restrict__ f_slab, const uint32_t *__restrict__ perm, long long count,
                                   ChunkPtrs chunk, long long L) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const long long j = perm[k];
    const long long q = j / L;
    f_slab[k] = __ldcv((const C *) chunk.p[q] + (j - q * L));
  }
}

size_t csize(const nfftcu_group_s *g) { return g->prec == NFFTCU_DOUBLE ? 16 : 8; }

int64_t chunk_begin(const nfftcu_group_s *g, int q) { const int64_t b = (int64_t) q * g->L; return b < g->M ? b : g->M; }

void all_wait_all(nfftcu_group_s *g) {   // every device's stream waits for what every device has enqueued so far
  for (int r = 0; r < g->P; r++) {
    cudaSetDevice(g->dev[r]);
    cudaEventRecord(g->ev[r], g->shard[r]->stream);
  }
  for (int r = 0; r < g->P; r++) {
    cudaSetDevice(g->dev[r]);
    for (int q = 0; q < g->P; q++)
      if (q != r) cudaStreamWaitEvent(g->shard[r]->stream, g->ev[q], 0);
  }
}

void mark(nfftcu_group_s *g, int r, int i) { cudaEventRecord(g->ev_t[4 * r + i], g->shard[r]->stream); }

int sync_all(nfftcu_group_s *g) {
  for (int r = 0; r < g->P; r++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
    NFFTCU_CUDA(cudaStreamSynchronize(g->shard[r]->stream));
  }
  for (int i = 0; i < 3; i++) g->ms[i] = 0.f;
  for (int r = 0; r < g->P; r++)
    for (int i = 0; i < 3; i++) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, g->ev_t[4 * r + i], g->ev_t[4 * r + i + 1]) == cudaSuccess && t > g->ms[i]) g->ms[i] = t;
    }
  cudaGetLastError();
  return NFFTCU_OK;
}

ChunkPtrs chunks_of(const nfftcu_group_s *g) {
  ChunkPtrs cp;
  memset(&cp, 0, sizeof(cp));
  for (int q = 0; q < g->P; q++) cp.p[q] = g->f_chunk[q];
  return cp;
}

int launch_blocks(const nfftcu_ctx *c, long long count) {
  long long b = (count + 255) / 256;
  const long long cap = (long long) c->sm_count * 16;
  if (b > cap) b = cap;
  return (int) (b < 1 ? 1 : b);
}

int group_trafo(nfftcu_group_s *g, const void *f_hat_host, void *f_host) {
  const size_t C = csize(g);
  // phase 1: inputs (may block on pageable memory: nothing that could wait for another device is queued yet).
  // f_hat crosses the host link ONCE: device r uploads slice r and the devices all-gather the slices over NVLink
  // (P - 1 peer copies of N_total / P coefficients into every device); uploading the whole f_hat to every device
  // made the 8-GPU trafo host-link-bound (8 x 268 MB at cfg4 against ~95 GB/s of aggregate host bandwidth).
  for (int r = 0; r < g->P; r++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
    mark(g, r, 0);
    const int64_t kb = g->N_total * r / g->P, ke = g->N_total * (r + 1) / g->P;
    if (ke > kb)
      NFFTCU_CUDA(cudaMemcpyAsync((char *) g->fhat[r] + C * (size_t) kb, (const char *) f_hat_host + C * (size_t) kb,
                                  C * (size_t) (ke - kb), cudaMemcpyHostToDevice, g->shard[r]->stream));
  }
  if (g->P > 1) {
    all_wait_all(g);
    for (int r = 0; r < g->P; r++) {
      NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
      for (int q = 1; q < g->P; q++) {   // staggered start: device r first pulls from r + 1
        const int s = (r + q) % g->P;
        const int64_t kb = g->N_total * s / g->P, ke = g->N_total * (s + 1) / g->P;
        if (ke > kb)
          NFFTCU_CUDA(cudaMemcpyAsync((char *) g->fhat[r] + C * (size_t) kb, (const char *) g->fhat[s] + C * (size_t) kb,
                                      C * (size_t) (ke - kb), cudaMemcpyDefault, g->shard[r]->stream));
      }
    }
  }
  for (int r = 0; r < g->P; r++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
    mark(g, r, 1);
  }
  // phase 2: D + F + B per device, then the all-to-all into caller order
  const ChunkPtrs cp = chunks_of(g);
  for (int r = 0; r < g->P; r++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
    nfftcu_ctx *c = g->shard[r];
    const long long cnt = g->begin[r + 1] - g->begin[r];
    NFFTCU_TRY(nfftcu_trafo_dev(c, g->fhat[r], g->f_slab[r]));
    if (cnt > 0) {
      if (g->prec == NFFTCU_DOUBLE)
        slab_scatter_kernel<double2><<<launch_blocks(c, cnt), 256, 0, c->stream>>>((const double2 *) g->f_slab[r], g->perm[r], cnt, cp, g->L);
      else
        slab_scatter_kernel<float2><<<launch_blocks(c, cnt), 256, 0, c->stream>>>((const float2 *) g->f_slab[r], g->perm[r], cnt, cp, g->L);
      c->launches++;
      NFFTCU_CUDA(cudaGetLastError());
    }
  }
  all_wait_all(g);
  // phase 3: every device delivers its caller-order chunk over its own host link
  for (int q = 0; q < g->P; q++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[q]));
    mark(g, q, 2);
    const int64_t b = chunk_begin(g, q), e = chunk_begin(g, q + 1);
    if (e > b)
      NFFTCU_CUDA(cudaMemcpyAsync((char *) f_host + C * (size_t) b, g->f_chunk[q], C * (size_t) (e - b), cudaMemcpyDeviceToHost,
                                  g->shard[q]->stream));
    mark(g, q, 3);
  }
  return sync_all(g);
}

int group_adjoint(nfftcu_group_s *g, const void *f_host, void *f_hat_host) {
  const size_t C = csize(g);
  for (int q = 0; q < g->P; q++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[q]));
    mark(g, q, 0);
    const int64_t b = chunk_begin(g, q), e = chunk_begin(g, q + 1);
    if (e > b)
      NFFTCU_CUDA(cudaMemcpyAsync(g->f_chunk[q], (const char *) f_host + C * (size_t) b, C * (size_t) (e - b), cudaMemcpyHostToDevice,
                                  g->shard[q]->stream));
    mark(g, q, 1);
  }
  all_wait_all(g);
  const ChunkPtrs cp = chunks_of(g);
  for (int r = 0; r < g->P; r++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
    nfftcu_ctx *c = g->shard[r];
    const long long cnt = g->begin[r + 1] - g->begin[r];
    if (cnt > 0) {
      if (g->prec == NFFTCU_DOUBLE)
        slab_gather_kernel<double2><<<launch_blocks(c, cnt), 256, 0, c->stream>>>((double2 *) g->f_slab[r], g->perm[r], cnt, cp, g->L);
      else
        slab_gather_kernel<float2><<<launch_blocks(c, cnt), 256, 0, c->stream>>>((float2 *) g->f_slab[r], g->perm[r], cnt, cp, g->L);
      c->launches++;
      NFFTCU_CUDA(cudaGetLastError());
    }
    const bool pruned = c->opt_fft_prune != 0 && !c->fft_no_prune;
    NFFTCU_TRY(stage_BT(c, g->f_slab[r]));
    NFFTCU_TRY(stage_F(c, +1, pruned));
    NFFTCU_TRY(peer_reduce_DT(c, nullptr));   // fused D^T + reduce-scatter: slice r of f_hat lands in the exchange buffer
  }
  for (int r = 0; r < g->P; r++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
    mark(g, r, 2);
    long long kb = 0, ke = 0;
    void *src = peer_slice_ptr(g->shard[r], &kb, &ke);
    if (ke > kb)
      NFFTCU_CUDA(cudaMemcpyAsync((char *) f_hat_host + C * (size_t) kb, src, C * (size_t) (ke - kb), cudaMemcpyDeviceToHost,
                                  g->shard[r]->stream));
    mark(g, r, 3);
  }
  return sync_all(g);
}

int group_refresh(nfftcu_group_s *g, const void *x_host, const void *in, void *out, bool forward, int *changed) {
  if (changed) *changed = 0;
  if (g->nodes_version == 0 || !g->sorter->x_fp_valid) {
    const int64_t before = g->nodes_version;
    NFFTCU_TRY(nfftcu_group_set_nodes(g, x_host));
    if (changed) *changed = g->nodes_version != before;
    return forward ? group_trafo(g, in, out) : group_adjoint(g, in, out);
  }
  // the devices run with the resident slabs while the host fingerprints x; a mismatch repeats the transform
  // (the fingerprint is taken after the transform has been enqueued and before the host waits for it)
  struct Hash { const void *x; size_t bytes; uint64_t fp; } h = {x_host, (size_t) (g->prec == NFFTCU_DOUBLE ? 8 : 4) * (size_t) g->M * g->d, 0};
  // the group transforms synchronise internally, so hash first on a helper thread and join afterwards
  std::thread th([&h] { h.fp = fingerprint(h.x, h.bytes); });
  int r = forward ? group_trafo(g, in, out) : group_adjoint(g, in, out);
  th.join();
  if (r != NFFTCU_OK) return r;
  if (h.fp != g->sorter->x_fp) {
    NFFTCU_TRY(nfftcu_group_set_nodes(g, x_host));
    if (changed) *changed = 1;
    r = forward ? group_trafo(g, in, out) : group_adjoint(g, in, out);
  }
  return r;
}

}  // namespace
}  // namespace nfftcu

using namespace nfftcu;

extern "C" {

int nfftcu_group_create(nfftcu_group **out, int precision, int d, const int64_t *N, const int64_t *n, int64_t m,
                        int64_t M, unsigned flags, const int *devices, int ndevices) {
  if (!out || !devices || ndevices < 1 || ndevices > NFFTCU_MAX_PEERS) {
    set_error("nfftcu_group_create: 1..%d devices expected, got %d", NFFTCU_MAX_PEERS, ndevices);
    return NFFTCU_EINVAL;
  }
  nfftcu_group_s *g = new nfftcu_group_s();
  g->P = ndevices;
  g->prec = precision;
  g->d = d;
  g->M = M;
  g->L = (M + ndevices - 1) / ndevices;
  if (g->L < 1) g->L = 1;
  g->dev.assign(devices, devices + ndevices);
  g->shard.assign(ndevices, nullptr);
  g->perm.assign(ndevices, nullptr);
  g->f_slab.assign(ndevices, nullptr);
  g->f_chunk.assign(ndevices, nullptr);
  g->fhat.assign(ndevices, nullptr);
  g->ev.assign(ndevices, nullptr);
  g->ev_t.assign(4 * ndevices, nullptr);
  g->begin.resize(ndevices + 1);
  for (int r = 0; r <= ndevices; r++) g->begin[r] = M / ndevices * r + (r < M % ndevices ? r : M % ndevices);
  const int status = [&]() -> int {
    for (int r = 0; r < ndevices; r++)
      for (int q = 0; q < r; q++)
        if (devices[q] == devices[r]) {
          set_error("nfftcu_group_create: device %d listed twice", devices[r]);
          return NFFTCU_EINVAL;
        }
    NFFTCU_TRY(create_ctx(&g->sorter, precision, d, N, n, m, M, flags | (1u << 11), devices[0], true));
    const size_t C = csize(g);
    for (int r = 0; r < ndevices; r++) {
      const int64_t Mr = g->begin[r + 1] - g->begin[r];
      // shards never need the reference order themselves: strip NFFT_SORT_NODES (bit 11) / BLOCKWISE (bit 12)
      NFFTCU_TRY(create_ctx(&g->shard[r], precision, d, N, n, m, Mr, flags & ~((1u << 11) | (1u << 12)), devices[r], false));
      nfftcu_ctx *c = g->shard[r];
      if (c->direct_only || c->grid2) {
        set_error("nfftcu_group_create: multi-GPU plans need a grid plan (N_t > m) without a split FFT axis");
        return NFFTCU_EINVAL;
      }
      g->N_total = c->N_total;
      NFFTCU_CUDA(cudaSetDevice(devices[r]));
      NFFTCU_CUDA(pool_malloc((void **) &g->perm[r], sizeof(uint32_t) * (size_t) (Mr > 0 ? Mr : 1)));
      NFFTCU_CUDA(pool_malloc(&g->f_slab[r], C * (size_t) (Mr > 0 ? Mr : 1)));
      NFFTCU_CUDA(pool_malloc(&g->f_chunk[r], C * (size_t) g->L));
      NFFTCU_CUDA(pool_malloc(&g->fhat[r], C * (size_t) c->N_total));
      NFFTCU_CUDA(cudaEventCreateWithFlags(&g->ev[r], cudaEventDisableTiming));
      for (int i = 0; i < 4; i++) NFFTCU_CUDA(cudaEventCreate(&g->ev_t[4 * r + i]));
    }
    return peer_attach_local(g->shard.data(), ndevices, /*all_outputs=*/false);
  }();
  if (status != NFFTCU_OK) {
    const std::string msg = nfftcu_last_error();
    nfftcu_group_destroy(g);
    set_error("%s", msg.c_str());
    return status;
  }
  *out = g;
  return NFFTCU_OK;
}

int nfftcu_group_destroy(nfftcu_group *g) {
  if (!g) return NFFTCU_OK;
  for (int r = 0; r < g->P; r++) {
    cudaSetDevice(g->dev[r]);
    if (g->shard[r] && g->shard[r]->stream) cudaStreamSynchronize(g->shard[r]->stream);
  }
  for (int r = 0; r < g->P; r++) {
    cudaSetDevice(g->dev[r]);
    void *bufs[] = {(void *) g->perm[r], g->f_slab[r], g->f_chunk[r], g->fhat[r]};
    for (void *p : bufs)
      if (p) pool_free(p);
    if (g->ev[r]) cudaEventDestroy(g->ev[r]);
    for (int i = 0; i < 4; i++)
      if (g->ev_t[4 * r + i]) cudaEventDestroy(g->ev_t[4 * r + i]);
  }
  for (int r = 0; r < g->P; r++) nfftcu_destroy(g->shard[r]);
  nfftcu_destroy(g->sorter);
  delete g;
  return NFFTCU_OK;
}

int nfftcu_group_set_nodes(nfftcu_group *g, const void *x_host) {
  if (!g) { set_error("null group"); return NFFTCU_EINVAL; }
  const int64_t before = nfftcu_nodes_version(g->sorter);
  NFFTCU_TRY(nfftcu_set_nodes(g->sorter, x_host));   // upload + reference sort on device 0 (fingerprint: no-op when unchanged)
  if (nfftcu_nodes_version(g->sorter) == before && g->nodes_version != 0) return NFFTCU_OK;
  const size_t rs = (size_t) (g->prec == NFFTCU_DOUBLE ? 8 : 4) * (size_t) g->d;
  for (int r = 0; r < g->P; r++) {
    const int64_t b = g->begin[r], e = g->begin[r + 1];
    NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
    if (e > b)
      NFFTCU_CUDA(cudaMemcpyAsync(g->perm[r], g->sorter->perm_ref + b, sizeof(uint32_t) * (size_t) (e - b), cudaMemcpyDefault,
                                  g->shard[r]->stream));
    // the slab is read straight out of device 0's sorted array (peer copy into the shard's own node buffer)
    NFFTCU_TRY(nfftcu_set_nodes_dev(g->shard[r], (const char *) g->sorter->x_sorted + rs * (size_t) b));
  }
  for (int r = 0; r < g->P; r++) {
    NFFTCU_CUDA(cudaSetDevice(g->dev[r]));
    NFFTCU_CUDA(cudaStreamSynchronize(g->shard[r]->stream));
  }
  g->nodes_version++;
  return NFFTCU_OK;
}

int64_t nfftcu_group_nodes_version(nfftcu_group *g) { return g ? g->nodes_version : 0; }

/* exact NDFT (nfft_trafo_direct / nfft_adjoint_direct) of a multi-device plan: device 0 holds all nodes (the sorter) */
int nfftcu_group_direct(nfftcu_group *g, int adjoint, const void *in_host, void *out_host) {
  if (!g || g->nodes_version == 0) { set_error("group transform called before nfftcu_group_set_nodes"); return NFFTCU_ESTATE; }
  return adjoint ? nfftcu_adjoint_direct(g->sorter, in_host, out_host) : nfftcu_trafo_direct(g->sorter, in_host, out_host);
}

int nfftcu_group_get_index_x(nfftcu_group *g, int64_t *index_x_host) {
  if (!g) { set_error("null group"); return NFFTCU_EINVAL; }
  return nfftcu_get_index_x(g->sorter, index_x_host);
}

int nfftcu_group_trafo(nfftcu_group *g, const void *f_hat_host, void *f_host) {
  if (!g || g->nodes_version == 0) { set_error("group transform called before nfftcu_group_set_nodes"); return NFFTCU_ESTATE; }
  return group_trafo(g, f_hat_host, f_host);
}
int nfftcu_group_adjoint(nfftcu_group *g, const void *f_host, void *f_hat_host) {
  if (!g || g->nodes_version == 0) { set_error("group transform called before nfftcu_group_set_nodes"); return NFFTCU_ESTATE; }
  return group_adjoint(g, f_host, f_hat_host);
}
int nfftcu_group_trafo_refresh(nfftcu_group *g, const void *x_host, const void *f_hat_host, void *f_host, int *changed) {
  if (!g) { set_error("null group"); return NFFTCU_EINVAL; }
  return group_refresh(g, x_host, f_hat_host, f_host, true, changed);
}
int nfftcu_group_adjoint_refresh(nfftcu_group *g, const void *x_host, const void *f_host, void *f_hat_host, int *changed) {
  if (!g) { set_error("null group"); return NFFTCU_EINVAL; }
  return group_refresh(g, x_host, f_host, f_hat_host, false, changed);
}
int nfftcu_group_size(nfftcu_group *g) { return g ? g->P : 0; }
nfftcu_ctx *nfftcu_group_ctx(nfftcu_group *g, int rank) { return (g && rank >= 0 && rank < g->P) ? g->shard[rank] : nullptr; }
int nfftcu_group_times(nfftcu_group *g, float ms[3]) {
  if (!g) { set_error("null group"); return NFFTCU_EINVAL; }
  for (int i = 0; i < 3; i++) ms[i] = g->ms[i];
  return NFFTCU_OK;
}

}  // extern "C"
