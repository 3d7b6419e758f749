// peer.cu -- the adjoint's cross-GPU reduction fused into D^T over NVLink / NVSwitch peer memory.
//
// Node-sharded adjoint: every GPU spreads its node slab into its OWN oversampled grid and runs its own
// backward FFT; the result is f_hat[k] = c(k) * sum_p grid_p[kappa(k)].  D^T is linear, so instead of
//     D^T on every GPU  ->  partial f_hat in HBM  ->  ncclAllReduce (separate kernel, 2 more passes over f_hat)
// rank r's D^T kernel reads the band corners of ALL P grids straight through peer pointers (P2P loads over
// NVLink, fixed summation order p = 0..P-1), multiplies by c once and stores slice r of f_hat -- k in
// [r N_total/P, (r+1) N_total/P) -- into the exchange buffer of every rank that wants the result (all-reduce
// semantics for the process-per-GPU driver, own buffer only for the in-process group whose ranks copy their
// slices to the host in parallel).  One kernel does D^T and the collective; a rank moves (P-1)/P * C*N_total bytes
// in and as much out, which is what a reduce-scatter + all-gather moves, without the partial f_hat round trip.
//
// Ordering between GPUs uses device-side flag barriers in peer memory (epoch counters, system-scope fences): one
// before the kernel (all backward FFTs finished) and one after it (all slices delivered, grids may be reused).
// The same code serves ranks in one process (direct peer pointers, cudaDeviceEnablePeerAccess) and ranks in
// different processes (CUDA IPC handles exchanged by the caller, e.g. over torch.distributed).
#include "common.cuh"

#include <string.h>

namespace nfftcu {

struct PeerState {
  int rank = 0, world = 1;
  bool ipc = false;                       // peers were opened from IPC handles (close them on detach)
  bool all_outputs = true;                // deliver every slice to every rank (all-reduce) or keep own slice only
  void *grid[NFFTCU_MAX_PEERS] = {nullptr};
  void *xchg[NFFTCU_MAX_PEERS] = {nullptr};
  unsigned *flag[NFFTCU_MAX_PEERS] = {nullptr};
  void *own_xchg = nullptr;               // N_total complex
  unsigned *own_flag = nullptr;           // NFFTCU_MAX_PEERS epochs + 1 error word
  unsigned epoch = 0;
};

namespace {

struct PeerPtrs {
  void *grid[NFFTCU_MAX_PEERS];
  void *out[NFFTCU_MAX_PEERS];
  unsigned *flag[NFFTCU_MAX_PEERS];
};

struct DGeomP {
  long long N[NFFTCU_MAX_D], n[NFFTCU_MAX_D];
  int d;
};
template <typename T> struct CPtrsP { const T *c[NFFTCU_MAX_D]; };

// Every rank runs this one-warp kernel on its stream: lane p tells peer p "rank `rank` reached epoch e" and then
// waits until peer p has told us the same.  ~4.5 s of spinning is treated as a lost peer: the error word is set
// and the kernel returns instead of hanging the GPU.
__global__ void peer_barrier_kernel(PeerPtrs pp, int rank, int world, unsigned epoch) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  volatile unsigned *dst = pp.flag[p] + rank;
  *dst = epoch;
  __threadfence_system();
  volatile unsigned *src = pp.flag[rank] + p;
  const long long t0 = clock64();
  while ((int) (*src - epoch) < 0) {
    if (clock64() - t0 > (1ll << 33)) {
      pp.flag[rank][NFFTCU_MAX_PEERS] = 1u;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

__device__ __forceinline__ double2 ld_peer(const double2 *p) { return __ldcv(p); }
__device__ __forceinline__ float2 ld_peer(const float2 *p) { return __ldcv(p); }

template <typename T>
__global__ void deconv_crop_reduce_kernel(PeerPtrs pp, DGeomP geo, CPtrsP<T> cp, long long k_begin,
                                          long long k_end, int world, int nout, int out_first) {
  typedef typename Cplx<T>::type C;
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long kl = k_begin + (long long) blockIdx.x * blockDim.x + threadIdx.x; kl < k_end; kl += stride) {
    long long rem = kl;
    long long ks[NFFTCU_MAX_D];
#pragma unroll 1
    for (int t = geo.d - 1; t >= 0; t--) {
      ks[t] = rem % geo.N[t];
      rem /= geo.N[t];
    }
    long long gi = 0;
    T w = (T) 1;
    for (int t = 0; t < geo.d; t++) {
      const long long k = ks[t] - geo.N[t] / 2;
      gi = gi * geo.n[t] + (k >= 0 ? k : geo.n[t] + k);
      w = (t == 0) ? cp.c[0][ks[0]] : w * cp.c[t][ks[t]];
    }
    T re = (T) 0, im = (T) 0;
    for (int p = 0; p < world; p++) {   // fixed order: every rank would get the same bits for the same k
      const C v = ld_peer((const C *) pp.grid[p] + gi);
      re += v.x;
      im += v.y;
    }
    const C r = make_c<T>(re * w, im * w);
    for (int q = 0; q < nout; q++) ((C *) pp.out[out_first + q])[kl] = r;
  }
  __threadfence_system();
}

int barrier(nfftcu_ctx *c, const PeerPtrs &pp) {
  PeerState *ps = c->peer;
  ps->epoch++;
  peer_barrier_kernel<<<1, 32, 0, c->stream>>>(pp, ps->rank, ps->world, ps->epoch);
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

PeerPtrs ptrs_of(const PeerState *ps) {
  PeerPtrs pp;
  memset(&pp, 0, sizeof(pp));
  for (int p = 0; p < ps->world; p++) {
    pp.grid[p] = ps->grid[p];
    pp.out[p] = ps->xchg[p];
    pp.flag[p] = ps->flag[p];
  }
  return pp;
}

int ensure_own(nfftcu_ctx *c) {
  if (!c->peer) c->peer = new PeerState();
  PeerState *ps = c->peer;
  if (c->direct_only || !c->grid || c->grid2) {
    set_error("peer reduce: needs a grid plan without a split FFT axis");
    return NFFTCU_ESTATE;
  }
  const size_t bytes = 2 * real_size(c) * (size_t) c->N_total;
  // plain cudaMalloc: IPC handles cover whole allocations, and these buffers must not be recycled by the pool
  // while a peer still has them mapped
  if (!ps->own_xchg) NFFTCU_CUDA(cudaMalloc(&ps->own_xchg, bytes ? bytes : 16));
  if (!ps->own_flag) {
    NFFTCU_CUDA(cudaMalloc((void **) &ps->own_flag, sizeof(unsigned) * (NFFTCU_MAX_PEERS + 1)));
    NFFTCU_CUDA(cudaMemset(ps->own_flag, 0, sizeof(unsigned) * (NFFTCU_MAX_PEERS + 1)));
  }
  return NFFTCU_OK;
}

}  // namespace

void peer_detach(nfftcu_ctx *c) {
  PeerState *ps = c->peer;
  if (!ps) return;
  if (ps->ipc) {
    for (int p = 0; p < ps->world; p++) {
      if (p == ps->rank) continue;
      if (ps->grid[p]) cudaIpcCloseMemHandle(ps->grid[p]);
      if (ps->xchg[p]) cudaIpcCloseMemHandle(ps->xchg[p]);
      if (ps->flag[p]) cudaIpcCloseMemHandle(ps->flag[p]);
    }
  }
  if (ps->own_xchg) cudaFree(ps->own_xchg);
  if (ps->own_flag) cudaFree(ps->own_flag);
  delete ps;
  c->peer = nullptr;
}

// the fused D^T + reduce on the plan's stream; `out` = where this rank's copy of the result goes (may be null when
// all_outputs is off and the caller reads the slice from the exchange buffer)
int peer_reduce_DT(nfftcu_ctx *c, void *f_hat_dev) {
  PeerState *ps = c->peer;
  if (!ps || ps->world < 1 || !ps->grid[ps->rank]) {
    set_error("peer reduce: plan is not attached to its peers (nfftcu_peer_attach)");
    return NFFTCU_ESTATE;
  }
  ps->grid[ps->rank] = c->grid;
  const PeerPtrs pp = ptrs_of(ps);
  NFFTCU_TRY(barrier(c, pp));   // every rank's backward FFT is complete
  DGeomP geo;
  geo.d = c->d;
  for (int t = 0; t < c->d; t++) { geo.N[t] = c->N[t]; geo.n[t] = c->n[t]; }
  const long long kb = c->N_total * ps->rank / ps->world, ke = c->N_total * (ps->rank + 1) / ps->world;
  const int threads = 256;
  long long blocks = (ke - kb + threads - 1) / threads;
  const long long cap = (long long) c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int nout = ps->all_outputs ? ps->world : 1, first = ps->all_outputs ? 0 : ps->rank;
  if (c->prec == NFFTCU_DOUBLE) {
    CPtrsP<double> cp;
    for (int t = 0; t < c->d; t++) cp.c[t] = (const double *) c->c_dev[t];
    deconv_crop_reduce_kernel<double><<<(unsigned) blocks, threads, 0, c->stream>>>(pp, geo, cp, kb, ke, ps->world, nout, first);
  } else {
    CPtrsP<float> cp;
    for (int t = 0; t < c->d; t++) cp.c[t] = (const float *) c->c_dev[t];
    deconv_crop_reduce_kernel<float><<<(unsigned) blocks, threads, 0, c->stream>>>(pp, geo, cp, kb, ke, ps->world, nout, first);
  }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  NFFTCU_TRY(barrier(c, pp));   // all slices delivered; peers' grids are free for the next spread
  if (f_hat_dev) {
    const size_t C = 2 * real_size(c);
    if (ps->all_outputs)
      NFFTCU_CUDA(cudaMemcpyAsync(f_hat_dev, ps->own_xchg, C * (size_t) c->N_total, cudaMemcpyDeviceToDevice, c->stream));
    else
      NFFTCU_CUDA(cudaMemcpyAsync((char *) f_hat_dev + C * (size_t) kb, (char *) ps->own_xchg + C * (size_t) kb,
                                  C * (size_t) (ke - kb), cudaMemcpyDeviceToDevice, c->stream));
  }
  return NFFTCU_OK;
}

void *peer_slice_ptr(nfftcu_ctx *c, long long *k_begin, long long *k_end) {
  PeerState *ps = c->peer;
  *k_begin = c->N_total * ps->rank / ps->world;
  *k_end = c->N_total * (ps->rank + 1) / ps->world;
  return (char *) ps->own_xchg + 2 * real_size(c) * (size_t) *k_begin;
}

// in-process ranks: direct pointers
int peer_attach_local(nfftcu_ctx **ctxs, int world, bool all_outputs) {
  if (world < 1 || world > NFFTCU_MAX_PEERS) {
    set_error("peer attach: world %d out of range [1,%d]", world, NFFTCU_MAX_PEERS);
    return NFFTCU_EINVAL;
  }
  for (int r = 0; r < world; r++) {
    NFFTCU_CUDA(cudaSetDevice(ctxs[r]->device));
    NFFTCU_TRY(ensure_own(ctxs[r]));
    for (int p = 0; p < world; p++) {
      if (p == r || ctxs[p]->device == ctxs[r]->device) continue;
      int can = 0;
      NFFTCU_CUDA(cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[p]->device));
      if (!can) {
        set_error("peer attach: device %d cannot access device %d", ctxs[r]->device, ctxs[p]->device);
        return NFFTCU_ESTATE;
      }
      cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[p]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) NFFTCU_CUDA(e);
      cudaGetLastError();
    }
  }
  for (int r = 0; r < world; r++) {
    PeerState *ps = ctxs[r]->peer;
    ps->rank = r;
    ps->world = world;
    ps->ipc = false;
    ps->all_outputs = all_outputs;
    for (int p = 0; p < world; p++) {
      ps->grid[p] = ctxs[p]->grid;
      ps->xchg[p] = ctxs[p]->peer->own_xchg;
      ps->flag[p] = ctxs[p]->peer->own_flag;
    }
  }
  return NFFTCU_OK;
}

}  // namespace nfftcu

using namespace nfftcu;

extern "C" {

int nfftcu_peer_export(nfftcu_ctx *c, void *handles) {
  if (!c || !handles) { set_error("nfftcu_peer_export: null argument"); return NFFTCU_EINVAL; }
  NFFTCU_CUDA(cudaSetDevice(c->device));
  NFFTCU_TRY(ensure_own(c));
  cudaIpcMemHandle_t h[3];
  NFFTCU_CUDA(cudaIpcGetMemHandle(&h[0], c->grid));
  NFFTCU_CUDA(cudaIpcGetMemHandle(&h[1], c->peer->own_xchg));
  NFFTCU_CUDA(cudaIpcGetMemHandle(&h[2], c->peer->own_flag));
  static_assert(sizeof(h) == NFFTCU_PEER_HANDLE_BYTES, "handle blob size");
  memcpy(handles, h, sizeof(h));
  return NFFTCU_OK;
}

int nfftcu_peer_attach(nfftcu_ctx *c, int rank, int world, const void *all_handles) {
  if (!c || !all_handles || world < 1 || world > NFFTCU_MAX_PEERS || rank < 0 || rank >= world) {
    set_error("nfftcu_peer_attach: bad argument (rank %d, world %d)", rank, world);
    return NFFTCU_EINVAL;
  }
  NFFTCU_CUDA(cudaSetDevice(c->device));
  NFFTCU_TRY(ensure_own(c));
  PeerState *ps = c->peer;
  ps->rank = rank;
  ps->world = world;
  ps->ipc = true;
  ps->all_outputs = true;
  const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *) all_handles;
  for (int p = 0; p < world; p++) {
    if (p == rank) {
      ps->grid[p] = c->grid;
      ps->xchg[p] = ps->own_xchg;
      ps->flag[p] = ps->own_flag;
      continue;
    }
    NFFTCU_CUDA(cudaIpcOpenMemHandle(&ps->grid[p], h[3 * p + 0], cudaIpcMemLazyEnablePeerAccess));
    NFFTCU_CUDA(cudaIpcOpenMemHandle(&ps->xchg[p], h[3 * p + 1], cudaIpcMemLazyEnablePeerAccess));
    NFFTCU_CUDA(cudaIpcOpenMemHandle((void **) &ps->flag[p], h[3 * p + 2], cudaIpcMemLazyEnablePeerAccess));
  }
  return NFFTCU_OK;
}

int nfftcu_peer_detach(nfftcu_ctx *c) {
  if (!c) return NFFTCU_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  peer_detach(c);
  return NFFTCU_OK;
}

// the fused D^T + reduce alone, on whatever the grids hold (profiling: the collective's own cost)
int nfftcu_peer_reduce_only(nfftcu_ctx *c, void *f_hat_dev) {
  if (!c || !c->peer) { set_error("nfftcu_peer_reduce_only: plan is not attached to its peers"); return NFFTCU_ESTATE; }
  NFFTCU_CUDA(cudaSetDevice(c->device));
  return peer_reduce_DT(c, f_hat_dev);
}

// 0 = fine; 1 = a flag barrier gave up waiting for a peer (result invalid)
int nfftcu_peer_error(nfftcu_ctx *c) {
  if (!c || !c->peer || !c->peer->own_flag) return 0;
  unsigned e = 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaMemcpy(&e, c->peer->own_flag + NFFTCU_MAX_PEERS, sizeof(unsigned), cudaMemcpyDeviceToHost);
  return (int) e;
}

}  // extern "C"
