// mempool.cu -- per-process caches of device and page-locked host buffers.
//
// Plan-per-coil callers (applications/mri/mri2d/reconstruct_data_2d.c runs one nfft_init_guru ... nfft_finalize per
// coil file; BASELINE configs[4]) create and destroy plans of identical shapes back to back, and cudaMalloc /
// cudaMallocHost / cudaFree(Host) dominate such a life cycle (tens of calls at 0.1 - 3 ms each).  Freed buffers are
// therefore kept in exact-size free lists (per device for device memory) up to a byte cap and handed out again.
// pool_free synchronises the device before a buffer becomes reusable, which is the guarantee cudaFree gives
// implicitly.  Env NFFT_B200_POOL_MB sets the cap per kind (default 4096, 0 disables the caches); nfftcu_pool_trim()
// returns everything cached to the driver (for processes that share the GPU with other CUDA users).
#include "common.cuh"

#include <stdlib.h>

#include <map>
#include <mutex>
#include <unordered_map>

namespace nfftcu {

namespace {

struct Pool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void *> free_list;   // (device, bytes) -> buffer
  std::unordered_map<void *, std::pair<int, size_t>> live;   // buffers handed out
  size_t cached = 0;
};

Pool g_dev, g_host;

size_t cap_bytes() {
  static const size_t cap = [] {
    const char *e = getenv("NFFT_B200_POOL_MB");
    const long long mb = e ? atoll(e) : 4096;
    return (size_t) (mb < 0 ? 0 : mb) << 20;
  }();
  return cap;
}

void flush(Pool &P, bool host) {
  for (auto &kv : P.free_list) {
    if (host) cudaFreeHost(kv.second);
    else cudaFree(kv.second);
  }
  P.free_list.clear();
  P.cached = 0;
}

cudaError_t pool_get(Pool &P, bool host, void **out, size_t bytes) {
  if (bytes == 0) bytes = 1;
  int dev = -1;
  if (!host) { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
  std::lock_guard<std::mutex> lock(P.mu);
  auto it = P.free_list.find({dev, bytes});
  if (it != P.free_list.end()) {
    *out = it->second;
    P.free_list.erase(it);
    P.cached -= bytes;
  } else {
    // page-locked memory is portable: every device of a sharded plan copies from / to it
    cudaError_t e = host ? cudaHostAlloc(out, bytes, cudaHostAllocPortable) : cudaMalloc(out, bytes);
    if (e != cudaSuccess) {   // make room and retry once
      cudaGetLastError();
      flush(P, host);
      e = host ? cudaHostAlloc(out, bytes, cudaHostAllocPortable) : cudaMalloc(out, bytes);
      if (e != cudaSuccess) return e;
    }
  }
  P.live[*out] = {dev, bytes};
  return cudaSuccess;
}

cudaError_t pool_put(Pool &P, bool host, void *p) {
  if (!p) return cudaSuccess;
  std::unique_lock<std::mutex> lock(P.mu);
  auto it = P.live.find(p);
  if (it == P.live.end()) {   // not ours (allocated before the pool existed or by the caller)
    lock.unlock();
    return host ? cudaFreeHost(p) : cudaFree(p);
  }
  const std::pair<int, size_t> key = it->second;
  P.live.erase(it);
  if (key.second > cap_bytes() / 4 || P.cached + key.second > cap_bytes()) {
    lock.unlock();
    return host ? cudaFreeHost(p) : cudaFree(p);
  }
  lock.unlock();
  // what cudaFree guarantees: nothing in flight still uses the buffer -- on the device that OWNS it, which in a
  // multi-GPU process need not be the current one
  int prev = -1;
  if (!host && cudaGetDevice(&prev) == cudaSuccess && prev != key.first) cudaSetDevice(key.first);
  else prev = -1;
  cudaError_t e = cudaDeviceSynchronize();
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) return e;
  lock.lock();
  P.free_list.insert({key, p});
  P.cached += key.second;
  return cudaSuccess;
}

}  // namespace

cudaError_t pool_malloc_bytes(void **p, size_t bytes) { return pool_get(g_dev, false, p, bytes); }
cudaError_t pool_free(void *p) { return pool_put(g_dev, false, p); }
cudaError_t pool_malloc_host_bytes(void **p, size_t bytes) { return pool_get(g_host, true, p, bytes); }
cudaError_t pool_free_host(void *p) { return pool_put(g_host, true, p); }
bool pool_owns_host(void *p) {
  std::lock_guard<std::mutex> lock(g_host.mu);
  return g_host.live.count(p) != 0;
}
void pool_trim() {
  { std::lock_guard<std::mutex> lock(g_dev.mu); flush(g_dev, false); }
  { std::lock_guard<std::mutex> lock(g_host.mu); flush(g_host, true); }
}

}  // namespace nfftcu
