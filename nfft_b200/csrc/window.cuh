// window.cuh -- per-node window setup shared by the interpolation (B) and spreading (B^T)
// kernels: tap origin u = floor(x*n) - m (uo, kernel/nfft/nfft.c:324-332 of the reference), the
// 2m+2 Kaiser-Bessel values psi[l] = phi(x - (u+l)/n) (nfft.c:4896-4908; precompute_psi
// 5838-5840) and the wrapped grid offsets ((u+l) mod n) * stride.
#pragma once

#include "common.cuh"

namespace nfftcu {

struct NodeGeom {
  long long n[NFFTCU_MAX_D];
  long long stride[NFFTCU_MAX_D];   // row-major element stride of each dimension
  double b[NFFTCU_MAX_D];
  double ws[NFFTCU_MAX_D];          // power-of-two window scale (ctx->wscale)
  int window;                       // window family
  double m2;                        // m*m
  int d;
  int m;
  int W;                            // 2m+2 taps per dimension
};

inline NodeGeom make_node_geom(const nfftcu_ctx *c) {
  NodeGeom g;
  g.d = c->d;
  g.m = (int) c->m;
  g.W = 2 * (int) c->m + 2;
  g.m2 = (double) c->m * (double) c->m;
  g.window = c->window;
  long long s = 1;
  for (int t = c->d - 1; t >= 0; t--) {
    g.n[t] = c->n[t];
    g.b[t] = c->b[t];
    g.ws[t] = c->wscale[t];
    g.stride[t] = s;
    s *= c->n[t];
  }
  return g;
}

// One warp fills, for its node, psi[t*W+l] and off[t*W+l] (t<d, l<W) in shared memory.
// `table` (may be null) is the node's precomputed row psi_table[(k*d+t)*W+l].
template <typename T>
__device__ __forceinline__ void warp_node_window(const T *__restrict__ xj, const NodeGeom &g,
                                                 const T *__restrict__ table, T *psi,
                                                 long long *off, int lane) {
  const int cnt = g.d * g.W;
  for (int i = lane; i < cnt; i += 32) {
    const int t = i / g.W, l = i - t * g.W;
    const T x = xj[t];
    const long long n = g.n[t];
    const long long u = cell_of(x, n) - g.m;
    if (table) psi[i] = table[i];
    else {
      const double dist = (double) x * (double) n - (double) (u + l);
      psi[i] = (T) window_phi(dist, g.m2, g.b[t], g.window, g.ws[t]);
    }
    long long idx = (u + l) % n;
    if (idx < 0) idx += n;
    off[i] = idx * g.stride[t];
  }
}

}  // namespace nfftcu
