// spread.cu -- B^T step of nfft_adjoint: g := 0; g[(u_j + l) mod n] += prod_t psi_t[l_t] * f_j.
//
// Replaces nfft_adjoint_1d/2d/3d_B (kernel/nfft/nfft.c:2585-2773, 3583-3805, 5126-5384 of the
// reference) with both of its compute flavours -- the `omp atomic` scatter (4393-4436) and the
// owner-computes "blockwise" slabs (4289-4388, 1345-1420) -- and the generic B_openmp_T
// (1974-2098).
//
// Kernel "generic" (any d, any m): one warp per node in processing (sorted) order; lanes run
// along the contiguous grid dimension so that the reductions of a warp land in one or two
// 128-byte lines, and are issued as fire-and-forget global reductions (RED.ADD) that resolve
// in L2.  Summation order therefore differs from the reference's; the result differs from
// it at the 1e-16 level per add (the reference's own atomic and blockwise variants differ
// from each other by <= 1.2e-15).
#include "common.cuh"
#include "window.cuh"

namespace nfftcu {

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }

template <typename T, int D>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
spread_generic_kernel(typename Cplx<T>::type *__restrict__ g, const T *__restrict__ xs,
                      const uint32_t *__restrict__ perm,
                      const typename Cplx<T>::type *__restrict__ f, long long M, NodeGeom geo,
                      const T *__restrict__ psi_table) {
  typedef typename Cplx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = geo.W, cnt = geo.d * W;
  long long *off = reinterpret_cast<long long *>(smem_raw) + (size_t) warp * cnt;
  T *psi = reinterpret_cast<T *>(smem_raw + sizeof(long long) * (size_t) kWarpsPerBlock * cnt) +
           (size_t) warp * cnt;
  const int LW = (W <= 16) ? 16 : 32;
  const int rows_per_pass = 32 / LW;
  const int sub = lane / LW, l_in = lane % LW;
  T *gr = reinterpret_cast<T *>(g);
  const long long nwarps = (long long) gridDim.x * kWarpsPerBlock;
  for (long long k = (long long) blockIdx.x * kWarpsPerBlock + warp; k < M; k += nwarps) {
    warp_node_window<T>(xs + k * geo.d, geo, psi_table ? psi_table + k * cnt : nullptr, psi, off,
                        lane);
    const C fj = f[perm[k]];
    __syncwarp();
    if (D == 1) {
      for (int l = lane; l < W; l += 32) {
        T *p = gr + 2 * off[l];
        red_add(p, psi[l] * fj.x);
        red_add(p + 1, psi[l] * fj.y);
      }
    } else if (D == 2) {
      for (int l0 = sub; l0 < W; l0 += rows_per_pass) {
        const long long o0 = off[l0];
        const T p0 = psi[l0];
        for (int l1 = l_in; l1 < W; l1 += LW) {
          const T w = p0 * psi[W + l1];
          T *p = gr + 2 * (o0 + off[W + l1]);
          red_add(p, w * fj.x);
          red_add(p + 1, w * fj.y);
        }
      }
    } else if (D == 3) {
      for (int l0 = 0; l0 < W; l0++) {
        const long long o0 = off[l0];
        const T p0 = psi[l0];
        for (int l1 = sub; l1 < W; l1 += rows_per_pass) {
          const long long o01 = o0 + off[W + l1];
          const T p01 = p0 * psi[W + l1];
          for (int l2 = l_in; l2 < W; l2 += LW) {
            const T w = p01 * psi[2 * W + l2];
            T *p = gr + 2 * (o01 + off[2 * W + l2]);
            red_add(p, w * fj.x);
            red_add(p + 1, w * fj.y);
          }
        }
      }
    } else {
      long long rows = 1;
      for (int t = 0; t < geo.d - 1; t++) rows *= W;
      const int last = (geo.d - 1) * W;
      for (long long row = sub; row < rows; row += rows_per_pass) {
        int dig[NFFTCU_MAX_D];
        long long rem = row;
        for (int t = geo.d - 2; t >= 0; t--) { dig[t] = (int) (rem % W); rem /= W; }
        long long o = 0;
        T w0 = (T) 1;
        for (int t = 0; t < geo.d - 1; t++) {
          o += off[t * W + dig[t]];
          w0 = (t == 0) ? psi[dig[0]] : w0 * psi[t * W + dig[t]];
        }
        for (int l = l_in; l < W; l += LW) {
          const T w = w0 * psi[last + l];
          T *p = gr + 2 * (o + off[last + l]);
          red_add(p, w * fj.x);
          red_add(p + 1, w * fj.y);
        }
      }
    }
    __syncwarp();
  }
}

template <typename T>
int run_generic(nfftcu_ctx *c, const void *f_dev) {
  typedef typename Cplx<T>::type C;
  const NodeGeom geo = make_node_geom(c);
  const int cnt = geo.d * geo.W;
  const size_t smem = (sizeof(long long) + sizeof(T)) * (size_t) kWarpsPerBlock * cnt;
  long long blocks = (c->M + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long) c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const dim3 grid((unsigned) blocks), block(kWarpsPerBlock * 32);
  C *g = (C *) c->grid;
  const T *xs = (const T *) c->x_sorted;
  const T *tab = (c->opt_psi_table && c->psi_table_valid) ? (const T *) c->psi_table : nullptr;
  const C *f = (const C *) f_dev;
  switch (c->d) {
    case 1: spread_generic_kernel<T, 1><<<grid, block, smem, c->stream>>>(g, xs, c->perm, f, c->M, geo, tab); break;
    case 2: spread_generic_kernel<T, 2><<<grid, block, smem, c->stream>>>(g, xs, c->perm, f, c->M, geo, tab); break;
    case 3: spread_generic_kernel<T, 3><<<grid, block, smem, c->stream>>>(g, xs, c->perm, f, c->M, geo, tab); break;
    default: spread_generic_kernel<T, 0><<<grid, block, smem, c->stream>>>(g, xs, c->perm, f, c->M, geo, tab); break;
  }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

}  // namespace

int stage_BT(nfftcu_ctx *c, const void *f_dev, bool slab_ok) {
  if (c->cur_batch > 1 && !(c->tile2_ready && c->opt_b_kernel != 1)) {
    // kernel families without a batch dimension: one right-hand side per launch, grid pointer moved along the batch
    const int K = c->cur_batch;
    const size_t C = 2 * real_size(c);
    char *g0 = (char *) c->grid;
    int r = NFFTCU_OK;
    c->cur_batch = 1;
    for (int k = 0; k < K && r == NFFTCU_OK; k++) {
      c->grid = g0 + C * (size_t) c->n_total * k;
      r = stage_BT(c, (const char *) f_dev + C * (size_t) c->M * k, slab_ok);
    }
    c->grid = g0;
    c->cur_batch = K;
    return r;
  }
  if (slab_ok && c->slab_on && c->cur_batch == 1) {
    // slab mode: the pruned backward F reads the window planes only (fft.cu), so only they are cleared
    const size_t plane = 2 * real_size(c) * (size_t) (c->n_total / c->n[0]);
    const long long w0 = c->slab_w0, wc = c->slab_wc, n0 = c->n[0];
    const long long first = w0 + wc <= n0 ? wc : n0 - w0;
    NFFTCU_CUDA(cudaMemsetAsync((char *) c->grid + plane * (size_t) w0, 0, plane * (size_t) first, c->stream));
    if (first < wc) NFFTCU_CUDA(cudaMemsetAsync(c->grid, 0, plane * (size_t) (wc - first), c->stream));
  } else {
    NFFTCU_CUDA(cudaMemsetAsync(c->grid, 0, 2 * real_size(c) * (size_t) c->n_total * (size_t) c->cur_batch, c->stream));
  }
  if (c->M == 0) return NFFTCU_OK;
  if (c->mma_ready) return mma3d_spread(c, f_dev);
  if (c->tile2_ready && c->opt_b_kernel != 1) return tile2d_spread(c, f_dev);
  if (c->tile_ready && c->opt_b_kernel != 1) return tile3d_spread(c, f_dev);
  if (!c->ref_sorted) {
    set_error("stage_BT: generic kernel needs the reference node order (set NFFTCU_OPT_B_KERNEL before set_nodes)");
    return NFFTCU_ESTATE;
  }
  return c->prec == NFFTCU_DOUBLE ? run_generic<double>(c, f_dev) : run_generic<float>(c, f_dev);
}

}  // namespace nfftcu
