// interp.cu -- B step of nfft_trafo: f_j = sum_l prod_t psi_t[l_t] * g[(u_j + l) mod n].
//
// Replaces nfft_trafo_1d/2d/3d_B + *_compute (kernel/nfft/nfft.c:2283-2445/2131-2153,
// 3221-3410/2927-3004, 4687-4914/4020-4265 of the reference) and the generic B_openmp_A
// (1172-1278).  The window is evaluated on the fly (or read from the optional per-node table,
// the PRE_PSI analogue) in double and rounded once to the plan precision.
//
// Kernel "generic" (any d, any m): one warp per node, nodes taken in processing (sorted) order so
// that neighbouring warps hit neighbouring grid lines in L2.  Lanes run along the contiguous
// grid dimension (16 lanes per tap row when 2m+2 <= 16, so two rows per pass), outer
// dimensions are plain loops, products are formed as (psi0*psi1)*psi2 like nfft.c:4048.
#include "common.cuh"
#include "window.cuh"

namespace nfftcu {

namespace {

constexpr int kWarpsPerBlock = 8;

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int D>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
interp_generic_kernel(const typename Cplx<T>::type *__restrict__ g, const T *__restrict__ xs,
                      const uint32_t *__restrict__ perm, typename Cplx<T>::type *__restrict__ f,
                      long long M, NodeGeom geo, const T *__restrict__ psi_table) {
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
  const long long nwarps = (long long) gridDim.x * kWarpsPerBlock;
  for (long long k = (long long) blockIdx.x * kWarpsPerBlock + warp; k < M; k += nwarps) {
    warp_node_window<T>(xs + k * geo.d, geo, psi_table ? psi_table + k * cnt : nullptr, psi, off,
                        lane);
    __syncwarp();
    T accr = (T) 0, acci = (T) 0;
    if (D == 1) {
      for (int l = lane; l < W; l += 32) {
        const C v = g[off[l]];
        accr += psi[l] * v.x;
        acci += psi[l] * v.y;
      }
    } else if (D == 2) {
      for (int l0 = sub; l0 < W; l0 += rows_per_pass) {
        const long long o0 = off[l0];
        const T p0 = psi[l0];
        for (int l1 = l_in; l1 < W; l1 += LW) {
          const C v = g[o0 + off[W + l1]];
          const T w = p0 * psi[W + l1];
          accr += w * v.x;
          acci += w * v.y;
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
            const C v = g[o01 + off[2 * W + l2]];
            const T w = p01 * psi[2 * W + l2];
            accr += w * v.x;
            acci += w * v.y;
          }
        }
      }
    } else {
      long long rows = 1;
      for (int t = 0; t < geo.d - 1; t++) rows *= W;
      const int last = (geo.d - 1) * W;
      for (long long row = sub; row < rows; row += rows_per_pass) {
        // decode row (digits base W over dims 0..d-2, dim d-2 fastest), weight in dim order
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
          const C v = g[o + off[last + l]];
          const T w = (geo.d == 1) ? psi[l] : w0 * psi[last + l];
          accr += w * v.x;
          acci += w * v.y;
        }
      }
    }
    accr = warp_sum(accr);
    acci = warp_sum(acci);
    if (lane == 0) f[perm[k]] = make_c<T>(accr, acci);
    __syncwarp();
  }
}

template <typename T>
__global__ void psi_table_kernel(const T *__restrict__ xs, T *__restrict__ table, long long M,
                                 NodeGeom geo) {
  const int cnt = geo.d * geo.W;
  const long long total = M * cnt;
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long k = i / cnt;
    const int r = (int) (i - k * cnt);
    const int t = r / geo.W, l = r - t * geo.W;
    const T x = xs[k * geo.d + t];
    const long long u = cell_of(x, geo.n[t]) - geo.m;
    const double dist = (double) x * (double) geo.n[t] - (double) (u + l);
    table[i] = (T) window_phi(dist, geo.m2, geo.b[t], geo.window, geo.ws[t]);
  }
}

template <typename T>
int run_generic(nfftcu_ctx *c, void *f_dev) {
  typedef typename Cplx<T>::type C;
  const NodeGeom geo = make_node_geom(c);
  const int cnt = geo.d * geo.W;
  const size_t smem = (sizeof(long long) + sizeof(T)) * (size_t) kWarpsPerBlock * cnt;
  long long blocks = (c->M + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long) c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const dim3 grid((unsigned) blocks), block(kWarpsPerBlock * 32);
  const C *g = (const C *) c->grid;
  const T *xs = (const T *) c->x_sorted;
  const T *tab = (c->opt_psi_table && c->psi_table_valid) ? (const T *) c->psi_table : nullptr;
  switch (c->d) {
    case 1: interp_generic_kernel<T, 1><<<grid, block, smem, c->stream>>>(g, xs, c->perm, (C *) f_dev, c->M, geo, tab); break;
    case 2: interp_generic_kernel<T, 2><<<grid, block, smem, c->stream>>>(g, xs, c->perm, (C *) f_dev, c->M, geo, tab); break;
    case 3: interp_generic_kernel<T, 3><<<grid, block, smem, c->stream>>>(g, xs, c->perm, (C *) f_dev, c->M, geo, tab); break;
    default: interp_generic_kernel<T, 0><<<grid, block, smem, c->stream>>>(g, xs, c->perm, (C *) f_dev, c->M, geo, tab); break;
  }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

}  // namespace

int stage_B(nfftcu_ctx *c, void *f_dev) {
  if (c->M == 0) return NFFTCU_OK;
  if (c->tile2_ready && c->opt_b_kernel != 1) return tile2d_interp(c, f_dev);   // batch = gridDim.y
  if (c->cur_batch > 1) {
    // the other kernel families take one right-hand side per launch: walk the batch with the grid pointer moved
    const int K = c->cur_batch;
    const size_t C = 2 * real_size(c);
    char *g0 = (char *) c->grid;
    int r = NFFTCU_OK;
    c->cur_batch = 1;
    for (int k = 0; k < K && r == NFFTCU_OK; k++) {
      c->grid = g0 + C * (size_t) c->n_total * k;
      r = stage_B(c, (char *) f_dev + C * (size_t) c->M * k);
    }
    c->grid = g0;
    c->cur_batch = K;
    return r;
  }
  if (c->mma_ready) return mma3d_interp(c, f_dev);
  if (c->tile_ready && c->opt_b_kernel != 1) return tile3d_interp(c, f_dev);
  if (!c->ref_sorted) {
    set_error("stage_B: generic kernel needs the reference node order (set NFFTCU_OPT_B_KERNEL before set_nodes)");
    return NFFTCU_ESTATE;
  }
  return c->prec == NFFTCU_DOUBLE ? run_generic<double>(c, f_dev) : run_generic<float>(c, f_dev);
}

int build_psi_table(nfftcu_ctx *c) {
  if (c->M == 0) return NFFTCU_OK;
  const NodeGeom geo = make_node_geom(c);
  const size_t bytes = real_size(c) * (size_t) c->M * geo.d * geo.W;
  if (!c->psi_table) NFFTCU_CUDA(pool_malloc(&c->psi_table, bytes));
  const int threads = 256;
  long long blocks = ((long long) c->M * geo.d * geo.W + threads - 1) / threads;
  const long long cap = (long long) c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (c->prec == NFFTCU_DOUBLE)
    psi_table_kernel<double><<<(unsigned) blocks, threads, 0, c->stream>>>(
        (const double *) c->x_sorted, (double *) c->psi_table, c->M, geo);
  else
    psi_table_kernel<float><<<(unsigned) blocks, threads, 0, c->stream>>>(
        (const float *) c->x_sorted, (float *) c->psi_table, c->M, geo);
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  c->psi_table_valid = true;
  return NFFTCU_OK;
}

}  // namespace nfftcu
