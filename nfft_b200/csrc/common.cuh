// common.cuh -- internal declarations shared by the translation units of libnfftcu.so.
// Not part of the public boundary (that is include/nfftcu.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/nfftcu.h"

namespace nfftcu {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define NFFTCU_CUDA(call)                                                                       \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ::nfftcu::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                    \
                          cudaGetErrorString(e__));                                             \
      return NFFTCU_ECUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define NFFTCU_TRY(call)                                                                        \
  do {                                                                                          \
    int r__ = (call);                                                                           \
    if (r__ != NFFTCU_OK) return r__;                                                           \
  } while (0)

// ---- cached allocations (mempool.cu): drop-in for cudaMalloc / cudaFree / cudaMallocHost / cudaFreeHost ----------
cudaError_t pool_malloc_bytes(void **p, size_t bytes);
cudaError_t pool_free(void *p);
cudaError_t pool_malloc_host_bytes(void **p, size_t bytes);
cudaError_t pool_free_host(void *p);
bool pool_owns_host(void *p);   // p is a live page-locked buffer handed out by pool_malloc_host
void pool_trim();               // give all cached (free) device and page-locked buffers back to the driver
template <typename T> inline cudaError_t pool_malloc(T **p, size_t bytes) { return pool_malloc_bytes((void **) p, bytes); }
template <typename T> inline cudaError_t pool_malloc_host(T **p, size_t bytes) { return pool_malloc_host_bytes((void **) p, bytes); }

// ---- scalar/complex type traits ----------------------------------------------------------------
template <typename T> struct Cplx;
template <> struct Cplx<double> { typedef double2 type; };
template <> struct Cplx<float> { typedef float2 type; };

template <typename T> __host__ __device__ inline typename Cplx<T>::type make_c(T re, T im);
template <> __host__ __device__ inline double2 make_c<double>(double re, double im) { return make_double2(re, im); }
template <> __host__ __device__ inline float2 make_c<float>(float re, float im) { return make_float2(re, im); }

// ---- plan context -----------------------------------------------------------------------------
// one shared-memory transform length
struct FftLine {
  int64_t len = 0;
  void *tw = nullptr;   // len complex (plan precision): exp(-2 pi i q / len)
  int kind = 0;         // 0: len==1, 1: radix-4/2 Stockham (power of two), 2: O(len^2) table DFT, 3: mixed radix, 4: Bluestein
  int nst = 0;          // kind 3: radix per stage
  int radix[24] = {0};
  // kind 4 (a prime factor > 61): chirp-z through a power-of-two convolution of length P >= 2 len - 1
  int64_t blu_P = 0;
  void *blu_chirp = nullptr;   // len complex: exp(-i pi k^2 / len)
  void *blu_bhat = nullptr;    // P complex: DFT_P of the wrapped conjugate chirp, scaled by 1/P
  void *blu_tw = nullptr;      // P complex: exp(-2 pi i q / P)
};
struct FftAxis {
  int64_t len = 0;
  bool split = false;   // four-step: len = sub1.len * sub2.len, twiddle W_len^q = twA[q >> tw_shift] * twB[q & mask]
  FftLine whole, sub1, sub2;
  void *twA = nullptr, *twB = nullptr;
  int tw_shift = 0;
};

}  // namespace nfftcu

namespace nfftcu {
struct PeerState;
// one set of plan-time tables of the tcgen05 kernels (tc5.cu)
struct Tc5Tables {
  void *batches = nullptr;          // uint2 per batch: first node, base | nb << 24 | flags
  uint32_t *batch_start = nullptr;  // units + 1 offsets into batches
  uint32_t *chunk_start = nullptr;  // units + 1 offsets into chunks
  void *chunks = nullptr;           // uint4 per chunk: tile, first batch, end batch
  void *images = nullptr;           // operand image of every batch
  size_t images_bytes = 0;
  long long units = 0, batch_cap = 0, chunk_cap = 0, nchunks = 0, nbatches = 0;
};
}

struct nfftcu_ctx_s {
  int prec = NFFTCU_DOUBLE;
  int d = 0;
  int device = 0;
  int64_t N[NFFTCU_MAX_D] = {0}, n[NFFTCU_MAX_D] = {0};
  int64_t m = 0, M = 0, N_total = 0, n_total = 0;
  unsigned flags = 0;
  bool direct_only = false;           // any N_t <= m or n_t <= 2m+2 (nfft.c:5658-5664)
  double b[NFFTCU_MAX_D] = {0}, sigma[NFFTCU_MAX_D] = {0};
  int window = NFFTCU_WINDOW_KAISER_BESSEL;   // window family (create flag NFFTCU_FLAG_GAUSSIAN)
  // Power-of-two factor folded into the device window values of dimension t (and out of c_dev[t]): fp32 plans use
  // 2^-round(log2 phi_hat_t(0)) so that grid values stay O(data) instead of O(1e11^d * data) -- the reference's
  // unscaled fp32 Kaiser-Bessel adjoint overflows to inf at cfg3 (M = 1e7, positive samples).  Exact (powers of
  // two), 1.0 for fp64 plans.
  double wscale[NFFTCU_MAX_D] = {1, 1, 1, 1, 1, 1, 1, 1};
  std::vector<double> c_host[NFFTCU_MAX_D];   // c_phi_inv in double
  void *c_dev[NFFTCU_MAX_D] = {nullptr};      // c_phi_inv in plan precision
  void *grid = nullptr;                       // batch_cap x n_total complex (slice 0 serves the single transforms)
  int batch_cap = 1;                          // right-hand sides the grid / staging buffers can hold (nfftcu_*_batch)
  int cur_batch = 1;                          // right-hand sides of the transform in flight: every stage reads it
  void *grid2 = nullptr;                      // second buffer, only for plans with a split (four-step) FFT axis
  bool fft_no_prune = false;                  // a split axis runs unpruned passes
  // slab mode (nodes_ready -> slab_detect): all taps of the resident nodes lie in the planes [slab_w0, slab_w0 + slab_wc)
  // (mod n_0) of the first axis; the pruned F passes and the B^T memset then leave every other plane alone (fft.cu)
  bool slab_on = false;
  long long slab_w0 = 0, slab_wc = 0;
  int opt_slab = 0;                           // NFFTCU_OPT_SLAB_FFT: 0 auto | 1 off
  nfftcu::FftAxis fft[NFFTCU_MAX_D];

  // nodes
  bool have_nodes = false;
  void *x_dev = nullptr;            // M*d reals, caller order
  void *x_stage = nullptr;          // upload target, compared against x_dev before re-sorting
  int *diff_flag = nullptr;         // device flag written by the compare kernel
  int64_t nodes_version = 0;
  void *x_sorted = nullptr;         // M*d reals, processing order
  uint32_t *perm = nullptr;         // processing order -> original node index
  void *keys_ref = nullptr;         // sorted reference keys (uint64), for index_x
  uint32_t *perm_ref = nullptr;     // reference permutation (== perm when node order is the reference key)
  void *psi_table = nullptr;        // optional: M * d * (2m+2) reals in processing order
  bool psi_table_valid = false;     // table was built for the current nodes (nodes_ready clears it)
  uint64_t x_fp = 0;                // fingerprint of the HOST array the resident nodes were uploaded from
  bool x_fp_valid = false;
  unsigned built_sig = 0;           // node_opts_signature at the time the node-dependent state was built
  bool parked_nodes = false;        // revived from the plan cache: node state of the previous owner, not yet adopted
  // tile-binned order for the 3-D pencil-sweep kernels (tile3d.cu)
  bool tile_ready = false;
  bool tile2_ready = false;         // tile_* hold the tile order of the 2-D kernels (tile2d.cu)
  bool mma_ready = false;           // tile_* hold the (tile, u2) order of the DMMA kernels (mma3d.cu)
  void *mma_batches = nullptr;      // uint2 per batch: first node, zlo | nb << 24 | last << 28
  uint32_t *mma_batch_start = nullptr;   // units+1 offsets into mma_batches
  uint32_t *mma_counts = nullptr;   // scratch: batches per unit
  long long mma_units = 0, mma_batch_cap = 0;
  uint32_t *mma_chunk_start = nullptr;   // units+1 offsets into mma_chunks
  void *mma_chunks = nullptr;       // uint4 per CTA: tile, first batch, end batch (runs of <= 384 batches of one unit)
  long long mma_nchunks = 0, mma_chunk_cap = 0;
  void *mma_images = nullptr;       // placed operand blocks (psi0, psi1, psi2: 3 KB) of every batch, built per node set when they fit
  size_t mma_images_bytes = 0;
  bool mma_images_ready = false;
  bool mma_images_tf32 = false;     // images hold packed fp32 / TF32 pairs (fp32 plans on the TF32 kernels)
  int opt_window_images = 0;        // 0 auto | 1 off | 2 on regardless of the memory budget
  // fp32 plans on the tcgen05 / TMEM kernels (tc5.cu): own batch tables (<= 16 nodes; interpolation: 8-aligned window base,
  // 24 slots; spreading: 16-aligned, 32 slots), chunk lists and operand images
  nfftcu::Tc5Tables tc5i, tc5s;
  uint32_t *tc5_counts = nullptr;   // scratch: batches / chunks per unit
  void *tc5_ft = nullptr;           // spreading: samples in tile order, M + 18 float2
  long long tc5_ft_cap = 0;
  long long tc5_counts_units = 0;
  bool tc5_ready = false;           // tc5i is valid for the current nodes: B runs on tc5_interp_kernel
  bool tc5s_ready = false;          // tc5s is valid: B^T runs on tc5_spread_kernel
  int opt_tc5 = 0;                  // NFFTCU_OPT_TC5: 0 auto (fp32, d = 3, m <= 6: B and B^T) | 1 off | 2 B only | 3 B and B^T
  bool ref_sorted = false;          // keys_ref / perm / x_sorted are valid for the current nodes
  void *tile_keys = nullptr;        // uint64 bin ids, sorted
  uint32_t *tile_perm = nullptr;    // tile order -> original node index
  void *tile_x = nullptr;           // M*3 reals in tile order
  uint32_t *bin_start = nullptr;    // nbins+1 offsets into the tile order
  void *tile_psi = nullptr;         // node records of the pencil kernels (padded window vectors, slab, f)
  bool tile_psi_valid = false;      // window part of the records matches the current nodes
  long long tile_nbins = 0;
  void *f_tile = nullptr;           // M complex: samples in tile order (gathered / to be scattered)
  // piecewise-polynomial window (kbpoly.cu): coef[(t*(deg+1)+k)*W + l]
  int kbpoly_deg = -1;
  int kbpoly_fit = -1;              // degree the fit needed (<= kbpoly_deg): Horner length of the DMMA kernels
  std::vector<double> kbpoly_host;
  void *kbpoly_dev = nullptr;
  void *sort_tmp = nullptr;         // scratch kept between set_nodes calls
  size_t sort_tmp_bytes = 0;

  nfftcu::PeerState *peer = nullptr;   // fused D^T + cross-GPU reduce (peer.cu)
  bool nodes_only = false;             // sorter of a multi-GPU group: no grid, no FFT plan, never transforms
  bool no_cache = false;               // nfftcu_destroy really destroys (plan cache eviction / trim)

  // staging buffers for the host-pointer API
  void *fhat_dev = nullptr;
  void *f_dev = nullptr;

  cudaStream_t stream = nullptr;
  cudaStream_t side_stream = nullptr;   // node refresh overlapped with a transform (api.cu: host_transform_refresh)
  int *h_flag = nullptr;                // pinned: result of the on-device node comparison
  cudaEvent_t ev_side = nullptr;
  bool own_stream = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t evk[2] = {nullptr, nullptr};   // around the main B / B^T kernel launch (opt_timing)
  float bkernel_ms = 0.f;
  bool evk_recorded = false;
  float stage_ms[3] = {0.f, 0.f, 0.f};
  int64_t launches = 0;

  int opt_timing = 0;
  int opt_psi_table = 0;
  int opt_b_kernel = 0;
  int opt_node_order = 0;
  int opt_b_flush = 0;
  int opt_fft_kernel = 0;           // 0 auto | 1 shared-memory Stockham only (fft_stockham_kernel)
  int opt_fft_prune = 1;            // band-pruned FFT passes + D without zero padding inside trafo / adjoint
  int sm_count = 148;
};

namespace nfftcu {

inline size_t real_size(const nfftcu_ctx *c) { return c->prec == NFFTCU_DOUBLE ? 8 : 4; }

// ---- stage entry points, one per translation unit ---------------------------------------------
int sort_nodes(nfftcu_ctx *c);                                      // sort.cu
int radix_sort_pairs(nfftcu_ctx *c, uint64_t *keys, uint32_t *vals, long long M, int bits);  // sort.cu
int gather_nodes(nfftcu_ctx *c, const uint32_t *perm, void *dst);   // sort.cu
constexpr int kKbPolyDeg = 16;                                      // stored degree of the window polynomials
int build_kb_poly(nfftcu_ctx *c);                                   // kbpoly.cu
bool tile3d_supported(const nfftcu_ctx *c);                         // tile3d.cu
int tile3d_bin_nodes(nfftcu_ctx *c);                                // tile3d.cu
int tile3d_interp(nfftcu_ctx *c, void *f_dev);                      // tile3d.cu
int tile3d_spread(nfftcu_ctx *c, const void *f_dev);                // tile3d.cu
bool tile2d_supported(const nfftcu_ctx *c);                         // tile2d.cu
int tile2d_bin_nodes(nfftcu_ctx *c);                                // tile2d.cu
int tile2d_interp(nfftcu_ctx *c, void *f_dev);                      // tile2d.cu
int tile2d_spread(nfftcu_ctx *c, const void *f_dev);                // tile2d.cu
bool mma3d_supported(const nfftcu_ctx *c);                          // mma3d.cu
int mma3d_bin_nodes(nfftcu_ctx *c);                                 // mma3d.cu
int mma3d_interp(nfftcu_ctx *c, void *f_dev);                       // mma3d.cu
int mma3d_spread(nfftcu_ctx *c, const void *f_dev);                 // mma3d.cu
int stage_D(nfftcu_ctx *c, const void *f_hat_dev, bool band_only = false);   // deconv.cu
int stage_DT(nfftcu_ctx *c, void *f_hat_dev);                       // deconv.cu
int fft_plan_axes(nfftcu_ctx *c);                                   // fft.cu
void fft_free_axes(nfftcu_ctx *c);                                  // fft.cu
int stage_F(nfftcu_ctx *c, int sign, bool pruned = false);          // fft.cu
int stage_B(nfftcu_ctx *c, void *f_dev);                            // interp.cu
int stage_BT(nfftcu_ctx *c, const void *f_dev, bool slab_ok = false);   // spread.cu; slab_ok: the caller runs the pruned F behind it
int build_psi_table(nfftcu_ctx *c);                                 // interp.cu
int ndft_trafo(nfftcu_ctx *c, const void *f_hat_dev, void *f_dev);  // ndft.cu
int ndft_adjoint(nfftcu_ctx *c, const void *f_dev, void *f_hat_dev);// ndft.cu
void peer_detach(nfftcu_ctx *c);                                    // peer.cu
int peer_attach_local(nfftcu_ctx **ctxs, int world, bool all_outputs);   // peer.cu
int peer_reduce_DT(nfftcu_ctx *c, void *f_hat_dev);                 // peer.cu
void *peer_slice_ptr(nfftcu_ctx *c, long long *k_begin, long long *k_end);   // peer.cu
int create_ctx(nfftcu_ctx **out, int precision, int d, const int64_t *N, const int64_t *n, int64_t m, int64_t M,
               unsigned flags, int device, bool nodes_only);        // api.cu
int nodes_ready(nfftcu_ctx *c);                                     // api.cu
int ensure_batch(nfftcu_ctx *c, int K);                             // api.cu: grow grid / f_tile for K right-hand sides
uint64_t fingerprint(const void *data, size_t bytes);               // api.cu
void plan_cache_clear();                                            // api.cu

// ---- Kaiser-Bessel window, evaluated in double for both precisions ----------------------------
// phi(t) with t = n*(x - l/n) the distance in grid units, s = m^2 - t^2:
//   s>0: sinh(b sqrt(s))/(pi sqrt(s)); s<0: sin(b sqrt(-s))/(pi sqrt(-s)); s==0: b/pi
// (include/infft.h:209-215 of the reference; NOT truncated outside |t|<=m).
__device__ __forceinline__ double kb_phi(double t, double m2, double b) {
  const double kInvPi = 0.31830988618379067153776752674502872;
  const double s = m2 - t * t;
  if (s > 0.0) {
    const double r = sqrt(s);
    return sinh(b * r) * kInvPi / r;
  }
  if (s < 0.0) {
    const double r = sqrt(-s);
    return sin(b * r) * kInvPi / r;
  }
  return b * kInvPi;
}

// window value at distance t (grid units) for the plan's window family, times the plan's power-of-two scale:
//   Kaiser-Bessel: kb_phi;   Gaussian (include/infft.h:154-157): exp(-t^2/b) / sqrt(pi b), b = 2 sigma/(2 sigma-1) m/pi
__device__ __forceinline__ double window_phi(double t, double m2, double b, int window, double scale) {
  if (window == NFFTCU_WINDOW_GAUSSIAN) return scale * exp(-t * t / b) * rsqrt(3.14159265358979323846264338327950288 * b);
  return scale * kb_phi(t, m2, b);
}

// c = floor(x*n) evaluated in the plan's precision exactly as the reference's uo()
// (nfft.c:324-332): one rounded multiply, then floor.
__device__ __forceinline__ long long cell_of(double x, long long n) {
  return (long long) floor(__dmul_rn(x, (double) n));
}
__device__ __forceinline__ long long cell_of(float x, long long n) {
  return (long long) floorf(__fmul_rn(x, (float) n));
}

}  // namespace nfftcu
