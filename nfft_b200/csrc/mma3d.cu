// mma3d.cu -- 3-D interpolation (B) and spreading (B^T) with the window contraction issued as tensor-core
// instructions: FP64 DMMA (mma.sync.m8n8k4.f64) for fp64 plans, 3xTF32 mma.sync.m16n8k8 for fp32 plans; the fast
// path for d = 3, m <= 6.
//
// Reference being replaced: nfft_trafo_3d_B / nfft_trafo_3d_compute (kernel/nfft/nfft.c:4687-4914,
// 4020-4265) and nfft_adjoint_3d_B with its atomic / blockwise compute variants (5126-5384,
// 4393-4436, 4289-4388).  Same arithmetic -- f_j = sum psi0 psi1 psi2 g, g += psi0 psi1 psi2 f_j -- in
// a different summation order.
//
// Why DMMA.  The taps of one node are a rank-1 (separable) weight on a (2m+2)^3 box; 2m+2 = 14 gives
// 5488 FP64 FMAs per node and direction, and the B200 FP64 pipe delivers the same 64 FMA/clk/SM
// whether it is fed by DFMA or by DMMA (profiles/r01w_microbench_dmma.txt: 36.9 TFLOP/s either way,
// the two share the pipe).  A DFMA tap loop needs one issue slot and 1/15 of a broadcast shared-memory
// load per 32 FMAs and ran at 30-44 % of the pipe (profiles/r01u); one DMMA carries 256 FMAs, takes its
// operands from registers and leaves the LSU and the issue slots almost idle.  The contraction becomes
// a small GEMM once nodes are processed in BATCHES of 8 that share a grid window:
//
//   tile (a,b) = (u0 / T, u1 / T), T = 17 - (2m+2)  => every tap box of the tile lies in a FOOTPRINT of
//   16 x 16 grid rows (pencils along z);  nodes of a tile are walked in ascending u2 (their lowest tap
//   in z), 8 at a time, and a batch only takes nodes with u2 in [zlo, zlo+2], zlo even, so that all its
//   taps fall in the WINDOW z in [zlo, zlo+16).
//
//   interpolation   V[z, node]  = sum_p G[p, z] * (psi0 psi1)[p, node]      p = 256 pencils, re / im separately
//                   f_node      = sum_z psi2[z, node] V[z, node]
//   spreading       G[p, z]    += sum_node (psi0 psi1)[p, node] * (psi2 f)[node, z]
//
// (The non-tensor FP64 work of the MMA warps is what the in-order warps have no room for, so both kernels keep it
// minimal: interpolation contracts the pencils first, with the row weights psi0 psi1 as the B operand -- 16
// products per lane and batch -- and applies psi2 in 8 FMAs, instead of contracting z first and forming 32
// weighted row sums per lane; spreading multiplies the samples into the B operand psi2 -- 8 products per lane --
// instead of the A operand psi0 psi1, whose fragments are then shared by the re / im accumulators.)
//
// A CTA of 4 MMA warps owns a tile; warp w owns footprint rows l0 = 4w..4w+3, all 16 l1 = 64 pencils.  The window
// lives in REGISTERS for the whole sweep of the tile along z: as A fragments for interpolation (64 doubles per lane,
// refilled two cells at a time from L2 behind the batch's DMMAs), as C accumulators for spreading (64 doubles per
// lane, retired two cells at a time with RED.ADD).  Window slots are circular (cell z lives in slot z mod 16), so
// sliding the window moves no registers.  Per batch a warp issues 64 DMMAs = 16384 FMAs for 8 nodes: lane
// efficiency (14/16)^3 = 67 % of the useful 5488 per node.
//
// Batches are formed once per node set (plan time, like the reference's precompute_psi): a table of
// (first node, count, window base) per batch, the batch range of every work unit, and the chunk list (runs of at
// most 384 batches of one unit: a CTA's load is bounded for clustered node sets).
//
// Warp specialisation.  A CTA has two warpgroups; setmaxnreg moves registers from the second to the MMA warps
// (200 / 56).  The second group delivers, per batch, the placed operand block psi0 / psi1 / psi2 -- zero-padded,
// offset in the footprint, circular slot in z -- into a ring of shared-memory stages guarded by full / empty
// mbarriers, and for interpolation adds up the partial sums the MMA warps release:
//   * FEEDER mode (default, when the images fit): the blocks were written once per node set by mma_images_kernel
//     (3 KB per batch) and one elected lane per batch issues a TMA bulk copy that completes the stage's full barrier;
//   * PRODUCER mode: the warps evaluate the window from the node coordinates -- the piecewise polynomials of
//     kbpoly.cu, a warp owns every 4th batch and runs its 12 Horner chains per lane interleaved.
//
// fp32 plans run the same structure on the TF32 tensor path (mma.sync.m16n8k8, 3xTF32 split; *_tf32_kernel below).
// Spreading has an optional flush through shared memory and TMA bulk reductions (cp.reduce.async.bulk .add.f64,
// FLUSH = 1, NFFTCU_OPT_B_FLUSH = 2); plain RED.ADD from the accumulator registers measured faster and is the default.
#include "common.cuh"
#include "mma3d.cuh"

namespace nfftcu {

#ifdef NFFTCU_DBG_CLOCKS
// timing probes (debug builds only, tools/dbg_clocks.py): clock64() section totals of MMA warp 0 of every CTA
__device__ unsigned long long g_dbg[16];
#endif

namespace {

constexpr int kNB = 8;        // nodes per batch
constexpr int kF = 16;        // footprint rows per axis, window slots
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int wrapi(int v, int n) {
  if (v < 0) v += n;
  if (v >= n) v -= n;
  if (v < 0 || v >= n) { v %= n; if (v < 0) v += n; }
  return v;
}

// volatile: keeps the issue order written in the kernels (independent accumulators interleaved; ptxas
// otherwise schedules the k-steps of one accumulator back to back and waits out the 26-cycle latency)
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

// c = floor(x * n) in the PLAN's precision (the binning keys use cell_of of common.cuh: same rounding);
// x is carried as a double that holds the stored value exactly
template <typename TS> __device__ __forceinline__ int cell_int(double x, int n);
template <> __device__ __forceinline__ int cell_int<double>(double x, int n) { return __double2int_rd(__dmul_rn(x, (double) n)); }
template <> __device__ __forceinline__ int cell_int<float>(double x, int n) { return __float2int_rd(__fmul_rn((float) x, (float) n)); }

__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(float *p, double v) { atomicAdd(p, (float) v); }

// ---- binning ---------------------------------------------------------------------------------------------
template <typename TS>
__global__ void mma_keys_kernel(const TS *__restrict__ x, uint64_t *__restrict__ keys,
                                uint32_t *__restrict__ vals, long long M, MmaParams P) {
  const long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int u0 = wrapi((int) (cell_of(x[3 * j], P.n0) - P.m), P.n0);
  const int u1 = wrapi((int) (cell_of(x[3 * j + 1], P.n1) - P.m), P.n1);
  const int u2 = wrapi((int) (cell_of(x[3 * j + 2], P.n2) - P.m), P.n2);
  const unsigned long long tile = (unsigned long long) (u0 / P.T) * P.NT1 + (u1 / P.T);
  keys[j] = tile * P.n2 + u2;
  vals[j] = (uint32_t) j;
}

// unit_start[u] = first position whose key >= first key of work unit u, u = 0..units
__global__ void mma_unit_bounds_kernel(const uint64_t *__restrict__ keys, uint32_t *__restrict__ unit_start,
                                       long long units, long long M, MmaParams P) {
  const long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (u > units) return;
  const long long tile = u / P.zseg, seg = u - tile * P.zseg;
  const uint64_t key = (uint64_t) tile * P.n2 + (uint64_t) ((long long) P.n2 * seg / P.zseg);
  long long lo = 0, hi = M;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  unit_start[u] = (uint32_t) lo;
}

template <typename C2>
__global__ void mma_gather_f_kernel(const C2 *__restrict__ f, const uint32_t *__restrict__ perm,
                                    C2 *__restrict__ ft, long long M) {
  const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (k < M) ft[k] = f[perm[k]];
}

// ---- batch table (plan time) ------------------------------------------------------------------------------
// entry.x = first node (tile order), entry.y = zlo | nb << 24 | last << 28
__device__ __forceinline__ int bt_zlo(uint2 e) { return (int) (e.y & 0xffffffu); }
__device__ __forceinline__ int bt_nb(uint2 e) { return (int) ((e.y >> 24) & 0xfu); }
__device__ __forceinline__ int bt_last(uint2 e) { return (int) ((e.y >> 28) & 1u); }

// One warp per work unit walks the unit's sorted nodes 32 at a time and cuts them into batches: a batch
// starts at the first unassigned node, zlo = its u2 rounded down to even, and takes up to 8 nodes with
// u2 <= zlo + 2.  FILL = false counts, FILL = true writes the entries at batch_start[unit].
template <bool FILL>
__global__ void mma_batches_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ unit_start,
                                   uint32_t *__restrict__ counts, const uint32_t *__restrict__ batch_start,
                                   uint2 *__restrict__ table, long long units, MmaParams P) {
  const long long unit = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (unit >= units) return;
  const long long k0 = unit_start[unit], k1 = unit_start[unit + 1];
  const uint64_t base = (uint64_t) (unit / P.zseg) * P.n2;
  const uint32_t out = FILL ? batch_start[unit] : 0;
  uint32_t nbat = 0;
  long long pos = k0;
  while (pos < k1) {
    const long long idx = pos + lane;
    const int u = idx < k1 ? (int) (keys[idx] - base) : 0x7fffffff;
    int o = 0;
    while (pos + o < k1) {
      if (o + kNB > 32 && pos + 32 < k1) break;   // the batch may extend beyond this chunk: reload from pos + o
      const int zlo = __shfl_sync(kFull, u, o) & ~1;
      const bool member = lane >= o && lane < o + kNB && u <= zlo + 2;
      const int nb = __popc(__ballot_sync(kFull, member));
      if (FILL && lane == 0) {
        const unsigned last = (pos + o + nb >= k1) ? 1u : 0u;
        table[out + nbat] = make_uint2((uint32_t) (pos + o), (unsigned) zlo | ((unsigned) nb << 24) | (last << 28));
      }
      nbat++;
      o += nb;
    }
    pos += o;
  }
  if (!FILL && lane == 0) counts[unit] = nbat;
}

// exclusive scan of counts[0..n) into out[0..n], single CTA
__global__ void mma_scan_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ out, long long n) {
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

// ---- work chunks (plan time) --------------------------------------------------------------------------------
// A CTA sweeps one chunk = a run of at most kChunkBatches consecutive batches of one work unit (tile, z segment).
// With evenly spread nodes a unit is one chunk; clustered node sets (radial / spiral trajectories: thousands of
// batches in the tiles around the centre) are cut into many, so that the load of a CTA is bounded.  A chunk is
// self-contained: interpolation loads the window of its first batch itself, spreading retires the whole window
// with reductions at its end.
constexpr int kChunkBatches = 384;

__global__ void mma_chunk_count_kernel(const uint32_t *__restrict__ batch_start, uint32_t *__restrict__ counts, long long units) {
  const long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= units) return;
  const uint32_t nb = batch_start[u + 1] - batch_start[u];
  counts[u] = (nb + kChunkBatches - 1) / kChunkBatches;
}

__global__ void mma_chunk_fill_kernel(const uint32_t *__restrict__ batch_start, const uint32_t *__restrict__ chunk_start,
                                      uint4 *__restrict__ chunks, long long units, int zseg) {
  const long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= units) return;
  const uint32_t b0 = batch_start[u], nb = batch_start[u + 1] - b0;
  const uint32_t cnt = chunk_start[u + 1] - chunk_start[u];
  if (cnt == 0) return;
  const uint32_t size = (nb + cnt - 1) / cnt;   // equal shares
  for (uint32_t k = 0; k < cnt; k++) {
    const uint32_t lo = b0 + k * size, hi = (k + 1 == cnt) ? b0 + nb : lo + size;
    chunks[chunk_start[u] + k] = make_uint4((uint32_t) (u / zseg), lo, hi, 0u);
  }
}

// ---- mbarrier helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// Both waits use the hardware-suspended form of try_wait (suspend-time hint): the warp is parked by the barrier unit and
// wakes when the phase completes, instead of polling.  An SM sub-partition has ONE dispatch port and a DMMA holds it
// for 16 cycles, so every polling instruction of a waiting warp is taken from the DMMA stream of its neighbours.
#ifndef NFFTCU_MMA_SUSPEND_NS
#define NFFTCU_MMA_SUSPEND_NS 0x989680
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity), "r"(NFFTCU_MMA_SUSPEND_NS) : "memory");
}

// non-blocking phase test: issued early, its latency overlaps the code up to the point where the result is needed
__device__ __forceinline__ uint32_t mbar_test(uint64_t *bar, int parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  return done;
}

// producer-side wait: the producers run up to kStages batches ahead and mostly wait
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, int parity) { mbar_wait(bar, parity); }

// ---- shared memory ----------------------------------------------------------------------------------------
// A ring of kStages operand blocks; producer warp w fills the stages of batches j = w (mod 4).
// ops[v][slot][node]: v = 0: psi0 placed in the footprint, 1: psi1 (interpolation) or psi1*f.re
// (spreading), 2: psi2 in circular window slots, 3: psi1*f.im (spreading only)
constexpr int kStages = 8;
constexpr int kCoefK = kKbPolyDeg + 2;   // powers 0..17 per tap: rows of 144 bytes, (even, odd) pairs 16-byte aligned
constexpr int kMmaRegs = 200, kProdRegs = 56;   // setmaxnreg: 128 * (200 + 56) = 2 CTAs per SM

template <int W, bool SPREAD>
struct Shared {
  double ops[kStages][SPREAD ? 4 : 3][kF][kNB];
  double red[kStages][SPREAD ? 1 : 4][8][4][4];   // interpolation: [MMA warp][nr][kq][re0, im0, re1, im1] partial sums
  uint2 meta[kStages];               // batch table entry of the stage's batch
  uint64_t full[kStages], empty[kStages];
  double coef[3 * W * kCoefK];       // [t][tap][power], powers padded with zeros to kCoefK
  unsigned rowoff[kF * kF];
};

// Producer warp pw (0..3) of the CTA: batches j = pw, pw+4, ... of the unit.  Lane = (node i = lane & 7,
// slot quarter qg = lane >> 3): 4 slots x 3 dimensions = 12 independent Horner chains per lane.
// TF32 = true (fp32 plans on the TF32 tensor path): the ring entries keep their 8-byte slots but hold fp32 data, so
// that the conversions happen once per value here and not once per MMA warp: psi2 as the (hi, lo) tf32 pair of the
// 3xTF32 split, every other operand as one float in the low word.
__device__ __forceinline__ double pack_tf32_pair(double v) {
  const unsigned hi = (__float_as_uint((float) v) + 0x1000u) & 0xffffe000u;   // see split_tf32
  const unsigned lo = __float_as_uint((float) (v - (double) __uint_as_float(hi)));
  return __hiloint2double((int) lo, (int) hi);   // low word = hi part, high word = lo part
}
__device__ __forceinline__ double pack_f32(double v) { return __hiloint2double(0, __float_as_int((float) v)); }

constexpr int kImgDoubles = 3 * kF * kNB;   // one batch image: psi0, psi1, psi2 placed = 3072 bytes

// interpolation: the batch in stage st has been released by all MMA warps -- add up the 128 per-lane partial sums
template <typename TS, int W, bool SPREAD>
__device__ __forceinline__ void finalize_batch(Shared<W, SPREAD> &S, int st, const uint32_t *__restrict__ perm,
                                               TS *__restrict__ f, int lane) {
  if (!SPREAD) {
    // output o = 2*node + comp lives at [w][nr][node >> 1][2*(node & 1) + comp] = 16 consecutive doubles per
    // (w, nr); lanes o and o+16 each sum 16 of the 32 partials
    const uint2 mt = S.meta[st];
    const double *rp = &S.red[st][0][0][0][0] + (lane & 15) + (lane >> 4) * 16 * 16;
    double v0 = 0.0, v1 = 0.0;
#pragma unroll
    for (int q = 0; q < 16; q += 2) { v0 += rp[q * 16]; v1 += rp[(q + 1) * 16]; }
    v0 += v1;
    v0 += __shfl_xor_sync(kFull, v0, 16);
    if (lane < 2 * kNB && (lane >> 1) < bt_nb(mt)) f[2 * (size_t) perm[mt.x + (lane >> 1)] + (lane & 1)] = (TS) v0;
    __syncwarp();
  }
}

// Feeder warp pw (0..3) of a CTA that runs on precomputed window images: per batch one elected lane hands the
// 3072-byte image to the TMA unit, which lands it in the ring stage and completes the stage's full barrier.
template <typename TS, int W, bool SPREAD>
__device__ __forceinline__ void feeder_loop(Shared<W, SPREAD> &S, const double *__restrict__ img,
                                            const uint32_t *__restrict__ perm, TS *__restrict__ f,
                                            const uint2 *__restrict__ table, int nbat, int pw, int lane) {
  for (int j = pw; j < nbat; j += 4) {
    const int st = j % kStages;
    if (j >= kStages) {
      mbar_wait(&S.empty[st], ((j / kStages) - 1) & 1);
      finalize_batch<TS, W, SPREAD>(S, st, perm, f, lane);
    }
    if (lane == 0) {
      S.meta[st] = table[j];
      // the MMA warps' (generic proxy) reads of the stage are ordered before this point by the empty barrier; order
      // them before the async-proxy write as well
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&S.full[st])),
                   "r"((unsigned) (kImgDoubles * sizeof(double))) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_addr(&S.ops[st][0][0][0])), "l"(img + (size_t) j * kImgDoubles),
                     "r"((unsigned) (kImgDoubles * sizeof(double))), "r"(smem_addr(&S.full[st])) : "memory");
    }
    __syncwarp();
  }
  if (!SPREAD) {
    for (int jj = pw; jj < nbat; jj += 4) {
      if (jj + kStages < nbat) continue;
      const int st = jj % kStages;
      mbar_wait(&S.empty[st], (jj / kStages) & 1);
      finalize_batch<TS, W, SPREAD>(S, st, perm, f, lane);
    }
  }
}

// IMAGE = true: plan-time build of the window images -- the same evaluation and placement, written to img[j] in
// global memory instead of the ring (no barriers, no finalisation; always the SPREAD = false operand set)
template <typename TS, int W, bool SPREAD, bool TF32 = false, bool IMAGE = false>
__device__ __forceinline__ void producer_loop(Shared<W, SPREAD> &S, const TS *__restrict__ xt,
                                              const typename Cplx<TS>::type *__restrict__ ft,
                                              const uint32_t *__restrict__ perm,
                                              TS *__restrict__ f, const uint2 *__restrict__ table,
                                              int nbat, int a, int bt, const MmaParams &P, int pw, int lane,
                                              double *__restrict__ img = nullptr) {
  const int i = lane & 7, qg = lane >> 3;
  auto finalize = [&](int st) { finalize_batch<TS, W, SPREAD>(S, st, perm, f, lane); };
  // the prefetched values stay in the storage type until they are used: a conversion right behind the load would
  // make the warp wait for the load at once
  auto load_nodes = [&](uint2 mt, TS &xv, TS &fv) {
    const int nb = bt_nb(mt);
    xv = (lane < 3 * nb) ? xt[3 * (size_t) mt.x + lane] : (TS) 0;
    if (SPREAD) fv = (lane < 2 * nb) ? reinterpret_cast<const TS *>(ft)[2 * (size_t) mt.x + lane] : (TS) 0;
  };
  if (pw >= nbat) return;
  uint2 mt = table[pw], mt_next = make_uint2(0, 0);
  TS xv, fv = 0, xv_next = 0, fv_next = 0;
  load_nodes(mt, xv, fv);
  if (pw + 4 < nbat) mt_next = table[pw + 4];
  for (int j = pw; j < nbat; j += 4) {
    const int st = j % kStages;
    // prefetch: nodes of this warp's next batch, table entry of the one after
    uint2 mt_next2 = make_uint2(0, 0);
    if (j + 4 < nbat) load_nodes(mt_next, xv_next, fv_next);
    if (j + 8 < nbat) mt_next2 = table[j + 8];
    if (!IMAGE && j >= kStages) {
      mbar_wait_sleep(&S.empty[st], ((j / kStages) - 1) & 1);
      finalize(st);
    }
    double *const ops = IMAGE ? img + (size_t) j * kImgDoubles : &S.ops[st][0][0][0];   // [v][slot][node]
    const int nb = bt_nb(mt);
    const bool live = i < nb;
    double fr = 0.0, fi = 0.0;
    if (SPREAD) { fr = (double) __shfl_sync(kFull, fv, 2 * i); fi = (double) __shfl_sync(kFull, fv, 2 * i + 1); }
#pragma unroll
    for (int t = 0; t < 3; t++) {
      const double x = (double) __shfl_sync(kFull, xv, 3 * i + t);
      const int nt = t == 0 ? P.n0 : t == 1 ? P.n1 : P.n2;
      const int c = cell_int<TS>(x, nt);
      const int u = wrapi(c - P.m, nt);
      const double y = 2.0 * (x * (double) nt - (double) c) - 1.0;
      const int lo = t == 0 ? u - P.T * a : t == 1 ? u - P.T * bt : u;   // psi_t[l] goes to slot (lo + l) & 15
      // p_l(y) = E_l(y^2) + y O_l(y^2): two half-length Horner chains, one 16-byte load per step.  The window is
      // even, so the mirrored tap is p_{W-1-l}(y) = p_l(-y) = E_l - y O_l: a lane evaluates taps l = 2qg, 2qg+1
      // (< W/2) and writes both; the lanes whose l >= W/2 write the 16-W zero slots behind the taps.
      const double2 *cft = reinterpret_cast<const double2 *>(S.coef + t * W * kCoefK);
      const double y2 = y * y;
      double E[2], O[2];
      int off[2];
#ifdef NFFTCU_DBG_DEG2
      const int ptop = 1;   // timing experiment only: wrong window
#else
      const int ptop = P.deg >> 1;
#endif
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const int l = 2 * qg + r;
        off[r] = (l < W / 2 ? l : 0) * (kCoefK / 2);
        const double2 c = cft[off[r] + ptop];
        E[r] = c.x;
        O[r] = c.y;
      }
#pragma unroll 2
      for (int p = ptop - 1; p >= 0; p--) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const double2 c = cft[off[r] + p];
          E[r] = fma(E[r], y2, c.x);
          O[r] = fma(O[r], y2, c.y);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const int l = 2 * qg + r;
        const bool tap = l < W / 2;
        const int zz = l - W / 2;
        if (!tap && 2 * zz >= kF - W) continue;
        const int sa = (tap ? lo + l : lo + W + 2 * zz) & (kF - 1);
        const int sb = (tap ? lo + W - 1 - l : lo + W + 2 * zz + 1) & (kF - 1);
        const double yo = y * O[r];
        const double va = (tap && live) ? E[r] + yo : 0.0, vb = (tap && live) ? E[r] - yo : 0.0;
        if (SPREAD && t == 1) {
          ops[(1 * kF + sa) * kNB + i] = TF32 ? pack_f32(va * fr) : va * fr;
          ops[(1 * kF + sb) * kNB + i] = TF32 ? pack_f32(vb * fr) : vb * fr;
          ops[(3 * kF + sa) * kNB + i] = TF32 ? pack_f32(va * fi) : va * fi;
          ops[(3 * kF + sb) * kNB + i] = TF32 ? pack_f32(vb * fi) : vb * fi;
        } else if (TF32) {
          ops[(t * kF + sa) * kNB + i] = t == 2 ? pack_tf32_pair(va) : pack_f32(va);
          ops[(t * kF + sb) * kNB + i] = t == 2 ? pack_tf32_pair(vb) : pack_f32(vb);
        } else {
          ops[(t * kF + sa) * kNB + i] = va;
          ops[(t * kF + sb) * kNB + i] = vb;
        }
      }
    }
    if (!IMAGE) {
      if (lane == 0) S.meta[st] = mt;
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.full[st]);
    }
    mt = mt_next; mt_next = mt_next2;
    xv = xv_next; fv = fv_next;
  }
  if (!SPREAD && !IMAGE) {
    for (int jj = pw; jj < nbat; jj += 4) {
      if (jj + kStages < nbat) continue;   // finished in the loop, when the stage was refilled
      const int st = jj % kStages;
      mbar_wait_sleep(&S.empty[st], (jj / kStages) & 1);
      finalize(st);
    }
  }
}

#define NFFTCU_MMA_PROLOGUE(SPREADV, TF32V)                                                                \
  constexpr int T = kF + 1 - W;                                                                       \
  extern __shared__ __align__(128) unsigned char smem_raw[];                                          \
  Shared<W, SPREADV> &S = *reinterpret_cast<Shared<W, SPREADV> *>(smem_raw);                          \
  const int tid = threadIdx.x;                                                                        \
  const int n2 = P.n2;                                                                                \
  const uint4 chunk = chunks[blockIdx.x];   /* tile, first batch, end batch */                        \
  const uint32_t b0 = chunk.y;                                                                        \
  const int nbat = (int) (chunk.z - b0);                                                              \
  if (nbat == 0) return;                                                                              \
  table += b0;                                                                                        \
  const int tile = (int) chunk.x;                                                                     \
  const int a = tile / P.NT1, bt = tile - a * P.NT1;                                                  \
  for (int i = tid; i < 3 * W * kCoefK; i += 256) {                                                   \
    const int t = i / (W * kCoefK), l = (i / kCoefK) % W, k = i % kCoefK;                             \
    S.coef[i] = k <= kKbPolyDeg ? poly[(t * (kKbPolyDeg + 1) + k) * W + l] : 0.0;                     \
  }                                                                                                   \
  {                                                                                                   \
    const int l0 = tid >> 4, l1 = tid & 15;                                                           \
    S.rowoff[tid] = (unsigned) ((wrapi(T * a + l0, P.n0) * (long long) P.n1 + wrapi(T * bt + l1, P.n1)) * n2); \
  }                                                                                                   \
  if (tid == 0) {                                                                                     \
    for (int st = 0; st < kStages; st++) { mbar_init(&S.full[st], 1); mbar_init(&S.empty[st], 4); }   \
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                                \
  }                                                                                                   \
  __syncthreads();                                                                                    \
  if (tid >= 128) {                                                                                   \
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProdRegs));                             \
    if (IMG) feeder_loop<TS, W, SPREADV>(S, P.img + (size_t) b0 * kImgDoubles, perm, f, table, nbat, (tid >> 5) - 4, tid & 31); \
    else producer_loop<TS, W, SPREADV, TF32V>(S, xt, ft, perm, f, table, nbat, a, bt, P, (tid >> 5) - 4, tid & 31); \
    return;                                                                                           \
  }                                                                                                   \
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kMmaRegs));                                \
  const int lane = tid & 31, warp = tid >> 5;                                                         \
  const int kq = lane & 3, nr = lane >> 2;                                                            \
  const unsigned *const rowoff_s = S.rowoff + 4 * warp * kF + nr;   /* group g: [(g >> 1) * kF + 8 * (g & 1)] */

// ---- interpolation ------------------------------------------------------------------------------------
template <typename TS, int W, bool IMG>
__global__ void __launch_bounds__(256, 2)
interp_mma_kernel(const typename Cplx<TS>::type *__restrict__ G, const TS *__restrict__ xt,
                  const uint32_t *__restrict__ perm, TS *__restrict__ f,
                  const uint4 *__restrict__ chunks, const uint2 *__restrict__ table,
                  const double *__restrict__ poly, MmaParams P) {
  const typename Cplx<TS>::type *const ft = nullptr;
  NFFTCU_MMA_PROLOGUE(false, false)

  // Contraction order: the pencils first, V[z, node] = sum_p G[p, z] w[p, node] with the row weights
  // w = psi0[l0] psi1[l1] as the B operand (16 products per lane and batch), then f_node = sum_z psi2[z, node] V[z, node]
  // in 8 FMAs per lane.  (The other order -- z first, then 32 weighted row sums per lane -- costs three times the
  // non-tensor FP64 instructions, and those are what the FP64 pipe has no room for.)  The warp owns the 64 pencils
  // p = 0..63 of its four footprint rows, l0 = 4 warp + (p >> 4), l1 = p & 15.  DMMA m8n8k4: rows = 8 window cells,
  // k = 4 pencils, columns = 8 nodes; k-step s covers pencils 4s .. 4s+3.
  // A[s][mt][re/im]: grid value of pencil 4s+kq at the window cell in slot 8*mt+nr (cell z sits in slot z mod 16)
  double A[16][2][2];
  int zwin = -1000;    // window base (even); the window holds cells [zwin, zwin+16)
  unsigned poff[16];   // grid offsets of the lane's pencils 4s+kq, kept in registers for the whole sweep
#pragma unroll
  for (int s = 0; s < 16; s++) poff[s] = S.rowoff[4 * warp * kF + 4 * s + kq];

  auto fill_all = [&](int zlo) {
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      int z = zlo + ((8 * mt + nr - zlo) & 15);
      if (z >= n2) z -= n2;
#pragma unroll
      for (int s = 0; s < 16; s++) {
        const typename Cplx<TS>::type v = G[poff[s] + z];
        A[s][mt][0] = (double) v.x;
        A[s][mt][1] = (double) v.y;
      }
    }
  };
  // pair (zp, zp+1), zp even: cell zp + (nr & 1) belongs to the lanes with (nr >> 1) == (zp & 7) >> 1, m-tile (zp >> 3) & 1
  auto load_pair = [&](int zp) {
    if ((nr >> 1) == ((zp & 7) >> 1)) {
      int z = zp + (nr & 1);
      if (z >= n2) z -= n2;
      if (((zp >> 3) & 1) == 0) {
#pragma unroll
        for (int s = 0; s < 16; s++) {
          const typename Cplx<TS>::type v = G[poff[s] + z];
          A[s][0][0] = (double) v.x;
          A[s][0][1] = (double) v.y;
        }
      } else {
#pragma unroll
        for (int s = 0; s < 16; s++) {
          const typename Cplx<TS>::type v = G[poff[s] + z];
          A[s][1][0] = (double) v.x;
          A[s][1][1] = (double) v.y;
        }
      }
    }
  };
  auto advance_to = [&](int zlo) {
    if (zlo - zwin >= kF || zwin < 0) {
      fill_all(zlo);
      zwin = zlo;
    } else {
      while (zwin < zlo) { load_pair(zwin + kF); zwin += 2; }
    }
  };

#ifdef NFFTCU_DBG_CLOCKS
  long long tw = 0, tm = 0, tr = 0, tt = 0, t_all = clock64();
#define DBG_T(var) const long long var = clock64()
#else
#define DBG_T(var)
#endif
  uint2 e_next = table[0];
  for (int j = 0; j < nbat; j++) {
    const int st = j % kStages;
    const int zlo = bt_zlo(e_next);
    const uint32_t ready = mbar_test(&S.full[st], (j / kStages) & 1);   // normally complete: the feeders run stages ahead
    if (j + 1 < nbat) e_next = table[j + 1];   // lands during this batch: the next window base
    if (zlo != zwin) advance_to(zlo);
    DBG_T(c0);
    if (!ready) mbar_wait(&S.full[st], (j / kStages) & 1);
    DBG_T(c1);

    // ---- row weights of node nr: w[s] = psi0[l0 = 4 warp + (s >> 2)] psi1[l1 = 4 (s & 3) + kq]
    double q0[4], q1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      q0[i] = S.ops[st][0][4 * warp + i][nr];
      q1[i] = S.ops[st][1][4 * i + kq][nr];
    }
    const double2 p2a = *reinterpret_cast<const double2 *>(&S.ops[st][2][nr][2 * kq]);       // psi2, slot nr, nodes 2kq, 2kq+1
    const double2 p2b = *reinterpret_cast<const double2 *>(&S.ops[st][2][8 + nr][2 * kq]);   // slot 8 + nr
    double c[2][2][2];   // [m-tile][re/im][col]: V[slot 8 mt + nr][node 2 kq + col]
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int cc = 0; cc < 2; cc++) c[mt][cc][0] = c[mt][cc][1] = 0.0;
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const double w = q0[s >> 2] * q1[s & 3];
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int cc = 0; cc < 2; cc++) dmma(c[mt][cc][0], c[mt][cc][1], A[s][mt][cc], w);
    }
    double accr0 = p2a.x * c[0][0][0], accr1 = p2a.y * c[0][0][1];
    double acci0 = p2a.x * c[0][1][0], acci1 = p2a.y * c[0][1][1];
    accr0 = fma(p2b.x, c[1][0][0], accr0); accr1 = fma(p2b.y, c[1][0][1], accr1);
    acci0 = fma(p2b.x, c[1][1][0], acci0); acci1 = fma(p2b.y, c[1][1][1], acci1);
    DBG_T(c2);
    // the window registers are free again: slide the window to the next batch now, so that the refill
    // loads fly while this batch is being reduced
    if (j + 1 < nbat && bt_zlo(e_next) != zwin) advance_to(bt_zlo(e_next));
    DBG_T(c3);
    // per-lane partial sums (over the lane's window cell pair and the warp's pencils) go to the ring; the producer /
    // feeder warp that refills the stage adds them up
    *reinterpret_cast<double2 *>(&S.red[st][warp][nr][kq][0]) = make_double2(accr0, acci0);
    *reinterpret_cast<double2 *>(&S.red[st][warp][nr][kq][2]) = make_double2(accr1, acci1);
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[st]);
#ifdef NFFTCU_DBG_CLOCKS
    { const long long c4 = clock64(); tw += c1 - c0; tm += c2 - c1; tr += c3 - c2; tt += c4 - c3; }
#endif
  }
#ifdef NFFTCU_DBG_CLOCKS
  if (warp == 0 && lane == 0) {
    atomicAdd(&g_dbg[0], (unsigned long long) tw); atomicAdd(&g_dbg[1], (unsigned long long) tm);
    atomicAdd(&g_dbg[2], (unsigned long long) tr); atomicAdd(&g_dbg[3], (unsigned long long) tt);
    atomicAdd(&g_dbg[4], (unsigned long long) (clock64() - t_all)); atomicAdd(&g_dbg[5], (unsigned long long) nbat);
  }
#endif
}

// ---- spreading ------------------------------------------------------------------------------------------
// FLUSH = 0: retired cells go to the grid with RED.ADD straight from the accumulator registers
// FLUSH = 1: staged per warp in shared memory as 128-byte runs and reduced into the grid by the TMA unit
constexpr int kStgRow = 9;   // staging row pitch in 16-byte cells: 8 cells + 1 pad (bank-conflict-free, 16-byte aligned)

template <typename TS, int W, int FLUSH, bool IMG>
__global__ void __launch_bounds__(256, 2)
spread_mma_kernel(typename Cplx<TS>::type *__restrict__ G, const TS *__restrict__ xt,
                  const typename Cplx<TS>::type *__restrict__ ft,
                  const uint4 *__restrict__ chunks, const uint2 *__restrict__ table,
                  const double *__restrict__ poly, MmaParams P) {
  const uint32_t *const perm = nullptr;
  TS *const f = nullptr;
  NFFTCU_MMA_PROLOGUE(true, false)
  // staging: [buffer][warp][64 rows][kStgRow] double2 behind the Shared block
  double2 *const stg_base = reinterpret_cast<double2 *>(smem_raw + ((sizeof(Shared<W, true>) + 127) & ~(size_t) 127));
  double2 *const stg_w = stg_base + (size_t) warp * 64 * kStgRow;

  double C[8][2][2][2];   // [group][re/im][n-tile][col]: accumulator of pencil (group, nr), slot 8*nt + 2*kq + col
#pragma unroll
  for (int g = 0; g < 8; g++)
#pragma unroll
    for (int cc = 0; cc < 2; cc++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) C[g][cc][nt][0] = C[g][cc][nt][1] = 0.0;

  // staging state (uniform across the warp)
  int sblk = -1;      // 8-cell block (wrapped z >> 3) being staged, -1: none
  int snext = 0;      // next pair position (0,2,4,6) of the block that has not been written
  bool sbusy = false; // the TMA unit may still be reading the staging rows (last flush not waited for)
  TS *const Gd = reinterpret_cast<TS *>(G);

  auto stage_store = [&](int pos, bool zero, int nt) {   // pair position pos (even) of the staged block
    if (sbusy) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      sbusy = false;
    }
    if (kq == (pos >> 1)) {
      double2 *dst = stg_w + (size_t) nr * kStgRow + pos;
#pragma unroll
      for (int g = 0; g < 8; g++) {
        double2 v0 = make_double2(0.0, 0.0), v1 = v0;
        if (!zero) {
          if (nt == 0) { v0 = make_double2(C[g][0][0][0], C[g][1][0][0]); v1 = make_double2(C[g][0][0][1], C[g][1][0][1]); }
          else { v0 = make_double2(C[g][0][1][0], C[g][1][1][0]); v1 = make_double2(C[g][0][1][1], C[g][1][1][1]); }
        }
        dst[(size_t) g * 8 * kStgRow] = v0;
        dst[(size_t) g * 8 * kStgRow + 1] = v1;
      }
    }
  };
  auto flush_block = [&]() {   // complete the staged block with zeros and hand its 64 rows to the TMA unit
    while (snext < 8) { stage_store(snext, true, 0); snext += 2; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const int row = 2 * lane + rr;   // warp-local row = g*8 + pencil
      const int g = row >> 3, pn = row & 7;
      const unsigned off = S.rowoff[(4 * warp + (g >> 1)) * kF + 8 * (g & 1) + pn];
      const double2 *src = stg_w + (size_t) row * kStgRow;
      void *dst = G + off + 8 * sblk;   // FLUSH = 1 is instantiated for TS = double only
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 128;"
                   ::"l"(dst), "r"(smem_addr(src)) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    sbusy = true;   // waited for before the rows are written again, normally a batch or more later
    sblk = -1;
    snext = 0;
  };
  unsigned goff[8];   // grid offsets of the lane's pencils (group g, nr), kept in registers for the whole sweep
#pragma unroll
  for (int g = 0; g < 8; g++) goff[g] = rowoff_s[(g >> 1) * kF + 8 * (g & 1)];
  // retire pair (zp, zp+1) of the window (zp even, unwrapped): n-tile (zp>>3)&1, lanes kq == (zp&7)>>1
  auto retire_pair = [&](int zp) {
    int zw = zp;
    if (zw >= n2) zw -= n2;
    const int nt = (zp >> 3) & 1;
    if (FLUSH == 1) {
      const int blk = zw >> 3, pos = zw & 7;
      if (blk != sblk) {
        if (sblk >= 0) flush_block();
        sblk = blk;
        snext = 0;
      }
      while (snext < pos) { stage_store(snext, true, 0); snext += 2; }
      stage_store(pos, false, nt);
      snext = pos + 2;
      if (snext == 8) flush_block();
    }
    if (kq == ((zp & 7) >> 1)) {
#pragma unroll
      for (int g = 0; g < 8; g++) {
        if (FLUSH == 0) {
          TS *dst = Gd + 2 * ((size_t) goff[g] + zw);
          if (nt == 0) {
            red_add(dst, C[g][0][0][0]); red_add(dst + 1, C[g][1][0][0]);
            red_add(dst + 2, C[g][0][0][1]); red_add(dst + 3, C[g][1][0][1]);
          } else {
            red_add(dst, C[g][0][1][0]); red_add(dst + 1, C[g][1][1][0]);
            red_add(dst + 2, C[g][0][1][1]); red_add(dst + 3, C[g][1][1][1]);
          }
        }
        if (nt == 0) { C[g][0][0][0] = C[g][1][0][0] = C[g][0][0][1] = C[g][1][0][1] = 0.0; }
        else { C[g][0][1][0] = C[g][1][1][0] = C[g][0][1][1] = C[g][1][1][1] = 0.0; }
      }
    }
  };

  int zwin = -1;
  uint2 e_next = table[0];
  for (int j = 0; j < nbat; j++) {
    const int st = j % kStages;
    const int zlo = bt_zlo(e_next);
    const uint32_t ready = mbar_test(&S.full[st], (j / kStages) & 1);   // result needed behind the window slide
    double fr[2] = {0.0, 0.0}, fi[2] = {0.0, 0.0};   // IMG: samples of nodes kq and 4 + kq (the image holds psi1, not psi1 f)
    if (IMG) {
#pragma unroll
      for (int ks = 0; ks < 2; ks++)
        if (4 * ks + kq < bt_nb(e_next)) {
          const typename Cplx<TS>::type v = ft[(size_t) e_next.x + 4 * ks + kq];
          fr[ks] = (double) v.x;
          fi[ks] = (double) v.y;
        }
    }
    if (j + 1 < nbat) e_next = table[j + 1];
    // ---- slide the window to zlo: the cells below it are final
    if (zwin < 0) zwin = zlo;
    if (zlo != zwin) {
      const int zend = (zlo - zwin >= kF) ? zwin + kF : zlo;
      for (int zp = zwin; zp < zend; zp += 2) retire_pair(zp);
      zwin = zlo;
    }
    // ---- G += (psi0 psi1 f) * psi2
    if (!ready) mbar_wait(&S.full[st], (j / kStages) & 1);
    double bf[2][2];   // [n-tile][k-step]
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int ks = 0; ks < 2; ks++) bf[nt][ks] = S.ops[st][2][8 * nt + nr][4 * ks + kq];
    double p1r[2][2], p1i[2][2];   // [half][k-step]: psi1 f.re, psi1 f.im (IMG: psi1 in p1r, the samples go into psi2 instead)
#pragma unroll
    for (int hh = 0; hh < 2; hh++)
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        p1r[hh][ks] = S.ops[st][1][8 * hh + nr][4 * ks + kq];
        p1i[hh][ks] = IMG ? 0.0 : S.ops[st][3][8 * hh + nr][4 * ks + kq];
      }
    // IMG: the sample of node k multiplies the B operand psi2[k][z] (8 products per lane) instead of the A operand
    // psi0 psi1 [row][k] (32 per lane): G += (psi0 psi1) * (psi2 f.re), (psi0 psi1) * (psi2 f.im)
    double bfi[2][2];
    if (IMG) {
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) { bfi[nt][ks] = bf[nt][ks] * fi[ks]; bf[nt][ks] *= fr[ks]; }
    }
    double p0[4][2];
#pragma unroll
    for (int h = 0; h < 4; h++) { p0[h][0] = S.ops[st][0][4 * warp + h][kq]; p0[h][1] = S.ops[st][0][4 * warp + h][4 + kq]; }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[st]);   // operands are in registers: the stage can be refilled
#pragma unroll
    for (int h = 0; h < 4; h++) {
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        const int g = 2 * h + hh;
        const double ar0 = p0[h][0] * p1r[hh][0], ar1 = p0[h][1] * p1r[hh][1];
        const double ai0 = IMG ? ar0 : p0[h][0] * p1i[hh][0], ai1 = IMG ? ar1 : p0[h][1] * p1i[hh][1];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          dmma(C[g][0][nt][0], C[g][0][nt][1], ar0, bf[nt][0]);
          dmma(C[g][1][nt][0], C[g][1][nt][1], ai0, IMG ? bfi[nt][0] : bf[nt][0]);
        }
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          dmma(C[g][0][nt][0], C[g][0][nt][1], ar1, bf[nt][1]);
          dmma(C[g][1][nt][0], C[g][1][nt][1], ai1, IMG ? bfi[nt][1] : bf[nt][1]);
        }
      }
    }
  }
  for (int zp = zwin; zp < zwin + kF; zp += 2) retire_pair(zp);
  if (FLUSH == 1) {
    if (sblk >= 0) flush_block();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ---- fp32 plans: the same contraction on the TF32 tensor path, split 3 ways ------------------------------------
// For nfftf_ plans the grid, the samples and the result are fp32, so the FP64 pipe is not needed for the
// contraction itself: mma.sync.m16n8k8 TF32 with fp32 accumulation runs at 277 TFLOP/s on this part
// (profiles/r03i_microbench_tf32.txt), 7.4x the DMMA rate.  TF32 keeps 11 significant bits, so every operand is split
// x = hi + lo (hi = tf32(x), lo = tf32(x - hi): 22 bits) and a product is formed as lo*hi + hi*lo + hi*hi
// (3 MMAs; the dropped lo*lo term is 2^-22 relative): 1.4e-7 rel-l2 on a 16-term window contraction against
// fp64, i.e. fp32 accuracy.  One m16n8k8 covers exactly a 2 x 2 block of the m8n8k4 tiles of the fp64 kernels
// (two pencil groups, two k4 steps), so the register window, the producers (window values still evaluated in
// double and rounded once), the operand ring and the batch table are shared with them.
// hi = x rounded to TF32 (nearest, on the 13 dropped mantissa bits, as two integer instructions instead of the
// multi-instruction cvt.rna.tf32 sequence); lo = x - hi is exact in fp32 and is handed to the tensor core as it is:
// an fp32 bit pattern is a valid TF32 operand whose low 13 mantissa bits the tensor core does not read.
__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3,
                                         unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int W, bool IMG>
__global__ void __launch_bounds__(256, 2)
interp_tf32_kernel(const float2 *__restrict__ G, const float *__restrict__ xt,
                   const uint32_t *__restrict__ perm, float *__restrict__ f,
                   const uint4 *__restrict__ chunks, const uint2 *__restrict__ table,
                   const double *__restrict__ poly, MmaParams P) {
  typedef float TS;
  const float2 *const ft = nullptr;
  NFFTCU_MMA_PROLOGUE(false, true)

  unsigned Ah[8][2][4], Al[8][2][4];   // [group][re/im][slot]: tf32 hi / lo of the grid value, layout of interp_mma_kernel
  int zwin = -1000;

  auto fill_all = [&](int zlo) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
      int z = zlo + ((4 * s + kq - zlo) & 15);
      if (z >= n2) z -= n2;
#pragma unroll
      for (int g = 0; g < 8; g++) {
        const float2 v = G[rowoff_s[(g >> 1) * kF + 8 * (g & 1)] + z];
        split_tf32(v.x, Ah[g][0][s], Al[g][0][s]);
        split_tf32(v.y, Ah[g][1][s], Al[g][1][s]);
      }
    }
  };
  // refill loads are parked as raw float2 and split into (hi, lo) only when the next batch starts: a conversion
  // right behind the load would make the warp wait out the L2 latency at once instead of during the reduction
  float2 pend[8];
  int pend_slot = -1;
  auto commit = [&]() {
    if (pend_slot < 0) return;
    switch (pend_slot) {
#define NFFTCU_COMMIT(SL)                                                                       \
      case SL:                                                                                  \
        _Pragma("unroll") for (int g = 0; g < 8; g++) {                                         \
          split_tf32(pend[g].x, Ah[g][0][SL], Al[g][0][SL]);                                    \
          split_tf32(pend[g].y, Ah[g][1][SL], Al[g][1][SL]);                                    \
        }                                                                                       \
        break;
      NFFTCU_COMMIT(0) NFFTCU_COMMIT(1) NFFTCU_COMMIT(2) NFFTCU_COMMIT(3)
#undef NFFTCU_COMMIT
    }
    pend_slot = -1;
  };
  auto load_pair = [&](int zp) {
    if ((kq >> 1) == ((zp >> 1) & 1)) {
      commit();   // a second pair for this lane before the first was used (window moved by more than 4 cells)
      int z = zp + (kq & 1);
      if (z >= n2) z -= n2;
#pragma unroll
      for (int g = 0; g < 8; g++) pend[g] = G[rowoff_s[(g >> 1) * kF + 8 * (g & 1)] + z];
      pend_slot = (zp >> 2) & 3;
    }
  };
  auto advance_to = [&](int zlo) {
    if (zlo - zwin >= kF || zwin < 0) {
      pend_slot = -1;   // the whole window is replaced
      fill_all(zlo);
      zwin = zlo;
    } else {
      while (zwin < zlo) { load_pair(zwin + kF); zwin += 2; }
    }
  };

  uint2 e_next = table[0];
  for (int j = 0; j < nbat; j++) {
    const int st = j % kStages;
    const int zlo = bt_zlo(e_next);
    const uint32_t ready = mbar_test(&S.full[st], (j / kStages) & 1);   // result needed behind the window slide
    if (j + 1 < nbat) e_next = table[j + 1];
    if (zlo != zwin) advance_to(zlo);
    if (!ready) mbar_wait(&S.full[st], (j / kStages) & 1);
    commit();
    // the refill for the NEXT batch goes out before this batch's MMAs (it lands in pend, not in the window): a slide
    // by 2 or 4 cells gives every lane at most one pair; larger slides are done behind the MMAs as before
    if (j + 1 < nbat) {
      const int dz = bt_zlo(e_next) - zwin;
      if (dz == 2 || dz == 4) {
        while (zwin < bt_zlo(e_next)) { load_pair(zwin + kF); zwin += 2; }
      }
    }

    unsigned bh[4], bl[4];   // psi2 at slot 4s+kq of node nr: (hi, lo) packed by the producer
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const uint2 v = *reinterpret_cast<const uint2 *>(&S.ops[st][2][4 * s + kq][nr]);
      bh[s] = v.x;
      bl[s] = v.y;
    }
    const uint4 q1a = *reinterpret_cast<const uint4 *>(&S.ops[st][1][nr][2 * kq]);        // floats of nodes 2kq, 2kq+1 in .x, .z
    const uint4 q1b = *reinterpret_cast<const uint4 *>(&S.ops[st][1][8 + nr][2 * kq]);
    const float p1ax = __uint_as_float(q1a.x), p1ay = __uint_as_float(q1a.z);
    const float p1bx = __uint_as_float(q1b.x), p1by = __uint_as_float(q1b.z);
    float accr0 = 0.f, accr1 = 0.f, acci0 = 0.f, acci1 = 0.f;
#pragma unroll
    for (int h = 0; h < 4; h++) {   // footprint row l0 = 4*warp + h: groups 2h (rows 0-7 of the m16 tile) and 2h+1
      float c[2][4];                // [re/im][group 2h: col0, col1; group 2h+1: col0, col1]
#pragma unroll
      for (int cc = 0; cc < 2; cc++) c[cc][0] = c[cc][1] = c[cc][2] = c[cc][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; kk++)
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
#ifndef NFFTCU_DBG_1MMA
          mma_tf32(c[cc], Al[2 * h][cc][2 * kk], Al[2 * h + 1][cc][2 * kk], Al[2 * h][cc][2 * kk + 1], Al[2 * h + 1][cc][2 * kk + 1],
                   bh[2 * kk], bh[2 * kk + 1]);
          mma_tf32(c[cc], Ah[2 * h][cc][2 * kk], Ah[2 * h + 1][cc][2 * kk], Ah[2 * h][cc][2 * kk + 1], Ah[2 * h + 1][cc][2 * kk + 1],
                   bl[2 * kk], bl[2 * kk + 1]);
#endif
          mma_tf32(c[cc], Ah[2 * h][cc][2 * kk], Ah[2 * h + 1][cc][2 * kk], Ah[2 * h][cc][2 * kk + 1], Ah[2 * h + 1][cc][2 * kk + 1],
                   bh[2 * kk], bh[2 * kk + 1]);
        }
      const uint4 q0 = *reinterpret_cast<const uint4 *>(&S.ops[st][0][4 * warp + h][2 * kq]);
      const float p0x = __uint_as_float(q0.x), p0y = __uint_as_float(q0.z);
      const float wa0 = p0x * p1ax, wa1 = p0y * p1ay, wb0 = p0x * p1bx, wb1 = p0y * p1by;
      accr0 = fmaf(wa0, c[0][0], accr0); accr1 = fmaf(wa1, c[0][1], accr1);
      acci0 = fmaf(wa0, c[1][0], acci0); acci1 = fmaf(wa1, c[1][1], acci1);
      accr0 = fmaf(wb0, c[0][2], accr0); accr1 = fmaf(wb1, c[0][3], accr1);
      acci0 = fmaf(wb0, c[1][2], acci0); acci1 = fmaf(wb1, c[1][3], acci1);
    }
    if (j + 1 < nbat && bt_zlo(e_next) != zwin) advance_to(bt_zlo(e_next));
    *reinterpret_cast<double2 *>(&S.red[st][warp][nr][kq][0]) = make_double2((double) accr0, (double) acci0);
    *reinterpret_cast<double2 *>(&S.red[st][warp][nr][kq][2]) = make_double2((double) accr1, (double) acci1);
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[st]);
  }
}

template <int W, bool IMG>
__global__ void __launch_bounds__(256, 2)
spread_tf32_kernel(float2 *__restrict__ G, const float *__restrict__ xt, const float2 *__restrict__ ft,
                   const uint4 *__restrict__ chunks, const uint2 *__restrict__ table,
                   const double *__restrict__ poly, MmaParams P) {
  typedef float TS;
  const uint32_t *const perm = nullptr;
  float *const f = nullptr;
  NFFTCU_MMA_PROLOGUE(true, true)

  float C[4][2][2][4];   // [row pair h][re/im][n-tile][group 2h: col0, col1; group 2h+1: col0, col1], slot 8*nt + 2*kq + col
#pragma unroll
  for (int h = 0; h < 4; h++)
#pragma unroll
    for (int cc = 0; cc < 2; cc++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) C[h][cc][nt][0] = C[h][cc][nt][1] = C[h][cc][nt][2] = C[h][cc][nt][3] = 0.f;

  // one 16-byte reduction per pair (atomicAdd(float4*)) measured slower than two 8-byte ones on B200 (BT 5.70 vs 5.43 ms)
  const bool vec4 = false;
  // retire pair (zp, zp+1) of the window (zp even, unwrapped): n-tile (zp>>3)&1, lanes kq == (zp&7)>>1
  auto retire_pair = [&](int zp) {
    int zw = zp;
    if (zw >= n2) zw -= n2;
    if (kq == ((zp & 7) >> 1)) {
#define NFFTCU_RETIRE(NT)                                                                                   \
      _Pragma("unroll") for (int g = 0; g < 8; g++) {                                                       \
        float2 *dst = G + rowoff_s[(g >> 1) * kF + 8 * (g & 1)] + zw;                                       \
        const int o = 2 * (g & 1);                                                                          \
        if (vec4) {   /* both cells of the pair in one 16-byte reduction */                                 \
          atomicAdd(reinterpret_cast<float4 *>(dst), make_float4(C[g >> 1][0][NT][o], C[g >> 1][1][NT][o],  \
                                                                 C[g >> 1][0][NT][o + 1], C[g >> 1][1][NT][o + 1])); \
        } else {                                                                                            \
          atomicAdd(dst, make_float2(C[g >> 1][0][NT][o], C[g >> 1][1][NT][o]));                            \
          atomicAdd(dst + 1, make_float2(C[g >> 1][0][NT][o + 1], C[g >> 1][1][NT][o + 1]));                \
        }                                                                                                   \
        C[g >> 1][0][NT][o] = C[g >> 1][1][NT][o] = C[g >> 1][0][NT][o + 1] = C[g >> 1][1][NT][o + 1] = 0.f; \
      }
      if (((zp >> 3) & 1) == 0) { NFFTCU_RETIRE(0) } else { NFFTCU_RETIRE(1) }
#undef NFFTCU_RETIRE
    }
  };

  int zwin = -1;
  uint2 e_next = table[0];
  for (int j = 0; j < nbat; j++) {
    const int st = j % kStages;
    const int zlo = bt_zlo(e_next);
    float fr[2] = {0.f, 0.f}, fi[2] = {0.f, 0.f};   // IMG: samples of nodes kq and 4 + kq
    if (IMG) {
#pragma unroll
      for (int ks = 0; ks < 2; ks++)
        if (4 * ks + kq < bt_nb(e_next)) {
          const float2 v = ft[(size_t) e_next.x + 4 * ks + kq];
          fr[ks] = v.x;
          fi[ks] = v.y;
        }
    }
    if (j + 1 < nbat) e_next = table[j + 1];
    if (zwin < 0) zwin = zlo;
    if (zlo != zwin) {
      const int zend = (zlo - zwin >= kF) ? zwin + kF : zlo;
      for (int zp = zwin; zp < zend; zp += 2) retire_pair(zp);
      zwin = zlo;
    }
    mbar_wait(&S.full[st], (j / kStages) & 1);
    unsigned bh[2][2], bl[2][2];   // [n-tile][k half]: psi2 of node 4*ks+kq at slot 8*nt+nr
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        const uint2 v = *reinterpret_cast<const uint2 *>(&S.ops[st][2][8 * nt + nr][4 * ks + kq]);
        bh[nt][ks] = v.x;
        bl[nt][ks] = v.y;
      }
    float p1r[2][2], p1i[2][2];    // [half][k half]: psi1 f.re, psi1 f.im (IMG: psi1 in p1r, the samples go into psi2 instead)
#pragma unroll
    for (int hh = 0; hh < 2; hh++)
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        p1r[hh][ks] = __uint_as_float(reinterpret_cast<const uint2 *>(&S.ops[st][1][8 * hh + nr][4 * ks + kq])->x);
        p1i[hh][ks] = IMG ? 0.f : __uint_as_float(reinterpret_cast<const uint2 *>(&S.ops[st][3][8 * hh + nr][4 * ks + kq])->x);
      }
    // IMG: the sample of node k multiplies the B operand psi2[k][z] (8 products and splits per lane) instead of the
    // A operand psi0 psi1 [row][k] (32 per lane); the A fragments are then shared by the re and im accumulators
    unsigned bih[2][2], bil[2][2];
    if (IMG) {
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          const float v2 = __uint_as_float(bh[nt][ks]) + __uint_as_float(bl[nt][ks]);
          split_tf32(v2 * fi[ks], bih[nt][ks], bil[nt][ks]);
          split_tf32(v2 * fr[ks], bh[nt][ks], bl[nt][ks]);
        }
    }
    float p0[4][2];
#pragma unroll
    for (int h = 0; h < 4; h++) {
      p0[h][0] = __uint_as_float(reinterpret_cast<const uint2 *>(&S.ops[st][0][4 * warp + h][kq])->x);
      p0[h][1] = __uint_as_float(reinterpret_cast<const uint2 *>(&S.ops[st][0][4 * warp + h][4 + kq])->x);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[st]);   // operands are in registers: the stage can be refilled
#pragma unroll
    for (int h = 0; h < 4; h++) {
      if (IMG) {
        // A fragment: a0 (group 2h, node kq), a1 (group 2h+1, node kq), a2 (group 2h, node 4+kq), a3 (group 2h+1, node 4+kq)
        unsigned ah[4], al[4];
        split_tf32(p0[h][0] * p1r[0][0], ah[0], al[0]);
        split_tf32(p0[h][0] * p1r[1][0], ah[1], al[1]);
        split_tf32(p0[h][1] * p1r[0][1], ah[2], al[2]);
        split_tf32(p0[h][1] * p1r[1][1], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          mma_tf32(C[h][0][nt], al[0], al[1], al[2], al[3], bh[nt][0], bh[nt][1]);
          mma_tf32(C[h][1][nt], al[0], al[1], al[2], al[3], bih[nt][0], bih[nt][1]);
          mma_tf32(C[h][0][nt], ah[0], ah[1], ah[2], ah[3], bl[nt][0], bl[nt][1]);
          mma_tf32(C[h][1][nt], ah[0], ah[1], ah[2], ah[3], bil[nt][0], bil[nt][1]);
          mma_tf32(C[h][0][nt], ah[0], ah[1], ah[2], ah[3], bh[nt][0], bh[nt][1]);
          mma_tf32(C[h][1][nt], ah[0], ah[1], ah[2], ah[3], bih[nt][0], bih[nt][1]);
        }
      } else {
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
          unsigned ah[4], al[4];
          split_tf32(p0[h][0] * (cc ? p1i[0][0] : p1r[0][0]), ah[0], al[0]);
          split_tf32(p0[h][0] * (cc ? p1i[1][0] : p1r[1][0]), ah[1], al[1]);
          split_tf32(p0[h][1] * (cc ? p1i[0][1] : p1r[0][1]), ah[2], al[2]);
          split_tf32(p0[h][1] * (cc ? p1i[1][1] : p1r[1][1]), ah[3], al[3]);
#pragma unroll
          for (int nt = 0; nt < 2; nt++) {
            mma_tf32(C[h][cc][nt], al[0], al[1], al[2], al[3], bh[nt][0], bh[nt][1]);
            mma_tf32(C[h][cc][nt], ah[0], ah[1], ah[2], ah[3], bl[nt][0], bl[nt][1]);
            mma_tf32(C[h][cc][nt], ah[0], ah[1], ah[2], ah[3], bh[nt][0], bh[nt][1]);
          }
        }
      }
    }
  }
  for (int zp = zwin; zp < zwin + kF; zp += 2) retire_pair(zp);
}

// ---- window images (plan time) ----------------------------------------------------------------------------------
// The window values of a node set do not change between transforms, and evaluating them in the kernels costs the
// producer warps ~11 000 cycles per batch on an FP64 pipe that the DMMA stream keeps busy (measured with clock64()
// probes: the MMA warps wait for operands 8 % of the time and the producers never idle).  When memory allows, the
// placed operand block of every batch -- psi0, psi1, psi2 as the ring stage holds them, 3 KB -- is therefore written
// once per node set (the reference's precompute_psi, nfft.c:5819-5844, stores psi per node for the same reason) and
// the kernels' producer warps become feeders: one TMA bulk copy per batch.  For spreading the image holds psi1 and
// the MMA warps multiply by the samples.  4.1 GB at cfg3; plans whose images would not fit keep the evaluating producers.
template <typename TS, int W, bool TF32>
__global__ void __launch_bounds__(128)
mma_images_kernel(const TS *__restrict__ xt, const uint4 *__restrict__ chunks, const uint2 *__restrict__ table,
                  const double *__restrict__ poly, double *__restrict__ img, MmaParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Shared<W, false> &S = *reinterpret_cast<Shared<W, false> *>(smem_raw);
  const int tid = threadIdx.x;
  const uint4 chunk = chunks[blockIdx.x];
  const uint32_t b0 = chunk.y;
  const int nbat = (int) (chunk.z - b0);
  if (nbat == 0) return;
  const int tile = (int) chunk.x;
  const int a = tile / P.NT1, bt = tile - a * P.NT1;
  for (int i = tid; i < 3 * W * kCoefK; i += 128) {
    const int t = i / (W * kCoefK), l = (i / kCoefK) % W, k = i % kCoefK;
    S.coef[i] = k <= kKbPolyDeg ? poly[(t * (kKbPolyDeg + 1) + k) * W + l] : 0.0;
  }
  __syncthreads();
  producer_loop<TS, W, false, TF32, true>(S, xt, nullptr, nullptr, nullptr, table + b0, nbat, a, bt, P, tid >> 5, tid & 31,
                                          img + (size_t) b0 * kImgDoubles);
}

MmaParams make_params(const nfftcu_ctx *c) { return mma3d_params(c); }

}  // namespace

MmaParams mma3d_params(const nfftcu_ctx *c) {
  MmaParams P;
  P.n0 = (int) c->n[0];
  P.n1 = (int) c->n[1];
  P.n2 = (int) c->n[2];
  P.m = (int) c->m;
  P.T = kF + 1 - (2 * P.m + 2);
  P.NT0 = (P.n0 + P.T - 1) / P.T;
  P.NT1 = (P.n1 + P.T - 1) / P.T;
  const long long tiles = (long long) P.NT0 * P.NT1;
  long long zseg = (8ll * c->sm_count + tiles - 1) / tiles;
  if (zseg < 1) zseg = 1;
  if (zseg > P.n2 / 16) zseg = P.n2 / 16;
  if (zseg < 1) zseg = 1;
  P.zseg = (int) zseg;
  P.deg = c->kbpoly_fit;
  // the images hold fp64 values or packed fp32 / TF32 pairs, fixed when they were built: use them only for that kernel family
  const bool want_tf32 = c->prec == NFFTCU_FLOAT && c->opt_b_kernel != 3;
  P.img = (c->mma_images_ready && c->mma_images_tf32 == want_tf32) ? (const double *) c->mma_images : nullptr;
  P.M = c->M;
  return P;
}

namespace {

template <int W, int FLUSH>
size_t spread_smem() {
  size_t b = (sizeof(Shared<W, true>) + 127) & ~(size_t) 127;
  if (FLUSH == 1) b += sizeof(double2) * 4 * 64 * kStgRow;
  return b;
}

template <typename TS, int W, bool IMG>
int launch_img(nfftcu_ctx *c, const void *f_in, void *f_out, bool spread, const MmaParams &P) {
  typedef typename Cplx<TS>::type C2;
  const unsigned grid = (unsigned) c->mma_nchunks;
  if (grid == 0) return NFFTCU_OK;
  const uint4 *chunks = (const uint4 *) c->mma_chunks;
  const TS *xt = (const TS *) c->tile_x;
  const double *poly = (const double *) c->kbpoly_dev;
  const uint2 *table = (const uint2 *) c->mma_batches;
  if (sizeof(TS) == 4 && c->opt_b_kernel != 3) {   // fp32 plans: TF32 tensor path (NFFTCU_OPT_B_KERNEL = 3 forces DMMA)
    if (!spread) {
      const size_t smem = sizeof(Shared<W, false>);
      NFFTCU_CUDA(cudaFuncSetAttribute(interp_tf32_kernel<W, IMG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
      interp_tf32_kernel<W, IMG><<<grid, 256, smem, c->stream>>>((const float2 *) c->grid, (const float *) c->tile_x, c->tile_perm,
                                                            (float *) f_out, chunks, table, poly, P);
      if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
      c->launches++;
    } else {
      const int kb = 256;
      mma_gather_f_kernel<float2><<<(unsigned) ((c->M + kb - 1) / kb), kb, 0, c->stream>>>(
          (const float2 *) f_in, c->tile_perm, (float2 *) c->f_tile, c->M);
      const size_t smem = sizeof(Shared<W, true>);
      NFFTCU_CUDA(cudaFuncSetAttribute(spread_tf32_kernel<W, IMG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
      spread_tf32_kernel<W, IMG><<<grid, 256, smem, c->stream>>>((float2 *) c->grid, (const float *) c->tile_x,
                                                            (const float2 *) c->f_tile, chunks, table, poly, P);
      if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
      c->launches += 2;
    }
    NFFTCU_CUDA(cudaGetLastError());
    return NFFTCU_OK;
  }
  if (!spread) {
    const size_t smem = sizeof(Shared<W, false>);
    NFFTCU_CUDA(cudaFuncSetAttribute(interp_mma_kernel<TS, W, IMG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
    interp_mma_kernel<TS, W, IMG><<<grid, 256, smem, c->stream>>>((const C2 *) c->grid, xt, c->tile_perm, (TS *) f_out,
                                                             chunks, table, poly, P);
    if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
    c->launches++;
  } else {
    const int kb = 256;
    mma_gather_f_kernel<C2><<<(unsigned) ((c->M + kb - 1) / kb), kb, 0, c->stream>>>(
        (const C2 *) f_in, c->tile_perm, (C2 *) c->f_tile, c->M);
    const bool bulk = !IMG && sizeof(TS) == 8 && (P.n2 % 8 == 0) && c->opt_b_flush == 2;
    if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
    if (bulk) {
      const size_t smem = spread_smem<W, 1>();
      NFFTCU_CUDA(cudaFuncSetAttribute(spread_mma_kernel<double, W, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      spread_mma_kernel<double, W, 1, false><<<grid, 256, smem, c->stream>>>(
          (double2 *) c->grid, (const double *) c->tile_x, (const double2 *) c->f_tile, chunks, table, poly, P);
    } else {
      const size_t smem = spread_smem<W, 0>();
      NFFTCU_CUDA(cudaFuncSetAttribute(spread_mma_kernel<TS, W, 0, IMG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      spread_mma_kernel<TS, W, 0, IMG><<<grid, 256, smem, c->stream>>>((C2 *) c->grid, xt, (const C2 *) c->f_tile,
                                                                  chunks, table, poly, P);
    }
    if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
    c->launches += 2;
  }
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

template <typename TS, int W>
int launch(nfftcu_ctx *c, const void *f_in, void *f_out, bool spread, const MmaParams &P) {
  return P.img ? launch_img<TS, W, true>(c, f_in, f_out, spread, P) : launch_img<TS, W, false>(c, f_in, f_out, spread, P);
}

// plan time: window images for the current node set (see mma_images_kernel), when they fit
template <typename TS, int W>
int build_images_w(nfftcu_ctx *c, const MmaParams &P, long long nbatches) {
  const size_t smem = sizeof(Shared<W, false>);
  const unsigned grid = (unsigned) c->mma_nchunks;
  double *img = (double *) c->mma_images;
  const bool tf32 = sizeof(TS) == 4 && c->opt_b_kernel != 3;
  (void) nbatches;
  if (tf32) {
    NFFTCU_CUDA(cudaFuncSetAttribute(mma_images_kernel<TS, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    mma_images_kernel<TS, W, true><<<grid, 128, smem, c->stream>>>((const TS *) c->tile_x, (const uint4 *) c->mma_chunks,
                                                                  (const uint2 *) c->mma_batches,
                                                                  (const double *) c->kbpoly_dev, img, P);
  } else {
    NFFTCU_CUDA(cudaFuncSetAttribute(mma_images_kernel<TS, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    mma_images_kernel<TS, W, false><<<grid, 128, smem, c->stream>>>((const TS *) c->tile_x, (const uint4 *) c->mma_chunks,
                                                                   (const uint2 *) c->mma_batches,
                                                                   (const double *) c->kbpoly_dev, img, P);
  }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

template <typename TS>
int build_images(nfftcu_ctx *c, const MmaParams &P, long long nbatches) {
  switch (2 * (int) c->m + 2) {
    case 6: return build_images_w<TS, 6>(c, P, nbatches);
    case 8: return build_images_w<TS, 8>(c, P, nbatches);
    case 10: return build_images_w<TS, 10>(c, P, nbatches);
    case 12: return build_images_w<TS, 12>(c, P, nbatches);
    case 14: return build_images_w<TS, 14>(c, P, nbatches);
    default: break;
  }
  return NFFTCU_EINVAL;
}

template <typename TS>
int dispatch(nfftcu_ctx *c, const void *f_in, void *f_out, bool spread) {
  const MmaParams P = make_params(c);
  switch (2 * (int) c->m + 2) {
    case 6: return launch<TS, 6>(c, f_in, f_out, spread, P);
    case 8: return launch<TS, 8>(c, f_in, f_out, spread, P);
    case 10: return launch<TS, 10>(c, f_in, f_out, spread, P);
    case 12: return launch<TS, 12>(c, f_in, f_out, spread, P);
    case 14: return launch<TS, 14>(c, f_in, f_out, spread, P);
    default: break;
  }
  set_error("mma3d: unsupported window cut-off m=%lld", (long long) c->m);
  return NFFTCU_EINVAL;
}

}  // namespace

bool mma3d_supported(const nfftcu_ctx *c) {
  if (c->d != 3 || c->direct_only) return false;
  if (c->m < 2 || c->m > 6 || c->kbpoly_fit < 0) return false;
  for (int t = 0; t < 3; t++)
    if (c->n[t] < kF || c->n[t] > 0x3fffff) return false;
  if (c->n[2] % 2 != 0) return false;
  if (c->n_total >= (1ll << 31)) return false;   // 32-bit row offsets
  const long long T = kF + 1 - (2 * c->m + 2);
  const long long units = ((c->n[0] + T - 1) / T) * ((c->n[1] + T - 1) / T) * 64;
  return units < (1ll << 31);
}

// processing order of the DMMA kernels: nodes sorted by (tile, u2), work-unit offsets
int mma3d_bin_nodes(nfftcu_ctx *c) {
  const long long M = c->M;
  c->mma_ready = false;
  if (M == 0) return NFFTCU_OK;
  const MmaParams P = make_params(c);
  const long long units = (long long) P.NT0 * P.NT1 * P.zseg;
  const long long nkeys = (long long) P.NT0 * P.NT1 * P.n2;
  if (!c->tile_keys) NFFTCU_CUDA(pool_malloc(&c->tile_keys, sizeof(uint64_t) * (size_t) M));
  if (!c->tile_perm) NFFTCU_CUDA(pool_malloc((void **) &c->tile_perm, sizeof(uint32_t) * (size_t) M));
  if (!c->tile_x) NFFTCU_CUDA(pool_malloc(&c->tile_x, real_size(c) * (size_t) M * 3));
  if (!c->f_tile) NFFTCU_CUDA(pool_malloc(&c->f_tile, 2 * real_size(c) * (size_t) M));
  if (!c->bin_start || c->tile_nbins != units) {
    if (c->bin_start) pool_free(c->bin_start);
    c->bin_start = nullptr;
    NFFTCU_CUDA(pool_malloc((void **) &c->bin_start, sizeof(uint32_t) * (size_t) (units + 1)));
    c->tile_nbins = units;
  }
  const int kb = 256;
  if (c->prec == NFFTCU_DOUBLE)
    mma_keys_kernel<double><<<(unsigned) ((M + kb - 1) / kb), kb, 0, c->stream>>>(
        (const double *) c->x_dev, (uint64_t *) c->tile_keys, c->tile_perm, M, P);
  else
    mma_keys_kernel<float><<<(unsigned) ((M + kb - 1) / kb), kb, 0, c->stream>>>(
        (const float *) c->x_dev, (uint64_t *) c->tile_keys, c->tile_perm, M, P);
  c->launches++;
  int bits = 0;
  while ((1ll << bits) < nkeys && bits < 62) bits++;
  NFFTCU_TRY(radix_sort_pairs(c, (uint64_t *) c->tile_keys, c->tile_perm, M, bits));
  NFFTCU_TRY(gather_nodes(c, c->tile_perm, c->tile_x));
  mma_unit_bounds_kernel<<<(unsigned) ((units + 1 + kb - 1) / kb), kb, 0, c->stream>>>(
      (const uint64_t *) c->tile_keys, c->bin_start, units, M, P);
  c->launches++;
  // fp32 plans: the tcgen05 kernels take over when their tables and images could be built (tc5.cu); the tables and
  // window images of this file are then only built for the direction that stays here
  c->tc5_ready = c->tc5s_ready = false;
  if (tc5_selected(c)) NFFTCU_TRY(tc5_build(c, P));
  c->mma_images_ready = false;
  if (!(c->tc5_ready && c->tc5s_ready)) {
    // batch table: count per unit, scan, fill
    if (c->mma_units != units) {
      if (c->mma_batch_start) pool_free(c->mma_batch_start);
      if (c->mma_counts) pool_free(c->mma_counts);
      if (c->mma_chunk_start) pool_free(c->mma_chunk_start);
      c->mma_batch_start = c->mma_counts = c->mma_chunk_start = nullptr;
      NFFTCU_CUDA(pool_malloc((void **) &c->mma_batch_start, sizeof(uint32_t) * (size_t) (units + 1)));
      NFFTCU_CUDA(pool_malloc((void **) &c->mma_counts, sizeof(uint32_t) * (size_t) units));
      c->mma_units = units;
    }
    const unsigned wgrid = (unsigned) ((units * 32 + kb - 1) / kb);
    mma_batches_kernel<false><<<wgrid, kb, 0, c->stream>>>((const uint64_t *) c->tile_keys, c->bin_start, c->mma_counts,
                                                           nullptr, nullptr, units, P);
    mma_scan_kernel<<<1, 1024, 0, c->stream>>>(c->mma_counts, c->mma_batch_start, units);
    uint32_t total = 0;
    NFFTCU_CUDA(cudaMemcpyAsync(&total, c->mma_batch_start + units, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
    if ((long long) total > c->mma_batch_cap) {
      if (c->mma_batches) pool_free(c->mma_batches);
      c->mma_batches = nullptr;
      c->mma_batch_cap = (long long) total + total / 8 + 1024;
      NFFTCU_CUDA(pool_malloc(&c->mma_batches, sizeof(uint2) * (size_t) c->mma_batch_cap));
    }
    mma_batches_kernel<true><<<wgrid, kb, 0, c->stream>>>((const uint64_t *) c->tile_keys, c->bin_start, nullptr,
                                                          c->mma_batch_start, (uint2 *) c->mma_batches, units, P);
    // chunks: count per unit (reusing the counts scratch), scan, fill
    if (!c->mma_chunk_start) NFFTCU_CUDA(pool_malloc((void **) &c->mma_chunk_start, sizeof(uint32_t) * (size_t) (units + 1)));
    const unsigned ugrid = (unsigned) ((units + kb - 1) / kb);
    mma_chunk_count_kernel<<<ugrid, kb, 0, c->stream>>>(c->mma_batch_start, c->mma_counts, units);
    mma_scan_kernel<<<1, 1024, 0, c->stream>>>(c->mma_counts, c->mma_chunk_start, units);
    uint32_t nchunks = 0;
    NFFTCU_CUDA(cudaMemcpyAsync(&nchunks, c->mma_chunk_start + units, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
    if ((long long) nchunks > c->mma_chunk_cap) {
      if (c->mma_chunks) pool_free(c->mma_chunks);
      c->mma_chunks = nullptr;
      c->mma_chunk_cap = (long long) nchunks + nchunks / 8 + 1024;
      NFFTCU_CUDA(pool_malloc(&c->mma_chunks, sizeof(uint4) * (size_t) c->mma_chunk_cap));
    }
    mma_chunk_fill_kernel<<<ugrid, kb, 0, c->stream>>>(c->mma_batch_start, c->mma_chunk_start, (uint4 *) c->mma_chunks, units,
                                                      P.zseg);
    c->mma_nchunks = nchunks;
    c->launches += 6;
    NFFTCU_CUDA(cudaGetLastError());
    // window images: on when they fit (NFFTCU_OPT_WINDOW_IMAGES: 0 auto | 1 off | 2 on regardless of the budget)
    c->mma_images_ready = false;
    if (c->opt_window_images != 1 && total > 0) {
      const size_t need = sizeof(double) * kImgDoubles * (size_t) total;
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      const bool have = c->mma_images && c->mma_images_bytes >= need;
      const bool fits = have || c->opt_window_images == 2 || (need <= free_b / 2 && need <= ((size_t) 48 << 30));
      if (fits) {
        if (!have) {
          if (c->mma_images) pool_free(c->mma_images);
          c->mma_images = nullptr;
          c->mma_images_bytes = 0;
          if (pool_malloc(&c->mma_images, need + need / 16) == cudaSuccess) c->mma_images_bytes = need + need / 16;
          else cudaGetLastError();   // no room after all: keep the evaluating producers
        }
        if (c->mma_images) {
          MmaParams Pi = P;
          Pi.img = nullptr;
          NFFTCU_TRY(c->prec == NFFTCU_DOUBLE ? build_images<double>(c, Pi, total) : build_images<float>(c, Pi, total));
          c->mma_images_tf32 = c->prec == NFFTCU_FLOAT && c->opt_b_kernel != 3;
          c->mma_images_ready = true;
        }
      }
    }
  }
  c->mma_ready = true;
  return NFFTCU_OK;
}

#ifdef NFFTCU_DBG_CLOCKS
extern "C" int nfftcu_debug_clocks(unsigned long long *out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_dbg, z, sizeof(z));
  return 0;
}
#endif

int mma3d_interp(nfftcu_ctx *c, void *f_dev) {
  if (c->tc5_ready) return tc5_interp(c, f_dev);
  return c->prec == NFFTCU_DOUBLE ? dispatch<double>(c, nullptr, f_dev, false) : dispatch<float>(c, nullptr, f_dev, false);
}

int mma3d_spread(nfftcu_ctx *c, const void *f_dev) {
  if (c->tc5s_ready) return tc5_spread(c, f_dev);
  return c->prec == NFFTCU_DOUBLE ? dispatch<double>(c, f_dev, nullptr, true) : dispatch<float>(c, f_dev, nullptr, true);
}

}  // namespace nfftcu
