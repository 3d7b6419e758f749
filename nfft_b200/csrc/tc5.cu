// tc5.cu -- 3-D interpolation (B) and spreading (B^T) of fp32 plans (nfftf_) on the 5th-generation tensor cores:
// tcgen05.mma kind::tf32 with the grid window AND the accumulators in tensor memory (TMEM).  d = 3, m <= 6.
//
// Reference being replaced: nfft_trafo_3d_B / nfft_adjoint_3d_B of the float build (kernel/nfft/nfft.c:4687-4914,
// 5126-5384, compute loops 4020-4265, 4289-4436).  Same arithmetic -- f_j = sum psi0 psi1 psi2 g, g += psi0 psi1 psi2 f_j
// -- in a different summation order, with every product carried as a 3xTF32 split (hi hi + lo hi + hi lo, fp32
// accumulation): 1.5e-7 rel-l2 against fp64 on the window contraction (tools/microbench6.cu), i.e. fp32 accuracy.
//
// Why not the register-window mma.sync kernels of mma3d.cu.  Those keep the sliding grid window in registers because
// mma.sync takes its operands from registers; on the fp32 path they are issue-bound (2550 warp instructions per 8-node
// batch).  tcgen05.mma takes A from TMEM: the window of a tile -- 512 rows (256 pencils x re / im) by a ring
// of 40 z-slots, as a (hi, lo) pair -- lives in 320 TMEM columns for the whole sweep, is refilled 8 cells at a time with
// tcgen05.st, and ONE thread issues the contraction.  Measured on B200 (profiles/r2l_microbench_tcgen05.txt): an
// M = 128, N = 16, K = 8 kind::tf32 MMA with A in TMEM retires every 17 cycles (A in shared memory: 78).
//
//   interpolation   T[row, node] = sum_z G[row, z] psi2[z, node]       row = (re/im, pencil p = 16 l0 + l1), z = 24 slots
//                   f_node       = sum_row (psi0[l0] psi1[l1])[node] T[row, node]           (epilogue, FFMA + shuffles)
//   spreading       G[row, z]   += sum_node (psi0 psi1 f)[row, node] psi2[node, z]            (accumulators = the window)
//
// Batches.  Nodes are in the (tile, u2) order of mma3d.cu.  An interpolation batch = up to 16 consecutive nodes of a tile
// whose taps lie in the 24 cells [base, base + 24), base = 8 floor(u2_first / 8): three k-steps of 8 slots, each starting
// at an 8-slot boundary of the ring (slot = z mod 40), so the k-steps never wrap and the A operand address is just a
// column offset.  psi2 is zero outside a node's taps, so whatever else the 24 slots hold only has to be finite.  (The
// spreading batches are cut with a 16-aligned base and 32 slots: see tc5_spread_kernel.)
//
// Per batch the plan-time IMAGE (5.2 KB, like the reference's PRE_PSI table) holds psi2 as the (hi, lo) B operand in the
// canonical K-major no-swizzle layout, psi0 / psi1 placed in the footprint, the batch entry and the output indices of
// its nodes; a feeder lane streams it into a shared-memory ring by TMA (cp.async.bulk + mbarrier expect_tx).
//
// tc5_interp_kernel: CTA = 14 warps, persistent over the chunk list, one CTA per SM (it owns all 512 TMEM columns):
//   warps 0-7   two epilogue groups (a group takes every second batch): tcgen05.ld of the batch's T (lane quarter
//               q = warp mod 4), row weights, 32-value butterfly reduction, f[perm[node]] stored by one warp per batch
//   warps 8-11  window refill: 8 cells x 2 pencils per thread from L2 (loaded one slide ahead), split into tf32
//               (hi, lo), tcgen05.st
//   warp  12    MMA issue (one elected lane): 3 k-steps x 3 split terms x 4 row blocks = 36 tcgen05.mma + 1 commit
//   warp  13    TMA feeder of the operand images
// Hand-offs: op_full / op_empty (image ring), a_ready (one phase per slide of 8 cells), acc_full (commit: T is complete)
// / acc_empty (T was read) are mbarriers; done_upto (shared-memory counter, written by the epilogue) tells the refill
// warps which batches have committed: a slot is only overwritten once every batch that still reads its old cell has
// (with 40 slots for a 24-slot span: bases at least 24 cells behind).
#include <stdlib.h>

#include "common.cuh"
#include "mma3d.cuh"

namespace nfftcu {

// timing probes (NFFT_B200_TC5_DBG & 8): cycle totals per role, summed over the CTAs; read with nfftcu_tc5_debug
__device__ unsigned long long g_t5dbg[32];

namespace {

constexpr int kN = 16;           // nodes per batch = N of the MMA
constexpr int kRing = 40;        // z-slots of the window ring: the 24 a batch reads + 16 the refill may already replace
constexpr int kSpan = 24;        // slots one batch contracts: 3 k-steps of 8
constexpr int kOpStages = 16;    // operand-image ring: 80 KB in flight per SM (the images stream from HBM, ~2 us away)
constexpr int kAcc = 3;          // accumulator stages (4 row blocks x 16 columns each)
constexpr int kEpi = 2;          // epilogue warpgroups; group g takes the batches with index = g (mod kEpi)
constexpr int kMmaWarps = 1;     // MMA-issuing warps.  (Two warps taking alternate batches, so that one's hand-shakes run under the
                                 // other's MMAs, measured no faster -- the refill and the epilogue then limit -- and are not wired.)
constexpr int kThreadsI = 32 * (4 * kEpi + 4 + kMmaWarps + 1);
constexpr int kImgBytes = 5248;  // psi2 hi 1536 | psi2 lo 1536 | psi0 1024 | psi1 1024 | output index of the 16 nodes 64 | batch entry 8 | pad
constexpr int kOffLo = 1536, kOffP0 = 3072, kOffP1 = 4096, kOffPerm = 5120, kOffEntry = 5184;
constexpr int kDone = 8;         // commit barriers in flight (acc_full ring)
constexpr int kSlides = 8;       // window hand-offs in flight (the hazards keep the refill warps within 2 slides of the MMAs)
constexpr int kBChunk = 256, kBGroup = 128;   // B operand: (slot/4)*256 + (node/8)*128 + (node%8)*16 + (slot%4)*4
constexpr int kColA = 64 * kAcc;  // TMEM: D stage s at 64 s + 16 b, A ring of row block b at kColA + 2 kRing b (+ kRing: lo) + slot
static_assert(kColA + 8 * kRing <= 512, "tensor memory has 512 columns");
// spreading (tc5_spread_kernel): batches with a 16-aligned base whose taps lie in [base, base + 32); the accumulators are a
// ring of 64 z-slots per row (256 TMEM columns), the A operand (psi0 psi1 f)[row, node] as a (hi, lo) pair takes 128 columns
// per stage, two stages
constexpr int kSpanS = 32;       // slots one spreading batch touches
constexpr int kRingS = 64;       // accumulator ring
constexpr int kOpStagesS = 12;   // operand-image ring of the spreading kernel
constexpr int kImgBytesS = 6272; // psi2 hi 2048 | psi2 lo 2048 | psi0 1024 | psi1 1024 | batch entry 8 | pad
constexpr int kOffLoS = 2048, kOffP0S = 4096, kOffP1S = 5120, kOffEntryS = 6144;
constexpr int kStageS = kImgBytesS + 256;   // + the batch's samples: 18 float2 from an even node index on (TMA needs 16-byte alignment)
constexpr int kOpG = 2;          // operand warpgroups of the spreading kernel; group g takes the batches = g (mod 2) and owns A stage g
constexpr int kThreadsS = 32 * (4 * kOpG + 4 + 2);
constexpr int kChunkBatches = 256;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int b_off(int n, int s) { return (s >> 2) * kBChunk + (n >> 3) * kBGroup + (n & 7) * 16 + (s & 3) * 4; }
// psi0 / psi1 rows of 16 floats; the four 16-byte chunks of row l are rotated by l >> 1 so that the eight rows a quarter
// warp reads in one LDS.128 phase fall into eight different bank groups
__device__ __forceinline__ int w_off(int l, int n) { return l * 64 + ((((n >> 2) + (l >> 1)) & 3) << 4) + (n & 3) * 4; }

// spreading: B[slot][node], K = nodes: (node/8)*1024 + ((node/4)&1)*512 + (slot/8)*128 + (slot%8)*16 + (node%4)*4
__device__ __forceinline__ int sb_off(int n, int s) { return (n >> 3) * 1024 + ((n >> 2) & 1) * 512 + (s >> 3) * 128 + (s & 7) * 16 + (n & 3) * 4; }

__device__ __forceinline__ int wrapi(int v, int n) {
  v %= n;
  if (v < 0) v += n;
  return v;
}

__device__ __forceinline__ int t5_base(uint2 e) { return (int) (e.y & 0xffffffu); }
__device__ __forceinline__ int t5_nb(uint2 e) { return (int) ((e.y >> 24) & 0x1fu); }

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Every hand-off of this kernel sits on a batch's critical path, so the waits poll (try_wait without a suspend-time
// hint returns within tens of cycles; the hinted form parks the warp and wakes it late).  NFFT_B200_TC5_WAIT=1 selects
// the hinted form for comparison.
template <bool HINT = false>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity) {
  if (HINT) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680) : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous; LBO = distance of the two core matrices of a k-step
// along K, SBO = distance of 8-row groups; descriptor version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t) ((saddr & 0x3ffff) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulate, A and B K-major
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
                 "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                 "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- batch table and chunks (plan time) ---------------------------------------------------------------------
// entry.x = first node (tile order), entry.y = base | nb << 24 | two_ksteps << 29.  One warp per work unit (see mma3d.cu) walks the unit's
// sorted nodes 32 at a time: a batch starts at the first unassigned node, base = its u2 rounded down to a multiple of
// ALIGN, and takes up to 16 nodes with u2 + W <= base + SPAN (interpolation: 8 / 24, spreading: 16 / 32).
template <bool FILL, int ALIGN, int SPAN>
__global__ void t5_batches_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ unit_start,
                                  uint32_t *__restrict__ counts, const uint32_t *__restrict__ batch_start,
                                  uint2 *__restrict__ table, long long units, MmaParams P) {
  const long long unit = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (unit >= units) return;
  const int W = 2 * P.m + 2;
  const long long k0 = unit_start[unit], k1 = unit_start[unit + 1];
  const uint64_t kbase = (uint64_t) (unit / P.zseg) * P.n2;
  const uint32_t out = FILL ? batch_start[unit] : 0;
  uint32_t nbat = 0;
  long long pos = k0;
  while (pos < k1) {
    const long long idx = pos + lane;
    const int u = idx < k1 ? (int) (keys[idx] - kbase) : 0x3fffffff;
    int o = 0;
    while (pos + o < k1) {
      if (o + kN > 32 && pos + 32 < k1) break;   // the batch may extend beyond these 32 nodes: reload from pos + o
      const int base = __shfl_sync(kFull, u, o) & ~(ALIGN - 1);
      const bool member = lane >= o && lane < o + kN && u + W <= base + SPAN;
      const int nb = __popc(__ballot_sync(kFull, member));
      // bit 29: the taps of every node end below base + 16 (the batch's last node has the largest u2): two k-steps suffice
      const int u_last = __shfl_sync(kFull, u, o + nb - 1);
      const unsigned two = (u_last + W <= base + 16) ? 1u : 0u;
      if (FILL && lane == 0) table[out + nbat] = make_uint2((uint32_t) (pos + o), (unsigned) base | ((unsigned) nb << 24) | (two << 29));
      nbat++;
      o += nb;
    }
    pos += o;
  }
  if (!FILL && lane == 0) counts[unit] = nbat;
}

// exclusive scan of counts[0..n) into out[0..n], single CTA
__global__ void t5_scan_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ out, long long n) {
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

__global__ void t5_chunk_count_kernel(const uint32_t *__restrict__ batch_start, uint32_t *__restrict__ counts, long long units) {
  const long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= units) return;
  const uint32_t nb = batch_start[u + 1] - batch_start[u];
  counts[u] = (nb + kChunkBatches - 1) / kChunkBatches;
}

// chunk = (tile, first batch, end batch): a run of at most kChunkBatches batches of one work unit
__global__ void t5_chunk_fill_kernel(const uint32_t *__restrict__ batch_start, const uint32_t *__restrict__ chunk_start,
                                     uint4 *__restrict__ chunks, long long units, int zseg) {
  const long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= units) return;
  const uint32_t b0 = batch_start[u], nb = batch_start[u + 1] - b0;
  const uint32_t cnt = chunk_start[u + 1] - chunk_start[u];
  if (cnt == 0) return;
  const uint32_t size = (nb + cnt - 1) / cnt;
  for (uint32_t k = 0; k < cnt; k++) {
    const uint32_t lo = b0 + k * size, hi = (k + 1 == cnt) ? b0 + nb : lo + size;
    chunks[chunk_start[u] + k] = make_uint4((uint32_t) (u / zseg), lo, hi, 0u);
  }
}

// ---- operand images (plan time) -------------------------------------------------------------------------------
struct WinParams {
  double b[3], ws[3], m2;
  int window;
};

// one warp per batch; the image buffer was cleared before, only the taps are written
template <bool SPREAD>
__global__ void __launch_bounds__(128)
t5_images_kernel(const float *__restrict__ xt, const uint32_t *__restrict__ perm, const uint4 *__restrict__ chunks,
                 const uint2 *__restrict__ table, unsigned char *__restrict__ img, MmaParams P, WinParams Wp) {
  const uint4 chunk = chunks[blockIdx.x];
  const int tile = (int) chunk.x;
  const int a = tile / P.NT1, bt = tile - a * P.NT1;
  const int W = 2 * P.m + 2;
  const int lane = threadIdx.x & 31;
  for (uint32_t b = chunk.y + (threadIdx.x >> 5); b < chunk.z; b += 4) {
    const uint2 e = table[b];
    const int base = t5_base(e), nb = t5_nb(e);
    unsigned char *im = img + (size_t) b * (SPREAD ? kImgBytesS : kImgBytes);
    if (!SPREAD && lane < 16) reinterpret_cast<uint32_t *>(im + kOffPerm)[lane] = lane < nb ? perm[e.x + lane] : 0u;
    if (lane == 16) *reinterpret_cast<uint2 *>(im + (SPREAD ? kOffEntryS : kOffEntry)) = e;
    for (int i = lane; i < nb * 3 * W; i += 32) {
      const int n = i / (3 * W), r = i - n * 3 * W, t = r / W, l = r - t * W;
      const float x = xt[3 * (size_t) (e.x + n) + t];
      const int nn = t == 0 ? P.n0 : (t == 1 ? P.n1 : P.n2);
      const long long u = cell_of(x, nn) - P.m;
      const double dist = (double) x * (double) nn - (double) (u + l);
      const float v = (float) window_phi(dist, Wp.m2, Wp.b[t], Wp.window, Wp.ws[t]);
      const int uw = wrapi((int) u, nn);
      if (t == 2) {
        const int s = uw - base + l;   // 0 .. 23 (spreading: 0 .. 31)
        const float hi = tf32_rna(v);
        *reinterpret_cast<float *>(im + (SPREAD ? sb_off(n, s) : b_off(n, s))) = hi;
        *reinterpret_cast<float *>(im + (SPREAD ? kOffLoS + sb_off(n, s) : kOffLo + b_off(n, s))) = v - hi;
      } else {
        const int row = uw - P.T * (t == 0 ? a : bt) + l;   // 0 .. 15
        *reinterpret_cast<float *>(im + (SPREAD ? (t == 0 ? kOffP0S : kOffP1S) : (t == 0 ? kOffP0 : kOffP1)) + w_off(row, n)) = v;
      }
    }
  }
}

// ---- interpolation --------------------------------------------------------------------------------------------
struct __align__(128) SmemI {
  unsigned char ops[kOpStages][kImgBytes];
  float red[kEpi][2][4][32];
  int done_upto[kMmaWarps];        // stream m: every batch = m (mod 2) up to done_upto[m] has committed (written by the epilogue
                                   // warps, polled by the refill warps)
  uint64_t op_full[kOpStages], op_empty[kOpStages];
  uint64_t acc_full[kDone];        // commit of batch j arrives on acc_full[j % kDone] (the accumulator stage is j % kAcc)
  uint64_t acc_empty[kAcc];
  uint64_t a_ready[kSlides];       // one phase per slide of 8 cells
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreadsI, 1)
tc5_interp_kernel(const float2 *__restrict__ G, float *__restrict__ f, const uint4 *__restrict__ chunks, int nchunks,
                  const uint2 *__restrict__ table, const unsigned char *__restrict__ img, MmaParams P, int dbg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemI &S = *reinterpret_cast<SmemI *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kOpStages; i++) { mbar_init(&S.op_full[i], 1); mbar_init(&S.op_empty[i], 4); }
    for (int i = 0; i < kDone; i++) mbar_init(&S.acc_full[i], 1);
    for (int i = 0; i < kSlides; i++) mbar_init(&S.a_ready[i], 4);
    for (int i = 0; i < kMmaWarps; i++) S.done_upto[i] = -1;
    for (int i = 0; i < kAcc; i++) mbar_init(&S.acc_empty[i], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 * kEpi + 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = S.tmem_base;
  const int n2 = P.n2;
  unsigned long long tdbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define CLK() ((dbg & 8) ? clock64() : 0ll)
  const long long t_begin = CLK();
  // every role walks the same sequence of batches: chunks blockIdx.x, + gridDim.x, ...; jg counts them.  The batch entry
  // (first node, base, count) and the output indices of its nodes travel inside the operand image, so no role has a
  // dependent global load in its loop.

  if (warp < 4 * kEpi) {
    // ===== epilogue: lane quarter q, row r = 32 q + lane of every row block; block b = 2 c + h holds component c (re / im)
    // of pencil p = 128 h + r, i.e. l0 = 8 h + 2 q + (lane >> 4), l1 = lane & 15
    const int q = warp & 3, grp = warp >> 2;
    const uint32_t lane_base = (uint32_t) (q * 32) << 16;
    const int l1 = lane & 15, l0a = 2 * q + (lane >> 4);
    long long jg = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      for (int j = 0; j < nbat; j++, jg++) {
        const int st = (int) (jg % kOpStages), s = (int) (jg % kAcc);
        if ((int) (jg % kEpi) != grp) continue;   // the other group's batch
        const long long c0 = CLK();
        mbar_wait(&S.acc_full[jg % kDone], (int) ((jg / kDone) & 1));
        const long long c1 = CLK();
        if (q == 0 && lane == 0) atomicMax(&S.done_upto[jg % kMmaWarps], (int) jg);   // a warp's commits complete in order
        tc_fence_after();
        float v[4][16];
#pragma unroll
        for (int b = 0; b < 4; b++) tmem_ld16(tb + lane_base + 64 * s + 16 * b, v[b]);
        mbar_wait(&S.op_full[st], (int) ((jg / kOpStages) & 1));   // long complete: makes the TMA writes visible here
        const unsigned char *op = S.ops[st];
        const int nb = t5_nb(*reinterpret_cast<const uint2 *>(op + kOffEntry));
        const uint32_t pj = reinterpret_cast<const uint32_t *>(op + kOffPerm)[lane & 15];
        float w[2][16];   // row weights (psi0[l0a] psi1[l1], psi0[l0a + 8] psi1[l1]) of the 16 nodes
#pragma unroll
        for (int cq = 0; cq < 4; cq++) {
          const float4 p1 = *reinterpret_cast<const float4 *>(op + kOffP1 + l1 * 64 + (((cq + (l1 >> 1)) & 3) << 4));
          const float4 pa = *reinterpret_cast<const float4 *>(op + kOffP0 + l0a * 64 + (((cq + (l0a >> 1)) & 3) << 4));
          const float4 pb = *reinterpret_cast<const float4 *>(op + kOffP0 + (l0a + 8) * 64 + (((cq + ((l0a + 8) >> 1)) & 3) << 4));
          w[0][4 * cq] = pa.x * p1.x; w[0][4 * cq + 1] = pa.y * p1.y; w[0][4 * cq + 2] = pa.z * p1.z; w[0][4 * cq + 3] = pa.w * p1.w;
          w[1][4 * cq] = pb.x * p1.x; w[1][4 * cq + 1] = pb.y * p1.y; w[1][4 * cq + 2] = pb.z * p1.z; w[1][4 * cq + 3] = pb.w * p1.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.op_empty[st]);   // the image stage was read
        tmem_ld_wait();
        const long long c2 = CLK();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.acc_empty[s]);   // T was read
        float acc[32];   // [c][n]
#pragma unroll
        for (int n = 0; n < 16; n++) {
          acc[n] = fmaf(w[1][n], v[1][n], w[0][n] * v[0][n]);
          acc[16 + n] = fmaf(w[1][n], v[3][n], w[0][n] * v[2][n]);
        }
        // butterfly: after the five steps lane L holds the warp's sum of value index L = 16 c + n
#pragma unroll
        for (int half = 16; half >= 1; half >>= 1) {
          const bool up = (lane & half) != 0;
#pragma unroll
          for (int i = 0; i < half; i++) {
            const float send = up ? acc[i] : acc[half + i];
            const float keep = up ? acc[half + i] : acc[i];
            acc[i] = keep + __shfl_xor_sync(kFull, send, half);
          }
        }
        const long long c3 = CLK();
        if ((dbg & 8) && warp == 0 && lane == 0) { tdbg[0] += c1 - c0; tdbg[1] += c2 - c1; tdbg[2] += c3 - c2; tdbg[3]++; }
        if (dbg & 1) continue;   // timing experiments only (NFFT_B200_TC5_DBG)
        float (*red)[32] = S.red[grp][(jg / kEpi) & 1];
        red[q][lane] = acc[0];
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if (q == (int) ((jg / kEpi) & 3)) {
          const float sum = (red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane]);
          const int n = lane & 15, c = lane >> 4;
          if (n < nb) f[2 * (size_t) pj + c] = sum;
        }
        if ((dbg & 8) && warp == 0 && lane == 0) tdbg[4] += clock64() - c3;
      }
    }
    if ((dbg & 8) && warp == 0 && lane == 0) { for (int i = 0; i < 5; i++) atomicAdd(&g_t5dbg[i], tdbg[i]); atomicAdd(&g_t5dbg[5], (unsigned long long) (clock64() - t_begin)); }
  } else if (warp < 4 * kEpi + 4) {
    // ===== window refill: thread (q, lane) owns row r = 32 q + lane of the four row blocks = pencils r and 128 + r.
    // The warps act per SLIDE (8 new cells), not per batch: they walk the batch table (32 entries per coalesced load) to
    // find the slides the batches need, and hand each slide to the MMA warp through a_ready[slide % kSlides].
    const int q = warp & 3;
    const uint32_t lane_base = (uint32_t) (q * 32) << 16;
    const int r = 32 * q + lane;
    long long jg0 = 0;     // index of the chunk's first batch in the CTA's batch sequence
    long long sg = 0;      // slides handed over so far
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      const int tile = (int) chunk.x;
      const int a = tile / P.NT1, bt = tile - a * P.NT1;
      unsigned rowoff[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int p = 128 * h + r;
        rowoff[h] = (unsigned) ((wrapi(P.T * a + (p >> 4), P.n0) * (long long) P.n1 + wrapi(P.T * bt + (p & 15), P.n1)) * n2);
      }
      // two cursors into the chunk's batch table, each with 32 bases cached in the lanes
      int nj0 = 0, nbase = lane < nbat ? t5_base(table[chunk.y + lane]) : 0x3fffffff;   // needs: batches [nj0, nj0 + 32)
      int hj0 = 0, hbase = nbase;                                                         // hazards
      int hz = 0;            // batches [0, hz) of the chunk are known to have committed
      int whi = 0;           // cells [.., whi) of the chunk's sweep are in the ring
      // the next 8 cells [pz, pz + 8) of the thread's two pencils, loaded from L2 ahead of their use (the loads do not
      // touch tensor memory, so they need not wait for a hazard)
      float4 pre[2][4];
      int pz = -1;
      auto prefetch = [&](int z0) {
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            int z = z0 + 2 * i;
            if (z >= n2) z %= n2;
            pre[h][i] = *reinterpret_cast<const float4 *>(G + rowoff[h] + z);   // cells z, z + 1 (z even, n2 even)
          }
        pz = z0;
      };
      for (int j = 0; j < nbat; j++) {
        if (j >= nj0 + 32) { nj0 += 32; nbase = nj0 + lane < nbat ? t5_base(table[chunk.y + nj0 + lane]) : 0x3fffffff; }
        const int base = __shfl_sync(kFull, nbase, j - nj0);
        if (j == 0) { whi = base; prefetch(base); }
        if (whi < base) whi = base;
        for (; whi < base + kSpan; whi += 8) {
          // cells below whi + 8 - kRing are replaced: every batch whose base lies below that may still read them and
          // must have committed; so must every batch of the earlier chunks (another tile)
          const int dead = whi + 8 - kRing;
          for (;;) {
            if (hz >= hj0 + 32) { hj0 += 32; hbase = hj0 + lane < nbat ? t5_base(table[chunk.y + hj0 + lane]) : 0x3fffffff; }
            if (hz >= j || __shfl_sync(kFull, hbase, hz - hj0) >= dead) break;
            hz++;
          }
          const long long need = jg0 + hz - 1;   // last batch that must have committed (-1: none)
          const long long c0 = CLK();
          if (need >= 0) {
            // batch `need` and everything before it: the newest batch of either MMA stream
            const volatile int *du0 = &S.done_upto[need % kMmaWarps], *du1 = &S.done_upto[(need + 1) % kMmaWarps];
            while ((long long) *du0 < need) { }
            while ((long long) *du1 < need - 1) { }
          }
          tc_fence_after();
          const long long c1 = CLK();
          if (pz != whi) prefetch(whi);
          const int slot = whi % kRing;
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (dbg & 2) break;
            float re[8], im[8];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float4 g = pre[h][i];
              re[2 * i] = g.x; im[2 * i] = g.y; re[2 * i + 1] = g.z; im[2 * i + 1] = g.w;
            }
            float hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { hi[i] = tf32_rna(re[i]); lo[i] = re[i] - hi[i]; }
            tmem_st8(tb + lane_base + kColA + 2 * kRing * h + slot, hi);
            tmem_st8(tb + lane_base + kColA + 2 * kRing * h + kRing + slot, lo);
#pragma unroll
            for (int i = 0; i < 8; i++) { hi[i] = tf32_rna(im[i]); lo[i] = im[i] - hi[i]; }
            tmem_st8(tb + lane_base + kColA + 2 * kRing * (2 + h) + slot, hi);
            tmem_st8(tb + lane_base + kColA + 2 * kRing * (2 + h) + kRing + slot, lo);
          }
          const long long c2 = CLK();
          prefetch(whi + 8);   // needed by this batch (initial fill) or by one of the next
          tmem_st_wait();
          const long long c3 = CLK();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.a_ready[sg % kSlides]);
          sg++;
          if ((dbg & 8) && q == 0 && lane == 0) { tdbg[0] += c1 - c0; tdbg[1] += c2 - c1; tdbg[2] += c3 - c2; tdbg[3]++; }
        }
      }
      jg0 += nbat;
    }
    if ((dbg & 8) && q == 0 && lane == 0) { for (int i = 0; i < 4; i++) atomicAdd(&g_t5dbg[8 + i], tdbg[i]); }
  } else if (warp < 4 * kEpi + 4 + kMmaWarps) {
    // ===== MMA issue: the whole warp runs the loop (uniform operands), one elected lane issues
    constexpr uint32_t idesc = make_idesc(128, kN);
    const int mw = warp - (4 * kEpi + 4);
    long long jg = 0, sg = 0, sw = 0;   // batches walked, slides needed so far, slides this warp has waited for
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      int whi = 0;
      for (int j = 0; j < nbat; j++, jg++) {
        const int st = (int) (jg % kOpStages), s = (int) (jg % kAcc);
        const long long c0 = CLK();
        mbar_wait(&S.op_full[st], (int) ((jg / kOpStages) & 1));
        const long long c1 = CLK();
        // the batch entry travels in the image (reading it from the batch table through a lane cache measured 0.3 ms
        // slower per launch); with more than one issuing warp the image of another warp's batch could be gone already
        const unsigned ey = reinterpret_cast<const uint2 *>(S.ops[st] + kOffEntry)->y;
        const int base = (int) (ey & 0xffffffu);
        const int nk = ((ey >> 29) & 1u) ? 2 : 3;
        // the slides the batches need, in the order the refill warps produce them (same rule as theirs)
        if (j == 0 || whi < base) whi = base;
        for (; whi < base + kSpan; whi += 8) sg++;
        mbar_wait(&S.acc_empty[s], (int) (((jg / kAcc) & 1) ^ 1));
        const long long c2 = CLK();
        for (; sw < sg; sw++) mbar_wait(&S.a_ready[sw % kSlides], (int) ((sw / kSlides) & 1));
        const long long c3 = CLK();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bsm = smem_u32(S.ops[st]);
#pragma unroll
          for (int i = 0; i < 3; i++) {
            if (i >= nk || ((dbg & 4) && i > 0)) break;
            const uint32_t kb = (uint32_t) ((base + 8 * i) % kRing);
            const uint64_t bh = make_desc(bsm + 2 * i * kBChunk, kBChunk, kBGroup);
            const uint64_t bl = make_desc(bsm + kOffLo + 2 * i * kBChunk, kBChunk, kBGroup);
#pragma unroll
            for (int term = 0; term < 3; term++)
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const uint32_t a_hi = tb + kColA + 2 * kRing * b + kb;
                mma_ts(tb + 64 * s + 16 * b, term == 0 ? a_hi + kRing : a_hi, term == 1 ? bl : bh, idesc, (i | term) ? 1u : 0u);
              }
          }
          mma_commit(&S.acc_full[jg % kDone]);
        }
        __syncwarp();
        if ((dbg & 8) && lane == 0) { tdbg[0] += c1 - c0; tdbg[1] += c2 - c1; tdbg[2] += c3 - c2; tdbg[3] += clock64() - c3; tdbg[4]++; }
      }
    }
    if ((dbg & 8) && lane == 0 && mw == 0) { for (int i = 0; i < 5; i++) atomicAdd(&g_t5dbg[16 + i], tdbg[i]); }
  } else {
    // ===== feeder: one lane streams the operand images into the ring
    long long jg = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      for (int j = 0; j < nbat; j++, jg++) {
        const int st = (int) (jg % kOpStages);
        const long long c0 = CLK();
        mbar_wait<true>(&S.op_empty[st], (int) (((jg / kOpStages) & 1) ^ 1));
        if ((dbg & 8) && lane == 0) tdbg[0] += clock64() - c0;
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&S.op_full[st])), "r"(kImgBytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(S.ops[st])), "l"(img + (size_t) (chunk.y + j) * kImgBytes), "r"(kImgBytes),
                         "r"(smem_u32(&S.op_full[st])) : "memory");
        }
        __syncwarp();
      }
    }
    if ((dbg & 8) && lane == 0) atomicAdd(&g_t5dbg[24], tdbg[0]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * kEpi + 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}


// ---- spreading ------------------------------------------------------------------------------------------------
// G[row, z] += sum_node (psi0 psi1 f)[row, node] psi2[node, z]: the accumulators ARE the window -- a ring of 64 z-slots per
// row in TMEM columns [0, 256) (row block b at 64 b), fp32, alive for the whole sweep of a tile.  Per batch (<= 16 nodes,
// taps in [base, base + 32), base a multiple of 16):
//   operand warps (2 groups x 4): A[row, node] = psi0[l0] psi1[l1] f_c[node] for the thread's four rows, split into tf32
//                 (hi, lo), written to the group's A stage (128 columns) with tcgen05.st; a_full
//   MMA warp:     2 k-steps (8 nodes each) x 3 split terms x 4 row blocks of M = 128, N = 32 (two N = 16 pieces when the 32
//                 slots wrap around the ring), accumulating; commits release the A stage and the image stage
//   retire warps: a block of 16 cells leaves the ring when every batch that touches it has completed: tcgen05.ld,
//                 red.global.add.v4.f32 (two cells of a pencil, zero quadruples skipped), columns cleared with tcgen05.st
//   feeder:       image (psi2 as the B operand with K = nodes, psi0 / psi1) and the batch's samples by TMA
constexpr int kColAS = 256;   // A stages: kColAS + 128 stage + 32 b (+16: lo) + node

struct __align__(128) SmemS {
  unsigned char ops[kOpStagesS][kStageS];
  int done_upto;                   // every batch <= done_upto has completed (published by the operand warps)
  int retired[4];                  // retire warp q: every block <= retired[q] (CTA-wide block sequence) is in the grid and cleared
  uint64_t op_full[kOpStagesS], op_empty[kOpStagesS];
  uint64_t a_full[kOpG], a_empty[kOpG];
  uint32_t tmem_base;
};

__global__ void t5_gather_f_kernel(const float2 *__restrict__ f, const uint32_t *__restrict__ perm, float2 *__restrict__ ft,
                                   long long M) {
  const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (k < M) ft[k] = f[perm[k]];
}

__global__ void __launch_bounds__(kThreadsS, 1)
tc5_spread_kernel(float2 *__restrict__ G, const float2 *__restrict__ ft, const uint4 *__restrict__ chunks, int nchunks,
                  const uint2 *__restrict__ table, const unsigned char *__restrict__ img, MmaParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemS &S = *reinterpret_cast<SmemS *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kWarpRetire = 4 * kOpG, kWarpMma = 4 * kOpG + 4;
  if (tid == 0) {
    for (int i = 0; i < kOpStagesS; i++) { mbar_init(&S.op_full[i], 1); mbar_init(&S.op_empty[i], 1); }
    for (int i = 0; i < kOpG; i++) { mbar_init(&S.a_full[i], 4); mbar_init(&S.a_empty[i], 1); }
    S.done_upto = -1;
    for (int i = 0; i < 4; i++) S.retired[i] = -1;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = S.tmem_base;
  const int n2 = P.n2;
  if (warp >= kWarpRetire && warp < kWarpMma) {   // clear the accumulator ring
    const uint32_t lane_base = (uint32_t) ((warp & 3) * 32) << 16;
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) z[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 16; c++) tmem_st16(tb + lane_base + 16 * c, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kWarpRetire) {
    // ===== operand warps: lane quarter q, row r = 32 q + lane of every row block; block b = 2 c + h is component c of
    // pencil p = 128 h + r, i.e. l0 = 8 h + 2 q + (lane >> 4), l1 = lane & 15
    const int q = warp & 3, grp = warp >> 2;
    const uint32_t lane_base = (uint32_t) (q * 32) << 16;
    const int l1 = lane & 15, l0a = 2 * q + (lane >> 4);
    long long jg = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      for (int j = 0; j < nbat; j++, jg++) {
        if ((int) (jg % kOpG) != grp) continue;   // the other group's batch
        const int st = (int) (jg % kOpStagesS);
        mbar_wait(&S.op_full[st], (int) ((jg / kOpStagesS) & 1));
        const unsigned char *op = S.ops[st];
        const uint2 e = *reinterpret_cast<const uint2 *>(op + kOffEntryS);
        const float2 *fs = reinterpret_cast<const float2 *>(op + kImgBytesS) + (e.x & 1u);   // samples of nodes e.x ...
        float w[2][16];   // row weights (psi0[l0a] psi1[l1], psi0[l0a + 8] psi1[l1]) of the 16 nodes
#pragma unroll
        for (int cq = 0; cq < 4; cq++) {
          const float4 p1 = *reinterpret_cast<const float4 *>(op + kOffP1S + l1 * 64 + (((cq + (l1 >> 1)) & 3) << 4));
          const float4 pa = *reinterpret_cast<const float4 *>(op + kOffP0S + l0a * 64 + (((cq + (l0a >> 1)) & 3) << 4));
          const float4 pb = *reinterpret_cast<const float4 *>(op + kOffP0S + (l0a + 8) * 64 + (((cq + ((l0a + 8) >> 1)) & 3) << 4));
          w[0][4 * cq] = pa.x * p1.x; w[0][4 * cq + 1] = pa.y * p1.y; w[0][4 * cq + 2] = pa.z * p1.z; w[0][4 * cq + 3] = pa.w * p1.w;
          w[1][4 * cq] = pb.x * p1.x; w[1][4 * cq + 1] = pb.y * p1.y; w[1][4 * cq + 2] = pb.z * p1.z; w[1][4 * cq + 3] = pb.w * p1.w;
        }
        float fr[16], fi[16];
#pragma unroll
        for (int n = 0; n < 16; n++) { const float2 v = fs[n]; fr[n] = v.x; fi[n] = v.y; }
        // the group's A stage: free once the MMAs of the group's previous batch have completed
        mbar_wait(&S.a_empty[grp], (int) (((jg / kOpG) & 1) ^ 1));
        if (q == 0 && lane == 0 && jg >= kOpG) atomicMax(&S.done_upto, (int) (jg - kOpG));   // MMAs complete in order
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            float hi[16], lo[16];
#pragma unroll
            for (int n = 0; n < 16; n++) {
              const float v = w[h][n] * (c ? fi[n] : fr[n]);
              hi[n] = tf32_rna(v);
              lo[n] = v - hi[n];
            }
            const uint32_t col = tb + lane_base + kColAS + 128 * grp + 32 * (2 * c + h);
            tmem_st16(col, hi);
            tmem_st16(col + 16, lo);
          }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.a_full[grp]);
      }
    }
    // the group's last batch: publish its completion (the retire warps wait for it)
    const long long J = jg;
    long long last = J - 1;
    while (last >= 0 && (int) (last % kOpG) != grp) last--;
    if (last >= 0) {
      mbar_wait(&S.a_empty[grp], (int) ((last / kOpG) & 1));
      if (q == 0 && lane == 0) atomicMax(&S.done_upto, (int) last);
    }
  } else if (warp < kWarpMma) {
    // ===== retire warps: thread (q, lane) owns row r = 32 q + lane of the four row blocks = pencils r and 128 + r
    const int q = warp & 3;
    const uint32_t lane_base = (uint32_t) (q * 32) << 16;
    const int r = 32 * q + lane;
    long long jg0 = 0;
    int goff = 0;   // block sequence number of the chunk's first block
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      const int tile = (int) chunk.x;
      const int a = tile / P.NT1, bt = tile - a * P.NT1;
      unsigned rowoff[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int p = 128 * h + r;
        rowoff[h] = (unsigned) ((wrapi(P.T * a + (p >> 4), P.n0) * (long long) P.n1 + wrapi(P.T * bt + (p & 15), P.n1)) * n2);
      }
      const int kfirst = t5_base(table[chunk.y]) >> 4, klast = (t5_base(table[chunk.z - 1]) >> 4) + 1;
      int cj0 = 0, cbase = lane < nbat ? t5_base(table[chunk.y + lane]) : 0x3fffffff;   // cursor: bases of batches [cj0, cj0 + 32)
      int jp = 0;   // batches [0, jp) of the chunk have base / 16 <= k
      for (int k = kfirst; k <= klast; k++) {
        for (;;) {
          if (jp >= cj0 + 32) { cj0 += 32; cbase = cj0 + lane < nbat ? t5_base(table[chunk.y + cj0 + lane]) : 0x3fffffff; }
          if (jp >= nbat || (__shfl_sync(kFull, cbase, jp - cj0) >> 4) > k) break;
          jp++;
        }
        // block k is touched by the batches with base / 16 in {k - 1, k}: all of [0, jp) must have completed
        const long long need = jg0 + jp - 1;
        {
          const volatile int *du = &S.done_upto;
          while ((long long) *du < need) { }
        }
        tc_fence_after();
        const uint32_t col = (uint32_t) ((16 * k) & (kRingS - 1));
        float v[4][16];
#pragma unroll
        for (int b = 0; b < 4; b++) tmem_ld16(tb + lane_base + 64 * b + col, v[b]);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float re0 = v[h][i], im0 = v[2 + h][i], re1 = v[h][i + 1], im1 = v[2 + h][i + 1];
            if (re0 != 0.f || im0 != 0.f || re1 != 0.f || im1 != 0.f) {
              int z = 16 * k + i;
              if (z >= n2) z %= n2;
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                           ::"l"(G + rowoff[h] + z), "f"(re0), "f"(im0), "f"(re1), "f"(im1) : "memory");
            }
          }
        float zr[16];
#pragma unroll
        for (int i = 0; i < 16; i++) zr[i] = 0.f;
#pragma unroll
        for (int b = 0; b < 4; b++) tmem_st16(tb + lane_base + 64 * b + col, zr);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile int *>(&S.retired[q]) = goff + (k - kfirst);
      }
      goff += klast - kfirst + 1;
      jg0 += nbat;
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issue: the whole warp runs the loop (uniform operands), one elected lane issues
    constexpr uint32_t idesc32 = make_idesc(128, 32), idesc16 = make_idesc(128, 16);
    long long jg = 0;
    int goff = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      const int klast = (t5_base(table[chunk.z - 1]) >> 4) + 1;   // needed at the end of the chunk only
      int kfirst = 0;
      for (int j = 0; j < nbat; j++, jg++) {
        const int st = (int) (jg % kOpStagesS), as = (int) (jg % kOpG);
        mbar_wait(&S.op_full[st], (int) ((jg / kOpStagesS) & 1));
        const uint2 e = *reinterpret_cast<const uint2 *>(S.ops[st] + kOffEntryS);
        const int base = t5_base(e), kb = base >> 4;
        if (j == 0) kfirst = kb;
        mbar_wait(&S.a_full[as], (int) ((jg / kOpG) & 1));
        // the ring positions of blocks kb, kb + 1 were those of kb - 4, kb - 3: they and every block of the earlier chunks
        // (another tile) must be in the grid and cleared
        const int need = (kb - 3 >= kfirst) ? goff + (kb - 3 - kfirst) : goff - 1;
        {
          const volatile int *rt = S.retired;
          while (min(min(rt[0], rt[1]), min(rt[2], rt[3])) < need) { }
        }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bsm = smem_u32(S.ops[st]);
          const uint32_t s0 = (uint32_t) (base & (kRingS - 1));
          const bool wraps = s0 + kSpanS > kRingS;   // base = 48 (mod 64): two pieces of 16 slots
          const int nks = t5_nb(e) > 8 ? 2 : 1;      // nodes 8 .. 15 absent: their k-step is all zero
          for (int ks = 0; ks < nks; ks++) {
            const uint64_t bh = make_desc(bsm + ks * 1024, 512, 128);
            const uint64_t bl = make_desc(bsm + kOffLoS + ks * 1024, 512, 128);
#pragma unroll
            for (int term = 0; term < 3; term++)
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const uint32_t a_hi = tb + kColAS + 128 * as + 32 * b + 8 * ks;
                const uint32_t aa = term == 0 ? a_hi + 16 : a_hi;
                const uint64_t bb = term == 1 ? bl : bh;
                if (!wraps) {
                  mma_ts(tb + 64 * b + s0, aa, bb, idesc32, 1u);
                } else {
                  mma_ts(tb + 64 * b + s0, aa, bb, idesc16, 1u);
                  mma_ts(tb + 64 * b, aa, bb + (uint64_t) (256 >> 4), idesc16, 1u);   // slot groups 2, 3 -> ring slots 0 .. 15
                }
              }
          }
          mma_commit(&S.a_empty[as]);
          mma_commit(&S.op_empty[st]);
        }
        __syncwarp();
      }
      goff += klast - kfirst + 1;
    }
  } else {
    // ===== feeder: image and samples of every batch into the ring
    long long jg = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      int cj0 = 0;
      uint32_t cx = lane < nbat ? table[chunk.y + lane].x : 0u;   // first nodes of batches [cj0, cj0 + 32)
      for (int j = 0; j < nbat; j++, jg++) {
        if (j >= cj0 + 32) { cj0 += 32; cx = cj0 + lane < nbat ? table[chunk.y + cj0 + lane].x : 0u; }
        const uint32_t first = __shfl_sync(kFull, cx, j - cj0);
        const int st = (int) (jg % kOpStagesS);
        mbar_wait<true>(&S.op_empty[st], (int) (((jg / kOpStagesS) & 1) ^ 1));
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&S.op_full[st])), "r"(kImgBytesS + 144) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(S.ops[st])), "l"(img + (size_t) (chunk.y + j) * kImgBytesS), "r"(kImgBytesS),
                         "r"(smem_u32(&S.op_full[st])) : "memory");
          // 18 samples from the even node index below `first`: 16-byte aligned source, covers first .. first + 15
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(S.ops[st] + kImgBytesS)), "l"(ft + (first & ~1u)), "r"(144),
                         "r"(smem_u32(&S.op_full[st])) : "memory");
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

}  // namespace

// NFFTCU_OPT_TC5: 0 auto (B and B^T on the tcgen05 kernels) | 1 off | 2 B only | 3 B and B^T
bool tc5_selected(const nfftcu_ctx *c) {
  // env NFFT_B200_TC5=0: the switch for callers of the plan API, who cannot reach nfftcu_set_option
  static const bool env_off = getenv("NFFT_B200_TC5") && atoi(getenv("NFFT_B200_TC5")) == 0;
  if (env_off && c->opt_tc5 == 0) return false;
  if (c->prec != NFFTCU_FLOAT || c->opt_tc5 == 1 || !mma3d_supported(c)) return false;
  if (!(c->opt_b_kernel == 0)) return false;   // an explicit kernel choice (generic / pencils / DMMA) wins
  return c->opt_tc5 != 1;
}

namespace {

// batch table, chunk list and (cleared) image buffer of one kernel: SPREAD selects the batch rule and the image size
template <bool SPREAD>
int build_tables(nfftcu_ctx *c, const MmaParams &P, Tc5Tables &T, bool *ok) {
  *ok = false;
  const long long units = (long long) P.NT0 * P.NT1 * P.zseg;
  const int kb = 256;
  if (c->tc5_counts_units != units) {
    if (c->tc5_counts) pool_free(c->tc5_counts);
    c->tc5_counts = nullptr;
    NFFTCU_CUDA(pool_malloc((void **) &c->tc5_counts, sizeof(uint32_t) * (size_t) units));
    c->tc5_counts_units = units;
  }
  if (T.units != units) {
    if (T.batch_start) pool_free(T.batch_start);
    if (T.chunk_start) pool_free(T.chunk_start);
    T.batch_start = T.chunk_start = nullptr;
    NFFTCU_CUDA(pool_malloc((void **) &T.batch_start, sizeof(uint32_t) * (size_t) (units + 1)));
    NFFTCU_CUDA(pool_malloc((void **) &T.chunk_start, sizeof(uint32_t) * (size_t) (units + 1)));
    T.units = units;
  }
  constexpr int ALIGN = SPREAD ? 16 : 8, SPAN = SPREAD ? kSpanS : kSpan;
  const unsigned wgrid = (unsigned) ((units * 32 + kb - 1) / kb);
  t5_batches_kernel<false, ALIGN, SPAN><<<wgrid, kb, 0, c->stream>>>((const uint64_t *) c->tile_keys, c->bin_start, c->tc5_counts,
                                                                     nullptr, nullptr, units, P);
  t5_scan_kernel<<<1, 1024, 0, c->stream>>>(c->tc5_counts, T.batch_start, units);
  uint32_t total = 0;
  NFFTCU_CUDA(cudaMemcpyAsync(&total, T.batch_start + units, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  if ((long long) total > T.batch_cap) {
    if (T.batches) pool_free(T.batches);
    T.batches = nullptr;
    T.batch_cap = (long long) total + total / 8 + 1024;
    NFFTCU_CUDA(pool_malloc(&T.batches, sizeof(uint2) * (size_t) T.batch_cap));
  }
  t5_batches_kernel<true, ALIGN, SPAN><<<wgrid, kb, 0, c->stream>>>((const uint64_t *) c->tile_keys, c->bin_start, nullptr,
                                                                    T.batch_start, (uint2 *) T.batches, units, P);
  const unsigned ugrid = (unsigned) ((units + kb - 1) / kb);
  t5_chunk_count_kernel<<<ugrid, kb, 0, c->stream>>>(T.batch_start, c->tc5_counts, units);
  t5_scan_kernel<<<1, 1024, 0, c->stream>>>(c->tc5_counts, T.chunk_start, units);
  uint32_t nchunks = 0;
  NFFTCU_CUDA(cudaMemcpyAsync(&nchunks, T.chunk_start + units, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  if ((long long) nchunks > T.chunk_cap) {
    if (T.chunks) pool_free(T.chunks);
    T.chunks = nullptr;
    T.chunk_cap = (long long) nchunks + nchunks / 8 + 1024;
    NFFTCU_CUDA(pool_malloc(&T.chunks, sizeof(uint4) * (size_t) T.chunk_cap));
  }
  t5_chunk_fill_kernel<<<ugrid, kb, 0, c->stream>>>(T.batch_start, T.chunk_start, (uint4 *) T.chunks, units, P.zseg);
  T.nchunks = nchunks;
  T.nbatches = total;
  c->launches += 6;
  NFFTCU_CUDA(cudaGetLastError());
  if (total == 0) { *ok = true; return NFFTCU_OK; }
  // operand images, always resident (the kernels have no evaluating fallback; a plan whose images do not fit keeps the
  // mma.sync kernels)
  const size_t need = (size_t) (SPREAD ? kImgBytesS : kImgBytes) * (size_t) total;
  if (!T.images || T.images_bytes < need) {
    if (T.images) pool_free(T.images);
    T.images = nullptr;
    T.images_bytes = 0;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (need + need / 16 > free_b / 2 || pool_malloc(&T.images, need + need / 16) != cudaSuccess) {
      cudaGetLastError();
      return NFFTCU_OK;   // *ok stays false: the caller falls back to the mma.sync path
    }
    T.images_bytes = need + need / 16;
  }
  NFFTCU_CUDA(cudaMemsetAsync(T.images, 0, need, c->stream));
  WinParams Wp;
  for (int t = 0; t < 3; t++) { Wp.b[t] = c->b[t]; Wp.ws[t] = c->wscale[t]; }
  Wp.m2 = (double) c->m * (double) c->m;
  Wp.window = c->window;
  t5_images_kernel<SPREAD><<<(unsigned) nchunks, 128, 0, c->stream>>>((const float *) c->tile_x, c->tile_perm, (const uint4 *) T.chunks,
                                                                      (const uint2 *) T.batches, (unsigned char *) T.images, P, Wp);
  c->launches += 2;
  NFFTCU_CUDA(cudaGetLastError());
  *ok = true;
  return NFFTCU_OK;
}

}  // namespace

int tc5_build(nfftcu_ctx *c, const MmaParams &P) {
  c->tc5_ready = c->tc5s_ready = false;
  NFFTCU_TRY(build_tables<false>(c, P, c->tc5i, &c->tc5_ready));
  if (c->tc5_ready && c->opt_tc5 != 2) {
    NFFTCU_TRY(build_tables<true>(c, P, c->tc5s, &c->tc5s_ready));
    if (c->tc5s_ready && c->tc5_ft_cap < c->M + 18) {
      // samples in tile order, padded: the feeder copies 18 samples from an even node index on, also for the last batch
      if (c->tc5_ft) pool_free(c->tc5_ft);
      c->tc5_ft = nullptr;
      c->tc5_ft_cap = 0;
      NFFTCU_CUDA(pool_malloc(&c->tc5_ft, sizeof(float2) * (size_t) (c->M + 18)));
      c->tc5_ft_cap = c->M + 18;
    }
    if (c->tc5s_ready) NFFTCU_CUDA(cudaMemsetAsync((char *) c->tc5_ft + sizeof(float2) * (size_t) c->M, 0, sizeof(float2) * 18, c->stream));
  }
  return NFFTCU_OK;
}

int tc5_interp(nfftcu_ctx *c, void *f_dev) {
  const Tc5Tables &T = c->tc5i;
  if (T.nchunks == 0) return NFFTCU_OK;
  const MmaParams P = mma3d_params(c);
  unsigned grid = (unsigned) c->sm_count;
  if ((long long) grid > T.nchunks) grid = (unsigned) T.nchunks;
  // the CTA owns all 512 TMEM columns of its SM: at least 120 KB of dynamic shared memory keep a second CTA off the SM,
  // which would otherwise sit in tcgen05.alloc until the first one exits
  const int kOneCtaSmem = sizeof(SmemI) > 120 * 1024 ? (int) sizeof(SmemI) : 120 * 1024;
  NFFTCU_CUDA(cudaFuncSetAttribute(tc5_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOneCtaSmem));
  static const int dbg = getenv("NFFT_B200_TC5_DBG") ? atoi(getenv("NFFT_B200_TC5_DBG")) : 0;
  if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
  tc5_interp_kernel<<<grid, kThreadsI, kOneCtaSmem, c->stream>>>((const float2 *) c->grid, (float *) f_dev,
                                                               (const uint4 *) T.chunks, (int) T.nchunks,
                                                               (const uint2 *) T.batches, (const unsigned char *) T.images, P, dbg);
  if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}


int tc5_spread(nfftcu_ctx *c, const void *f_dev) {
  const Tc5Tables &T = c->tc5s;
  if (T.nchunks == 0) return NFFTCU_OK;
  const MmaParams P = mma3d_params(c);
  const int kb = 256;
  t5_gather_f_kernel<<<(unsigned) ((c->M + kb - 1) / kb), kb, 0, c->stream>>>((const float2 *) f_dev, c->tile_perm,
                                                                             (float2 *) c->tc5_ft, c->M);
  unsigned grid = (unsigned) c->sm_count;
  if ((long long) grid > T.nchunks) grid = (unsigned) T.nchunks;
  const int smem = sizeof(SmemS) > 120 * 1024 ? (int) sizeof(SmemS) : 120 * 1024;   // one CTA per SM (see tc5_interp)
  NFFTCU_CUDA(cudaFuncSetAttribute(tc5_spread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
  tc5_spread_kernel<<<grid, kThreadsS, smem, c->stream>>>((float2 *) c->grid, (const float2 *) c->tc5_ft, (const uint4 *) T.chunks,
                                                          (int) T.nchunks, (const uint2 *) T.batches, (const unsigned char *) T.images, P);
  if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
  c->launches += 2;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

}  // namespace nfftcu

// debugging aid (tools/tc5_check.py): the probe totals of the last launches, cleared on read
extern "C" int nfftcu_tc5_debug(unsigned long long *out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, nfftcu::g_t5dbg, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(nfftcu::g_t5dbg, z, sizeof(z));
  return 0;
}
