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
// of 32 z-slots, as a (hi, lo) pair -- lives in 256 TMEM columns for the whole sweep, is refilled 8 cells at a time with
// tcgen05.st, and ONE thread issues the contraction.  Measured on B200 (profiles/r2l_microbench_tcgen05.txt): an
// M = 128, N = 16, K = 8 kind::tf32 MMA with A in TMEM retires every 17 cycles (A in shared memory: 78).
//
//   interpolation   T[row, node] = sum_z G[row, z] psi2[z, node]       row = (re/im, pencil p = 16 l0 + l1), z = 24 slots
//                   f_node       = sum_row (psi0[l0] psi1[l1])[node] T[row, node]           (epilogue, FFMA + shuffles)
//   spreading       G[row, z]   += sum_node (psi0 psi1 f)[row, node] psi2[node, z]            (accumulators = the window)
//
// Batches.  Nodes are in the (tile, u2) order of mma3d.cu.  A batch = up to 16 consecutive nodes of a tile whose taps
// lie in the 24 cells [base, base + 24), base = 8 floor(u2_first / 8): three k-steps of 8 slots, each starting at an
// 8-slot boundary of the ring (slot = z mod 32), so the k-steps never wrap and the A operand address is just a column
// offset.  psi2 is zero outside a node's taps, so whatever else the 24 slots hold only has to be finite.
//
// Per batch the plan-time IMAGE (5 KB, like the reference's PRE_PSI table) holds psi2 as the (hi, lo) B operand in the
// canonical K-major no-swizzle layout and psi0 / psi1 placed in the footprint; a feeder lane streams it into a
// shared-memory ring by TMA (cp.async.bulk + mbarrier expect_tx).
//
// CTA = 10 warps, persistent over the chunk list, one CTA per SM (it owns all 512 TMEM columns):
//   warps 0-3  epilogue: tcgen05.ld of the batch's T (lane quarter q = warp), row weights, 32-value butterfly reduction,
//              f[perm[node]] stored by one warp per batch
//   warps 4-7  window refill: 8 cells x 2 pencils per thread from L2, split into tf32 (hi, lo), tcgen05.st
//   warp  8    MMA issue (one elected lane): 3 k-steps x 3 split terms x 4 row blocks = 36 tcgen05.mma + 1 commit
//   warp  9    TMA feeder of the operand images
// Hand-offs are mbarriers only: op_full / op_empty (image ring), a_ready (window covers the batch), acc_full (commit:
// T is complete) / acc_empty (T was read).  The refill warps run up to three batches ahead; a slot is only overwritten
// once every batch that still reads its old cell has committed (bases at least 16 cells behind).
#include "common.cuh"
#include "mma3d.cuh"

namespace nfftcu {

namespace {

constexpr int kN = 16;           // nodes per batch = N of the MMA
constexpr int kRing = 32;        // z-slots of the window ring
constexpr int kSpan = 24;        // slots one batch contracts: 3 k-steps of 8
constexpr int kOpStages = 6;     // operand-image ring
constexpr int kAcc = 4;          // accumulator stages (4 row blocks x 16 columns each)
constexpr int kImgBytes = 5120;  // psi2 hi 1536 | psi2 lo 1536 | psi0 1024 | psi1 1024
constexpr int kOffLo = 1536, kOffP0 = 3072, kOffP1 = 4096;
constexpr int kBChunk = 256, kBGroup = 128;   // B operand: (slot/4)*256 + (node/8)*128 + (node%8)*16 + (slot%4)*4
constexpr int kColA = 256;       // TMEM: D stage s at 64 s + 16 b, A ring of row block b at 256 + 64 b (+32: lo) + slot
constexpr int kChunkBatches = 256;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kOneCtaSmem = 120 * 1024;   // dynamic shared memory requested only to make the kernel one CTA per SM

__device__ __forceinline__ int b_off(int n, int s) { return (s >> 2) * kBChunk + (n >> 3) * kBGroup + (n & 7) * 16 + (s & 3) * 4; }
// psi0 / psi1 rows of 16 floats; the four 16-byte chunks of row l are rotated by l >> 1 so that the eight rows a quarter
// warp reads in one LDS.128 phase fall into eight different bank groups
__device__ __forceinline__ int w_off(int l, int n) { return l * 64 + ((((n >> 2) + (l >> 1)) & 3) << 4) + (n & 3) * 4; }

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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680) : "memory");
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- batch table and chunks (plan time) ---------------------------------------------------------------------
// entry.x = first node (tile order), entry.y = base | nb << 24.  One warp per work unit (see mma3d.cu) walks the unit's
// sorted nodes 32 at a time: a batch starts at the first unassigned node, base = its u2 rounded down to a multiple of 8,
// and takes up to 16 nodes with u2 + W <= base + 24.
template <bool FILL>
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
      const int base = __shfl_sync(kFull, u, o) & ~7;
      const bool member = lane >= o && lane < o + kN && u + W <= base + kSpan;
      const int nb = __popc(__ballot_sync(kFull, member));
      if (FILL && lane == 0) table[out + nbat] = make_uint2((uint32_t) (pos + o), (unsigned) base | ((unsigned) nb << 24));
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
__global__ void __launch_bounds__(128)
t5_images_kernel(const float *__restrict__ xt, const uint4 *__restrict__ chunks, const uint2 *__restrict__ table,
                 unsigned char *__restrict__ img, MmaParams P, WinParams Wp) {
  const uint4 chunk = chunks[blockIdx.x];
  const int tile = (int) chunk.x;
  const int a = tile / P.NT1, bt = tile - a * P.NT1;
  const int W = 2 * P.m + 2;
  const int lane = threadIdx.x & 31;
  for (uint32_t b = chunk.y + (threadIdx.x >> 5); b < chunk.z; b += 4) {
    const uint2 e = table[b];
    const int base = t5_base(e), nb = t5_nb(e);
    unsigned char *im = img + (size_t) b * kImgBytes;
    for (int i = lane; i < nb * 3 * W; i += 32) {
      const int n = i / (3 * W), r = i - n * 3 * W, t = r / W, l = r - t * W;
      const float x = xt[3 * (size_t) (e.x + n) + t];
      const int nn = t == 0 ? P.n0 : (t == 1 ? P.n1 : P.n2);
      const long long u = cell_of(x, nn) - P.m;
      const double dist = (double) x * (double) nn - (double) (u + l);
      const float v = (float) window_phi(dist, Wp.m2, Wp.b[t], Wp.window, Wp.ws[t]);
      const int uw = wrapi((int) u, nn);
      if (t == 2) {
        const int s = uw - base + l;   // 0 .. 23
        const float hi = tf32_rna(v);
        *reinterpret_cast<float *>(im + b_off(n, s)) = hi;
        *reinterpret_cast<float *>(im + kOffLo + b_off(n, s)) = v - hi;
      } else {
        const int row = uw - P.T * (t == 0 ? a : bt) + l;   // 0 .. 15
        *reinterpret_cast<float *>(im + (t == 0 ? kOffP0 : kOffP1) + w_off(row, n)) = v;
      }
    }
  }
}

// ---- interpolation --------------------------------------------------------------------------------------------
struct __align__(128) SmemI {
  unsigned char ops[kOpStages][kImgBytes];
  float red[kAcc][4][32];
  uint64_t op_full[kOpStages], op_empty[kOpStages];
  uint64_t acc_full[kAcc], acc_empty[kAcc];
  uint64_t a_ready[4];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(320, 1)
tc5_interp_kernel(const float2 *__restrict__ G, const uint32_t *__restrict__ perm, float *__restrict__ f,
                  const uint4 *__restrict__ chunks, int nchunks, const uint2 *__restrict__ table,
                  const unsigned char *__restrict__ img, MmaParams P) {
  __shared__ SmemI S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kOpStages; i++) { mbar_init(&S.op_full[i], 1); mbar_init(&S.op_empty[i], 4); }
    for (int i = 0; i < kAcc; i++) { mbar_init(&S.acc_full[i], 1); mbar_init(&S.acc_empty[i], 4); }
    for (int i = 0; i < 4; i++) mbar_init(&S.a_ready[i], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = S.tmem_base;
  const int n2 = P.n2;

  if (warp < 4) {
    // ===== epilogue: lane quarter q = warp, row r = 32 q + lane of every row block; block b = 2 c + h holds component
    // c (re / im) of pencil p = 128 h + r, i.e. l0 = 8 h + 2 q + (lane >> 4), l1 = lane & 15
    const int q = warp;
    const uint32_t lane_base = (uint32_t) (q * 32) << 16;
    const int l1 = lane & 15, l0a = 2 * q + (lane >> 4);
    long long jg = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      for (int j = 0; j < nbat; j++, jg++) {
        const int st = (int) (jg % kOpStages), s = (int) (jg & 3);
        const uint2 e = table[chunk.y + j];
        mbar_wait(&S.acc_full[s], (int) ((jg >> 2) & 1));
        tc_fence_after();
        float v[4][16];
#pragma unroll
        for (int b = 0; b < 4; b++) tmem_ld16(tb + lane_base + 64 * s + 16 * b, v[b]);
        mbar_wait(&S.op_full[st], (int) ((jg / kOpStages) & 1));   // long complete: makes the TMA writes visible here
        const unsigned char *op = S.ops[st];
        float acc[32];   // [c][n]
#pragma unroll
        for (int cq = 0; cq < 4; cq++) {
          const float4 p1 = *reinterpret_cast<const float4 *>(op + kOffP1 + l1 * 64 + (((cq + (l1 >> 1)) & 3) << 4));
          const float4 pa = *reinterpret_cast<const float4 *>(op + kOffP0 + l0a * 64 + (((cq + (l0a >> 1)) & 3) << 4));
          const float4 pb = *reinterpret_cast<const float4 *>(op + kOffP0 + (l0a + 8) * 64 + (((cq + ((l0a + 8) >> 1)) & 3) << 4));
          const float w1[4] = {p1.x, p1.y, p1.z, p1.w}, wa[4] = {pa.x, pa.y, pa.z, pa.w}, wb[4] = {pb.x, pb.y, pb.z, pb.w};
          if (cq == 0) tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int n = 4 * cq + i;
            const float ka = wa[i] * w1[i], kb = wb[i] * w1[i];
            acc[n] = fmaf(kb, v[1][n], ka * v[0][n]);
            acc[16 + n] = fmaf(kb, v[3][n], ka * v[2][n]);
          }
        }
        // T was read, the operand stage was read
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&S.acc_empty[s]); mbar_arrive(&S.op_empty[st]); }
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
        S.red[s][q][lane] = acc[0];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (q == (int) (jg & 3)) {
          const float sum = (S.red[s][0][lane] + S.red[s][1][lane]) + (S.red[s][2][lane] + S.red[s][3][lane]);
          const int n = lane & 15, c = lane >> 4;
          if (n < t5_nb(e)) f[2 * (size_t) perm[e.x + n] + c] = sum;
        }
      }
    }
  } else if (warp < 8) {
    // ===== window refill: thread (q, lane) owns row r = 32 q + lane of the four row blocks = pencils r and 128 + r
    const int q = warp - 4;
    const uint32_t lane_base = (uint32_t) (q * 32) << 16;
    const int r = 32 * q + lane;
    long long jg = 0;
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
      int whi = 0, base1 = 0, base2 = 0;   // cells [.., whi) are loaded; bases of the two previous batches
      for (int j = 0; j < nbat; j++, jg++) {
        const int base = t5_base(table[chunk.y + j]);
        if (j == 0) {
          if (jg > 0) mbar_wait(&S.acc_full[(jg - 1) & 3], (int) (((jg - 1) >> 2) & 1));   // every earlier batch has committed
          whi = base;
        } else {
          if (jg >= 3) mbar_wait(&S.acc_full[(jg - 3) & 3], (int) (((jg - 3) >> 2) & 1));   // at most three batches ahead
          // cells below base - 8 are overwritten: every batch that still reads them must have committed
          if (base1 < base - 8) mbar_wait(&S.acc_full[(jg - 1) & 3], (int) (((jg - 1) >> 2) & 1));
          else if (j >= 2 && base2 < base - 8) mbar_wait(&S.acc_full[(jg - 2) & 3], (int) (((jg - 2) >> 2) & 1));
          if (whi < base) whi = base;
        }
        base2 = base1;
        base1 = base;
        bool stored = false;
        for (; whi < base + kSpan; whi += 8) {
          const int slot = whi & (kRing - 1);
#pragma unroll
          for (int h = 0; h < 2; h++) {
            float re[8], im[8];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              int z = whi + 2 * i;
              if (z >= n2) z %= n2;
              const float4 g = *reinterpret_cast<const float4 *>(G + rowoff[h] + z);   // cells z, z + 1 (z even, n2 even)
              re[2 * i] = g.x; im[2 * i] = g.y; re[2 * i + 1] = g.z; im[2 * i + 1] = g.w;
            }
            float hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { hi[i] = tf32_rna(re[i]); lo[i] = re[i] - hi[i]; }
            tmem_st8(tb + lane_base + kColA + 64 * h + slot, hi);
            tmem_st8(tb + lane_base + kColA + 64 * h + 32 + slot, lo);
#pragma unroll
            for (int i = 0; i < 8; i++) { hi[i] = tf32_rna(im[i]); lo[i] = im[i] - hi[i]; }
            tmem_st8(tb + lane_base + kColA + 64 * (2 + h) + slot, hi);
            tmem_st8(tb + lane_base + kColA + 64 * (2 + h) + 32 + slot, lo);
          }
          stored = true;
        }
        if (stored) tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.a_ready[jg & 3]);
      }
    }
  } else if (warp == 8) {
    // ===== MMA issue: the whole warp runs the loop (uniform operands), one elected lane issues
    constexpr uint32_t idesc = make_idesc(128, kN);
    long long jg = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      uint2 e_next = nbat > 0 ? table[chunk.y] : make_uint2(0, 0);
      for (int j = 0; j < nbat; j++, jg++) {
        const int st = (int) (jg % kOpStages), s = (int) (jg & 3);
        const int base = t5_base(e_next);
        if (j + 1 < nbat) e_next = table[chunk.y + j + 1];
        mbar_wait(&S.op_full[st], (int) ((jg / kOpStages) & 1));
        mbar_wait(&S.acc_empty[s], (int) (((jg >> 2) & 1) ^ 1));
        mbar_wait(&S.a_ready[jg & 3], (int) ((jg >> 2) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bsm = smem_u32(S.ops[st]);
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const uint32_t kb = (uint32_t) (base + 8 * i) & (kRing - 1);
            const uint64_t bh = make_desc(bsm + 2 * i * kBChunk, kBChunk, kBGroup);
            const uint64_t bl = make_desc(bsm + kOffLo + 2 * i * kBChunk, kBChunk, kBGroup);
#pragma unroll
            for (int term = 0; term < 3; term++)
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const uint32_t a_hi = tb + kColA + 64 * b + kb;
                mma_ts(tb + 64 * s + 16 * b, term == 0 ? a_hi + 32 : a_hi, term == 1 ? bl : bh, idesc, (i | term) ? 1u : 0u);
              }
          }
          mma_commit(&S.acc_full[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== feeder: one lane streams the operand images into the ring
    long long jg = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const uint4 chunk = chunks[ch];
      const int nbat = (int) (chunk.z - chunk.y);
      for (int j = 0; j < nbat; j++, jg++) {
        const int st = (int) (jg % kOpStages);
        mbar_wait(&S.op_empty[st], (int) (((jg / kOpStages) & 1) ^ 1));
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

}  // namespace

// NFFTCU_OPT_TC5: 0 auto | 1 off | 2 on
bool tc5_selected(const nfftcu_ctx *c) {
  if (c->prec != NFFTCU_FLOAT || c->opt_tc5 == 1 || !mma3d_supported(c)) return false;
  if (!(c->opt_b_kernel == 0)) return false;   // an explicit kernel choice (generic / pencils / DMMA) wins
  return c->opt_tc5 == 2 || c->opt_tc5 == 0;
}

int tc5_build(nfftcu_ctx *c, const MmaParams &P) {
  c->tc5_ready = false;
  const long long units = (long long) P.NT0 * P.NT1 * P.zseg;
  const int kb = 256;
  if (c->tc5_units != units) {
    if (c->tc5_batch_start) pool_free(c->tc5_batch_start);
    if (c->tc5_counts) pool_free(c->tc5_counts);
    if (c->tc5_chunk_start) pool_free(c->tc5_chunk_start);
    c->tc5_batch_start = c->tc5_counts = c->tc5_chunk_start = nullptr;
    NFFTCU_CUDA(pool_malloc((void **) &c->tc5_batch_start, sizeof(uint32_t) * (size_t) (units + 1)));
    NFFTCU_CUDA(pool_malloc((void **) &c->tc5_counts, sizeof(uint32_t) * (size_t) units));
    NFFTCU_CUDA(pool_malloc((void **) &c->tc5_chunk_start, sizeof(uint32_t) * (size_t) (units + 1)));
    c->tc5_units = units;
  }
  const unsigned wgrid = (unsigned) ((units * 32 + kb - 1) / kb);
  t5_batches_kernel<false><<<wgrid, kb, 0, c->stream>>>((const uint64_t *) c->tile_keys, c->bin_start, c->tc5_counts, nullptr,
                                                        nullptr, units, P);
  t5_scan_kernel<<<1, 1024, 0, c->stream>>>(c->tc5_counts, c->tc5_batch_start, units);
  uint32_t total = 0;
  NFFTCU_CUDA(cudaMemcpyAsync(&total, c->tc5_batch_start + units, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  if ((long long) total > c->tc5_batch_cap) {
    if (c->tc5_batches) pool_free(c->tc5_batches);
    c->tc5_batches = nullptr;
    c->tc5_batch_cap = (long long) total + total / 8 + 1024;
    NFFTCU_CUDA(pool_malloc(&c->tc5_batches, sizeof(uint2) * (size_t) c->tc5_batch_cap));
  }
  t5_batches_kernel<true><<<wgrid, kb, 0, c->stream>>>((const uint64_t *) c->tile_keys, c->bin_start, nullptr, c->tc5_batch_start,
                                                       (uint2 *) c->tc5_batches, units, P);
  const unsigned ugrid = (unsigned) ((units + kb - 1) / kb);
  t5_chunk_count_kernel<<<ugrid, kb, 0, c->stream>>>(c->tc5_batch_start, c->tc5_counts, units);
  t5_scan_kernel<<<1, 1024, 0, c->stream>>>(c->tc5_counts, c->tc5_chunk_start, units);
  uint32_t nchunks = 0;
  NFFTCU_CUDA(cudaMemcpyAsync(&nchunks, c->tc5_chunk_start + units, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  NFFTCU_CUDA(cudaStreamSynchronize(c->stream));
  if ((long long) nchunks > c->tc5_chunk_cap) {
    if (c->tc5_chunks) pool_free(c->tc5_chunks);
    c->tc5_chunks = nullptr;
    c->tc5_chunk_cap = (long long) nchunks + nchunks / 8 + 1024;
    NFFTCU_CUDA(pool_malloc(&c->tc5_chunks, sizeof(uint4) * (size_t) c->tc5_chunk_cap));
  }
  t5_chunk_fill_kernel<<<ugrid, kb, 0, c->stream>>>(c->tc5_batch_start, c->tc5_chunk_start, (uint4 *) c->tc5_chunks, units, P.zseg);
  c->tc5_nchunks = nchunks;
  c->tc5_nbatches = total;
  c->launches += 6;
  NFFTCU_CUDA(cudaGetLastError());
  if (total == 0) { c->tc5_ready = true; return NFFTCU_OK; }
  // operand images: 5 KB per batch, always resident (the kernels have no evaluating fallback; a plan whose images do not
  // fit keeps the mma.sync kernels)
  const size_t need = (size_t) kImgBytes * (size_t) total;
  if (!c->tc5_images || c->tc5_images_bytes < need) {
    if (c->tc5_images) pool_free(c->tc5_images);
    c->tc5_images = nullptr;
    c->tc5_images_bytes = 0;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (need + need / 16 > free_b / 2 || pool_malloc(&c->tc5_images, need + need / 16) != cudaSuccess) {
      cudaGetLastError();
      return NFFTCU_OK;   // tc5_ready stays false: the caller falls back to the mma.sync path
    }
    c->tc5_images_bytes = need + need / 16;
  }
  NFFTCU_CUDA(cudaMemsetAsync(c->tc5_images, 0, need, c->stream));
  WinParams Wp;
  for (int t = 0; t < 3; t++) { Wp.b[t] = c->b[t]; Wp.ws[t] = c->wscale[t]; }
  Wp.m2 = (double) c->m * (double) c->m;
  Wp.window = c->window;
  t5_images_kernel<<<(unsigned) nchunks, 128, 0, c->stream>>>((const float *) c->tile_x, (const uint4 *) c->tc5_chunks,
                                                              (const uint2 *) c->tc5_batches, (unsigned char *) c->tc5_images, P, Wp);
  c->launches += 2;
  NFFTCU_CUDA(cudaGetLastError());
  c->tc5_ready = true;
  return NFFTCU_OK;
}

int tc5_interp(nfftcu_ctx *c, void *f_dev) {
  if (c->tc5_nchunks == 0) return NFFTCU_OK;
  const MmaParams P = mma3d_params(c);
  unsigned grid = (unsigned) c->sm_count;
  if ((long long) grid > c->tc5_nchunks) grid = (unsigned) c->tc5_nchunks;
  // the CTA owns all 512 TMEM columns of its SM: the (unused) dynamic shared memory keeps a second CTA off the SM, which
  // would otherwise sit in tcgen05.alloc until the first one exits
  NFFTCU_CUDA(cudaFuncSetAttribute(tc5_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOneCtaSmem));
  if (c->opt_timing) cudaEventRecord(c->evk[0], c->stream);
  tc5_interp_kernel<<<grid, 320, kOneCtaSmem, c->stream>>>((const float2 *) c->grid, c->tile_perm, (float *) f_dev,
                                                 (const uint4 *) c->tc5_chunks, (int) c->tc5_nchunks,
                                                 (const uint2 *) c->tc5_batches, (const unsigned char *) c->tc5_images, P);
  if (c->opt_timing) { cudaEventRecord(c->evk[1], c->stream); c->evk_recorded = true; }
  c->launches++;
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}

}  // namespace nfftcu
