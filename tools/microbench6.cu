// tools/microbench6.cu -- tcgen05 (5th-generation tensor core) probe for the fp32 window contraction on B200:
//   1. correctness of tcgen05.mma.cta_group::1.kind::tf32 with the operand layouts tc5.cu uses
//      (K-major, no swizzle; A = a RING of 32 z-slots per row whose k-steps start at any 8-slot boundary, B = the
//      window values of 16 nodes), accumulators in TMEM, read back with tcgen05.ld.32x32b;
//   2. accuracy of the 3xTF32 split (a_hi b_hi + a_lo b_hi + a_hi b_lo, fp32 accumulate in TMEM) against fp64;
//   3. issue / execution rate of M = 128, N = 16 (and 32) MMAs from one thread, and the TMEM read rate of 4 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench6 tools/microbench6.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, int parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a wrong descriptor must not hang the box
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, int parity) {
  for (long long i = 0; i < 20000000ll; i++) if (mbar_try(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes (128 contiguous bytes); LBO = byte distance of the two core
  // matrices of one k-step along K, SBO = byte distance of 8-row groups along M / N; version 1 (Blackwell)
  return (uint64_t) ((saddr & 0x3ffff) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

constexpr int kASlots = 32, kARow8 = 1024, kAChunk = 128;   // A: (row/8)*1024 + (slot/4)*128 + (row%8)*16 + (slot%4)*4
constexpr int kBChunk = 256, kBGroup = 128;                  // B: (slot/4)*256 + (n/8)*128 + (n%8)*16 + (slot%4)*4

__device__ __forceinline__ int a_off(int row, int slot) { return (row >> 3) * kARow8 + (slot >> 2) * kAChunk + (row & 7) * 16 + (slot & 3) * 4; }
__device__ __forceinline__ int b_off(int n, int slot) { return (slot >> 2) * kBChunk + (n >> 3) * kBGroup + (n & 7) * 16 + (slot & 3) * 4; }

// mode 0: operands as given (one term); mode 1: 3xTF32 split.  A: [128][32] row-major floats, B: [16][24], q = ring
// rotation (the window starts at slot 8 q).  D out: [128][16]
__global__ void __launch_bounds__(128) k_correct(const float *A, const float *B, float *D, int q, int mode, int *status) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char *Ahi = sm, *Alo = sm + 16384, *Bhi = sm + 32768, *Blo = sm + 32768 + 2048;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * kASlots; i += 128) {
    const int row = i / kASlots, slot = i % kASlots;
    const float v = A[i];
    const float hi = mode ? tf32_rna(v) : v;
    *reinterpret_cast<float *>(Ahi + a_off(row, slot)) = hi;
    *reinterpret_cast<float *>(Alo + a_off(row, slot)) = v - hi;
  }
  for (int i = tid; i < 16 * 24; i += 128) {
    const int n = i / 24, slot = i % 24;
    const float v = B[i];
    const float hi = mode ? tf32_rna(v) : v;
    *reinterpret_cast<float *>(Bhi + b_off(n, slot)) = hi;
    *reinterpret_cast<float *>(Blo + b_off(n, slot)) = v - hi;
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 16);
    uint32_t acc = 0;
    for (int i = 0; i < 3; i++) {
      const int kb = (8 * q + 8 * i) & 31;
      const uint64_t ah = make_desc(smem_u32(Ahi) + (kb >> 2) * kAChunk, kAChunk, kARow8);
      const uint64_t al = make_desc(smem_u32(Alo) + (kb >> 2) * kAChunk, kAChunk, kARow8);
      const uint64_t bh = make_desc(smem_u32(Bhi) + 2 * i * kBChunk, kBChunk, kBGroup);
      const uint64_t bl = make_desc(smem_u32(Blo) + 2 * i * kBChunk, kBChunk, kBGroup);
      if (mode) {
        mma_tf32(tb, al, bh, idesc, acc); acc = 1;
        mma_tf32(tb, ah, bl, idesc, acc);
      }
      mma_tf32(tb, ah, bh, idesc, acc); acc = 1;
    }
    mma_commit(&bar);
  }
  const bool ok = mbar_wait_bounded(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) { if (tid == 0) *status = 1; }
  else {
    uint32_t v[16];
    tmem_ld16(tb + ((uint32_t) (warp * 32) << 16), v);
    tmem_ld_wait();
    for (int n = 0; n < 16; n++) D[tid * 16 + n] = __uint_as_float(v[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(32));
}

// rate: one thread issues `reps` groups of 12 MMAs (4 accumulator blocks x 3 k-steps, M = 128, N = NN), one commit per
// group; the other warps wait.  cyc[0] = cycles from the first issue to the completion of the last group, cyc[1] =
// cycles the issuing thread spent issuing.
template <int NN>
__global__ void __launch_bounds__(128) k_rate(int reps, long long *cyc, int *status) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 8192) / 4; i += 128) reinterpret_cast<float *>(sm)[i] = 1.0f + 1e-3f * (i & 63);
  if (tid == 0) { for (int i = 0; i < 4; i++) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  bool ok = true;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, NN);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
#pragma unroll
      for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const uint64_t a = make_desc(smem_u32(sm) + mb * 16384 + 2 * i * kAChunk, kAChunk, kARow8);
          const uint64_t b = make_desc(smem_u32(sm) + 65536 + 2 * i * kBChunk * (NN / 16), kBChunk * (NN / 16), kBGroup);
          mma_tf32(tb + mb * NN, a, b, idesc, (r | i) ? 1u : 0u);
        }
      if (r >= 4) ok = ok && mbar_wait_bounded(&bar[r & 3], ((r >> 2) - 1) & 1);   // at most 4 groups in flight
      mma_commit(&bar[r & 3]);
    }
    const long long t1 = clock64();
    for (int r = reps - 4; r < reps; r++) ok = ok && mbar_wait_bounded(&bar[r & 3], (r >> 2) & 1);
    const long long t2 = clock64();
    cyc[0] = t2 - t0;
    cyc[1] = t1 - t0;
    if (!ok) *status = 2;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(128));
}

// TMEM read rate: 4 warps, each reads its 32 lanes x 64 columns `reps` times (4 x 32x32b.x16)
__global__ void __launch_bounds__(128) k_ldtm(int reps, long long *cyc, float *sink) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base + ((uint32_t) (warp * 32) << 16);
  uint32_t z[16];
  for (int i = 0; i < 16; i++) z[i] = 0;
  for (int c = 0; c < 4; c++)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(tb + 16 * c), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]),
                    "r"(z[8]), "r"(z[9]), "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  float s = 0;
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    uint32_t v0[16], v1[16], v2[16], v3[16];
    tmem_ld16(tb, v0); tmem_ld16(tb + 16, v1); tmem_ld16(tb + 32, v2); tmem_ld16(tb + 48, v3);
    tmem_ld_wait();
    for (int i = 0; i < 16; i++) s += __uint_as_float(v0[i] ^ v1[i] ^ v2[i] ^ v3[i]);
  }
  const long long t1 = clock64();
  sink[tid] = s;
  if (tid == 0) cyc[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(64));
}


// ---- A operand from TMEM (TS form): the grid window lives in TMEM (lane = row, one column per z-slot), written by the
// four warps that own the lane quarters with tcgen05.st
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// kind::i8 (signed 8-bit operands, s32 accumulate, K = 32 per MMA = the same 32 bytes per row and k-step as kind::tf32):
// rate only -- the gate for an Ozaki-style fp64 emulation on the integer tensor path
__device__ __forceinline__ uint32_t make_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)) : "memory");
}

// TMEM columns: [0,16) D, [32,64) A_hi ring, [64,96) A_lo ring
__global__ void __launch_bounds__(128) k_correct_ts(const float *A, const float *B, float *D, int q, int mode, int *status) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char *Bhi = sm, *Blo = sm + 2048;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16 * 24; i += 128) {
    const int n = i / 24, slot = i % 24;
    const float v = B[i];
    const float hi = mode ? tf32_rna(v) : v;
    *reinterpret_cast<float *>(Bhi + b_off(n, slot)) = hi;
    *reinterpret_cast<float *>(Blo + b_off(n, slot)) = v - hi;
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  const uint32_t lane_base = (uint32_t) (warp * 32) << 16;
  for (int slot = 0; slot < 32; slot += 2) {   // thread = row tid
    const float v0 = A[tid * 32 + slot], v1 = A[tid * 32 + slot + 1];
    const float h0 = mode ? tf32_rna(v0) : v0, h1 = mode ? tf32_rna(v1) : v1;
    tmem_st2(tb + lane_base + 32 + slot, h0, h1);
    tmem_st2(tb + lane_base + 64 + slot, v0 - h0, v1 - h1);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 16);
    uint32_t acc = 0;
    for (int i = 0; i < 3; i++) {
      const int kb = (8 * q + 8 * i) & 31;
      const uint64_t bh = make_desc(smem_u32(Bhi) + 2 * i * kBChunk, kBChunk, kBGroup);
      const uint64_t bl = make_desc(smem_u32(Blo) + 2 * i * kBChunk, kBChunk, kBGroup);
      if (mode) {
        mma_tf32_ts(tb, tb + 64 + kb, bh, idesc, acc); acc = 1;
        mma_tf32_ts(tb, tb + 32 + kb, bl, idesc, acc);
      }
      mma_tf32_ts(tb, tb + 32 + kb, bh, idesc, acc); acc = 1;
    }
    mma_commit(&bar);
  }
  const bool ok = mbar_wait_bounded(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) { if (tid == 0) *status = 1; }
  else {
    uint32_t v[16];
    tmem_ld16(tb + lane_base, v);
    tmem_ld_wait();
    for (int n = 0; n < 16; n++) D[tid * 16 + n] = __uint_as_float(v[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(128));
}

// rate, A from TMEM: groups of 36 MMAs (4 row blocks x 3 k-steps x 3 terms) like one batch of tc5.cu; TMEM columns
// [0, 4 NN) D, [256, 512) A (4 row blocks x (hi, lo) x 32 slots)
template <int NN, int ORDER, int KIND = 0>
__global__ void __launch_bounds__(128) k_rate_ts(int reps, long long *cyc, int *status) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384 / 4; i += 128) reinterpret_cast<float *>(sm)[i] = 1.0f + 1e-3f * (i & 63);
  if (tid == 0) { for (int i = 0; i < 4; i++) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  const uint32_t lane_base = (uint32_t) (warp * 32) << 16;
  for (int c = 0; c < 256; c += 2) tmem_st2(tb + lane_base + 256 + c, 1.0f + 1e-3f * c, 0.5f);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  bool ok = true;
  if (warp == 0) {   // the whole warp runs the loop (uniform operands); one elected lane issues
    const uint32_t idesc = KIND ? make_idesc_i8(128, NN) : make_idesc(128, NN);
    const uint64_t b0 = make_desc(smem_u32(sm), kBChunk * (NN / 16), kBGroup);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      if (r >= 4) ok = ok && mbar_wait_bounded(&bar[r & 3], ((r >> 2) - 1) & 1);
      const uint32_t kb0 = (uint32_t) (8 * r) & 31;
      constexpr uint32_t kLo = (6 * kBChunk * (NN / 16)) >> 4;
      const bool leader = elect_one();
      if (leader) {
      if (ORDER == 0) {          // the three terms of one (k-step, row block) back to back: dependent accumulations
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const uint32_t kb = (kb0 + 8 * i) & 31;
          const uint64_t bh = b0 + (uint64_t) ((2 * i * kBChunk * (NN / 16)) >> 4);
          const uint64_t bl = bh + kLo;
#pragma unroll
          for (int mb = 0; mb < 4; mb++) {
            const uint32_t a_hi = tb + 256 + mb * 64 + kb, a_lo = a_hi + 32, d = tb + mb * NN;
            mma_tf32_ts(d, a_lo, bh, idesc, i ? 1u : 0u);
            mma_tf32_ts(d, a_hi, bl, idesc, 1u);
            mma_tf32_ts(d, a_hi, bh, idesc, 1u);
          }
        }
      } else {                   // row block innermost: consecutive MMAs go to different accumulators
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const uint32_t kb = (kb0 + 8 * i) & 31;
          const uint64_t bh = b0 + (uint64_t) ((2 * i * kBChunk * (NN / 16)) >> 4);
          const uint64_t bl = bh + kLo;
#pragma unroll
          for (int term = 0; term < 3; term++)
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
              const uint32_t a_hi = tb + 256 + mb * 64 + kb, a_lo = a_hi + 32, d = tb + mb * NN;
              if (KIND) mma_i8_ts(d, term == 0 ? a_lo : a_hi, term == 1 ? bl : bh, idesc, (i | term) ? 1u : 0u);
              else mma_tf32_ts(d, term == 0 ? a_lo : a_hi, term == 1 ? bl : bh, idesc, (i | term) ? 1u : 0u);
            }
        }
      }
      mma_commit(&bar[r & 3]);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    for (int r = reps - 4; r < reps; r++) ok = ok && mbar_wait_bounded(&bar[r & 3], (r >> 2) & 1);
    const long long t2 = clock64();
    if (tid == 0) { cyc[0] = t2 - t0; cyc[1] = t1 - t0; if (!ok) *status = 2; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}


// issue rate with NW issuing warps (each its own accumulators and commit barriers, the same A ring): does a second issuing
// thread raise the ~16-cycle cadence of one thread's tcgen05.mma stream?
template <int NN, int NW>
__global__ void __launch_bounds__(128) k_rate_multi(int reps, long long *cyc, int *status) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar[4][4];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384 / 4; i += 128) reinterpret_cast<float *>(sm)[i] = 1.0f + 1e-3f * (i & 63);
  if (tid == 0) { for (int w = 0; w < 4; w++) for (int i = 0; i < 4; i++) mbar_init(&bar[w][i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  const uint32_t lane_base = (uint32_t) (warp * 32) << 16;
  for (int c = 0; c < 256; c += 2) tmem_st2(tb + lane_base + 256 + c, 1.0f + 1e-3f * c, 0.5f);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  bool ok = true;
  const long long t0 = clock64();
  if (warp < NW) {
    const uint32_t idesc = make_idesc(128, NN);
    const uint64_t b0 = make_desc(smem_u32(sm), kBChunk * (NN / 16), kBGroup);
    constexpr uint32_t kLo = (6 * kBChunk * (NN / 16)) >> 4;
    for (int r = 0; r < reps; r++) {
      if (r >= 4) ok = ok && mbar_wait_bounded(&bar[warp][r & 3], ((r >> 2) - 1) & 1);
      const uint32_t kb0 = (uint32_t) (8 * r) & 31;
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const uint32_t kb = (kb0 + 8 * i) & 31;
          const uint64_t bh = b0 + (uint64_t) ((2 * i * kBChunk * (NN / 16)) >> 4);
          const uint64_t bl = bh + kLo;
#pragma unroll
          for (int term = 0; term < 3; term++)
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
              const uint32_t a_hi = tb + 256 + mb * 64 + kb, a_lo = a_hi + 32, d = tb + warp * 4 * NN + mb * NN;
              mma_tf32_ts(d, term == 0 ? a_lo : a_hi, term == 1 ? bl : bh, idesc, (i | term) ? 1u : 0u);
            }
        }
        mma_commit(&bar[warp][r & 3]);
      }
      __syncwarp();
    }
    for (int r = reps - 4; r < reps; r++) ok = ok && mbar_wait_bounded(&bar[warp][r & 3], (r >> 2) & 1);
  }
  __syncthreads();
  if (tid == 0) { cyc[0] = clock64() - t0; }
  if (!ok) *status = 3;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}

static float tf32_trunc_host(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  int *status; long long *cyc; float *sink;
  CK(cudaMallocManaged(&status, 4)); CK(cudaMallocManaged(&cyc, 16)); CK(cudaMalloc(&sink, 4096));
  *status = 0;
  std::vector<float> A(128 * 32), B(16 * 24), D(128 * 16);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  const size_t smem = 32768 + 4096;
  CK(cudaFuncSetAttribute(k_correct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  srand(7);
  for (int mode = 0; mode < 2; mode++)
    for (int q = 0; q < 4; q += (mode ? 3 : 1)) {
      // mode 0: tf32-exact inputs (exact products); mode 1: full fp32 inputs with a Kaiser-Bessel-like dynamic range in B
      for (auto &v : A) { v = (float) rand() / RAND_MAX - 0.5f; if (!mode) v = tf32_trunc_host(v); }
      for (size_t i = 0; i < B.size(); i++) {
        float v = (float) rand() / RAND_MAX;
        if (mode) v *= expf(-0.25f * (float) ((i % 24) - 11) * ((i % 24) - 11));
        else v = tf32_trunc_host(v);
        B[i] = v;
      }
      CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(dD, 0xff, D.size() * 4));
      k_correct<<<1, 128, smem>>>(dA, dB, dD, q, mode, status);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double num = 0, den = 0, worst = 0;
      for (int r = 0; r < 128; r++)
        for (int n = 0; n < 16; n++) {
          double ref = 0;
          for (int s = 0; s < 24; s++) ref += (double) A[r * 32 + ((8 * q + s) & 31)] * (double) B[n * 24 + s];
          const double e = (double) D[r * 16 + n] - ref;
          num += e * e; den += ref * ref;
          if (fabs(e) > worst) worst = fabs(e);
        }
      printf("correct mode=%d q=%d status=%d rel_l2=%.3e max_abs_err=%.3e (D[0]=%g D[17]=%g)\n", mode, q, *status,
             sqrt(num / den), worst, D[0], D[17]);
      if (*status) { printf("barrier timeout: MMA never completed\n"); return 1; }
    }
  {
    // single term on full-fp32 inputs: does the tensor core truncate or round its tf32 operands?
    for (auto &v : A) v = (float) rand() / RAND_MAX - 0.5f;
    for (auto &v : B) v = (float) rand() / RAND_MAX;
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    k_correct<<<1, 128, smem>>>(dA, dB, dD, 0, 0, status);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double nt = 0, nf = 0, den = 0;
    for (int r = 0; r < 128; r++)
      for (int n = 0; n < 16; n++) {
        double rt = 0, rf = 0;
        for (int s = 0; s < 24; s++) {
          rt += (double) tf32_trunc_host(A[r * 32 + s]) * (double) tf32_trunc_host(B[n * 24 + s]);
          rf += (double) A[r * 32 + s] * (double) B[n * 24 + s];
        }
        nt += (D[r * 16 + n] - rt) * (D[r * 16 + n] - rt); nf += (D[r * 16 + n] - rf) * (D[r * 16 + n] - rf); den += rf * rf;
      }
    printf("raw fp32 operands, one term: rel_l2 vs truncated-operand model %.3e, vs exact %.3e\n", sqrt(nt / den), sqrt(nf / den));
  }

  CK(cudaFuncSetAttribute(k_correct_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096));
  for (int mode = 0; mode < 2; mode++)
    for (int q = 0; q < 4; q += 3) {
      for (auto &v : A) { v = (float) rand() / RAND_MAX - 0.5f; if (!mode) v = tf32_trunc_host(v); }
      for (size_t i = 0; i < B.size(); i++) {
        float v = (float) rand() / RAND_MAX;
        if (mode) v *= expf(-0.25f * (float) ((i % 24) - 11) * ((i % 24) - 11));
        else v = tf32_trunc_host(v);
        B[i] = v;
      }
      CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(dD, 0xff, D.size() * 4));
      k_correct_ts<<<1, 128, 4096>>>(dA, dB, dD, q, mode, status);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double num = 0, den = 0;
      for (int r = 0; r < 128; r++)
        for (int n = 0; n < 16; n++) {
          double ref = 0;
          for (int s2 = 0; s2 < 24; s2++) ref += (double) A[r * 32 + ((8 * q + s2) & 31)] * (double) B[n * 24 + s2];
          const double e = (double) D[r * 16 + n] - ref;
          num += e * e; den += ref * ref;
        }
      printf("A-from-TMEM correct mode=%d q=%d status=%d rel_l2=%.3e (D[0]=%g D[17]=%g)\n", mode, q, *status, sqrt(num / den), D[0], D[17]);
      if (*status) { printf("barrier timeout (TS)\n"); return 1; }
    }
#define RATE_TS(NN, ORD) do { \
    CK(cudaFuncSetAttribute(k_rate_ts<NN, ORD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384)); \
    for (int pass = 0; pass < 2; pass++) { k_rate_ts<NN, ORD><<<1, 128, 16384>>>(2000, cyc, status); CK(cudaDeviceSynchronize()); } \
    printf("rate A-from-TMEM M=128 N=%d order=%d: %.1f cycles per MMA executed, %.1f issued (36 per batch), status %d\n", NN, ORD, \
           (double) cyc[0] / (36.0 * 2000), (double) cyc[1] / (36.0 * 2000), *status); } while (0)
  RATE_TS(16, 0); RATE_TS(16, 1); RATE_TS(32, 0); RATE_TS(32, 1); RATE_TS(64, 0); RATE_TS(64, 1);
#define RATE_I8(NN) do { \
    CK(cudaFuncSetAttribute(k_rate_ts<NN, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384)); \
    for (int pass = 0; pass < 2; pass++) { k_rate_ts<NN, 1, 1><<<1, 128, 16384>>>(2000, cyc, status); CK(cudaDeviceSynchronize()); } \
    printf("rate kind::i8 A-from-TMEM M=128 N=%d K=32: %.1f cycles per MMA, status %d\n", NN, (double) cyc[0] / (36.0 * 2000), *status); } while (0)
  RATE_I8(16); RATE_I8(32); RATE_I8(64);
#define RATE_MULTI(NN, NW) do { \
    CK(cudaFuncSetAttribute(k_rate_multi<NN, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384)); \
    for (int pass = 0; pass < 2; pass++) { k_rate_multi<NN, NW><<<1, 128, 16384>>>(2000, cyc, status); CK(cudaDeviceSynchronize()); } \
    printf("rate %d issuing warps, M=128 N=%d: %.1f cycles per MMA aggregate (%d MMAs), status %d\n", NW, NN, \
           (double) cyc[0] / (36.0 * 2000 * NW), 36 * 2000 * NW, *status); } while (0)
  RATE_MULTI(16, 1); RATE_MULTI(16, 2); RATE_MULTI(16, 4); RATE_MULTI(32, 2);
  const size_t smem_rate = 65536 + 8192;
  CK(cudaFuncSetAttribute(k_rate<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_rate));
  CK(cudaFuncSetAttribute(k_rate<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_rate));
  for (int pass = 0; pass < 2; pass++) {
    const int reps = 2000;
    k_rate<16><<<1, 128, smem_rate>>>(reps, cyc, status);
    CK(cudaDeviceSynchronize());
    if (pass) printf("rate N=16: %.1f cycles per MMA (M=128,K=8) executed, %.1f cycles per MMA issued, status %d\n",
                     (double) cyc[0] / (12.0 * reps), (double) cyc[1] / (12.0 * reps), *status);
    k_rate<32><<<1, 128, smem_rate>>>(reps, cyc, status);
    CK(cudaDeviceSynchronize());
    if (pass) printf("rate N=32: %.1f cycles per MMA executed, %.1f issued, status %d\n",
                     (double) cyc[0] / (12.0 * reps), (double) cyc[1] / (12.0 * reps), *status);
  }
  for (int pass = 0; pass < 2; pass++) {
    const int reps = 4000;
    k_ldtm<<<1, 128>>>(reps, cyc, sink);
    CK(cudaDeviceSynchronize());
    if (pass) printf("tcgen05.ld 4 warps x (32 lanes x 64 columns): %.1f cycles per round = %.1f B/clk/SM\n",
                     (double) cyc[0] / reps, 4.0 * 32 * 64 * 4 / ((double) cyc[0] / reps));
  }
  printf("done\n");
  return 0;
}
