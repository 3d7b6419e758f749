// tools/microbench.cu -- B200 rates that decide the B / B^T kernel design (not part of the product):
// FP64 FMA issue rate, shared-memory LDS.128 / broadcast LDS.64 bandwidth, global RED.ADD.F64
// throughput (coalesced rows vs scattered), shared-memory atomicAdd(double).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void dfma_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void ffma_kernel(float *out, int iters) {
  float a0 = threadIdx.x * 1e-9f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float b = 1.0000001f, c = 1e-9f;
  for (int i = 0; i < iters; i++) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// each lane reads its own 16 B (conflict-free LDS.128), 8 independent loads per iteration
__global__ void lds128_kernel(double *out, int iters) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  double2 acc = make_double2(0, 0);
  int idx = threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double2 v = sm[(idx + k * 256) & 4095];
      acc.x += v.x; acc.y += v.y;
    }
    idx = (idx + 32) & 4095;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

// pure load-issue test: no FP64 adds on the critical path (xor of the bits)
__global__ void lds128_nofp_kernel(unsigned long long *out, int iters) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  unsigned long long acc = 0;
  int idx = threadIdx.x;
  const ulonglong2 *s = (const ulonglong2 *) sm;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const ulonglong2 v = s[(idx + k * 256) & 4095];
      acc ^= v.x ^ v.y;
    }
    idx = (idx + 32) & 4095;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// all lanes of a warp read the same 8 B / 16 B (broadcast)
__global__ void lds_bcast_kernel(unsigned long long *out, int iters, int wide) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  unsigned long long acc = 0;
  int idx = threadIdx.x >> 5;
  const ulonglong2 *s = (const ulonglong2 *) sm;
  const unsigned long long *s1 = (const unsigned long long *) sm;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (wide) { const ulonglong2 v = s[(idx + k * 64) & 4095]; acc ^= v.x ^ v.y; }
      else { acc ^= s1[(idx + k * 64) & 8191]; }
    }
    idx = (idx + 8) & 4095;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// global reductions: warp w of the grid adds to a row of `row` doubles starting at a pseudo-random
// row of a big array (coalesced within the warp), or every lane to its own random address.
__global__ void red_kernel(double *g, long long nrows, int rowlen, int iters, int scattered) {
  const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long s = warp * 0x9E3779B97F4A7C15ull + 12345;
  for (int i = 0; i < iters; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    long long r = (long long) ((s >> 20) % (unsigned long long) nrows);
    if (scattered) {
      unsigned long long s2 = s ^ (lane * 0xD6E8FEB86659FD93ull);
      s2 = s2 * 6364136223846793005ull + 1442695040888963407ull;
      r = (long long) ((s2 >> 20) % (unsigned long long) nrows);
      atomicAdd(&g[r * rowlen + (lane % rowlen)], 1.0);
    } else if (lane < rowlen) {
      atomicAdd(&g[r * rowlen + lane], 1.0);
    }
  }
}

// same but sorted-like locality: consecutive warps hit consecutive rows (+ small jitter)
__global__ void red_local_kernel(double *g, long long nrows, int rowlen, int iters) {
  const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < iters; i++) {
    const long long r = (warp * 3 + i * 17) % nrows;
    if (lane < rowlen) atomicAdd(&g[r * rowlen + lane], 1.0);
  }
}

__global__ void smem_atomic_kernel(double *out, int iters) {
  extern __shared__ double smd[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) smd[i] = 0;
  __syncthreads();
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x;
  for (int i = 0; i < iters; i++) {
    s = s * 1664525u + 1013904223u;
    atomicAdd(&smd[(s >> 8) & 8191], 1.0);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = smd[0];
}

__global__ void smem_atomic_f32_kernel(float *out, int iters) {
  extern __shared__ float smf[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) smf[i] = 0;
  __syncthreads();
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x;
  for (int i = 0; i < iters; i++) {
    s = s * 1664525u + 1013904223u;
    atomicAdd(&smf[(s >> 8) & 16383], 1.0f);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = smf[0];
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  {
    const int iters = 20000, blocks = sms * 4, threads = 256;
    float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
    double fmas = (double) blocks * threads * iters * 8;
    printf("DFMA: %.1f GFMA/s = %.2f TFLOP/s fp64 (%.1f FMA/clk/SM at 1.9 GHz)\n", fmas / ms / 1e6, 2 * fmas / ms / 1e9, fmas / ms / 1e6 / sms / 1.9);
    ms = time_ms([&] { ffma_kernel<<<blocks, threads>>>((float *) out, iters); });
    printf("FFMA: %.1f GFMA/s = %.2f TFLOP/s fp32\n", fmas / ms / 1e6, 2 * fmas / ms / 1e9);
  }
  {
    const int iters = 4000, blocks = sms * 2, threads = 256;
    const size_t sm = 4096 * 16;
    CK(cudaFuncSetAttribute(lds128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
    CK(cudaFuncSetAttribute(lds128_nofp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
    CK(cudaFuncSetAttribute(lds_bcast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
    float ms = time_ms([&] { lds128_kernel<<<blocks, threads, sm>>>(out, iters); });
    double bytes = (double) blocks * threads * iters * 8 * 16;
    printf("LDS.128 + 2 DADD: %.1f TB/s aggregate = %.1f B/clk/SM at 1.9 GHz\n", bytes / ms / 1e9, bytes / ms / 1e6 / sms / 1.9);
    ms = time_ms([&] { lds128_nofp_kernel<<<blocks, threads, sm>>>((unsigned long long *) out, iters); });
    printf("LDS.128 (no fp): %.1f TB/s aggregate = %.1f B/clk/SM at 1.9 GHz\n", bytes / ms / 1e9, bytes / ms / 1e6 / sms / 1.9);
    ms = time_ms([&] { lds_bcast_kernel<<<blocks, threads, sm>>>((unsigned long long *) out, iters, 0); });
    double insts = (double) blocks * threads / 32 * iters * 8;
    printf("LDS.64 broadcast: %.1f G warp-instr/s = %.2f warp-instr/clk/SM\n", insts / ms / 1e6, insts / ms / 1e6 / sms / 1.9);
    ms = time_ms([&] { lds_bcast_kernel<<<blocks, threads, sm>>>((unsigned long long *) out, iters, 1); });
    printf("LDS.128 broadcast: %.1f G warp-instr/s = %.2f warp-instr/clk/SM\n", insts / ms / 1e6, insts / ms / 1e6 / sms / 1.9);
  }
  {
    const long long nrows = 1 << 20;   // x 16 doubles = 128 MB
    const int rowlen = 16;
    double *g; CK(cudaMalloc(&g, sizeof(double) * nrows * rowlen));
    CK(cudaMemset(g, 0, sizeof(double) * nrows * rowlen));
    const int iters = 2000, blocks = sms * 8, threads = 256;
    const double warps = (double) blocks * threads / 32;
    float ms = time_ms([&] { red_kernel<<<blocks, threads>>>(g, nrows, rowlen, iters, 0); });
    printf("RED.F64 coalesced rows of 16 (random rows): %.1f G doubles/s (%.2f G rows/s)\n", warps * iters * 16 / ms / 1e6, warps * iters / ms / 1e6);
    ms = time_ms([&] { red_kernel<<<blocks, threads>>>(g, nrows, 14, iters, 0); });
    printf("RED.F64 rows of 14 at stride 14 (unaligned, random rows): %.1f G doubles/s\n", warps * iters * 14 / ms / 1e6);
    ms = time_ms([&] { red_local_kernel<<<blocks, threads>>>(g, nrows, rowlen, iters); });
    printf("RED.F64 coalesced rows of 16 (neighbouring rows, L2-local): %.1f G doubles/s\n", warps * iters * 16 / ms / 1e6);
    ms = time_ms([&] { red_kernel<<<blocks, threads>>>(g, nrows, rowlen, iters / 4, 1); });
    printf("RED.F64 scattered lanes: %.1f G doubles/s\n", warps * (iters / 4) * 32 / ms / 1e6);
    cudaFree(g);
  }
  {
    const int iters = 20000, blocks = sms * 2, threads = 256;
    CK(cudaFuncSetAttribute(smem_atomic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(smem_atomic_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    float ms = time_ms([&] { smem_atomic_kernel<<<blocks, threads, 65536>>>(out, iters); });
    double ops = (double) blocks * threads * iters;
    printf("shared atomicAdd(double), random addr: %.1f G ops/s = %.2f ops/clk/SM\n", ops / ms / 1e6, ops / ms / 1e6 / sms / 1.9);
    ms = time_ms([&] { smem_atomic_f32_kernel<<<blocks, threads, 65536>>>((float *) out, iters); });
    printf("shared atomicAdd(float), random addr: %.1f G ops/s = %.2f ops/clk/SM\n", ops / ms / 1e6, ops / ms / 1e6 / sms / 1.9);
  }
  return 0;
}
