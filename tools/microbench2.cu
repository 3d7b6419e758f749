// tools/microbench2.cu -- how many warps per SM sub-partition does the FP64 pipe need?
// DFMA throughput with 32 independent accumulators per thread vs resident warps, and the same
// with 15 broadcast LDS.64 per 60 DFMA (the tile3d node-loop shape).
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void dfma_ilp(double *out, int iters) {
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; k++) a[k] = threadIdx.x * 1e-9 + k;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 32; k++) a[k] = fma(a[k], b, c);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 32; k++) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void node_loop_shape(double *out, int iters) {
  __shared__ double pad[8][48];
  for (int i = threadIdx.x; i < 8 * 48; i += blockDim.x) (&pad[0][0])[i] = 1e-3 * i;
  __syncthreads();
  double ar[2][15], ai[2][15];
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int k = 0; k < 15; k++) { ar[j][k] = 0; ai[j][k] = 0; }
  const int l0 = threadIdx.x / 16 % 16, l1 = threadIdx.x % 16;
  for (int it = 0; it < iters; it++) {
    const double *pd = pad[it & 7];
    const double w0 = pd[l0] * pd[16 + l1], w1 = pd[(l0 + 8) & 15] * pd[16 + l1];
    const double fr = pd[3], fi = pd[5];
    const double a0 = w0 * fr, b0 = w0 * fi, a1 = w1 * fr, b1 = w1 * fi;
#pragma unroll
    for (int k = 0; k < 15; k++) {
      const double p = pd[32 + k];
      ar[0][k] += a0 * p; ai[0][k] += b0 * p; ar[1][k] += a1 * p; ai[1][k] += b1 * p;
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int k = 0; k < 15; k++) s += ar[j][k] + ai[j][k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 2048));
  for (int threads = 128; threads <= 1024; threads *= 2) {
    const int iters = 4000;
    float ms = time_ms([&] { dfma_ilp<<<sms, threads>>>(out, iters); });
    double fmas = (double) sms * threads * iters * 32;
    printf("dfma_ilp   %4d thr/SM (%d warps/SMSP): %.1f FMA/clk/SM @1.965GHz\n", threads, threads / 128, fmas / ms / 1e6 / sms / 1.965);
  }
  for (int threads = 128; threads <= 512; threads *= 2) {
    const int iters = 20000;
    float ms = time_ms([&] { node_loop_shape<<<sms, threads>>>(out, iters); });
    double fmas = (double) sms * threads * iters * 60;
    printf("node_shape %4d thr/SM (%d warps/SMSP): %.1f FMA/clk/SM, %.0f clk per node-iteration\n", threads, threads / 128, fmas / ms / 1e6 / sms / 1.965, ms * 1e-3 * 1.965e9 / iters);
  }
  for (int blocks = 1; blocks <= 2; blocks++) {
    const int iters = 20000;
    float ms = time_ms([&] { node_loop_shape<<<sms * blocks, 128>>>(out, iters); });
    printf("node_shape 128 thr x %d CTA/SM: %.0f clk per node-iteration per CTA\n", blocks, ms * 1e-3 * 1.965e9 / iters);
  }
  return 0;
}
