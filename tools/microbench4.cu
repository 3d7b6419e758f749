// tools/microbench4.cu -- FP64 tensor-core (DMMA, mma.sync f64) rate and latency on B200, alone and
// mixed with DFMA.  Decides whether the B / B^T window contraction should be issued as DMMA.
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double *c, const double *a, const double *b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ILP independent accumulator tiles per warp, n iterations
template <int ILP>
__global__ void k884(double *out, long long *cyc, int n) {
  double c[ILP][2];
#pragma unroll
  for (int k = 0; k < ILP; k++) { c[k][0] = threadIdx.x; c[k][1] = k; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) dmma884(c[k][0], c[k][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void k1688(double *out, long long *cyc, int n) {
  double c[ILP][4];
#pragma unroll
  for (int k = 0; k < ILP; k++) for (int q = 0; q < 4; q++) c[k][q] = threadIdx.x + q;
  double a[4], b[2];
  for (int q = 0; q < 4; q++) a[q] = 1.0 + threadIdx.x * 1e-9 * q;
  for (int q = 0; q < 2; q++) b[q] = 1e-9 * threadIdx.x + q;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) dmma1688(c[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) for (int q = 0; q < 4; q++) s += c[k][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void k16816(double *out, long long *cyc, int n) {
  double c[ILP][4];
#pragma unroll
  for (int k = 0; k < ILP; k++) for (int q = 0; q < 4; q++) c[k][q] = threadIdx.x + q;
  double a[8], b[4];
  for (int q = 0; q < 8; q++) a[q] = 1.0 + threadIdx.x * 1e-9 * q;
  for (int q = 0; q < 4; q++) b[q] = 1e-9 * threadIdx.x + q;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) dmma16816(c[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) for (int q = 0; q < 4; q++) s += c[k][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// mixed: per iteration ILP DMMA(884) + NF DFMA
template <int ILP, int NF>
__global__ void kmix(double *out, long long *cyc, int n) {
  double c[ILP > 0 ? ILP : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
  for (int k = 0; k < ILP; k++) { c[k][0] = threadIdx.x; c[k][1] = k; }
#pragma unroll
  for (int k = 0; k < NF; k++) f[k] = k + threadIdx.x;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) dmma884(c[k][0], c[k][1], a, b);
#pragma unroll
    for (int k = 0; k < NF; k++) f[k] = fma(f[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += c[k][0] + c[k][1];
#pragma unroll
  for (int k = 0; k < NF; k++) s += f[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double *out; long long *cyc, h;
  cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8);
  const int n = 20000;
  printf("device %s, %d SMs\n", p.name, sms);
  // latency: 1 warp, 1 chain
  k884<1><<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA m8n8k4 dependent chain: %.2f cycles\n", (double) h / n);
  k1688<1><<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA m16n8k8 dependent chain: %.2f cycles\n", (double) h / n);
  k16816<1><<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA m16n8k16 dependent chain: %.2f cycles\n", (double) h / n);
  for (int warps : {1, 2, 4}) {
    k884<8><<<1, 32 * warps>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA m8n8k4 ilp=8 warps=%d on one SM: %.2f cycles per DMMA per warp\n", warps, (double) h / n / 8);
  }
  // throughput, whole chip
  for (int wps : {4, 8, 16, 32}) {
    float ms = timeit([&] { k884<8><<<sms, 32 * wps>>>(out, cyc, n); });
    double fma = (double) sms * wps * n * 8 * 256;
    printf("m8n8k4   ilp8 %2d warps/SM: %.1f GFMA/s = %.2f TFLOP/s (%.1f FMA/clk/SM @1.965)\n", wps, fma / ms * 1e-6, 2 * fma / ms * 1e-9, fma / ms * 1e-6 / sms / 1.965);
    ms = timeit([&] { k884<2><<<sms, 32 * wps>>>(out, cyc, n); });
    fma = (double) sms * wps * n * 2 * 256;
    printf("m8n8k4   ilp2 %2d warps/SM: %.2f TFLOP/s\n", wps, 2 * fma / ms * 1e-9);
    ms = timeit([&] { k1688<4><<<sms, 32 * wps>>>(out, cyc, n); });
    fma = (double) sms * wps * n * 4 * 1024;
    printf("m16n8k8  ilp4 %2d warps/SM: %.2f TFLOP/s\n", wps, 2 * fma / ms * 1e-9);
    ms = timeit([&] { k16816<4><<<sms, 32 * wps>>>(out, cyc, n); });
    fma = (double) sms * wps * n * 4 * 2048;
    printf("m16n8k16 ilp4 %2d warps/SM: %.2f TFLOP/s\n", wps, 2 * fma / ms * 1e-9);
  }
  // mixed DMMA + DFMA, 8 warps/SM
  {
    const int wps = 8;
    float m0 = timeit([&] { kmix<8, 0><<<sms, 32 * wps>>>(out, cyc, n); });
    float m8 = timeit([&] { kmix<8, 8><<<sms, 32 * wps>>>(out, cyc, n); });
    float m16 = timeit([&] { kmix<8, 16><<<sms, 32 * wps>>>(out, cyc, n); });
    float m32 = timeit([&] { kmix<8, 32><<<sms, 32 * wps>>>(out, cyc, n); });
    float f32 = timeit([&] { kmix<0, 32><<<sms, 32 * wps>>>(out, cyc, n); });
    printf("mixed per-iteration (8 DMMA884 + k DFMA), 8 warps/SM: k=0 %.3f ms, k=8 %.3f, k=16 %.3f, k=32 %.3f; 32 DFMA alone %.3f ms\n",
           m0, m8, m16, m32, f32);
  }
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
