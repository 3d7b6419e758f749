// tools/microbench3.cu -- dependent-chain latencies on B200: DFMA, FFMA, LDS->DFMA, shared atomics
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void dfma_chain(double *out, long long *cyc, int n, int ilp) {
  double a[8];
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 1e-9 + k;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  if (ilp == 1) for (int i = 0; i < n; i++) a[0] = fma(a[0], b, c);
  else if (ilp == 2) for (int i = 0; i < n; i++) { a[0] = fma(a[0], b, c); a[1] = fma(a[1], b, c); }
  else if (ilp == 4) for (int i = 0; i < n; i++) { a[0] = fma(a[0], b, c); a[1] = fma(a[1], b, c); a[2] = fma(a[2], b, c); a[3] = fma(a[3], b, c); }
  else for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < 8; k++) s += a[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ffma_chain(float *out, long long *cyc, int n) {
  float a = threadIdx.x * 1e-9f; const float b = 1.0000001f, c = 1e-9f;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = fmaf(a, b, c);
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_dfma_chain(double *out, long long *cyc, int n) {
  __shared__ double sm[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = (i + 1) % 256;
  __syncthreads();
  double a = 0; int idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { double v = sm[idx & 255]; a = fma(a, 1.0000001, v); idx = (int) v; }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
  const int n = 10000;
  for (int ilp : {1, 2, 4, 8}) {
    for (int warps : {1, 4}) {
      dfma_chain<<<1, 32 * warps>>>(out, cyc, n, ilp); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("DFMA ilp=%d warps=%d: %.2f cycles per loop iteration (%.2f per DFMA)\n", ilp, warps, (double) h / n, (double) h / n / ilp);
    }
  }
  ffma_chain<<<1, 32>>>((float *) out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("FFMA dependent chain: %.2f cycles\n", (double) h / n);
  lds_dfma_chain<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS->DFMA->index dependent chain: %.2f cycles per iteration\n", (double) h / n);
  return 0;
}
