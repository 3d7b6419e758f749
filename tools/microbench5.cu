// tools/microbench5.cu -- legacy tensor path for fp32 data on B200: mma.sync m16n8k8 TF32 (fp32 accumulate) rate and
// latency, and the accuracy of the 3xTF32 split (a_hi b_hi + a_hi b_lo + a_lo b_hi) against an fp64 dot product.
// Decides whether the fp32 (nfftf_) window contraction should leave the FP64 DMMA kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void mma_tf32(float *c, const unsigned *a, const unsigned *b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ unsigned to_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

template <int ILP>
__global__ void ktf32(float *out, long long *cyc, int n) {
  float c[ILP][4];
#pragma unroll
  for (int k = 0; k < ILP; k++) for (int q = 0; q < 4; q++) c[k][q] = threadIdx.x + q;
  unsigned a[4], b[2];
  for (int q = 0; q < 4; q++) a[q] = to_tf32(1.0f + threadIdx.x * 1e-3f * q);
  for (int q = 0; q < 2; q++) b[q] = to_tf32(1e-3f * threadIdx.x + q);
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) mma_tf32(c[k], a, b);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) for (int q = 0; q < 4; q++) s += c[k][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// accuracy: one warp, C[16x8] = A[16xK] B[Kx8], K = 16 (two k8 steps), values with a wide dynamic range like the
// Kaiser-Bessel window (1 ... 1e12); 3xTF32 vs plain TF32 vs fp32 FMA vs fp64
__global__ void kacc(const float *A, const float *B, float *C3, float *C1) {   // A row-major 16x16, B col-major (k,n) 16x8
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  float c3[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  for (int ks = 0; ks < 2; ks++) {
    float av[4] = {A[g * 16 + 8 * ks + t], A[(g + 8) * 16 + 8 * ks + t], A[g * 16 + 8 * ks + t + 4], A[(g + 8) * 16 + 8 * ks + t + 4]};
    float bv[2] = {B[g * 16 + 8 * ks + t], B[g * 16 + 8 * ks + t + 4]};
    unsigned ah[4], al[4], bh[2], bl[2];
    for (int q = 0; q < 4; q++) { ah[q] = to_tf32(av[q]); al[q] = to_tf32(av[q] - __uint_as_float(ah[q])); }
    for (int q = 0; q < 2; q++) { bh[q] = to_tf32(bv[q]); bl[q] = to_tf32(bv[q] - __uint_as_float(bh[q])); }
    mma_tf32(c3, al, bh);
    mma_tf32(c3, ah, bl);
    mma_tf32(c3, ah, bh);
    mma_tf32(c1, ah, bh);
  }
  C3[g * 8 + 2 * t] = c3[0]; C3[g * 8 + 2 * t + 1] = c3[1]; C3[(g + 8) * 8 + 2 * t] = c3[2]; C3[(g + 8) * 8 + 2 * t + 1] = c3[3];
  C1[g * 8 + 2 * t] = c1[0]; C1[g * 8 + 2 * t + 1] = c1[1]; C1[(g + 8) * 8 + 2 * t] = c1[2]; C1[(g + 8) * 8 + 2 * t + 1] = c1[3];
}

template <int ILP>
void rate(int warps, int sms) {
  float *out; long long *cyc;
  cudaMalloc(&out, sizeof(float) * 1024 * 1024); cudaMalloc(&cyc, 8);
  const int n = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  ktf32<ILP><<<sms, warps * 32>>>(out, cyc, 100);
  cudaEventRecord(e0);
  ktf32<ILP><<<sms, warps * 32>>>(out, cyc, n);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double fma = (double) sms * warps * ILP * n * 16 * 8 * 8;
  printf("tf32 m16n8k8 ilp%d %2d warps/SM: %.1f TFLOP/s (%.0f FMA/clk/SM @1.965), %.2f cycles per MMA per warp\n", ILP, warps,
         2 * fma / (ms * 1e-3) * 1e-12, fma / (ms * 1e-3) / sms / 1.965e9, (double) c / ((double) n * ILP));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  rate<1>(1, 1);
  rate<4>(1, 1);
  rate<8>(1, 1);
  for (int w : {4, 8, 16, 32}) { rate<4>(w, sms); rate<8>(w, sms); }
  // accuracy
  float hA[256], hB[128]; double ref[128];
  srand(1);
  double e3 = 0, e1 = 0, ef = 0, nrm = 0;
  for (int rep = 0; rep < 50; rep++) {
    for (int i = 0; i < 256; i++) hA[i] = (float) ((rand() / (double) RAND_MAX - 0.5) * 1e-33);
    for (int n = 0; n < 8; n++) for (int k = 0; k < 16; k++) hB[n * 16 + k] = (float) exp(28.0 * (1.0 - pow((k - 7.3 - 0.05 * n) / 8.0, 2)));
    float *dA, *dB, *d3, *d1; cudaMalloc(&dA, 1024); cudaMalloc(&dB, 512); cudaMalloc(&d3, 512); cudaMalloc(&d1, 512);
    cudaMemcpy(dA, hA, 1024, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 512, cudaMemcpyHostToDevice);
    kacc<<<1, 32>>>(dA, dB, d3, d1);
    float h3[128], h1[128]; cudaMemcpy(h3, d3, 512, cudaMemcpyDeviceToHost); cudaMemcpy(h1, d1, 512, cudaMemcpyDeviceToHost);
    for (int r = 0; r < 16; r++) for (int n = 0; n < 8; n++) {
      double s = 0; float sf = 0;
      for (int k = 0; k < 16; k++) { s += (double) hA[r * 16 + k] * (double) hB[n * 16 + k]; sf = fmaf(hA[r * 16 + k], hB[n * 16 + k], sf); }
      ref[r * 8 + n] = s;
      e3 += (h3[r * 8 + n] - s) * (h3[r * 8 + n] - s); e1 += (h1[r * 8 + n] - s) * (h1[r * 8 + n] - s); ef += (sf - s) * (sf - s); nrm += s * s;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(d3); cudaFree(d1);
  }
  printf("rel l2 error of a 16-term window contraction vs fp64: 3xTF32 %.3e, 1xTF32 %.3e, fp32 FMA %.3e\n", sqrt(e3 / nrm), sqrt(e1 / nrm), sqrt(ef / nrm));
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
