// peaks.cu -- roofline denominators measured on the device the plan runs on (bench.py calls this in-process
// every run and records what it used): the FP64 tensor-core (DMMA m8n8k4) rate the fp64 B / B^T kernels are
// bound by, the legacy mma.sync TF32 rate of the fp32 kernels, and a plain device copy as an HBM cross-check
// of MEASURED_PEAKS.json.  Register-only kernels, independent accumulators (ILP 8), best of `reps` launches.
#include "common.cuh"

namespace nfftcu {
namespace {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma_tf32(float *c, const unsigned *a, const unsigned *b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void peak_dmma_kernel(double *out, int n) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; k++) { c[k][0] = threadIdx.x; c[k][1] = k; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) dmma884(c[k][0], c[k][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
  out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void peak_tf32_kernel(float *out, int n) {
  float c[8][4];
#pragma unroll
  for (int k = 0; k < 8; k++)
    for (int q = 0; q < 4; q++) c[k][q] = threadIdx.x + q;
  unsigned a[4], b[2];
  for (int q = 0; q < 4; q++) a[q] = __float_as_uint(1.0f + threadIdx.x * 0.001f * q) & 0xffffe000u;
  for (int q = 0; q < 2; q++) b[q] = __float_as_uint(0.001f * threadIdx.x + q) & 0xffffe000u;
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) mma_tf32(c[k], a, b);
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++)
    for (int q = 0; q < 4; q++) s += c[k][q];
  out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void peak_copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16) {
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
}

template <typename F> float best_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

}  // namespace
}  // namespace nfftcu

using namespace nfftcu;

extern "C" int nfftcu_measure_peaks(int device, double *fp64_tensor_tflops, double *tf32_mma_sync_tflops,
                                    double *copy_gbs) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("nfftcu_measure_peaks: no usable CUDA device %d (count %d)", device, ndev);
    return NFFTCU_ENODEV;
  }
  NFFTCU_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  NFFTCU_CUDA(cudaGetDeviceProperties(&prop, device));
  const int sms = prop.multiProcessorCount, warps = 16, n = 4000, reps = 5;
  void *out = nullptr;
  NFFTCU_CUDA(cudaMalloc(&out, (size_t) sms * warps * 32 * sizeof(double)));
  if (fp64_tensor_tflops) {
    const float ms = best_ms([&] { peak_dmma_kernel<<<sms, warps * 32>>>((double *) out, n); }, reps);
    *fp64_tensor_tflops = 2.0 * (double) sms * warps * n * 8 * 256 / (ms * 1e-3) * 1e-12;
  }
  if (tf32_mma_sync_tflops) {
    const float ms = best_ms([&] { peak_tf32_kernel<<<sms, warps * 32>>>((float *) out, n); }, reps);
    *tf32_mma_sync_tflops = 2.0 * (double) sms * warps * n * 8 * (16 * 8 * 8) / (ms * 1e-3) * 1e-12;
  }
  cudaFree(out);
  if (copy_gbs) {
    const size_t bytes = (size_t) 1 << 30;
    void *a = nullptr, *b = nullptr;
    NFFTCU_CUDA(cudaMalloc(&a, bytes));
    if (cudaMalloc(&b, bytes) != cudaSuccess) {
      cudaFree(a);
      set_error("nfftcu_measure_peaks: cudaMalloc failed");
      return NFFTCU_ENOMEM;
    }
    cudaMemset(a, 1, bytes);
    const float ms = best_ms([&] { peak_copy_kernel<<<sms * 8, 512>>>((const uint4 *) a, (uint4 *) b, bytes / 16); }, reps);
    *copy_gbs = 2.0 * (double) bytes / (ms * 1e-3) * 1e-9;
    cudaFree(a);
    cudaFree(b);
  }
  NFFTCU_CUDA(cudaGetLastError());
  return NFFTCU_OK;
}
