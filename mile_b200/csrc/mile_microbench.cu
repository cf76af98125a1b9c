// mile_microbench.cu -- measured FP32 CUDA-core peak for the roofline denominator.
// MEASURED_PEAKS.json carries HBM GB/s and bf16 tensor TFLOP/s only; the narrow-MLP configs are
// bound by the FP32 FMA pipe, so bench.py measures that peak live.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mile_b200.h"

template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) fma_peak_kernel(float* out, int iters, float a, float b) {
  // 16 independent accumulator chains per thread
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  if (VARIANT == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
  } else {
    // packed fp32x2 FMA (sm_100+: fma.rn.f32x2 -> SASS FFMA2)
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(av), "l"(bv));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;  // never true; keeps the loop alive
}

extern "C" int mile_measure_fp32_peak(int32_t device, int32_t variant, double* tflops_out) {
  if (!tflops_out) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  float* out = nullptr;
  if (cudaMalloc(&out, 4) != cudaSuccess) return -1;
  const int blocks = prop.multiProcessorCount * 2, threads = 1024, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    if (variant == 0) fma_peak_kernel<0><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    else fma_peak_kernel<1><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return -1; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  *tflops_out = best;
  return 0;
}
