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

// register-operand tensor instruction of the narrow-MLP evaluator (mile_mma.cuh): 8 independent accumulator chains per warp
__global__ void __launch_bounds__(1024, 1) mma_peak_kernel(float* out, int iters) {
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + 0.001f * (threadIdx.x + i));
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + 0.001f * (threadIdx.x + i));
  float c[8][4];
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  if (s == 12345.678f) out[0] = s;
}

// variant 0: scalar FFMA, 1: packed FFMA2, 2: mma.sync m16n8k8 tf32 (TFLOP/s of tf32 products, 2*16*8*8 per warp instruction)
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
    else if (variant == 1) fma_peak_kernel<1><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    else mma_peak_kernel<<<blocks, threads>>>(out, iters / 4);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return -1; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = variant == 2 ? 2048.0 * 8.0 * (double)(iters / 4) * (double)blocks * (threads / 32)
                                      : 2.0 * 16.0 * (double)iters * (double)blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  *tflops_out = best;
  return 0;
}
