// mma_microbench.cu -- issue rate and dependent latency of the register-operand tensor instruction
// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 on sm_100a (the instruction a register-chained
// narrow-MLP evaluator would use), next to FFMA.  Decides whether the 16-wide layers of the BASELINE regression
// configs are better served by 3xTF32 register MMAs than by CUDA-core FMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_mb mma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ILP independent accumulator chains per warp; each chain is a sequence of dependent MMAs
template <int ILP>
__global__ void __launch_bounds__(1024, 1) k_mma(int iters, float* out, long long* cyc) {
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + 0.001f * (threadIdx.x + i));
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + 0.001f * (threadIdx.x + i));
  float c[ILP][4];
  for (int j = 0; j < ILP; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < ILP; ++j) mma_tf32(c[j], a, b);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0.f;
  for (int j = 0; j < ILP; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  if (s == 1.2345f) out[0] = s;
}

// C fragment fed back as the next A fragment (the layer-to-layer chaining pattern), with a relu between
__global__ void __launch_bounds__(1024, 1) k_chain(int iters, float* out, long long* cyc) {
  unsigned b[2];
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.01f + 0.0001f * (threadIdx.x + i));
  float c0[4] = {1.f, 1.f, 1.f, 1.f}, c1[4] = {1.f, 1.f, 1.f, 1.f};
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      // two n-tiles of the previous layer = two k-steps of the next one
      unsigned a0[4] = {__float_as_uint(fmaxf(c0[0], 0.f)), __float_as_uint(fmaxf(c0[2], 0.f)), __float_as_uint(fmaxf(c0[1], 0.f)), __float_as_uint(fmaxf(c0[3], 0.f))};
      unsigned a1[4] = {__float_as_uint(fmaxf(c1[0], 0.f)), __float_as_uint(fmaxf(c1[2], 0.f)), __float_as_uint(fmaxf(c1[1], 0.f)), __float_as_uint(fmaxf(c1[3], 0.f))};
      float n0[4] = {0.f, 0.f, 0.f, 0.f}, n1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_tf32(n0, a0, b); mma_tf32(n1, a0, b);
      mma_tf32(n0, a1, b); mma_tf32(n1, a1, b);
      for (int i = 0; i < 4; ++i) { c0[i] = n0[i] + 1.f; c1[i] = n1[i] + 1.f; }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
  if (s == 1.2345f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(1024, 1) k_ffma(int iters, float* out, long long* cyc) {
  float a = 1.0f + 0.001f * threadIdx.x, b = 0.5f;
  float c[ILP];
  for (int j = 0; j < ILP; ++j) c[j] = (float)j;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < ILP; ++j) c[j] = fmaf(a, c[j], b);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0.f;
  for (int j = 0; j < ILP; ++j) s += c[j];
  if (s == 1.2345f) out[0] = s;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 8 * 148);
  const int iters = 2000;
  long long h[148];
  auto report = [&](const char* name, int threads, int per_iter, double flop_per_op) {
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double ops_per_warp = (double)iters * per_iter;
    const double cyc_per_op_warp = (double)h[0] / ops_per_warp;                 // as seen by one warp
    const double warps = threads / 32.0;
    const double sm_ops_per_cyc = ops_per_warp * warps / (double)h[0];
    printf("%-34s threads %4d : %.2f cycles per op per warp, %.3f warp-ops/cycle/SM, %.0f FLOP/cycle/SM\n", name, threads,
           cyc_per_op_warp, sm_ops_per_cyc, sm_ops_per_cyc * flop_per_op);
  };
  for (int threads : {32, 128, 256, 512, 1024}) {
    k_mma<1><<<148, threads>>>(iters, out, cyc); report("mma.m16n8k8.tf32 dependent chain", threads, 8, 2048.0);
    k_mma<2><<<148, threads>>>(iters, out, cyc); report("mma.m16n8k8.tf32 ILP 2", threads, 16, 2048.0);
    k_mma<4><<<148, threads>>>(iters, out, cyc); report("mma.m16n8k8.tf32 ILP 4", threads, 32, 2048.0);
    k_mma<8><<<148, threads>>>(iters, out, cyc); report("mma.m16n8k8.tf32 ILP 8", threads, 64, 2048.0);
    k_chain<<<148, threads>>>(iters, out, cyc); report("layer chain (4 mma + relu) / 4", threads, 32, 2048.0);
    k_ffma<1><<<148, threads>>>(iters, out, cyc); report("FFMA dependent chain", threads, 8, 64.0);
    k_ffma<8><<<148, threads>>>(iters, out, cyc); report("FFMA ILP 8", threads, 64, 64.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
