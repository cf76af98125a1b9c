// lds_microbench.cu -- cost (SM cycles per warp-wide LDS.128 / STS.128) of the shared-memory access
// patterns the tile GEMMs use.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_mb lds_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) k(int pattern, int iters, float* out, long long* cyc, int stride) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int off;  // float offset, multiple of 4
  switch (pattern) {
    case 0: off = 0; break;                                       // all lanes same address
    case 1: off = (lane & 3) * 4; break;                          // 4 distinct contiguous (W pattern)
    case 2: off = (lane >> 2) * stride; break;                    // 8 distinct rows, 4 lanes each (A pattern)
    case 3: off = lane * stride; break;                           // 32 distinct rows (thread-per-row)
    case 4: off = (lane & 1) * 4; break;                          // 2 distinct contiguous
    case 5: off = (lane >> 4) * stride; break;                    // 2 distinct rows (half-warps)
    case 6: off = (lane >> 3) * stride; break;                    // 4 distinct rows (quarter-warps)
    case 7: off = (lane & 7) * 4; break;                          // 8 distinct contiguous = 128B
    case 8: off = ((lane >> 2) & 1) * stride + (lane & 3) * 4 + (lane >> 3) * 2 * stride; break; // store pattern: 4 jt x 8 q
    default: off = lane * 4; break;                               // 32 distinct contiguous (512B)
  }
  off += warp * 8;  // different warps, different base (keeps 16B alignment)
  const float4* p = reinterpret_cast<const float4*>(sm + off);
  float4 acc = make_float4(0, 0, 0, 0);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      float4 v = p[u * 64];  // +1KB per unrolled load, same bank pattern
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc.x == 1.2345f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 8 * 148);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  const char* names[] = {"uniform", "4 contiguous (W)", "8 rows x4 lanes (A)", "32 rows", "2 contiguous", "2 rows (half-warps)",
                         "4 rows (quarter-warps)", "8 contiguous 128B", "4jt x 8q rows", "32 contiguous 512B"};
  for (int threads : {256, 1024}) {
    for (int stride : {12, 16, 20}) {
      for (int pat = 0; pat < 10; ++pat) {
        k<<<148, threads, 100 * 1024>>>(pat, iters, out, cyc, stride);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = (double)h[0] / ((double)iters * 16 * (threads / 32));
        printf("threads %4d stride %2d pattern %d %-26s : %.2f SM-cycles per warp LDS.128 (incl. 4 FADD)\n", threads, stride, pat,
               names[pat], c);
      }
    }
  }
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
