// tma_probe.cu -- does cuTensorMapEncodeTiled accept the 5-D "core matrix" view of a row-major fp32 matrix
// (dims {4 cols, 8 rows, C/4, R/8, batch} with non-monotonic strides) and does a box land in shared memory in the
// canonical no-swizzle UMMA order?   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int c2, int c3, int c4) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t sbar = (uint32_t)__cvta_generic_to_shared(&bar), sdst = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int bytes = 4 * 8 * 4 * 2 * 4;   // box {4,8,4,2,1} floats
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(sdst), "l"(&tm), "r"(0), "r"(0), "r"(c2), "r"(c3), "r"(c4), "r"(sbar) : "memory");
  }
  uint32_t done = 0;
  for (long spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(sbar), "r"(0) : "memory");
    if (spin > (1L << 22)) { if (threadIdx.x == 0) printf("TIMEOUT\n"); return; }
  }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = sm[i];
}

int main() {
  const int R = 40, C = 32, NB = 2;   // matrix [NB][R][C] row-major, element value = b*10000 + r*100 + c
  std::vector<float> h((size_t)NB * R * C);
  for (int b = 0; b < NB; ++b) for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[((size_t)b * R + r) * C + c] = b * 10000 + r * 100 + c;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 4096);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[5] = {4, 8, (cuuint64_t)C / 4, (cuuint64_t)R / 8, NB};
  cuuint64_t strides[4] = {(cuuint64_t)C * 4, 16, (cuuint64_t)8 * C * 4, (cuuint64_t)R * C * 4};
  cuuint32_t box[5] = {4, 8, 4, 2, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  probe<<<1, 128, 4096>>>(tm, o, /*c2 = col quad*/ 2, /*c3 = row group*/ 1, /*batch*/ 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s\n", cudaGetErrorString(e));
  std::vector<float> out(256);
  cudaMemcpy(out.data(), o, 1024, cudaMemcpyDeviceToHost);
  // expected canonical order: [rg 2][cq 4][r8 8][c4 4] with rows 8..23, cols 8..23 of batch 1
  int bad = 0;
  for (int rg = 0; rg < 2; ++rg) for (int cq = 0; cq < 4; ++cq) for (int r8 = 0; r8 < 8; ++r8) for (int c4 = 0; c4 < 4; ++c4) {
    const float want = 10000 + (8 + rg * 8 + r8) * 100 + (8 + cq * 4 + c4);
    const float got = out[((rg * 4 + cq) * 8 + r8) * 4 + c4];
    if (want != got) { if (bad < 5) printf("mismatch rg%d cq%d r%d c%d want %.0f got %.0f\n", rg, cq, r8, c4, want, got); ++bad; }
  }
  printf("first 8: %.0f %.0f %.0f %.0f %.0f %.0f %.0f %.0f\nmismatches: %d\n", out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7], bad);
  return 0;
}
