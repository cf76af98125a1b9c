// mile_wide.cuh -- HBM-resident layer-by-layer path for networks that do not fit the shared-memory kernels
// (wide / large-d configs such as the 4x256 complexity-ablation shape, d = 201 218).
//
// The per-chain value_and_grad becomes a sequence of chain-batched GEMMs with fused epilogues
//   forward   A_{l+1}[c] = act(A_l[c] W_l[c] + b_l[c])                 [N x in] [in x out]
//   backward  D_{l-1}[c] = act'(A_l[c]) .* (D_l[c] W_l[c]^T)           [N x out][out x in]
//   dW_l[c]   = A_l[c]^T D_l[c]  (split-K over the rows, two-pass deterministic reduce),  db_l = 1^T D_l
// followed by the same integrator-only kernel the data-sharded variant uses (mile_sharded.cuh).
// The GEMM core here is a plain FP32 SIMT tile kernel (128x128x16 tiles, 8x8 register micro-tiles): exact fp32
// like the reference.  The tcgen05 / TMEM version of this core (3xTF32 split) is the planned replacement
// (DESIGN.md section 4.4); the orchestration, epilogues and parity tests stay.
#pragma once
#include "mile_device.cuh"

struct GemmArgs {
  const float* A; long a_batch; long sam, sak;     // A(m,k) = A[b*a_batch + m*sam + k*sak]
  const float* B; long b_batch; long sbk, sbn;     // B(k,n) = B[b*b_batch + k*sbk + n*sbn]
  float* C; long c_batch; long ldc;                // C(m,n) = C[b*c_batch + s*c_slice + m*ldc + n]
  const float* A_lo; const float* B_lo; float* C_lo;   // tf32 remainders (v - tf32(v)) for the tcgen05 v2 core; C_lo optional output
  int M, N, K, kslices; long c_slice;
  int epi;                                         // 0 none | 1 +bias, act | 2 +bias | 3 * act'(aux value)
  const float* bias; long bias_batch;
  const float* aux; long aux_batch; long ldaux;
  int act, nbatch;
  float* csum;                                     // TMA core, epi 3: optional column-sum partials of C (see Tc2Args)
};

__device__ __forceinline__ float act_value(int act, float z) {
  float a, da;
  act_eval(act, z, a, da);
  return a;
}
// act' from the activation VALUE (valid for identity / relu / sigmoid / tanh / leaky_relu)
__device__ __forceinline__ float act_deriv_from_value(int act, float a) {
  switch (act) {
    case MILE_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case MILE_ACT_SIGMOID: return a * (1.f - a);
    case MILE_ACT_TANH: return 1.f - a * a;
    case MILE_ACT_LEAKY_RELU: return a >= 0.f ? 1.f : 0.01f;
    default: return 1.f;
  }
}

#define WG_BM 128
#define WG_BN 128
#define WG_BK 16

__global__ void __launch_bounds__(256) wide_gemm_kernel(const GemmArgs g) {
  __shared__ float As[WG_BK][WG_BM + 4];
  __shared__ float Bs[WG_BK][WG_BN + 4];
  const int b = blockIdx.z / g.kslices, ks = blockIdx.z % g.kslices;
  const int m0 = blockIdx.y * WG_BM, n0 = blockIdx.x * WG_BN;
  const int kper = ((g.K + g.kslices - 1) / g.kslices + WG_BK - 1) / WG_BK * WG_BK;
  const int kbeg = ks * kper, kend = min(g.K, kbeg + kper);
  const float* A = g.A + (long)b * g.a_batch;
  const float* B = g.B + (long)b * g.b_batch;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // loader index maps: consecutive threads follow the unit-stride dimension of each operand
  const bool a_kfast = g.sak == 1, b_nfast = g.sbn == 1;
  for (int k0 = kbeg; k0 < kend; k0 += WG_BK) {
#pragma unroll
    for (int e = 0; e < (WG_BM * WG_BK) / 256; ++e) {
      const int idx = tid + e * 256;
      const int kk = a_kfast ? (idx % WG_BK) : (idx / WG_BM), mm = a_kfast ? (idx / WG_BK) : (idx % WG_BM);
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < kend) ? __ldg(A + (long)m * g.sam + (long)k * g.sak) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < (WG_BN * WG_BK) / 256; ++e) {
      const int idx = tid + e * 256;
      const int kk = b_nfast ? (idx / WG_BN) : (idx % WG_BK), nn = b_nfast ? (idx % WG_BN) : (idx / WG_BK);
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < kend) ? __ldg(B + (long)k * g.sbk + (long)n * g.sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WG_BK; ++kk) {
      float a[8], bb[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* C = g.C + (long)b * g.c_batch + (long)ks * g.c_slice;
  const float* bias = g.bias ? g.bias + (long)b * g.bias_batch : nullptr;
  const float* aux = g.aux ? g.aux + (long)b * g.aux_batch : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.epi == 1) v = act_value(g.act, v + bias[n]);
      else if (g.epi == 2) v = v + bias[n];
      else if (g.epi == 3) v = v * act_deriv_from_value(g.act, aux[(long)m * g.ldaux + n]);
      C[(long)m * g.ldc + n] = v;
      if (g.C_lo) g.C_lo[(long)b * g.c_batch + (long)ks * g.c_slice + (long)m * g.ldc + n] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
  }
}

// =====================================================================================================
// Skinny shapes (first layer K = n_features, output layer N = n_outputs): pure HBM streaming, one pass over the large
// operand, no tile padding.  A 128x128 GEMM tile would be >90% empty for them.
// =====================================================================================================
#define WS_KMAX 16
#define WS_NMAX 8

// C[M x N] = epi(A[M x K] B[K x N]) with K <= 16: thread = (row, 4 consecutive columns); B and bias live in shared memory.
// grid (ceil(M/64), ceil(N/256), batch), 256 threads.
__global__ void __launch_bounds__(256) wide_smallk_kernel(const GemmArgs g) {
  __shared__ __align__(16) float Bs[WS_KMAX][256];
  __shared__ __align__(16) float bs[256];
  const int b = blockIdx.z, n0 = blockIdx.y * 256, tid = threadIdx.x;
  const float* A = g.A + (long)b * g.a_batch;
  const float* B = g.B + (long)b * g.b_batch;
  const float* bias = g.bias ? g.bias + (long)b * g.bias_batch : nullptr;
  const float* aux = g.aux ? g.aux + (long)b * g.aux_batch : nullptr;
  float* C = g.C + (long)b * g.c_batch;
  for (int e = tid; e < WS_KMAX * 256; e += 256) {
    const int k = e >> 8, nn = e & 255;
    Bs[k][nn] = (k < g.K && n0 + nn < g.N) ? __ldg(B + (long)k * g.sbk + (long)(n0 + nn) * g.sbn) : 0.f;
  }
  bs[tid] = (bias && n0 + tid < g.N) ? __ldg(bias + n0 + tid) : 0.f;
  __shared__ __align__(16) float As[64][WS_KMAX];
  {
    const int m0 = blockIdx.x * 64;
    for (int e = tid; e < 64 * WS_KMAX; e += 256) {
      const int r = e / WS_KMAX, k = e % WS_KMAX;
      As[r][k] = (m0 + r < g.M && k < g.K) ? __ldg(A + (long)(m0 + r) * g.sam + (long)k * g.sak) : 0.f;
    }
  }
  __syncthreads();
  const int tx = tid & 63, ty = tid >> 6, n = n0 + tx * 4;
  if (n >= g.N) return;
  const bool vec = n + 3 < g.N && (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                   (g.epi != 3 || ((g.ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0));
  const int mend = min(g.M, (int)(blockIdx.x + 1) * 64);
  // The thread's [K][4] slab of B stays in registers for all of its rows and the CTA's 64 rows of A are staged once in
  // shared memory (coalesced), so a row costs 4 broadcast LDS.128 + 4 K FMAs + one 16-byte store: the kernel is then
  // bound by writing C.  (Before: 16 LDS.128 of B + 16 uniform global loads of A per row, 100 us for a 100 MB output,
  // profiles/r1k_launches_wide_4x256.csv.)
  float wreg[WS_KMAX][4];
#pragma unroll
  for (int k = 0; k < WS_KMAX; ++k) {
    const float4 w = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    wreg[k][0] = w.x; wreg[k][1] = w.y; wreg[k][2] = w.z; wreg[k][3] = w.w;
  }
  for (int m = blockIdx.x * 64 + ty; m < mend; m += 4) {
    float a[WS_KMAX];
    const float4* ar = reinterpret_cast<const float4*>(&As[m - blockIdx.x * 64][0]);
#pragma unroll
    for (int k4 = 0; k4 < WS_KMAX / 4; ++k4) { const float4 t = ar[k4]; a[4 * k4] = t.x; a[4 * k4 + 1] = t.y; a[4 * k4 + 2] = t.z; a[4 * k4 + 3] = t.w; }
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WS_KMAX; ++k) {
      if (k < g.K) {
        o[0] = fmaf(a[k], wreg[k][0], o[0]); o[1] = fmaf(a[k], wreg[k][1], o[1]);
        o[2] = fmaf(a[k], wreg[k][2], o[2]); o[3] = fmaf(a[k], wreg[k][3], o[3]);
      }
    }
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.epi == 3) {
      if (vec) { const float4 t = *reinterpret_cast<const float4*>(aux + (long)m * g.ldaux + n); x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w; }
      else for (int e = 0; e < 4; ++e) if (n + e < g.N) x[e] = aux[(long)m * g.ldaux + n + e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (g.epi == 1) o[e] = act_value(g.act, o[e] + bs[tx * 4 + e]);
      else if (g.epi == 2) o[e] = o[e] + bs[tx * 4 + e];
      else if (g.epi == 3) o[e] = o[e] * act_deriv_from_value(g.act, x[e]);
    }
    if (vec) *reinterpret_cast<float4*>(C + (long)m * g.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
    else for (int e = 0; e < 4; ++e) if (n + e < g.N) C[(long)m * g.ldc + n + e] = o[e];
  }
}

// C[M x N] = epi(A[M x K] B[K x N]) with N <= 8: one warp per row, lanes stride K (float4 when A rows are contiguous and
// aligned), warp-shuffle reduction.  B [K x N] in shared memory (K <= 1024).  grid (ceil(M/64), 1, batch), 256 threads.
__global__ void __launch_bounds__(256) wide_smalln_kernel(const GemmArgs g) {
  __shared__ float Bs[1024 * WS_NMAX];
  const int b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* A = g.A + (long)b * g.a_batch;
  const float* B = g.B + (long)b * g.b_batch;
  const float* bias = g.bias ? g.bias + (long)b * g.bias_batch : nullptr;
  const float* aux = g.aux ? g.aux + (long)b * g.aux_batch : nullptr;
  float* C = g.C + (long)b * g.c_batch;
  for (int e = tid; e < g.K * WS_NMAX; e += 256) {
    const int k = e / WS_NMAX, j = e % WS_NMAX;
    Bs[e] = j < g.N ? __ldg(B + (long)k * g.sbk + (long)j * g.sbn) : 0.f;
  }
  __syncthreads();
  const bool vec = g.sak == 1 && (g.sam & 3) == 0 && (g.K & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const int mend = min(g.M, (int)(blockIdx.x + 1) * 64);
  // 4 rows per warp pass: all loads of the pass are issued before the first use (this kernel is a pure HBM stream)
  for (int mb = blockIdx.x * 64 + warp * 4; mb < mend; mb += 32) {
    float acc[4][WS_NMAX];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < WS_NMAX; ++j) acc[r][j] = 0.f;
    if (vec) {
      for (int k4 = lane; k4 < (g.K >> 2); k4 += 64) {
        float4 a[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            a[r][h] = (mb + r < mend && k4 + 32 * h < (g.K >> 2))
                          ? __ldg(reinterpret_cast<const float4*>(A + (long)(mb + r) * g.sam) + k4 + 32 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (k4 + 32 * h >= (g.K >> 2)) continue;
          const float* w = Bs + (k4 + 32 * h) * 4 * WS_NMAX;
#pragma unroll
          for (int j = 0; j < WS_NMAX; ++j) {
            if (j >= g.N) continue;
            const float w0 = w[j], w1 = w[WS_NMAX + j], w2 = w[2 * WS_NMAX + j], w3 = w[3 * WS_NMAX + j];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r][j] += a[r][h].x * w0 + a[r][h].y * w1 + a[r][h].z * w2 + a[r][h].w * w3;
          }
        }
      }
    } else {
      for (int k = lane; k < g.K; k += 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (mb + r >= mend) continue;
          const float a = __ldg(A + (long)(mb + r) * g.sam + (long)k * g.sak);
#pragma unroll
          for (int j = 0; j < WS_NMAX; ++j) acc[r][j] = fmaf(a, Bs[k * WS_NMAX + j], acc[r][j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < WS_NMAX; ++j) {
      if (j >= g.N) continue;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[r][j] += __shfl_xor_sync(0xffffffffu, acc[r][j], o);
    }
    // lane (r * 8 + j) writes output (row mb + r, column j)
    const int r = lane >> 3, j = lane & 7, m = mb + r;
    if (j < g.N && m < mend) {
      float v = 0.f;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int jj = 0; jj < WS_NMAX; ++jj) if (rr == r && jj == j) v = acc[rr][jj];
      if (g.epi == 1) v = act_value(g.act, v + bias[j]);
      else if (g.epi == 2) v = v + bias[j];
      else if (g.epi == 3) v = v * act_deriv_from_value(g.act, aux[(long)m * g.ldaux + j]);
      C[(long)m * g.ldc + j] = v;
    }
  }
}

// Row-reduction products for the parameter gradients of skinny layers and for every bias gradient:
//   part[b][slice][...] = sum_{r in slice} Wd(r, t) * S(r, q),   t < WD (thread = column of the wide operand), q < s <= 16
// `S == nullptr` means S == 1 (column sums of Wd: the bias gradient).  small_is_row: output index q * WD + t (S indexes the
// rows of dW) else t * s + q.  grid (nslices, ceil(WD/256), batch), 256 threads; deterministic two-pass with
// wide_slice_reduce_kernel.
struct RowReduceArgs {
  const float* Wd; long wd_batch, wd_ld; int WD;
  const float* S; long s_batch, s_ld; int s; int small_is_row;
  long rows; int nslices;
  float* part; long p_batch, p_slice;
};

__global__ void __launch_bounds__(256) wide_rowreduce_kernel(const RowReduceArgs g) {
  const int t = blockIdx.y * 256 + threadIdx.x, b = blockIdx.z, sl = blockIdx.x;
  if (t >= g.WD) return;
  const long per = (g.rows + g.nslices - 1) / g.nslices, r0 = sl * per, r1 = min(g.rows, r0 + per);
  const float* W = g.Wd + (long)b * g.wd_batch + t;
  const float* S = g.S ? g.S + (long)b * g.s_batch : nullptr;
  float acc[WS_KMAX];
#pragma unroll
  for (int q = 0; q < WS_KMAX; ++q) acc[q] = 0.f;
  long r = r0;
  if (!S) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (; r + 15 < r1; r += 16) {
      float w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = __ldg(W + (r + i) * g.wd_ld);
      s0 += (w[0] + w[4]) + (w[8] + w[12]); s1 += (w[1] + w[5]) + (w[9] + w[13]);
      s2 += (w[2] + w[6]) + (w[10] + w[14]); s3 += (w[3] + w[7]) + (w[11] + w[15]);
    }
    for (; r < r1; ++r) s0 += __ldg(W + r * g.wd_ld);
    acc[0] = (s0 + s1) + (s2 + s3);
  } else {
    for (; r + 7 < r1; r += 8) {
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = __ldg(W + (r + i) * g.wd_ld);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < WS_KMAX; ++q)
          if (q < g.s) acc[q] = fmaf(w[i], __ldg(S + (r + i) * g.s_ld + q), acc[q]);
    }
    for (; r < r1; ++r) {
      const float w = __ldg(W + r * g.wd_ld);
#pragma unroll
      for (int q = 0; q < WS_KMAX; ++q)
        if (q < g.s) acc[q] = fmaf(w, __ldg(S + r * g.s_ld + q), acc[q]);
    }
  }
  float* out = g.part + (long)b * g.p_batch + (long)sl * g.p_slice;
  const int s = S ? g.s : 1;
#pragma unroll
  for (int q = 0; q < WS_KMAX; ++q)
    if (q < s) out[g.small_is_row ? (long)q * g.WD + t : (long)t * s + q] = acc[q];
}

// sum the split-K slices: dst[b*dst_batch + e] = sum_s src[b*src_batch + s*slice + e]   (fixed order)
__global__ void wide_slice_reduce_kernel(const float* __restrict__ src, long src_batch, long slice, int kslices,
                                         float* __restrict__ dst, long dst_batch, long n, int nbatch) {
  const long total = n * nbatch;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long b = t / n, e = t % n;
    // four interleaved partial sums so that the slice loads are in flight together (148 dependent L2 loads per
    // element took 15 us for a 1 MB reduction); fixed association, so the result is run-to-run deterministic
    const float* p = src + b * src_batch + e;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
#pragma unroll 2
    for (; k + 3 < kslices; k += 4) {
      s0 += p[(long)k * slice]; s1 += p[(long)(k + 1) * slice]; s2 += p[(long)(k + 2) * slice]; s3 += p[(long)(k + 3) * slice];
    }
    for (; k < kslices; ++k) s0 += p[(long)k * slice];
    dst[b * dst_batch + e] = (s0 + s1) + (s2 + s3);
  }
}

// per (chain, row): log-likelihood term and d/d(out) (probabilistic.py:93-109); block partial sums of ll
__global__ void __launch_bounds__(256) wide_loglik_kernel(DevModel M, const float* __restrict__ out, float* __restrict__ dout,
                                                          const void* __restrict__ y, long N, long N8, float* __restrict__ llpart) {
  __shared__ __align__(16) float red[64];
  int phase = 0;
  const int c = blockIdx.y, K = M.dims[M.NL];
  const long r = blockIdx.x * (long)blockDim.x + threadIdx.x;
  float ll = 0.f;
  if (r < N) {
    const float* o = out + ((long)c * N8 + r) * K;
    float* dd = dout + ((long)c * N8 + r) * K;
    if (M.task == MILE_TASK_REGRESSION) {
      const float yv = reinterpret_cast<const float*>(y)[r], mu = o[0], s = o[1];
      const float e = expf(s), sigma = fminf(fmaxf(e, 1e-6f), 1e6f);
      const float inside = (e > 1e-6f && e < 1e6f) ? 1.f : 0.f;
      const float s2 = sigma * sigma, res = yv - mu, q = res * res / s2;
      ll = (logf(6.283185307179586f * s2) + q) / -2.f;
      float dmu = res / s2, ds = (q - 1.f) * inside;
      if (isnan(ll)) { ll = 0.f; dmu = 0.f; ds = 0.f; }
      dd[0] = dmu * M.n_batches; dd[1] = ds * M.n_batches;
      for (int k = 2; k < K; ++k) dd[k] = 0.f;
    } else {
      const int yi = reinterpret_cast<const int*>(y)[r];
      float m = o[0];
      for (int k = 1; k < K; ++k) m = fmaxf(m, o[k]);
      float se = 0.f;
      for (int k = 0; k < K; ++k) se += expf(o[k] - m);
      ll = o[yi] - (m + logf(se));
      const bool bad = isnan(ll);
      for (int k = 0; k < K; ++k) dd[k] = bad ? 0.f : (-expf(o[k] - m) / se + (k == yi ? 1.f : 0.f)) * M.n_batches;
      if (bad) ll = 0.f;
    }
  }
  float v[1] = {ll};
  block_sum<1, 256>(v, red, phase);
  if (threadIdx.x == 0) llpart[(long)c * gridDim.x + blockIdx.x] = v[0];
}

// Online logsumexp fold of forward outputs out [n][Nt][K] into the per-chain state (same update as lppd_fold, mile_kernel.cuh)
__global__ void __launch_bounds__(256) wide_lppd_fold_kernel(DevModel M, const float* __restrict__ out, const void* __restrict__ yt,
                                                             float* __restrict__ lm, float* __restrict__ ls, int n, long Nt) {
  const int K = M.dims[M.NL];
  const long total = (long)n * Nt;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx % Nt;
    const float lp = pointwise_lppd_row(M, out + idx * K, yt, r);
    const float m = lm[idx], s = ls[idx];
    const float nm = fmaxf(m, lp);
    const float safe = isfinite(nm) ? nm : 0.f;
    lm[idx] = nm;
    ls[idx] = s * expf((isfinite(m) ? m : -INFINITY) - safe) + expf(lp - safe);
  }
}

#define WF_CLUSTER 8
// Every parameter-gradient block of an evaluation (kernel / bias of each layer) is left by its producer as `nslices`
// partial sums (split-K slices of the dW GEMMs, per-CTA partials of the streaming kernels, per-row-group column sums of
// the delta epilogues).  The finalize kernel adds them in slice order -- ONE pass instead of a slice-reduction launch
// behind every producer (8 launches and 20 % of the kernel time of an evaluation, profiles/r2n_launches_wide_4x256.csv).
#define WJ_MAX 32
struct WideJobs {
  int n;
  const float* src[WJ_MAX]; long src_batch[WJ_MAX]; long slice[WJ_MAX];
  int nslices[WJ_MAX], dst_off[WJ_MAX], len[WJ_MAX];
};

// gl[c][i] = sum of the partials of element i + w * prior gradient;  gl[c][d] = n_batches * sum(ll partials) + w * log prior
// grid (NB, chains): the d elements of a chain are strided over the NB x 1024 threads of its CTAs (the partials of the
// 4x256 config are 89 MB: the pass needs every SM).  The two per-chain scalars go through per-CTA partials; the CTA that
// arrives last (atomic ticket) adds them in CTA order, so the value does not depend on which CTA that was.
__global__ void __launch_bounds__(1024) wide_finalize_kernel(DevModel M, const float* __restrict__ theta, float* __restrict__ gl,
                                                            const float* __restrict__ llpart, int nblk, float prior_weight,
                                                            float* __restrict__ fin_part, unsigned int* __restrict__ fin_count,
                                                            const __grid_constant__ WideJobs J) {
  __shared__ __align__(16) float red[256];
  __shared__ int is_last;
  const int G = (int)gridDim.x, rank = (int)blockIdx.x;
  int phase = 0;
  const int c = blockIdx.y, d = M.d;
  const float loc = M.prior_loc, sc = M.prior_scale, s2 = sc * sc;
  const float lognorm = M.prior == MILE_PRIOR_NORMAL ? logf(6.283185307179586f * s2) : logf(2.f * sc);
  float v[2] = {0.f, 0.f};
  const int lane = threadIdx.x & 31;
  for (int j = 0; j < J.n; ++j) {
    const float* __restrict__ src = J.src[j] + (long)c * J.src_batch[j];
    const long slice = J.slice[j];
    const int ns = J.nslices[j], len = J.len[j], off = J.dst_off[j];
    // TPE threads (a power of two <= 32, consecutive lanes) share one element: lane t adds slices t, t + TPE, ... and the
    // group finishes with an xor-shuffle tree.  Many-slice jobs (row-group column sums: 4 per 128-row block, per-CTA
    // partials of the streaming kernels) would otherwise be a serial chain of hundreds of L2 round trips on a handful of
    // threads.  The association is fixed by (ns, TPE): run-to-run deterministic.
    int tpe = 1;
    while (tpe < 32 && tpe * 8 <= ns) tpe <<= 1;
    if (tpe == 1 && (len & 3) == 0 && (slice & 3) == 0 && (J.src_batch[j] & 3) == 0 && (reinterpret_cast<uintptr_t>(J.src[j]) & 15) == 0) {
      // few slices, long block (split-K partials of a weight matrix): four consecutive elements per thread, 16-byte loads
      for (int e4 = rank * 1024 + threadIdx.x; e4 < (len >> 2); e4 += G * 1024) {
        const float4* p = reinterpret_cast<const float4*>(src) + e4;
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        int k = 0;
#pragma unroll 4
        for (; k + 1 < ns; k += 2) {
          const float4 a = p[(long)k * (slice >> 2)], b = p[(long)(k + 1) * (slice >> 2)];
          s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w; s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
        }
        if (k < ns) { const float4 a = p[(long)k * (slice >> 2)]; s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w; }
        const float sum4[4] = {s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = off + 4 * e4 + q;
          const float dlt = theta[(long)c * d + i] - loc;
          float pv, pg;
          if (M.prior == MILE_PRIOR_NORMAL) { pv = (lognorm + dlt * dlt / s2) / -2.f; pg = -dlt / s2; }
          else { pv = -lognorm - fabsf(dlt) / sc; pg = -((dlt > 0.f) - (dlt < 0.f)) / sc; }
          gl[(long)c * (d + 1) + i] = sum4[q] + pg * prior_weight;
          v[0] += pv;
        }
      }
      continue;
    }
    const int sub = lane & (tpe - 1);
    const int per_pass = (G * 1024) / tpe;
    const int npass = (len + per_pass - 1) / per_pass;
    for (int ps = 0; ps < npass; ++ps) {      // (uniform trip count: the shuffles below need whole warps)
      const int e = ps * per_pass + (rank * 1024 + (int)threadIdx.x) / tpe;
      const bool on = e < len;
      const float* p = src + (on ? e : 0);
      float s0 = 0.f, s1 = 0.f, s2_ = 0.f, s3 = 0.f;
      int k = sub;
      if (on) {
#pragma unroll 2
        for (; k + 3 * tpe < ns; k += 4 * tpe) {
          s0 += p[(long)k * slice]; s1 += p[(long)(k + tpe) * slice]; s2_ += p[(long)(k + 2 * tpe) * slice]; s3 += p[(long)(k + 3 * tpe) * slice];
        }
        for (; k < ns; k += tpe) s0 += p[(long)k * slice];
      }
      float sum = (s0 + s1) + (s2_ + s3);
      for (int o = tpe >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (on && sub == 0) {
        const int i = off + e;
        const float dlt = theta[(long)c * d + i] - loc;
        float pv, pg;
        if (M.prior == MILE_PRIOR_NORMAL) { pv = (lognorm + dlt * dlt / s2) / -2.f; pg = -dlt / s2; }
        else { pv = -lognorm - fabsf(dlt) / sc; pg = -((dlt > 0.f) - (dlt < 0.f)) / sc; }
        gl[(long)c * (d + 1) + i] = sum + pg * prior_weight;
        v[0] += pv;
      }
    }
  }
  for (int i = rank * 1024 + threadIdx.x; i < nblk; i += G * 1024) v[1] += llpart[(long)c * nblk + i];
  block_sum<2, 1024>(v, red, phase);
  if (threadIdx.x == 0) {
    fin_part[((long)c * G + rank) * 2] = v[0]; fin_part[((long)c * G + rank) * 2 + 1] = v[1];
    __threadfence();
    is_last = atomicAdd(fin_count + c, 1u) == (unsigned int)(G - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    float p = 0.f, l = 0.f;
    for (int r = 0; r < G; ++r) {
      p += *reinterpret_cast<volatile float*>(fin_part + ((long)c * G + r) * 2);
      l += *reinterpret_cast<volatile float*>(fin_part + ((long)c * G + r) * 2 + 1);
    }
    gl[(long)c * (d + 1) + d] = l * M.n_batches + p * prior_weight;
    fin_count[c] = 0u;      // ready for the next launch (stream order)
  }
}

// =====================================================================================================
// First layer (K = n_features <= 16).  Forward  C = act(A B + bias):  the thread's [K][4] slab of B and its bias stay in
// registers, the CTA's 128 rows of A are staged once in shared memory, a row costs KQ broadcast LDS.128 + 4 K FMAs + one
// 16-byte store (the generic wide_smallk_kernel spent 41 % of the issue slots for 0.9 TB/s: runtime activation switch,
// predicated K loop, 64-row CTAs at two CTAs per SM; ncu in profiles/r2v_ncu_wide_small_kernels.txt).
// grid (ceil(M/128), ceil(N/256), batch), 256 threads; requires A rows contiguous (sak == 1), N % 4 == 0, aligned C.
// =====================================================================================================
template <int KQ, bool RELU>
__global__ void __launch_bounds__(256) wide_first_kernel(const GemmArgs g) {
  constexpr int KP = 4 * KQ, ROWS = 128;
  __shared__ __align__(16) float As[ROWS][KP];
  const int b = blockIdx.z, n0 = blockIdx.y * 256, m0 = blockIdx.x * ROWS, tid = threadIdx.x;
  const float* A = g.A + (long)b * g.a_batch;
  const float* B = g.B + (long)b * g.b_batch;
  const float* bias = g.bias + (long)b * g.bias_batch;
  float* C = g.C + (long)b * g.c_batch;
  for (int e = tid; e < ROWS * KP; e += 256) {
    const int r = e / KP, k = e % KP;
    As[r][k] = (m0 + r < g.M && k < g.K) ? __ldg(A + (long)(m0 + r) * g.sam + k) : 0.f;
  }
  const int tx = tid & 63, ty = tid >> 6, n = n0 + tx * 4;
  float wreg[KP][4], bs[4];
#pragma unroll
  for (int k = 0; k < KP; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) wreg[k][e] = (k < g.K && n + e < g.N) ? __ldg(B + (long)k * g.sbk + (long)(n + e) * g.sbn) : 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) bs[e] = n + e < g.N ? __ldg(bias + n + e) : 0.f;
  __syncthreads();
  if (n >= g.N) return;
  const int mend = min(g.M - m0, ROWS);
#pragma unroll 2
  for (int r = ty; r < mend; r += 4) {
    float o[4] = {bs[0], bs[1], bs[2], bs[3]};
#pragma unroll
    for (int q = 0; q < KQ; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(&As[r][4 * q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[e] = fmaf(a.x, wreg[4 * q][e], o[e]); o[e] = fmaf(a.y, wreg[4 * q + 1][e], o[e]);
        o[e] = fmaf(a.z, wreg[4 * q + 2][e], o[e]); o[e] = fmaf(a.w, wreg[4 * q + 3][e], o[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = RELU ? fmaxf(o[e], 0.f) : act_value(g.act, o[e]);
    *reinterpret_cast<float4*>(C + (long)(m0 + r) * g.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Weight gradient of a layer with few inputs:  part[b][slice][i * WD + j] = sum_{r in slice} S(r, i) Wd(r, j),  i < K <= 4 KQ.
// Thread = 4 columns j of the wide operand (one LDG.128 per row, four rows in flight), S rows staged in shared memory;
// the four row phases of a CTA are added through shared memory.  grid (nslices, ceil(WD/256), batch), 256 threads;
// requires WD % 4 == 0, wd_ld % 4 == 0, aligned Wd.
template <int KQ>
__global__ void __launch_bounds__(256) wide_dw_small_kernel(const RowReduceArgs g) {
  constexpr int KP = 4 * KQ, CH = 128;
  __shared__ __align__(16) float Ss[CH][KP];
  __shared__ __align__(16) float red[KP][256];
  const int b = blockIdx.z, sl = blockIdx.x, tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const int j = blockIdx.y * 256 + tx * 4;
  const long per = (g.rows + g.nslices - 1) / g.nslices, r0 = sl * per, r1 = min(g.rows, r0 + per);
  const float* W = g.Wd + (long)b * g.wd_batch + j;
  const float* S = g.S + (long)b * g.s_batch;
  float acc[KP][4];
#pragma unroll
  for (int k = 0; k < KP; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[k][e] = 0.f;
  const bool on = j < g.WD;
  for (long c0 = r0; c0 < r1; c0 += CH) {
    const int nr = (int)min((long)CH, r1 - c0);
    __syncthreads();
    for (int e = tid; e < CH * KP; e += 256) {
      const int r = e / KP, k = e % KP;
      Ss[r][k] = (r < nr && k < g.s) ? __ldg(S + (c0 + r) * g.s_ld + k) : 0.f;
    }
    __syncthreads();
    if (on) {
      for (int r = ty; r < nr; r += 16) {
        float4 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          w[u] = r + 4 * u < nr ? __ldg(reinterpret_cast<const float4*>(W + (c0 + r + 4 * u) * g.wd_ld)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (r + 4 * u >= nr) break;
#pragma unroll
          for (int q = 0; q < KQ; ++q) {
            const float4 sv = *reinterpret_cast<const float4*>(&Ss[r + 4 * u][4 * q]);
            const float s4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              acc[4 * q + kk][0] = fmaf(s4[kk], w[u].x, acc[4 * q + kk][0]); acc[4 * q + kk][1] = fmaf(s4[kk], w[u].y, acc[4 * q + kk][1]);
              acc[4 * q + kk][2] = fmaf(s4[kk], w[u].z, acc[4 * q + kk][2]); acc[4 * q + kk][3] = fmaf(s4[kk], w[u].w, acc[4 * q + kk][3]);
            }
          }
        }
      }
    }
  }
  // row phases 1..3 -> shared memory one after the other, phase 0 adds them in order
#pragma unroll 1
  for (int p = 1; p < 4; ++p) {
    if (ty == p) {
#pragma unroll
      for (int k = 0; k < KP; ++k) *reinterpret_cast<float4*>(&red[k][tx * 4]) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
    }
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        const float4 o = *reinterpret_cast<const float4*>(&red[k][tx * 4]);
        acc[k][0] += o.x; acc[k][1] += o.y; acc[k][2] += o.z; acc[k][3] += o.w;
      }
    }
    __syncthreads();
  }
  if (ty == 0 && on) {
    float* out = g.part + (long)b * g.p_batch + (long)sl * g.p_slice;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      if (k >= g.s) break;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (j + e < g.WD) out[(long)k * g.WD + j + e] = acc[k][e];
    }
  }
}

// =====================================================================================================
// Fused output layer (IN = 128 NV wide last hidden activation, K <= KP outputs): ONE pass over a_last does the head
// forward, the log-likelihood and d/d(out) (probabilistic.py:93-109), the partial sums of dW_head / db_head / ll, and
// writes the delta of the last hidden layer  D = (dout W_head^T) .* act'(a_last)  together with its column sums (the bias
// gradient of that layer).  Replaces five streaming passes over the same 100 MB activation matrix (head forward, log-lik,
// dW_head row-reduction, db_head row-reduction, delta back-propagation: 326 us of a 1.87 ms evaluation in
// profiles/r2n_launches_wide_4x256.csv) by one that reads a_last and writes D once.
// One warp per row: lane owns columns {128 v + 4 lane + e}.  grid (nblk, chains), 256 threads; CTA = contiguous row range.
// Partials per (chain, CTA): [IN*K] dW_head (theta order: in-major), [K] db_head, [IN] column sums of D, [1] ll.
// =====================================================================================================
struct HeadArgs {
  const float* a_last; long a_batch;        // [n][N8][IN]
  float* D; long d_batch;                   // [n][N8][IN]
  const float* theta; int d, kern_off, bias_off;
  const void* y; long N; int rows_per_cta;
  float* pW; float* pb; float* pcol; float* pll;     // partial areas: [n][nblk][IN*K], [n][nblk][K], [n][nblk][IN], [n][nblk]
};

template <int NV, int KP>
__global__ void __launch_bounds__(256) wide_head_kernel(DevModel M, const HeadArgs g) {
  constexpr int IN = 128 * NV, NW = 8, RU = 4;        // RU rows of a warp in flight together
  extern __shared__ __align__(16) float hsm[];
  float* Wh = hsm;                    // [KP][IN] (transposed: a lane's four columns are one conflict-free LDS.128)
  float* bh = Wh + IN * KP;           // [KP]
  float* scr = bh + KP;               // [NW][IN*KP + IN + KP + 1] cross-warp reduction
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.y, K = M.dims[M.NL], nblk = gridDim.x;
  const float* th = g.theta + (long)c * g.d;
  for (int e = tid; e < IN * KP; e += 256) { const int k = e / IN, i = e % IN; Wh[e] = k < K ? th[g.kern_off + i * K + k] : 0.f; }
  if (tid < KP) bh[tid] = tid < K ? th[g.bias_off + tid] : 0.f;
  __syncthreads();
  const float* A = g.a_last + (long)c * g.a_batch;
  float* D = g.D + (long)c * g.d_batch;
  // two outputs: the lane's slab of W_head lives in registers; more: it is re-read from shared memory per row
  constexpr bool WREG = KP == 2;
  float4 wreg[WREG ? NV : 1][KP];
  if (WREG) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < KP; ++k) wreg[v][k] = reinterpret_cast<const float4*>(Wh + k * IN + v * 128)[lane];
  }
  auto wld = [&](int v, int k) -> float4 { return WREG ? wreg[WREG ? v : 0][k] : reinterpret_cast<const float4*>(Wh + k * IN + v * 128)[lane]; };
  float accW[NV * 4][KP], accC[NV * 4], accB[KP], ll_acc = 0.f;
#pragma unroll
  for (int i = 0; i < NV * 4; ++i) { accC[i] = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) accW[i][k] = 0.f; }
#pragma unroll
  for (int k = 0; k < KP; ++k) accB[k] = 0.f;
  const long r_beg = (long)blockIdx.x * g.rows_per_cta, r_end = min(g.N, r_beg + g.rows_per_cta);
  for (long rb = r_beg + warp * RU; rb < r_end; rb += NW * RU) {
    float4 a4[RU][NV];
#pragma unroll
    for (int u = 0; u < RU; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v)
        a4[u][v] = rb + u < r_end ? __ldg(reinterpret_cast<const float4*>(A + (rb + u) * IN + v * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long r = rb + u;
      if (r >= r_end) break;           // (warp-uniform)
      float a[NV * 4];
#pragma unroll
      for (int v = 0; v < NV; ++v) { a[4 * v] = a4[u][v].x; a[4 * v + 1] = a4[u][v].y; a[4 * v + 2] = a4[u][v].z; a[4 * v + 3] = a4[u][v].w; }
      float o[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) o[k] = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const float4 w = wld(v, k);
          o[k] = fmaf(a[4 * v], w.x, o[k]); o[k] = fmaf(a[4 * v + 1], w.y, o[k]);
          o[k] = fmaf(a[4 * v + 2], w.z, o[k]); o[k] = fmaf(a[4 * v + 3], w.w, o[k]);
        }
#pragma unroll
      for (int k = 0; k < KP; ++k) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) o[k] += __shfl_xor_sync(0xffffffffu, o[k], s);
        o[k] += bh[k];
      }
      // log-likelihood term and d/d(out): same arithmetic as wide_loglik_kernel (every lane, redundantly)
      float dd[KP], ll;
#pragma unroll
      for (int k = 0; k < KP; ++k) dd[k] = 0.f;
      if (M.task == MILE_TASK_REGRESSION) {
        const float yv = reinterpret_cast<const float*>(g.y)[r], mu = o[0], sg = o[1];
        const float ex = expf(sg), sigma = fminf(fmaxf(ex, 1e-6f), 1e6f);
        const float inside = (ex > 1e-6f && ex < 1e6f) ? 1.f : 0.f;
        const float s2 = sigma * sigma, res = yv - mu, q = res * res / s2;
        ll = (logf(6.283185307179586f * s2) + q) / -2.f;
        float dmu = res / s2, ds = (q - 1.f) * inside;
        if (isnan(ll)) { ll = 0.f; dmu = 0.f; ds = 0.f; }
        dd[0] = dmu * M.n_batches; dd[1] = ds * M.n_batches;
      } else {
        const int yi = reinterpret_cast<const int*>(g.y)[r];
        float m = o[0];
#pragma unroll
        for (int k = 1; k < KP; ++k) if (k < K) m = fmaxf(m, o[k]);
        float se = 0.f, oy = 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) if (k < K) { se += expf(o[k] - m); if (k == yi) oy = o[k]; }
        ll = oy - (m + logf(se));
        const bool bad = isnan(ll);
#pragma unroll
        for (int k = 0; k < KP; ++k) if (k < K) dd[k] = bad ? 0.f : (-expf(o[k] - m) / se + (k == yi ? 1.f : 0.f)) * M.n_batches;
        if (bad) ll = 0.f;
      }
      ll_acc += ll;
#pragma unroll
      for (int k = 0; k < KP; ++k) accB[k] += dd[k];
      float dl[NV * 4];
#pragma unroll
      for (int i = 0; i < NV * 4; ++i) dl[i] = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const float4 w = wld(v, k);
          dl[4 * v] = fmaf(dd[k], w.x, dl[4 * v]); dl[4 * v + 1] = fmaf(dd[k], w.y, dl[4 * v + 1]);
          dl[4 * v + 2] = fmaf(dd[k], w.z, dl[4 * v + 2]); dl[4 * v + 3] = fmaf(dd[k], w.w, dl[4 * v + 3]);
#pragma unroll
          for (int e = 0; e < 4; ++e) accW[4 * v + e][k] = fmaf(a[4 * v + e], dd[k], accW[4 * v + e][k]);
        }
#pragma unroll
      for (int i = 0; i < NV * 4; ++i) { dl[i] *= act_deriv_from_value(M.act, a[i]); accC[i] += dl[i]; }
#pragma unroll
      for (int v = 0; v < NV; ++v)
        reinterpret_cast<float4*>(D + r * IN + v * 128)[lane] = make_float4(dl[4 * v], dl[4 * v + 1], dl[4 * v + 2], dl[4 * v + 3]);
    }
  }
  // cross-warp sums (warp order), then this CTA's partials
  constexpr int SW = IN * KP + IN + KP + 1;
  float* my = scr + warp * SW;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = v * 128 + lane * 4 + e;
#pragma unroll
      for (int k = 0; k < KP; ++k) my[col * KP + k] = accW[4 * v + e][k];
      my[IN * KP + col] = accC[4 * v + e];
    }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < KP; ++k) my[IN * KP + IN + k] = accB[k];
    my[IN * KP + IN + KP] = ll_acc;
  }
  __syncthreads();
  const long pidx = (long)c * nblk + blockIdx.x;
  for (int e = tid; e < SW; e += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += scr[w * SW + e];
    if (e < IN * KP) { const int i = e / KP, k = e % KP; if (k < K) g.pW[pidx * (IN * K) + i * K + k] = s; }
    else if (e < IN * KP + IN) g.pcol[pidx * IN + (e - IN * KP)] = s;
    else if (e < IN * KP + IN + KP) { const int k = e - IN * KP - IN; if (k < K) g.pb[pidx * K + k] = s; }
    else g.pll[pidx] = s;
  }
}

// =====================================================================================================
// tcgen05 / TMEM GEMM core for the large contractions of the wide path (K >= 32, N >= 64): 3xTF32 split
// (a = a_hi + a_lo, b = b_hi + b_lo;  a b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32 accumulation in TMEM) so
// the result keeps fp32-level accuracy (the parity bar is 1e-5).  Same GemmArgs / epilogues as the SIMT kernel.
// One CTA computes a 128 x 256 output tile: all threads stage the operand k-blocks into shared memory in the
// canonical K-major, no-swizzle UMMA layout (core matrix = 8 rows x 16 B; LBO = 128 B between the K halves,
// SBO = 1024 B between 8-row groups), one elected thread issues tcgen05.mma kind::tf32 (M=128, N=256, K=8),
// completion is tracked with tcgen05.commit -> mbarrier, and warps 0-3 drain the accumulator with tcgen05.ld.
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor / InstrDescriptor).
// =====================================================================================================
#define TC_BM 128
#define TC_BN 256
#define TC_BK 32
#define TC_SMEM_BYTES ((2 * TC_BM * TC_BK + 2 * TC_BN * TC_BK) * 4 + 64)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
  // start address >> 4 | LBO (128 B) >> 4 << 16 | SBO (1024 B) >> 4 << 32 | version 1 << 46 | SWIZZLE_NONE
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}

// canonical K-major / no-swizzle byte offset of element (row r, k) inside a [rows x TC_BK] tf32 tile
__device__ __forceinline__ uint32_t umma_off(int r, int k) {
  return (uint32_t)((((r & 7) + (r >> 3) * 64 + (k >> 2) * 8) << 4) + ((k & 3) << 2));
}

// bounded wait on the MMA-completion mbarrier: a wrong descriptor must end in a trap, never in a hung GPU
__device__ __forceinline__ void tc_wait(uint64_t* mbar, uint32_t parity) {
  uint32_t done = 0;
  for (long spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    if (spin > (1L << 24)) __trap();
  }
}

// Staging of a [ROWS x TC_BK] operand k-block in two halves so that the global loads of block i+1 overlap the
// MMAs of block i: tc_load_tile pulls the block into registers, tc_store_tile splits it into tf32 hi / lo parts
// and writes the canonical layout.  Every warp store instruction fills exactly one 128-byte core matrix
// (8 rows x 16 B): conflict-free in shared memory, and the global side reads whole 32-byte sectors whichever
// dimension of the operand is contiguous.
template <int ROWS>
struct TcRegs { float v[ROWS * TC_BK / 256]; };

template <int ROWS>
__device__ __forceinline__ bool tc_vec_ok(const float* src, long s_row, long s_k) {
  return s_k == 1 && (s_row & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
}

template <int ROWS>
__device__ __forceinline__ void tc_load_tile(TcRegs<ROWS>& R, const float* __restrict__ src, long s_row, long s_k, int row0,
                                             int rows_valid, int k0, int k_end, bool vec) {
  if (vec) {   // k contiguous and 16-byte aligned: float4 = 4 k's of one row
#pragma unroll
    for (int it = 0; it < ROWS * (TC_BK / 4) / 256; ++it) {
      const int idx = threadIdx.x + it * 256;
      const int r8 = idx & 7, kq = (idx >> 3) & 7, rg = idx >> 6;
      const int gr = row0 + rg * 8 + r8, gk = k0 + kq * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rows_valid) {
        if (gk + 3 < k_end) v = __ldg(reinterpret_cast<const float4*>(src + (long)gr * s_row + gk));
        else {
          if (gk < k_end) v.x = __ldg(src + (long)gr * s_row + gk);
          if (gk + 1 < k_end) v.y = __ldg(src + (long)gr * s_row + gk + 1);
          if (gk + 2 < k_end) v.z = __ldg(src + (long)gr * s_row + gk + 2);
        }
      }
      R.v[it * 4] = v.x; R.v[it * 4 + 1] = v.y; R.v[it * 4 + 2] = v.z; R.v[it * 4 + 3] = v.w;
    }
  } else {     // one core matrix (8 rows x 4 k) per warp instruction; the contiguous source dimension varies fastest
    const bool kfast = s_k == 1;
    const int lane5 = threadIdx.x & 31;
    const int r8 = kfast ? (lane5 >> 2) : (lane5 & 7), kk = kfast ? (lane5 & 3) : (lane5 >> 3);
#pragma unroll
    for (int u = 0; u < ROWS * TC_BK / 256; ++u) {
      const int cm = (threadIdx.x >> 5) + u * 8;
      const int kq = cm & 7, rg = cm >> 3;
      const int gr = row0 + rg * 8 + r8, gk = k0 + kq * 4 + kk;
      R.v[u] = (gr < rows_valid && gk < k_end) ? __ldg(src + (long)gr * s_row + (long)gk * s_k) : 0.f;
    }
  }
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }  // what the tf32 datapath keeps

template <int ROWS>
__device__ __forceinline__ void tc_store_tile(const TcRegs<ROWS>& R, char* hi, char* lo, bool kfast, bool vec) {
  if (vec) {
#pragma unroll
    for (int it = 0; it < ROWS * (TC_BK / 4) / 256; ++it) {
      const int idx = threadIdx.x + it * 256;
      const int r8 = idx & 7, kq = (idx >> 3) & 7, rg = idx >> 6;
      const float4 v = make_float4(R.v[it * 4], R.v[it * 4 + 1], R.v[it * 4 + 2], R.v[it * 4 + 3]);
      const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      const uint32_t o = (uint32_t)((r8 + rg * 64 + kq * 8) << 4);
      *reinterpret_cast<float4*>(hi + o) = h;
      *reinterpret_cast<float4*>(lo + o) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    }
  } else {
    const int lane5 = threadIdx.x & 31;
    const int r8 = kfast ? (lane5 >> 2) : (lane5 & 7), kk = kfast ? (lane5 & 3) : (lane5 >> 3);
#pragma unroll
    for (int u = 0; u < ROWS * TC_BK / 256; ++u) {
      const int cm = (threadIdx.x >> 5) + u * 8;
      const int kq = cm & 7, rg = cm >> 3;
      const float h = tf32_hi(R.v[u]);
      const uint32_t o = umma_off(rg * 8 + r8, kq * 4 + kk);
      *reinterpret_cast<float*>(hi + o) = h;
      *reinterpret_cast<float*>(lo + o) = R.v[u] - h;
    }
  }
}

__global__ void __launch_bounds__(256, 2) wide_gemm_tc_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) char tsm[];
  char* a_hi = tsm;
  char* a_lo = a_hi + TC_BM * TC_BK * 4;
  char* b_hi = a_lo + TC_BM * TC_BK * 4;
  char* b_lo = b_hi + TC_BN * TC_BK * 4;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + TC_BN * TC_BK * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z / g.kslices, ks = blockIdx.z % g.kslices;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  const int kper = ((g.K + g.kslices - 1) / g.kslices + TC_BK - 1) / TC_BK * TC_BK;
  const int kbeg = ks * kper, kend = min(g.K, kbeg + kper);
  const float* A = g.A + (long)b * g.a_batch;
  const float* B = g.B + (long)b * g.b_batch;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(TC_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  uint32_t parity = 0;
  bool first = true;
  const bool a_vec = tc_vec_ok<TC_BM>(A, g.sam, g.sak), b_vec = tc_vec_ok<TC_BN>(B, g.sbn, g.sbk);
  const bool a_kfast = g.sak == 1, b_kfast = g.sbk == 1;
  TcRegs<TC_BM> ra;
  TcRegs<TC_BN> rb;
  tc_load_tile<TC_BM>(ra, A, g.sam, g.sak, m0, g.M, kbeg, kend, a_vec);
  tc_load_tile<TC_BN>(rb, B, g.sbn, g.sbk, n0, g.N, kbeg, kend, b_vec);
  for (int k0 = kbeg; k0 < kend; k0 += TC_BK) {
    if (!first) {   // the previous k-block's MMAs must have consumed the shared tiles
      tc_wait(mbar, parity);
      parity ^= 1;
    }
    tc_store_tile<TC_BM>(ra, a_hi, a_lo, a_kfast, a_vec);
    tc_store_tile<TC_BN>(rb, b_hi, b_lo, b_kfast, b_vec);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
#pragma unroll
      for (int j = 0; j < TC_BK / 8; ++j) {   // one MMA covers K = 8 tf32 = two 16-byte K halves = 256 B further
        const uint32_t ko = j * 256;
        umma_tf32(tmem, umma_desc_kmajor(sa_hi + ko), umma_desc_kmajor(sb_lo + ko), idesc, (first && j == 0) ? 0u : 1u);
        umma_tf32(tmem, umma_desc_kmajor(sa_lo + ko), umma_desc_kmajor(sb_hi + ko), idesc, 1u);
        umma_tf32(tmem, umma_desc_kmajor(sa_hi + ko), umma_desc_kmajor(sb_hi + ko), idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(mbar)) : "memory");
    }
    first = false;
    if (k0 + TC_BK < kend) {   // global loads of the next k-block fly while the tensor core works on this one
      tc_load_tile<TC_BM>(ra, A, g.sam, g.sak, m0, g.M, k0 + TC_BK, kend, a_vec);
      tc_load_tile<TC_BN>(rb, B, g.sbn, g.sbk, n0, g.N, k0 + TC_BK, kend, b_vec);
    }
  }
  tc_wait(mbar, parity);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // Epilogue through shared memory (the operand stages are free now): warps 0-3 drain 32 accumulator columns at a
  // time from TMEM into a [128][36] tile (conflict-free float4 rows), then ALL warps apply bias / activation / act'
  // and write float4s with 8 lanes covering one 128-byte row segment -> fully coalesced global stores and aux loads.
  {
    float* C = g.C + (long)b * g.c_batch + (long)ks * g.c_slice;
    float* Clo = g.C_lo ? g.C_lo + (long)b * g.c_batch + (long)ks * g.c_slice : nullptr;
    const float* bias = g.bias ? g.bias + (long)b * g.bias_batch : nullptr;
    const float* aux = g.aux ? g.aux + (long)b * g.aux_batch : nullptr;
    float* tile = reinterpret_cast<float*>(tsm);          // 2 x [128][36] floats (double buffered)
    constexpr int TS = 36;
#pragma unroll 1
    for (int c0 = 0; c0 < TC_BN; c0 += 32) {
      float* tb = tile + ((c0 >> 5) & 1) * (TC_BM * TS);
      if (warp < 4) {   // warp w owns TMEM lanes [32w, 32w+32) == tile rows 32w + lane
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
            "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        float* trow = tb + (warp * 32 + lane) * TS;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(trow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                             __uint_as_float(r[j + 3]));
      }
      __syncthreads();
#pragma unroll 1   // keep the epilogue compact: the activation switch unrolled 8x thrashes the instruction cache
      for (int it = 0; it < (TC_BM * 8) / 256; ++it) {
        const int idx = tid + it * 256, row = idx >> 3, cq = idx & 7;
        const int m = m0 + row, n = n0 + c0 + cq * 4;
        if (m >= g.M || n >= g.N) continue;
        const float4 t4 = *reinterpret_cast<const float4*>(tb + row * TS + cq * 4);
        float o[4] = {t4.x, t4.y, t4.z, t4.w};
        float* crow = C + (long)m * g.ldc + n;
        const bool full4 = n + 3 < g.N, al = full4 && (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (n + e < g.N) {
            if (g.epi == 1) o[e] = act_value(g.act, o[e] + bias[n + e]);
            else if (g.epi == 2) o[e] = o[e] + bias[n + e];
            else if (g.epi == 3) o[e] = o[e] * act_deriv_from_value(g.act, aux[(long)m * g.ldaux + n + e]);
          }
        }
        if (al) {
          *reinterpret_cast<float4*>(crow) = make_float4(o[0], o[1], o[2], o[3]);
          if (Clo) *reinterpret_cast<float4*>(Clo + (long)m * g.ldc + n) =
              make_float4(o[0] - tf32_hi(o[0]), o[1] - tf32_hi(o[1]), o[2] - tf32_hi(o[2]), o[3] - tf32_hi(o[3]));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < g.N) { crow[e] = o[e]; if (Clo) Clo[(long)m * g.ldc + n + e] = o[e] - tf32_hi(o[e]); }
        }
      }
      // (double buffered: the tile written two chunks ago is only overwritten after the next __syncthreads)
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TC_BN));
}

// =====================================================================================================
// tcgen05 GEMM, version 2: TMA-fed, warp-specialised, mbarrier pipeline (K-major operands only).
//   warp 0 (one lane)  TMA producer: A, B, B_lo boxes of one k-block per stage -> full barrier (expect_tx)
//   warps 2-5          converters: A_lo = A - tf32(A) computed in shared memory (element-wise, so the 128-byte swizzle
//                      is irrelevant), fence.proxy.async, arrive on the stage's "converted" barrier; after the main
//                      loop the same warps run the epilogue (tcgen05.ld -> per-warp smem tile -> coalesced stores)
//   warp 1 (one lane)  MMA issuer: per 8-wide k-step three tcgen05.mma kind::tf32 M128 N256 K8 (a_hi*b_lo, a_lo*b_hi,
//                      a_hi*b_hi; the tf32 datapath truncates, so v itself serves as "hi"); tcgen05.commit -> empty
// Activations / deltas therefore live in HBM ONCE (no remainder tensors): the minimum traffic of a launch is
// A + C.  Only the (small, L2-resident) weights carry a pre-packed remainder copy.
// Operands are K-major [rows x 32 fp32] tiles = one 128-byte swizzle atom per row: 3-D tensor maps {K, rows, batch},
// box {32, rows, 1}, CU_TENSOR_MAP_SWIZZLE_128B; UMMA descriptors SWIZZLE_128B, SBO = 1024 B, k-step = +32 B in the atom.
// (MN-major tf32 operands did not give correct results with the no-swizzle descriptors, so the dW GEMMs
// (K = data rows) stay on the v1 core, which transposes while staging.)
// =====================================================================================================
#include <cuda.h>

#define T2_BM 128
#define T2_BN 256
#define T2_BK 32
#define T2_STAGES 2
#define T2_A_BYTES (T2_BM * T2_BK * 4)
#define T2_B_BYTES (T2_BN * T2_BK * 4)
#define T2_STAGE_BYTES (2 * T2_A_BYTES + 2 * T2_B_BYTES)
#define T2_TS 20   // epilogue tile row stride (floats): 16 columns + 4 pad, float4-aligned, conflict-free row writes
#define T2_SMEM_BYTES (T2_STAGES * T2_STAGE_BYTES + 8 * 32 * T2_TS * 4 + 256)

struct Tc2Args {
  CUtensorMap a_hi, b_hi, b_lo;
  int M, N, K, kslices, nbatch;
  float* C; long c_batch, c_slice, ldc;
  int epi, act;
  const float* bias; long bias_batch;
  const float* aux; long aux_batch, ldaux;
  // EPI 3 only (optional): column sums of the produced delta tile = partial bias gradient of the layer below, one partial
  // per (batch, m-block, 32-row group): csum[((b * nmb + mb) * 4 + q) * N + n].  Fused here, the bias gradients cost no
  // extra pass over the 100 MB delta matrices (wide_rowreduce_kernel: 21 us + a slice reduction per layer before).
  float* csum;
};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (long spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1L << 24)) __trap();   // a wrong descriptor must end in a trap, never in a hung GPU
  }
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// L2 prefetch of a TMA box (no shared-memory destination): issued a tile / several k-blocks ahead by the producer so that
// the later cp.async.bulk.tensor of the same box is an L2 hit.  With two 96 KB stages the producer can run only ONE
// k-block (1536 tensor-pipe cycles) ahead of the MMAs -- less than an HBM round trip plus the conversion pass, which left
// the tensor pipe idle ~40 % of every k-block (profiles/r1i_ncu_wide_tc2.txt: 58 % active).  OUTCOME: no gain (72 -> 72 us,
// profiles/r2x_launches_wide_4x256.csv), removed from the kernels again; the helper stays for the record.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// K-major SWIZZLE_128B descriptor: LBO unused (1), SBO = 1024 B between 8-row groups, layout type 2
// MN-major tf32 operands: the only shared-memory layout the tensor core accepts is "128-byte swizzle with 32-byte atoms"
// (UMMA layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): canonical ((4,8,m),(4,k)):((1,4,LBO),(32,SBO)) in
// elements.  The operand tile is a row of [32 K-rows x 128 B] TMA boxes, one per 32 MN-elements: LBO = 4096 B between
// boxes, SBO = 512 B between the 4-row K-groups inside a box.  One K = 8 instruction consumes two K-groups: k-step = +1024 B.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

#define T2_EPI_WARPS 8
#define T2_CVT_WARPS 4
#define T2_THREADS (64 + 32 * T2_CVT_WARPS + 32 * T2_EPI_WARPS)

// One output element of the epilogue.  EPI: 0 plain, 1 act(v + bias), 2 v + bias, 3 v * act'(aux).  RELU = compile-time
// fast path; otherwise the runtime activation switch.
template <int EPI, bool RELU>
__device__ __forceinline__ float t2_epi(float v, float bias, float aux, int act) {
  if (EPI == 1) { v += bias; return RELU ? fmaxf(v, 0.f) : act_value(act, v); }
  if (EPI == 2) return v + bias;
  if (EPI == 3) return RELU ? (aux > 0.f ? v : 0.f) : v * act_deriv_from_value(act, aux);
  return v;
}

// Epilogue of one 128 x 256 accumulator tile for one epilogue warp (TMEM lane quarter q, column half `half`): 8 passes of
// 16 columns, tcgen05.ld -> per-warp shared tile (transposition: a lane holds a row, the stores want a lane per column
// group) -> bias / activation / act' -> 64-byte row segments.  (A software-pipelined variant -- bias slab hoisted, aux
// and tcgen05.ld of pass cc+1 in flight during pass cc -- measured 10 % SLOWER on the act' flavour and equal on the
// others, profiles/r3c_launches_wide_4x256.csv: the epilogue is not on the critical path.)
#define T2_LD16(r, taddr)                                                                                                  \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                        \
      : "r"(taddr))

template <int EPI, bool RELU, class WaitFn>
__device__ __forceinline__ void t2_epilogue_tile(const Tc2Args& g, uint32_t tmem, int acc, float* tile, int q, int half, int lane,
                                                 int pb, int pks, int pm0, int pn0, int nmb, bool have_k, WaitFn wait_full) {
  constexpr int TS = T2_TS, NP = T2_BN / 32;
  float* C = g.C + (long)pb * g.c_batch + (long)pks * g.c_slice;
  const float* bias = g.bias ? g.bias + (long)pb * g.bias_batch : nullptr;
  const float* aux = g.aux ? g.aux + (long)pb * g.aux_batch : nullptr;
  const bool fast = (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && pn0 + T2_BN <= g.N &&
                    (EPI != 3 || ((g.ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0));
  const int cq = lane & 3, rsub = lane >> 2;
  const int nbase = pn0 + half * (T2_BN / 2) + cq * 4;
  const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * T2_BN + half * (T2_BN / 2));
  if (fast) {
    wait_full();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int cc = 0; cc < NP; ++cc) {     // 16 accumulator columns per pass
      uint32_t r[16];
      T2_LD16(r, tbase + (uint32_t)(cc * 16));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __syncwarp();   // the previous pass has been read out of the tile
      float* trow = tile + lane * TS;
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(trow + j) = have_k ? make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                     __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      // 4 lanes cover one 64-byte row segment: sector-aligned stores and aux loads, no per-element branches
      const int n = nbase + cc * 16;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (EPI == 1 || EPI == 2) b4 = make_float4(__ldg(bias + n), __ldg(bias + n + 1), __ldg(bias + n + 2), __ldg(bias + n + 3));   // (theta offsets are not 16-byte aligned)
      float4 x4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = pm0 + q * 32 + i * 8 + rsub;
        x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI == 3 && m < g.M) x4[i] = *reinterpret_cast<const float4*>(aux + (long)m * g.ldaux + n);
      }
      float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = i * 8 + rsub, m = pm0 + q * 32 + row;
        const float4 t4 = *reinterpret_cast<const float4*>(tile + row * TS + cq * 4);
        float4 o4;
        o4.x = t2_epi<EPI, RELU>(t4.x, b4.x, x4[i].x, g.act); o4.y = t2_epi<EPI, RELU>(t4.y, b4.y, x4[i].y, g.act);
        o4.z = t2_epi<EPI, RELU>(t4.z, b4.z, x4[i].z, g.act); o4.w = t2_epi<EPI, RELU>(t4.w, b4.w, x4[i].w, g.act);
        if (m < g.M) {
          *reinterpret_cast<float4*>(C + (long)m * g.ldc + n) = o4;
          if (EPI == 3) { cs.x += o4.x; cs.y += o4.y; cs.z += o4.z; cs.w += o4.w; }
        }
      }
      if (EPI == 3 && g.csum) {   // column sums over the 32 rows of this warp's lane quarter (fixed shuffle order)
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
          cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
          cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
        }
        if (lane < 4 && pm0 < g.M)
          *reinterpret_cast<float4*>(g.csum + (((long)pb * nmb + pm0 / T2_BM) * 4 + q) * g.N + n) = cs;
      }
    }
  } else {
    wait_full();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int cc = 0; cc < NP; ++cc) {
      uint32_t r[16];
      T2_LD16(r, tbase + (uint32_t)(cc * 16));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __syncwarp();
      float* trow = tile + lane * TS;
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(trow + j) = have_k ? make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                     __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      const int n = nbase + cc * 16;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int row = i * 8 + rsub, m = pm0 + q * 32 + row;
        if (m >= g.M) continue;
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
          if (n + e >= g.N) break;
          const float v = tile[row * TS + cq * 4 + e];
          C[(long)m * g.ldc + n + e] = t2_epi<EPI, RELU>(v, (EPI == 1 || EPI == 2) ? bias[n + e] : 0.f,
                                                         EPI == 3 ? aux[(long)m * g.ldaux + n + e] : 0.f, g.act);
        }
      }
    }
  }
}

// Persistent: CTA i works on tiles i, i + gridDim.x, ... (tile = (batch*kslice, m-block); N <= 256 is one n-block per
// tile row, larger N adds n-blocks).  Two TMEM accumulators (2 x 256 columns) let the epilogue of tile t overlap the
// main loop of tile t+1:
//   tmem_full[a]  MMA -> epilogue  (accumulator a complete)
//   tmem_empty[a] epilogue -> MMA  (accumulator a drained; count = T2_EPI_WARPS)
// MN = true: both operands are MN-major in global memory (A(m,k) at A[k*lda + m], B(n,k) at B[k*ldb + n]) -- the weight
// gradients dW = a^T delta with K = data rows.  Each operand tile is then a row of {32 MN-floats x 32 K-rows} TMA boxes,
// neither operand has a pre-packed remainder, and the converter warps produce A_lo and B_lo.
template <int EPI, bool RELU, bool MN = false>
__global__ void __launch_bounds__(T2_THREADS, 1) wide_gemm_tc2_kernel(const __grid_constant__ Tc2Args g) {
  extern __shared__ __align__(1024) char sm2[];
  const uint32_t sbase = smem_u32(sm2);
  // stage memory, then the epilogue tiles (T2_EPI_WARPS x [32][36] floats), then the barriers
  constexpr int TS = T2_TS;
  float* tiles = reinterpret_cast<float*>(sm2 + T2_STAGES * T2_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm2 + T2_STAGES * T2_STAGE_BYTES + T2_EPI_WARPS * 32 * TS * 4);
  // full[s] = bar0 + 8 s, empty[s] = +8 (S + s), converted[s] = +8 (2S + s), tmem_full[a] = +8 (3S + a), tmem_empty[a] = +8 (3S + 2 + a)
  const uint32_t bar0 = smem_u32(bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * T2_STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nmb = (g.M + T2_BM - 1) / T2_BM, nnb = (g.N + T2_BN - 1) / T2_BN;
  const int ntiles = nmb * nnb * g.nbatch * g.kslices;
  const int kper = ((g.K + g.kslices - 1) / g.kslices + T2_BK - 1) / T2_BK * T2_BK;
  if (tid == 0) {
    for (int s = 0; s < 3 * T2_STAGES + 4; ++s) {
      const int cnt = (s >= 2 * T2_STAGES && s < 3 * T2_STAGES) ? T2_CVT_WARPS : (s >= 3 * T2_STAGES + 2 ? T2_EPI_WARPS : 1);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * s), "r"(cnt));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * T2_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // tile index -> (batch b, k-slice ks, m0, n0, k range)
  auto tile_coords = [&](int t, int& b, int& ks, int& m0, int& n0, int& kbeg, int& nkb) {
    const int nb = t % nnb; t /= nnb;
    const int mb = t % nmb; t /= nmb;
    ks = t % g.kslices; b = t / g.kslices;
    m0 = mb * T2_BM; n0 = nb * T2_BN;
    kbeg = ks * kper;
    const int kend = min(g.K, kbeg + kper);
    nkb = kend > kbeg ? (kend - kbeg + T2_BK - 1) / T2_BK : 0;
  };
  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer ----
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int b, ks, m0, n0, kbeg, nkb;
        tile_coords(t, b, ks, m0, n0, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % T2_STAGES, ph = (it / T2_STAGES) & 1;
          mbar_wait(bar0 + 8 * (T2_STAGES + s), ph ^ 1);
          const uint32_t full = bar0 + 8 * s, st = sbase + s * T2_STAGE_BYTES;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full),
                       "r"((uint32_t)(MN ? T2_A_BYTES + T2_B_BYTES : T2_A_BYTES + 2 * T2_B_BYTES)) : "memory");
          const int k0 = kbeg + kb * T2_BK;
          if (MN) {
#pragma unroll
            for (int j = 0; j < T2_BM / 32; ++j) tma_load_3d(st + j * 4096, &g.a_hi, m0 + 32 * j, k0, b, full);
#pragma unroll
            for (int j = 0; j < T2_BN / 32; ++j) tma_load_3d(st + 2 * T2_A_BYTES + j * 4096, &g.b_hi, n0 + 32 * j, k0, b, full);
          } else {
            tma_load_3d(st, &g.a_hi, k0, m0, b, full);
            tma_load_3d(st + 2 * T2_A_BYTES, &g.b_hi, k0, n0, b, full);
            tma_load_3d(st + 2 * T2_A_BYTES + T2_B_BYTES, &g.b_lo, k0, n0, b, full);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---- MMA issuer ----
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24) |
                             (MN ? ((1u << 15) | (1u << 16)) : 0u);   // bits 15 / 16: A / B are MN-major
      int it = 0, tl = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
        int b, ks, m0, n0, kbeg, nkb;
        tile_coords(t, b, ks, m0, n0, kbeg, nkb);
        const int acc = tl & 1, aph = (tl >> 1) & 1;
        mbar_wait(bar0 + 8 * (3 * T2_STAGES + 2 + acc), aph ^ 1);      // accumulator drained by the epilogue of tile tl-2
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem + (uint32_t)(acc * T2_BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % T2_STAGES, ph = (it / T2_STAGES) & 1;
          mbar_wait(bar0 + 8 * (2 * T2_STAGES + s), ph);     // converted => the TMA boxes have landed too
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = sbase + s * T2_STAGE_BYTES;
          const uint32_t ahi = st, alo = st + T2_A_BYTES, bhi = st + 2 * T2_A_BYTES, blo = bhi + T2_B_BYTES;
#pragma unroll
          for (int j = 0; j < T2_BK / 8; ++j) {   // K = 8 tf32 = 32 B further inside the 128-byte swizzle atom
            const uint32_t acc0 = (kb == 0 && j == 0) ? 0u : 1u;
            if (MN) {
              const uint32_t o = j * 1024;   // next 8-row K-group of every box
              umma_tf32(tacc, umma_desc_mn_sw128(ahi + o), umma_desc_mn_sw128(blo + o), idesc, acc0);
              umma_tf32(tacc, umma_desc_mn_sw128(alo + o), umma_desc_mn_sw128(bhi + o), idesc, 1u);
              umma_tf32(tacc, umma_desc_mn_sw128(ahi + o), umma_desc_mn_sw128(bhi + o), idesc, 1u);
            } else {
              const uint32_t o = j * 32;
              umma_tf32(tacc, umma_desc_sw128(ahi + o), umma_desc_sw128(blo + o), idesc, acc0);
              umma_tf32(tacc, umma_desc_sw128(alo + o), umma_desc_sw128(bhi + o), idesc, 1u);
              umma_tf32(tacc, umma_desc_sw128(ahi + o), umma_desc_sw128(bhi + o), idesc, 1u);
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * (T2_STAGES + s)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * (3 * T2_STAGES + acc)) : "memory");
      }
    }
  } else if (warp < 2 + T2_CVT_WARPS) {
    // ---- converter warps: remainder tile of A for every stage ----
    const int ct = tid - 64;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int b, ks, m0, n0, kbeg, nkb;
      tile_coords(t, b, ks, m0, n0, kbeg, nkb);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % T2_STAGES, ph = (it / T2_STAGES) & 1;
        mbar_wait(bar0 + 8 * s, ph);
        const uint32_t ahi = sbase + s * T2_STAGE_BYTES, alo = ahi + T2_A_BYTES;
#pragma unroll
        for (int i = 0; i < T2_A_BYTES / 16 / (32 * T2_CVT_WARPS); ++i) {
          const uint32_t off = (uint32_t)(ct + i * 32 * T2_CVT_WARPS) * 16u;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ahi + off));
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(alo + off), "f"(v.x - tf32_hi(v.x)), "f"(v.y - tf32_hi(v.y)),
                       "f"(v.z - tf32_hi(v.z)), "f"(v.w - tf32_hi(v.w)) : "memory");
        }
        if (MN) {   // no pre-packed remainder for B either
          const uint32_t bhi = ahi + 2 * T2_A_BYTES, blo = bhi + T2_B_BYTES;
#pragma unroll
          for (int i = 0; i < T2_B_BYTES / 16 / (32 * T2_CVT_WARPS); ++i) {
            const uint32_t off = (uint32_t)(ct + i * 32 * T2_CVT_WARPS) * 16u;
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(bhi + off));
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blo + off), "f"(v.x - tf32_hi(v.x)), "f"(v.y - tf32_hi(v.y)),
                         "f"(v.z - tf32_hi(v.z)), "f"(v.w - tf32_hi(v.w)) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the async proxy
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * (2 * T2_STAGES + s)) : "memory");
      }
    }
  } else {
    // ---- epilogue warps: drain accumulator (tile & 1) while the main loop of the next tile runs ----
    const int ew = warp - 2 - T2_CVT_WARPS;
    const int q = warp & 3, half = ew >> 2;          // TMEM lane quarter (fixed by warp id) and column half of this warp
    float* tile = tiles + ew * (32 * TS);
    int tl = 1;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      int pb, pks, pm0, pn0, pkbeg, pnkb;
      tile_coords(t, pb, pks, pm0, pn0, pkbeg, pnkb);
      {
        const int el = tl - 1, acc = el & 1, aph = (el >> 1) & 1;
        mbar_wait(bar0 + 8 * (3 * T2_STAGES + acc), aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* C = g.C + (long)pb * g.c_batch + (long)pks * g.c_slice;
        const float* bias = g.bias ? g.bias + (long)pb * g.bias_batch : nullptr;
        const float* aux = g.aux ? g.aux + (long)pb * g.aux_batch : nullptr;
        const bool fast = (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && pn0 + T2_BN <= g.N &&
                          (EPI != 3 || ((g.ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0));
#pragma unroll 1
        for (int cc = 0; cc < T2_BN / 32; ++cc) {     // 16 accumulator columns per pass
          const int c0 = half * (T2_BN / 2) + cc * 16;
          uint32_t r[16];
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * T2_BN + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          __syncwarp();   // the previous pass has been read out of the tile
          float* trow = tile + lane * TS;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(trow + j) = pnkb > 0 ? make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                        __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
          __syncwarp();
          const int cq = lane & 3, n = pn0 + c0 + cq * 4;
          if (fast) {   // 4 lanes cover one 64-byte row segment: sector-aligned stores and aux loads, no per-element branches
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI == 1 || EPI == 2) b4 = make_float4(__ldg(bias + n), __ldg(bias + n + 1), __ldg(bias + n + 2), __ldg(bias + n + 3));   // (theta offsets are not 16-byte aligned)
            float4 x4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = pm0 + q * 32 + i * 8 + (lane >> 2);
              x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (EPI == 3 && m < g.M) x4[i] = *reinterpret_cast<const float4*>(aux + (long)m * g.ldaux + n);
            }
            float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = i * 8 + (lane >> 2), m = pm0 + q * 32 + row;
              const float4 t4 = *reinterpret_cast<const float4*>(tile + row * TS + cq * 4);
              float4 o4;
              o4.x = t2_epi<EPI, RELU>(t4.x, b4.x, x4[i].x, g.act); o4.y = t2_epi<EPI, RELU>(t4.y, b4.y, x4[i].y, g.act);
              o4.z = t2_epi<EPI, RELU>(t4.z, b4.z, x4[i].z, g.act); o4.w = t2_epi<EPI, RELU>(t4.w, b4.w, x4[i].w, g.act);
              if (m < g.M) {
                *reinterpret_cast<float4*>(C + (long)m * g.ldc + n) = o4;
                if (EPI == 3) { cs.x += o4.x; cs.y += o4.y; cs.z += o4.z; cs.w += o4.w; }
              }
            }
            if (EPI == 3 && g.csum) {   // column sums over the 32 rows of this warp's lane quarter (fixed shuffle order)
#pragma unroll
              for (int o = 4; o <= 16; o <<= 1) {
                cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
                cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
              }
              if (lane < 4 && pm0 < g.M)
                *reinterpret_cast<float4*>(g.csum + (((long)pb * nmb + pm0 / T2_BM) * 4 + q) * g.N + n) = cs;
            }
          } else {
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int row = i * 8 + (lane >> 2), m = pm0 + q * 32 + row;
              if (m >= g.M) continue;
#pragma unroll 1
              for (int e = 0; e < 4; ++e) {
                if (n + e >= g.N) break;
                const float v = tile[row * TS + cq * 4 + e];
                C[(long)m * g.ldc + n + e] = t2_epi<EPI, RELU>(v, (EPI == 1 || EPI == 2) ? bias[n + e] : 0.f,
                                                               EPI == 3 ? aux[(long)m * g.ldaux + n + e] : 0.f, g.act);
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * (3 * T2_STAGES + 2 + acc)) : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * T2_BN));
}

// =====================================================================================================
// tcgen05 GEMM, version 3: CTA PAIRS (cta_group::2) for the K-major products (forward  act(a W + b)  and backward
// (delta W^T) .* act').  Why: in the one-CTA kernel every tcgen05.mma M128 N256 K8 reads 4 KB of A and 8 KB of B from
// shared memory, three of them per product (3xTF32), and every k-block brings 64 KB of weights (hi + lo) into the SM: per
// k-block 144 KB of operand reads + 80 KB of TMA writes + 32 KB of conversion + the epilogue staging = ~290 KB against
// 128 B/clk -> ~2250 cycles for 1536 cycles of tensor work (measured 2640, tensor pipe 58 % busy,
// profiles/r1i_ncu_wide_tc2.txt; an L2 prefetch of the operands changed nothing, profiles/r2x_*: not latency).
// A pair of SMs shares one 256-row tile: each CTA stages ITS 128 rows of A and ITS half (128 of the 256 output columns)
// of the weights; one tcgen05.mma.cta_group::2 M256 N256 K8 issued by the leader reads A from both and the two weight
// halves from both.  Per SM that halves the weight bytes on every path (TMA writes 48 KB, operand reads 96 KB per
// k-block) and shrinks a stage to 64 KB, so three stages fit.
//   per CTA : warp 0 TMA producer (own A rows, own half of W_hi / W_lo) -> local full[s]
//             warps 2-5 converters (A_lo = A - tf32(A) in shared memory) -> arrive on the LEADER's conv[s]
//             warps 6-13 epilogue of the CTA's own 128 accumulator rows (TMEM lanes) -> arrive on the LEADER's tmem_empty[a]
//   leader  : warp 1 issues the MMAs once conv[s] has both CTAs' arrivals; tcgen05.commit multicasts to empty[s] /
//             tmem_full[a] of both CTAs
// =====================================================================================================
#define X2_STAGES 3
#define X2_B_BYTES (128 * T2_BK * 4)                       // this CTA's half of the weight tile
#define X2_STAGE_BYTES (2 * T2_A_BYTES + 2 * X2_B_BYTES)   // A, A_lo, W_hi half, W_lo half = 64 KB
#define X2_SMEM_BYTES (X2_STAGES * X2_STAGE_BYTES + T2_EPI_WARPS * 32 * T2_TS * 4 + 256)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t local_addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local_addr));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local barrier whose arrivals come from both CTAs of the pair
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (long spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1L << 24)) __trap();
  }
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((unsigned short)3) : "memory");
}

template <int EPI, bool RELU>
__global__ void __launch_bounds__(T2_THREADS, 1) wide_gemm_tc2x_kernel(const __grid_constant__ Tc2Args g) {
  extern __shared__ __align__(1024) char sm2[];
  const uint32_t sbase = smem_u32(sm2);
  constexpr int TS = T2_TS, S = X2_STAGES;
  float* tiles = reinterpret_cast<float*>(sm2 + S * X2_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm2 + S * X2_STAGE_BYTES + T2_EPI_WARPS * 32 * TS * 4);
  // full[s] = bar0 + 8 s, empty[s] = +8 (S + s), conv[s] = +8 (2S + s), tmem_full[a] = +8 (3S + a), tmem_empty[a] = +8 (3S + 2 + a)
  const uint32_t bar0 = smem_u32(bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nmb2 = (g.M + 2 * T2_BM - 1) / (2 * T2_BM), nnb = (g.N + T2_BN - 1) / T2_BN;
  const int nmb = (g.M + T2_BM - 1) / T2_BM;
  const int ntiles = nmb2 * nnb * g.nbatch;
  const int nkb = (g.K + T2_BK - 1) / T2_BK;
  if (tid == 0) {
    for (int s = 0; s < 3 * S + 4; ++s) {
      const int cnt = (s >= 2 * S && s < 3 * S) ? 2 * T2_CVT_WARPS : (s >= 3 * S + 2 ? 2 * T2_EPI_WARPS : 1);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * s), "r"(cnt));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * T2_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // both CTAs' barriers are initialised before anybody arrives remotely
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t lead0 = mapa_rank0(bar0);      // the leader's barrier block in the shared::cluster window
  auto tile_coords = [&](int t, int& b, int& m0, int& n0) {
    const int nb = t % nnb; t /= nnb;
    const int mb2 = t % nmb2; b = t / nmb2;
    m0 = mb2 * 2 * T2_BM + rank * T2_BM; n0 = nb * T2_BN;
  };
  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer: this CTA's rows of A, this CTA's half of the weights ----
      int it = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        int b, m0, n0;
        tile_coords(t, b, m0, n0);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S, ph = (it / S) & 1;
          mbar_wait_cluster(bar0 + 8 * (S + s), ph ^ 1);
          const uint32_t full = bar0 + 8 * s, st = sbase + s * X2_STAGE_BYTES;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((uint32_t)(T2_A_BYTES + 2 * X2_B_BYTES)) : "memory");
          const int k0 = kb * T2_BK;
          tma_load_3d(st, &g.a_hi, k0, m0, b, full);
          tma_load_3d(st + 2 * T2_A_BYTES, &g.b_hi, k0, n0 + rank * 128, b, full);
          tma_load_3d(st + 2 * T2_A_BYTES + X2_B_BYTES, &g.b_lo, k0, n0 + rank * 128, b, full);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {   // ---- MMA issuer (leader CTA only) ----
      // D = F32, A = B = TF32, both K-major, N = 256, M = 256 (128 rows per CTA)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)((2 * T2_BM) >> 4) << 24);
      int it = 0, tl = 0;
      for (int t = pair; t < ntiles; t += npairs, ++tl) {
        const int acc = tl & 1, aph = (tl >> 1) & 1;
        mbar_wait_cluster(bar0 + 8 * (3 * S + 2 + acc), aph ^ 1);      // accumulator drained by BOTH epilogues of tile tl-2
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem + (uint32_t)(acc * T2_BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S, ph = (it / S) & 1;
          mbar_wait_cluster(bar0 + 8 * (2 * S + s), ph);     // both CTAs: boxes landed and remainder tile written
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = sbase + s * X2_STAGE_BYTES;
          const uint32_t ahi = st, alo = st + T2_A_BYTES, bhi = st + 2 * T2_A_BYTES, blo = bhi + X2_B_BYTES;
#pragma unroll
          for (int j = 0; j < T2_BK / 8; ++j) {
            const uint32_t acc0 = (kb == 0 && j == 0) ? 0u : 1u, o = j * 32;
            umma_tf32_2cta(tacc, umma_desc_sw128(ahi + o), umma_desc_sw128(blo + o), idesc, acc0);
            umma_tf32_2cta(tacc, umma_desc_sw128(alo + o), umma_desc_sw128(bhi + o), idesc, 1u);
            umma_tf32_2cta(tacc, umma_desc_sw128(ahi + o), umma_desc_sw128(bhi + o), idesc, 1u);
          }
          umma_commit_2cta(bar0 + 8 * (S + s));
        }
        umma_commit_2cta(bar0 + 8 * (3 * S + acc));
      }
    }
  } else if (warp < 2 + T2_CVT_WARPS) {
    // ---- converter warps: remainder tile of this CTA's A rows, then tell the leader ----
    const int ct = tid - 64;
    int it = 0;
    for (int t = pair; t < ntiles; t += npairs) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % S, ph = (it / S) & 1;
        mbar_wait(bar0 + 8 * s, ph);
        const uint32_t ahi = sbase + s * X2_STAGE_BYTES, alo = ahi + T2_A_BYTES;
#pragma unroll
        for (int i = 0; i < T2_A_BYTES / 16 / (32 * T2_CVT_WARPS); ++i) {
          const uint32_t off = (uint32_t)(ct + i * 32 * T2_CVT_WARPS) * 16u;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ahi + off));
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(alo + off), "f"(v.x - tf32_hi(v.x)), "f"(v.y - tf32_hi(v.y)),
                       "f"(v.z - tf32_hi(v.z)), "f"(v.w - tf32_hi(v.w)) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead0 + 8 * (2 * S + s));
      }
    }
  } else {
    // ---- epilogue warps: this CTA's 128 accumulator rows of tile tl while the main loop of tile tl+1 runs ----
    const int ew = warp - 2 - T2_CVT_WARPS;
    const int q = warp & 3, half = ew >> 2;
    float* tile = tiles + ew * (32 * TS);
    int tl = 0;
    for (int t = pair; t < ntiles; t += npairs, ++tl) {
      int pb, pm0, pn0;
      tile_coords(t, pb, pm0, pn0);
      const int acc = tl & 1, aph = (tl >> 1) & 1;
      t2_epilogue_tile<EPI, RELU>(g, tmem, acc, tile, q, half, lane, pb, 0, pm0, pn0, nmb, true,
                                  [&]() { mbar_wait_cluster(bar0 + 8 * (3 * S + acc), aph); });
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead0 + 8 * (3 * S + 2 + acc));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * T2_BN));
}

// lo = v - tf32(v) of the layer weights into 16-byte aligned packed buffers (theta's own offsets / chain stride are not
// TMA-aligned): wpk_hi[c][off_l + e] = W_l[e], wpk_lo = W_l[e] - tf32(W_l[e]).  One launch for all layers that run on the
// TMA core: grid (x, layer).
struct PackArgs {
  int n_layers, kern_off[13], IN[13], OUT[13];
  long pack_off[13];
};
__global__ void __launch_bounds__(256) wide_pack_weights_kernel(const float* __restrict__ theta, int d, const __grid_constant__ PackArgs P,
                                                                float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ hiT,
                                                                float* __restrict__ loT, long pack_stride) {
  // W [IN x OUT] row-major (K-major B of the backward GEMM) and W^T [OUT x IN] (K-major B of the forward GEMM);
  // 32 x 32 tiles through shared memory so that both orientations are written in 128-byte rows.
  // grid (tiles, chains, layers), 256 threads = 32 x 8
  __shared__ float tl[32][33];
  const int L = blockIdx.z, IN = P.IN[L], OUT = P.OUT[L], b = blockIdx.y;
  const int tj = (OUT + 31) / 32, ti = (IN + 31) / 32;
  if ((int)blockIdx.x >= ti * tj) return;
  const int i0 = (blockIdx.x / tj) * 32, j0 = (blockIdx.x % tj) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* W = theta + (long)b * d + P.kern_off[L];
  const long base = (long)b * pack_stride + P.pack_off[L];
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;
    float v = 0.f;
    if (i < IN && j < OUT) {
      v = W[(long)i * OUT + j];
      hi[base + (long)i * OUT + j] = v;
      lo[base + (long)i * OUT + j] = v - tf32_hi(v);
    }
    tl[r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, i = i0 + tx;
    if (i < IN && j < OUT) {
      const float v = tl[tx][r];
      hiT[base + (long)j * IN + i] = v;
      loT[base + (long)j * IN + i] = v - tf32_hi(v);
    }
  }
}
