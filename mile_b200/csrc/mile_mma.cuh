// mile_mma.cuh -- register-chained tensor evaluator for the narrow regression MLPs (hidden width 16, Gaussian
// head: the airfoil / bikesharing / protein / 1024-chain configs of BASELINE.json).
//
// Why this shape (measured on B200, profiles/r2a_mma_microbench.txt): mma.sync.m16n8k8 tf32 issues once per 8 cycles
// per SM sub-partition (512 FMA/clk/SM = 4x the FFMA rate) with a 20-cycle dependent latency, and its operands are
// plain registers.  A 16-wide layer is two 8-column accumulator tiles, and the C fragment of one layer IS the A
// fragment of the next once the contraction index is relabelled (k = t <-> column 2t, k = t+4 <-> column 2t+1; the
// weight fragments are stored with the same relabelling), so one warp carries 16 rows through every layer forward and
// backward without shared memory, barriers or pipeline fill.  fp32 accuracy comes from the 3xTF32 split
// (a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi, 1e-6 relative).  tcgen05 is the wrong tool here: its smallest tile is
// 128 rows x 16 columns with the accumulator in TMEM, i.e. a shared-memory / TMEM round trip plus an mbarrier wait per
// 16-wide layer (the wide 4x256 path, mile_wide.cuh, is where it pays).
//
// Per 16-row tile and warp:  fwd l=0..NH-1 (MMA)  ->  head 16->2 + Gaussian log-likelihood + delta (FMA, exact fp32)
//   -> for l = NH-1..0:  dW_l^T += delta_l^T a_l  (MMA, K = rows; both operands transposed through a private
//   shared-memory patch)  and  delta_{l-1} = (delta_l W_l^T) * relu'(a_l)  (MMA, register-chained).
// Weight-gradient accumulators stay in registers for all of the warp's tiles; one cross-warp sum per evaluation.
//
// Restates the same arithmetic as grad_eval (mile_kernel.cuh): probabilistic.py:92-138 through basic.py:41-61,
// differentiated by hand.
#pragma once
#include "mile_fast.cuh"

// tf32 split by truncation: the tensor core reads the upper 19 bits of a 32-bit operand register (ptxas itself leaves the
// low 13 bits of a cvt.rna.tf32 result unmasked), so x serves as its own "hi" part and lo = x - trunc(x) is exact in fp32.
// Two instructions per value (LOP3 + FADD) instead of the nine that cvt.rna.tf32.f32 expands to on sm_100a.  Dropped
// terms (lo * lo and the truncation of lo) are <= 2^-20 relative per product.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
__device__ __forceinline__ void hmma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
struct AFrag { uint32_t hi[4], lo[4]; };
// A fragment (rows g / g+8, k = t / t+4) from four fp32 values in fragment order
__device__ __forceinline__ void make_afrag(AFrag& A, float v0, float v1, float v2, float v3) {
  split_tf32(v0, A.hi[0], A.lo[0]); split_tf32(v1, A.hi[1], A.lo[1]);
  split_tf32(v2, A.hi[2], A.lo[2]); split_tf32(v3, A.hi[3], A.lo[3]);
}
// 3xTF32 product, small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const AFrag& A, uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  hmma_tf32(c, A.lo, bh0, bh1);
  hmma_tf32(c, A.hi, bl0, bl1);
  hmma_tf32(c, A.hi, bh0, bh1);
}

template <int NL, int FP, int NT_>
struct MmaGE {
  static constexpr int NT = NT_;
  static constexpr int NW = NT / 32;
  static constexpr int H = 16;
  static constexpr int NH = NL - 1;              // hidden layers (all H wide); layer NH is the 16 -> 2 head
  static constexpr int KS0 = FP / 8;             // k-steps of layer 0
  // ---- weight-fragment image: groups of 4 floats (b0_hi, b1_hi, b0_lo, b1_lo), one group per (k-step, n-tile, lane)
  static constexpr int GRP_F0 = 0;                                  // layer 0 forward: KS0 * 2 * 32 groups
  static constexpr int GRP_F1 = KS0 * 64;                           // layers 1..NH-1 forward: 128 groups each
  static constexpr int GRP_B1 = GRP_F1 + (NH - 1) * 128;            // layers 1..NH-1 backward (W^T as the B operand)
  static constexpr int NGROUPS = GRP_B1 + (NH - 1) * 128;
  static constexpr int IMG_HEAD = NGROUPS * 4;                      // head: W[16][2] row-major, then b[2]
  static constexpr int IMG_FLOATS = IMG_HEAD + 40;
  // ---- per-warp transposition patches for the weight-gradient operands
  static constexpr int PT_S = 20;                                   // delta^T  [16 columns][16 rows + 4]
  static constexpr int PA_S = 24;                                   // a_l      [16 rows][16 columns + 8]
  static constexpr int PATCH_FLOATS = 16 * PT_S + 16 * PA_S;
  // ---- cross-warp reduction scratch (aliases the patches): MMA accumulators in fragment order + the FMA-path sums
  static constexpr int NACC0 = KS0 * 4;                             // dW_0^T: 16 x FP
  static constexpr int NACC = NACC0 + (NH - 1) * 8;                 // + dW_l^T 16 x 16
  static constexpr int NSMALL = 8 + 4 * NH + 2 + 1;                 // head kernel, hidden biases, head bias, log-lik
  static constexpr int SCR_WARP = NACC * 32 + NSMALL * 4;
  // one private region per warp, used first as patches and then as reduction scratch: no cross-warp aliasing, so no
  // block barrier is needed between a warp's last tile and its scratch writes
  static constexpr int RSTRIDE = ((PATCH_FLOATS > SCR_WARP ? PATCH_FLOATS : SCR_WARP) + 3) / 4 * 4;
  static constexpr int REGION_FLOATS = NW * RSTRIDE;
  static constexpr int TILE_FLOATS = IMG_FLOATS + REGION_FLOATS;
  // ---- index maps built once per launch (prepare): image group -> two flat parameter indices; scratch slot -> flat
  static constexpr int AUX_INTS = 2 * NGROUPS + NACC * 32 + NSMALL * 4;
  // largest parameter count this evaluator serves (F <= FP) and the elements one lane of the integrator warp owns
  static constexpr int DMAX = FP * H + H + (NH - 1) * (H * H + H) + 2 * H + 2;
  static constexpr int EMAX = (DMAX + 31) / 32;

  static __device__ __forceinline__ int* aux(Ctx& c) { return reinterpret_cast<int*>(c.aux); }

  static __device__ __forceinline__ void prepare(Ctx& c) {
    const DevModel& M = c.P.M;
    int* imap = aux(c);
    int* gmapA = imap + 2 * NGROUPS;
    int* gmapS = gmapA + NACC * 32;
    const int F = M.dims[0];
    for (int grp = threadIdx.x; grp < NGROUPS; grp += NT) {
      int s0, s1;
      if (grp < GRP_F1) {
        const int ks = grp >> 6, nt = (grp >> 5) & 1, ln = grp & 31, g = ln >> 2, t = ln & 3;
        const int i0 = 8 * ks + t, i1 = i0 + 4, j = 8 * nt + g;
        s0 = i0 < F ? M.kern_off[0] + i0 * H + j : -1;
        s1 = i1 < F ? M.kern_off[0] + i1 * H + j : -1;
      } else {
        const int idx = grp - GRP_F1, which = idx >> 7, rem = idx & 127;
        const int ks = rem >> 6, nt = (rem >> 5) & 1, ln = rem & 31, g = ln >> 2, t = ln & 3;
        if (which < NH - 1) {           // forward layer l: B[k][n] = W_l[in = 8ks + 2t (+1)][out = 8nt + g]
          const int l = which + 1, i0 = 8 * ks + 2 * t, j = 8 * nt + g;
          s0 = M.kern_off[l] + i0 * H + j; s1 = s0 + H;
        } else {                        // backward layer l: B[k][n] = W_l[in = 8nt + g][out = 8ks + 2t (+1)]
          const int l = which - (NH - 1) + 1, j0 = 8 * ks + 2 * t, i = 8 * nt + g;
          s0 = M.kern_off[l] + i * H + j0; s1 = s0 + 1;
        }
      }
      imap[2 * grp] = s0; imap[2 * grp + 1] = s1;
    }
    // accumulator register r of lane ln -> flat gradient index.  dW_l^T C fragment: m = out j = g (+8), n = in i = 8nt + 2t (+1)
    for (int o = threadIdx.x; o < NACC * 32; o += NT) {
      const int r = o >> 5, ln = o & 31, g = ln >> 2, t = ln & 3;
      int l, rr;
      if (r < NACC0) { l = 0; rr = r; } else { l = 1 + (r - NACC0) / 8; rr = (r - NACC0) % 8; }
      const int nt = rr >> 2, cc = rr & 3;
      const int j = g + ((cc >> 1) ? 8 : 0), i = 8 * nt + 2 * t + (cc & 1);
      gmapA[o] = i < M.dims[l] ? M.kern_off[l] + i * H + j : -1;
    }
    for (int o = threadIdx.x; o < NSMALL * 4; o += NT) {
      const int k = o >> 2, t = o & 3;
      int dst;
      if (k < 8) {                      // head kernel: k = (nt*2 + cc)*2 + w  ->  W_head[i = 8nt + 2t + cc][w]
        const int w = k & 1, cc = (k >> 1) & 1, nt = k >> 2;
        dst = M.kern_off[NH] + (8 * nt + 2 * t + cc) * 2 + w;
      } else if (k < 8 + 4 * NH) {      // hidden biases: k' = nt*2 + cc -> b_l[8nt + 2t + cc]
        const int kk = k - 8, l = kk >> 2, q = kk & 3;
        dst = M.bias_off[l] + 8 * (q >> 1) + 2 * t + (q & 1);
      } else if (k < 8 + 4 * NH + 2) {  // head bias (all lanes of a quad hold the same sum: lane t = 0 reports)
        dst = t == 0 ? M.bias_off[NH] + (k - 8 - 4 * NH) : -1;
      } else {                          // log-likelihood partial
        dst = t == 0 ? c.P.dS : -1;
      }
      gmapS[o] = dst;
    }
  }

  // ---- one gradient evaluation over this CTA's rows [r0, r1) ------------------------------------------------
  static __device__ __forceinline__ void run(Ctx& c, long r0, long r1, float* gpart) {
    const KParams& P = c.P;
    const DevModel& M = P.M;
    float* img = c.tile;
    float* region = c.tile + IMG_FLOATS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int* imap = aux(c);
    const int* gmapA = imap + 2 * NGROUPS;
    const int* gmapS = gmapA + NACC * 32;
    // 1. weight-fragment image of the current position (hi / lo tf32 parts), head weights in fp32
    for (int grp = tid; grp < NGROUPS; grp += NT) {
      const int s0 = imap[2 * grp], s1 = imap[2 * grp + 1];
      const float v0 = s0 >= 0 ? c.th[s0] : 0.f, v1 = s1 >= 0 ? c.th[s1] : 0.f;
      uint32_t h0, l0, h1, l1;
      split_tf32(v0, h0, l0); split_tf32(v1, h1, l1);
      reinterpret_cast<uint4*>(img)[grp] = make_uint4(h0, h1, l0, l1);
    }
    if (tid < 34) img[IMG_HEAD + tid] = c.th[tid < 32 ? M.kern_off[NH] + tid : M.bias_off[NH] + (tid - 32)];
    __syncthreads();
    PROF(0);

    const long nrows = r1 > r0 ? r1 - r0 : 0;
    const int ntiles = (int)((nrows + 15) >> 4);
    const int sx = M.sA[0];
    const float nb = M.n_batches;
    const float4* img4 = reinterpret_cast<const float4*>(img);
    float* PT = region + warp * RSTRIDE;
    float* PA = PT + 16 * PT_S;

    float accW[NH][2][4], accHead[8], db[NH][4], dbHead[2] = {0.f, 0.f}, ll_acc = 0.f;   // accW[0] uses n-tiles < KS0 only
#pragma unroll
    for (int l = 0; l < NH; ++l)
#pragma unroll
      for (int i = 0; i < 8; ++i) accW[l][i >> 2][i & 3] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) accHead[i] = 0.f;
#pragma unroll
    for (int l = 0; l < NH; ++l) { db[l][0] = 0.f; db[l][1] = 0.f; db[l][2] = 0.f; db[l][3] = 0.f; }

#pragma unroll 1
    for (int mt = warp; mt < ntiles; mt += NW) {
      const int row_a = mt * 16 + g, row_b = row_a + 8;
      float a[NH][2][4];                         // a[l][nt][*] = activation a_{l+1}, C-fragment layout
      // ---- layer 0: A operand straight from the X slice (rows beyond the shard are zero / clamped and masked below)
      {
        float cf[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float b0 = c.th[M.bias_off[0] + 8 * nt + 2 * t], b1 = c.th[M.bias_off[0] + 8 * nt + 2 * t + 1];
          cf[nt][0] = b0; cf[nt][1] = b1; cf[nt][2] = b0; cf[nt][3] = b1;
        }
        float xv[KS0][4];
        if (P.resident) {            // shared-memory slice (explicit LDS: a pointer that may be global or shared compiles to generic loads)
          const float* xa = c.xbuf + row_a * sx + t;
#pragma unroll
          for (int ks = 0; ks < KS0; ++ks) {
            xv[ks][0] = xa[8 * ks]; xv[ks][1] = xa[8 * sx + 8 * ks]; xv[ks][2] = xa[8 * ks + 4]; xv[ks][3] = xa[8 * sx + 8 * ks + 4];
          }
        } else {
          const long ca = row_a < nrows ? row_a : nrows - 1, cb = row_b < nrows ? row_b : nrows - 1;
          const float* __restrict__ xa = P.X + (r0 + ca) * sx + t;
          const float* __restrict__ xb = P.X + (r0 + cb) * sx + t;
#pragma unroll
          for (int ks = 0; ks < KS0; ++ks) {
            xv[ks][0] = __ldg(xa + 8 * ks); xv[ks][1] = __ldg(xb + 8 * ks); xv[ks][2] = __ldg(xa + 8 * ks + 4); xv[ks][3] = __ldg(xb + 8 * ks + 4);
          }
        }
#pragma unroll
        for (int ks = 0; ks < KS0; ++ks) {
          AFrag A;
          make_afrag(A, xv[ks][0], xv[ks][1], xv[ks][2], xv[ks][3]);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const float4 B = img4[GRP_F0 + (ks * 2 + nt) * 32 + lane];
            mma3(cf[nt], A, __float_as_uint(B.x), __float_as_uint(B.y), __float_as_uint(B.z), __float_as_uint(B.w));
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) a[0][nt][e] = fmaxf(cf[nt][e], 0.f);
      }
      // ---- hidden layers 1..NH-1: the C fragments of a_l are the A fragments (c0, c2, c1, c3) of the next layer
#pragma unroll
      for (int l = 1; l < NH; ++l) {
        float cf[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float b0 = c.th[M.bias_off[l] + 8 * nt + 2 * t], b1 = c.th[M.bias_off[l] + 8 * nt + 2 * t + 1];
          cf[nt][0] = b0; cf[nt][1] = b1; cf[nt][2] = b0; cf[nt][3] = b1;
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          AFrag A;
          make_afrag(A, a[l - 1][ks][0], a[l - 1][ks][2], a[l - 1][ks][1], a[l - 1][ks][3]);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const float4 B = img4[GRP_F1 + (l - 1) * 128 + (ks * 2 + nt) * 32 + lane];
            mma3(cf[nt], A, __float_as_uint(B.x), __float_as_uint(B.y), __float_as_uint(B.z), __float_as_uint(B.w));
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) a[l][nt][e] = fmaxf(cf[nt][e], 0.f);
      }
      // ---- head 16 -> 2 in fp32 FMAs: every lane owns columns {2t, 2t+1, 8+2t, 9+2t} of rows g and g+8
      const float4 w01 = *reinterpret_cast<const float4*>(img + IMG_HEAD + 4 * t);        // W[2t][0..1], W[2t+1][0..1]
      const float4 w23 = *reinterpret_cast<const float4*>(img + IMG_HEAD + 16 + 4 * t);   // W[8+2t][..], W[9+2t][..]
      const float (&an)[2][4] = a[NH - 1];
      float mu[2], sg[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mu[r] = an[0][2 * r] * w01.x + an[0][2 * r + 1] * w01.z + an[1][2 * r] * w23.x + an[1][2 * r + 1] * w23.z;
        sg[r] = an[0][2 * r] * w01.y + an[0][2 * r + 1] * w01.w + an[1][2 * r] * w23.y + an[1][2 * r + 1] * w23.w;
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
        for (int r = 0; r < 2; ++r) { mu[r] += __shfl_xor_sync(0xffffffffu, mu[r], o); sg[r] += __shfl_xor_sync(0xffffffffu, sg[r], o); }
      }
      const float hb0 = img[IMG_HEAD + 32], hb1 = img[IMG_HEAD + 33];
      float dmu[2], ds[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long row = r ? row_b : row_a;
        dmu[r] = 0.f; ds[r] = 0.f;
        if (row < nrows) {
          const float m_ = mu[r] + hb0, s_ = sg[r] + hb1;
          const float yv = reinterpret_cast<const float*>(P.y)[r0 + row];
          // sigma = clip(exp(s), 1e-6, 1e6) (probabilistic.py:100), written through sc = clip(s, ln 1e-6, ln 1e6)
          const float kLnClip = 13.815510557964274f;
          const float sc = fminf(fmaxf(s_, -kLnClip), kLnClip);
          const float inside = (s_ > -kLnClip && s_ < kLnClip) ? 1.f : 0.f;
          const float inv_s2 = expf(-2.f * sc), res = yv - m_, q = res * res * inv_s2;
          float ll = (MILE_LOG_2PI + 2.f * sc + q) * -0.5f;
          float d0 = res * inv_s2, d1 = (q - 1.f) * inside;
          if (isnan(ll)) { ll = 0.f; d0 = 0.f; d1 = 0.f; }   // jnp.nansum
          ll_acc += ll;
          dmu[r] = d0 * nb; ds[r] = d1 * nb;
        }
      }
      dbHead[0] += dmu[0] + dmu[1]; dbHead[1] += ds[0] + ds[1];
      float dl[2][4];                            // delta_l (pre-activation), C-fragment layout; starts at l = NH-1
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const float w0 = nt == 0 ? (cc ? w01.z : w01.x) : (cc ? w23.z : w23.x);
          const float w1 = nt == 0 ? (cc ? w01.w : w01.y) : (cc ? w23.w : w23.y);
          const float a0 = an[nt][cc], a1 = an[nt][2 + cc];
          accHead[(nt * 2 + cc) * 2 + 0] += a0 * dmu[0] + a1 * dmu[1];
          accHead[(nt * 2 + cc) * 2 + 1] += a0 * ds[0] + a1 * ds[1];
          dl[nt][cc] = a0 > 0.f ? fmaf(dmu[0], w0, ds[0] * w1) : 0.f;
          dl[nt][2 + cc] = a1 > 0.f ? fmaf(dmu[1], w0, ds[1] * w1) : 0.f;
        }
      // ---- backward: l = NH-1 .. 0
#pragma unroll
      for (int l = NH - 1; l >= 0; --l) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) { db[l][nt * 2] += dl[nt][0] + dl[nt][2]; db[l][nt * 2 + 1] += dl[nt][1] + dl[nt][3]; }
        // transposed patch of delta_l (PT[column][row]) and row-major patch of a_l (PA[row][column])
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) PT[(8 * nt + 2 * t + (e & 1)) * PT_S + g + ((e >> 1) ? 8 : 0)] = dl[nt][e];
        if (l >= 1) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            *reinterpret_cast<float2*>(PA + g * PA_S + 8 * nt + 2 * t) = make_float2(a[l - 1][nt][0], a[l - 1][nt][1]);
            *reinterpret_cast<float2*>(PA + (g + 8) * PA_S + 8 * nt + 2 * t) = make_float2(a[l - 1][nt][2], a[l - 1][nt][3]);
          }
        }
        __syncwarp();
        // dW_l^T [out j][in i] += sum_r delta_l[r][j] a_l[r][i]:  A[m = j][k = r] from PT, B[k = r][n = i] from PA / X
        // (the tensor core accumulates with round-toward-zero: a chain over all of the warp's tiles would bias the sum by
        //  ~n * 2^-24, so every tile starts from zero and is added to the running sum with a rounded FADD)
        float cw[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          AFrag A;
          make_afrag(A, PT[g * PT_S + 8 * ks + t], PT[(g + 8) * PT_S + 8 * ks + t], PT[g * PT_S + 8 * ks + t + 4],
                     PT[(g + 8) * PT_S + 8 * ks + t + 4]);
          const int nti = l == 0 ? KS0 : 2;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            if (nt < nti) {
              float v0, v1;
              if (l == 0) {
                if (P.resident) {
                  v0 = c.xbuf[(mt * 16 + 8 * ks + t) * sx + 8 * nt + g];
                  v1 = c.xbuf[(mt * 16 + 8 * ks + t + 4) * sx + 8 * nt + g];
                } else {
                  const long q0 = mt * 16 + 8 * ks + t, q1 = q0 + 4;
                  v0 = __ldg(P.X + (r0 + (q0 < nrows ? q0 : nrows - 1)) * sx + 8 * nt + g);
                  v1 = __ldg(P.X + (r0 + (q1 < nrows ? q1 : nrows - 1)) * sx + 8 * nt + g);
                }
              } else {
                v0 = PA[(8 * ks + t) * PA_S + 8 * nt + g];
                v1 = PA[(8 * ks + t + 4) * PA_S + 8 * nt + g];
              }
              uint32_t bh0, bl0, bh1, bl1;
              split_tf32(v0, bh0, bl0); split_tf32(v1, bh1, bl1);
              mma3(cw[nt], A, bh0, bh1, bl0, bl1);
            }
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          if (nt < (l == 0 ? KS0 : 2)) {
#pragma unroll
            for (int e = 0; e < 4; ++e) accW[l][nt][e] += cw[nt][e];
          }
        if (l >= 1) {
          // delta_{l-1} = (delta_l W_l^T) * relu'(a_l)
          float nd[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            AFrag A;
            make_afrag(A, dl[ks][0], dl[ks][2], dl[ks][1], dl[ks][3]);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const float4 B = img4[GRP_B1 + (l - 1) * 128 + (ks * 2 + nt) * 32 + lane];
              mma3(nd[nt], A, __float_as_uint(B.x), __float_as_uint(B.y), __float_as_uint(B.z), __float_as_uint(B.w));
            }
          }
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) dl[nt][e] = a[l - 1][nt][e] > 0.f ? nd[nt][e] : 0.f;
        }
        __syncwarp();
      }
    }

    PROF(1);
    // ---- cross-warp sum of the accumulators: every warp that owned tiles writes its sums into its own region
    {
      const int nact = ntiles < NW ? ntiles : NW;
      if (warp < nact) {
        // FMA-path sums: reduce over the 8 row groups g (lane bits 2..4); lanes g = 0 report
        float sm[NSMALL];
#pragma unroll
        for (int i = 0; i < 8; ++i) sm[i] = accHead[i];
#pragma unroll
        for (int l = 0; l < NH; ++l)
#pragma unroll
          for (int q = 0; q < 4; ++q) sm[8 + 4 * l + q] = db[l][q];
        sm[8 + 4 * NH] = dbHead[0]; sm[8 + 4 * NH + 1] = dbHead[1]; sm[8 + 4 * NH + 2] = ll_acc;
#pragma unroll
        for (int i = 0; i < NSMALL; ++i) {
          float v = sm[i];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          sm[i] = v;
        }
        float* scr = region + warp * RSTRIDE;
#pragma unroll
        for (int i = 0; i < NACC0; ++i) scr[i * 32 + lane] = accW[0][i >> 2][i & 3];
#pragma unroll
        for (int l = 1; l < NH; ++l)
#pragma unroll
          for (int i = 0; i < 8; ++i) scr[(NACC0 + (l - 1) * 8 + i) * 32 + lane] = accW[l][i >> 2][i & 3];
        if (g == 0) {
#pragma unroll
          for (int i = 0; i < NSMALL; ++i) scr[NACC * 32 + i * 4 + t] = sm[i];
        }
      }
      __syncthreads();
      for (int o = tid; o < SCR_WARP; o += NT) {
        const int dst = o < NACC * 32 ? gmapA[o] : gmapS[o - NACC * 32];
        if (dst >= 0) {
          // all partials of this slot in flight at once (predicated), then summed in warp order
          float pv[NW];
#pragma unroll
          for (int w = 0; w < NW; ++w) pv[w] = w < nact ? region[w * RSTRIDE + o] : 0.f;
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < NW; ++w) s += pv[w];
          gpart[dst] = dst == P.dS ? s * nb : s;
        }
      }
      // no trailing barrier: every caller synchronises (block or cluster barrier) before gpart is consumed, and the
      // regions are next written after further barriers
      PROF(5);
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Step loop for the tensor evaluator (MODE_SAMPLE / MODE_TUNE).  Same arithmetic as the loop of mile_mclmc_kernel,
// restructured after two ncu captures (profiles/r2d_*, r2g_*):
//   * in the generic loop all 16 warps execute every scalar chain of a B-step (sqrt, divide, expm1, log1p, rsqrt; ~100
//     dependent instructions) and draw the refresh noise in line (Philox + Box-Muller, ~150 instructions per element):
//     85 % of the issued instructions were this redundant or in-line work;
//   * a first restructuring with ONE integrator warp owning all d elements removed the redundancy but issued ~2700
//     dependent instructions per evaluation from a single warp (23 elements per lane): slower.
// This version keeps the ELEMENT sweeps on all threads (<= 2 elements per thread), runs each SCALAR chain on warp 0
// only and publishes the coefficients through shared memory, and lets the other 15 warps draw the NEXT step's noise
// into a double-buffered shared array during exactly those windows, so the draws cost no time on the critical path.
// Block reductions use a shuffle butterfly over the per-warp partials (36 instead of 80 instructions for four sums).
// Cluster exchange = reduce-scatter + all-gather over DSMEM (each CTA sums one slice of the partial gradients in rank
// order, then every CTA gathers the summed slices: 2 x d/G remote floats per CTA instead of (G-1) x d).  The
// flagged-word exchange through L2 (any G <= 16, cooperative launch) and G = 1 are the other two cases.
// ------------------------------------------------------------------------------------------------------------
// ---- DSMEM push exchange (cluster of G <= 16 CTAs per chain) -------------------------------------------------------------
// st.async: a 4-byte store into another CTA's shared memory that signals the DESTINATION's mbarrier with the bytes written;
// the receiver sleeps on its own barrier (mbarrier.try_wait) until the expected byte count has arrived.  One DSMEM store
// latency per hop and no cluster-wide barrier (cluster.sync costs ~380 cycles and the pull form needs two of them plus two
// dependent remote-load round trips per evaluation).
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(__float_as_uint(v)),
               "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void xbar_arm(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (long spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1L << 22)) __trap();     // a lost store must not hang the GPU
  }
}

template <int NV, int NT>
__device__ __forceinline__ void block_sum_bfly(float (&v)[NV], float* scratch, int& phase) {
  constexpr int NWARPS = NT / 32;
  static_assert(NWARPS == 16, "butterfly over 16 per-warp partials");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* buf = scratch + phase * (4 * NWARPS);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float s = warp_sum(v[k]);
    if (lane == 0) buf[k * NWARPS + warp] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float s = buf[k * NWARPS + (lane & 15)];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    v[k] = s;
  }
  phase ^= 1;
}

template <class GE>
__global__ void __launch_bounds__(GE::NT, 1) mile_mma_step_kernel(const __grid_constant__ KParams P) {
  constexpr int NT = GE::NT;
  extern __shared__ __align__(16) float smem[];
  const DevModel& M = P.M;
  cg::cluster_group cluster = cg::this_cluster();
  Ctx c(P);
  c.G = P.G;
  c.rank = c.G > 1 ? (P.sync_mode ? (int)(blockIdx.x % c.G) : (int)cluster.block_rank()) : 0;
  c.chain = blockIdx.x / c.G;
  c.phase = 0; c.lead = threadIdx.x == 0;
  c.wp = smem + P.off_wp; c.th = smem + P.off_th; c.uu = smem + P.off_u; c.gg = smem + P.off_g;
  c.thb = smem + P.off_thb; c.ub = smem + P.off_ub; c.gb = smem + P.off_gb; c.gpart = smem + P.off_gpart;
  c.avgx = smem + P.off_avgx; c.avgx2 = smem + P.off_avgx2; c.pmap = reinterpret_cast<int*>(smem + P.off_pmap);
  c.red = smem + P.off_red; c.red2 = c.red + 128; c.phase2 = 0; c.tile = smem + P.off_tile;
  c.xstream = c.tile + M.tile_floats;
  c.xbuf = P.resident ? smem + P.off_x : c.xstream;
  c.aux = smem + P.off_aux;
  float* zbuf = smem + P.off_z;        // [2 steps][2 slots][dS]
  float* gslice = smem + P.off_gs;     // this CTA's slice of the summed gradient (cluster exchange)
  float* sc = c.red + 128;             // scalars published by warp 0: ae, au, dK increment, nu
  const int d = M.d, ch = c.chain, dS = P.dS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tune = P.mode == MODE_TUNE;
  const int nslot = P.refresh_mode ? 2 : 1;
  const int i0 = tid, i1 = tid + NT;   // the (at most) two elements this thread owns: d <= GE::DMAX <= 2 NT
  const bool has0 = i0 < d, has1 = i1 < d;

  // ---- prologue ---------------------------------------------------------------------------------------------
  for (int i = tid; i < M.psize; i += NT) c.wp[i] = 0.f;
  build_pmap<NT>(M, c.pmap, dS);
  GE::prepare(c);
  for (int i = tid; i < d; i += NT) {
    c.th[i] = P.theta[(long)ch * d + i]; c.uu[i] = P.u[(long)ch * d + i]; c.gg[i] = P.grad[(long)ch * d + i];
    if (tune) { c.avgx[i] = P.avg_x[(long)ch * d + i]; c.avgx2[i] = P.avg_x2[(long)ch * d + i]; }
  }
  const long per = (P.N + c.G - 1) / c.G;
  const long r0 = per * c.rank < P.N ? per * c.rank : P.N;
  const long r1 = (r0 + per) < P.N ? (r0 + per) : P.N;
  if (P.resident) {
    const int sx = M.sA[0];
    const long nv4 = (r1 - r0) * (sx >> 2), np4 = (long)P.rows_res * (sx >> 2);
    const float4* s4 = reinterpret_cast<const float4*>(P.X + r0 * sx);
    float4* d4 = reinterpret_cast<float4*>(c.xbuf);
    for (long i = tid; i < np4; i += NT) d4[i] = i < nv4 ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // noise of step s_local into its buffer; `part` of `nparts` (the draws are spread over the idle windows of a step)
  auto draw_noise = [&](long s_local, int t0, int tstride, int part, int nparts) {
    float* zs = zbuf + (s_local & 1) * (2 * dS);
    const int npair = (d + 1) >> 1;            // one Philox call yields the normals of elements 2p and 2p + 1
    int j = 0;
    for (int k = t0; k < nslot * npair; k += tstride, ++j) {
      if (j % nparts != part) continue;
      const int slot = k >= npair ? 1 : 0, p2 = (k - slot * npair) * 2;
      if (P.z) {
        zs[slot * dS + p2] = noise_at(P, ch, s_local, slot, nslot, p2);
        if (p2 + 1 < d) zs[slot * dS + p2 + 1] = noise_at(P, ch, s_local, slot, nslot, p2 + 1);
      } else {
        float z0, z1;
        philox_normal2(P.seed, (uint32_t)(P.chain_base + ch), (uint64_t)(P.step_base + s_local), (uint32_t)slot + 1u, (uint32_t)(p2 >> 1), z0, z1);
        zs[slot * dS + p2] = z0;
        if (p2 + 1 < d) zs[slot * dS + p2 + 1] = z1;
      }
    }
  };
  if (P.n_steps > 0) draw_noise(0, tid, NT, 0, 1);
  // push exchange: recv[G][SL] + gfull[dS + 4] live in the gslice area; two mbarriers behind the published scalars
  const bool push = c.G > 1 && !P.sync_mode;
  const uint32_t xb1 = (uint32_t)__cvta_generic_to_shared(c.red + 160), xb2 = xb1 + 8;
  float* recv = gslice;
  float* gfull = gslice + (dS + 16);
  if (push) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xb1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xb2));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();      // every CTA's barriers exist before the first remote store
  }
  __syncthreads();

  // ---- chain scalars, uniform across the block -------------------------------------------------------------------
  float g2 = 0.f, ug = 0.f, nf = 0.f, lp = P.lp[ch], lp_old = 0.f, dK = 0.f;
  float eps = tune ? P.t_eps[ch] : P.eps[ch];
  const float Lc = tune ? P.t_L[ch] : P.L[ch];
  TuneRegs tr{0.f, 0.f, INFINITY, 0.f};
  if (tune) { tr.time = P.t_time[ch]; tr.xavg = P.t_xavg[ch]; tr.epsmax = P.t_epsmax[ch]; tr.wtot = P.t_wtot[ch]; }
  if (P.carry_valid) { g2 = P.carry[2 * ch]; ug = P.carry[2 * ch + 1]; }
  else {
    float v[2] = {0.f, 0.f};
    if (has0) { v[0] += c.gg[i0] * c.gg[i0]; v[1] += c.uu[i0] * c.gg[i0]; }
    if (has1) { v[0] += c.gg[i1] * c.gg[i1]; v[1] += c.uu[i1] * c.gg[i1]; }
    block_sum_bfly<2, NT>(v, c.red, c.phase);
    g2 = v[0]; ug = v[1];
  }
  const float b1 = 0.1931833275037836f, b2 = 1.f - 2.f * 0.1931833275037836f;
  const float loc = M.prior_loc, scl = M.prior_scale, s2 = scl * scl, inv_s2 = 1.f / s2, inv_scl = 1.f / scl;
  const float lognorm = M.prior == MILE_PRIOR_NORMAL ? logf(6.283185307179586f * s2) : logf(2.f * scl);
  const int n_evals = 2 * P.n_steps;
  PROF_DECL;
#pragma unroll 1
  for (int e = 0; e < n_evals; ++e) {
    const int h = e & 1, s = e >> 1;
    const float* zs = zbuf + (s & 1) * (2 * dS);
    if (h == 0) {
      lp_old = lp; dK = 0.f;
      if (tune) {
        if (has0) { c.thb[i0] = c.th[i0]; c.ub[i0] = c.uu[i0]; c.gb[i0] = c.gg[i0]; }
        if (has1) { c.thb[i1] = c.th[i1]; c.ub[i1] = c.uu[i1]; c.gb[i1] = c.gg[i1]; }
      }
      if (P.refresh_mode) refresh_momentum<NT>(c, 0.5f * eps, Lc, s, 0, nslot, ug, zs);
    }
    // ---- scalar chain of B(b1 | 1 - 2 b1) on warp 0; the other warps draw a part of the next step's noise ----------
    if (warp == 0) {
      float ae, au;
      const float dk = esh_coeffs(d, eps, h == 0 ? b1 : b2, g2, ug, ae, au);
      if (lane == 0) { sc[0] = ae; sc[1] = au; sc[2] = dk; }
    } else if (s + 1 < P.n_steps) {
      draw_noise(s + 1, tid - 32, NT - 32, h, 3);
    }
    __syncthreads();
    {
      const float ae = sc[0], au = sc[1], st = eps * 0.5f;
      dK += sc[2];
      if (has0) { const float un = ae * c.gg[i0] + au * c.uu[i0]; c.uu[i0] = un; c.th[i0] += st * un; }
      if (has1) { const float un = ae * c.gg[i1] + au * c.uu[i1]; c.uu[i1] = un; c.th[i1] += st * un; }
    }
    PROF(8);
    __syncthreads();                       // the new position is visible to the evaluator
    float* gp = c.gpart + (e & 1) * (dS + 4);
    GE::run(c, r0, r1, gp);
    PROF(10);
    // ---- exchange: afterwards gsum[0..d) = sum over the chain's CTAs of the likelihood gradient, gp[ll_idx] = sum of ll
    const float* gsum = c.gg;
    int ll_idx = dS + 1;
    if (c.G == 1) {
      __syncthreads();
      gsum = gp; ll_idx = dS;
    } else if (P.sync_mode) {
      const unsigned int xflag = P.xbase + (unsigned int)e + 1u;
      float2* slab = P.xchg + ((long)ch * 2 + (e & 1)) * c.G * (dS + 4);
      float2* mine = slab + c.rank * (dS + 4);
      __syncthreads();
      for (int i = tid; i <= dS; i += NT) ll_store(mine + i, gp[i], xflag);
      PROF(13);
      // reduce-scatter + all-gather through L2: every CTA sums ONE slice of the G partials (G flagged words per element of
      // its slice) and republishes the sums; everybody then reads d + 1 summed words.  The all-to-all form (every CTA
      // polls all G x d words) costs ~770 LSU wavefronts per poll and CTA at G = 12, this form ~60 + ~45.
      const int stride_g = dS + 4;
      const int SL = (dS + 1 + c.G - 1) / c.G;
      float2* sums = P.xchg2 + ((long)ch * 2 + (e & 1)) * (dS + 4);
      for (int j = tid; j < SL; j += NT) {
        const int idx = c.rank * SL + j;
        if (idx <= dS) ll_store(sums + idx, ll_sum(slab + idx, stride_g, c.G, xflag), xflag);
      }
      if (has1) {
        float s0, s1;
        ll_sum_pair(sums + i0, sums + i1, 0, 1, xflag, s0, s1);
        c.gg[i0] = s0; c.gg[i1] = s1;
      } else if (has0) {
        c.gg[i0] = ll_sum(sums + i0, 0, 1, xflag);
      }
      if (tid == NT - 1) gp[dS + 1] = ll_sum(sums + dS, 0, 1, xflag);
      __syncthreads();
    } else {
      // reduce-scatter + all-gather over DSMEM by PUSH; slices of SL elements over [0, dS] (element dS = log-likelihood).
      //   hop 1: every CTA stores element i of its partial into the owner's recv[rank][i - owner SL]  (owner = i / SL)
      //   owner: waits for G x |slice| x 4 bytes on its barrier 1, adds the G partials in rank order
      //   hop 2: the owner stores each sum into every CTA's gfull[i]; everybody waits for (dS + 1) x 4 bytes on barrier 2
      // Buffer reuse is safe without further barriers: a CTA can only start hop 1 of evaluation e+1 after it passed
      // barrier 2 of e, i.e. after EVERY owner finished reading its recv slots of e; an owner can only start hop 2 of e+1
      // after it received hop 1 of e+1 from everybody, i.e. after everybody consumed gfull of e.
      const int SL = (dS + 1 + c.G - 1) / c.G;
      const int my0 = c.rank * SL;
      const int mySL = my0 > dS ? 0 : (my0 + SL > dS + 1 ? dS + 1 - my0 : SL);
      const uint32_t par = (uint32_t)(e & 1);
      __syncthreads();                     // this CTA's partial is complete
      if (tid == 0) { xbar_arm(xb1, (uint32_t)(c.G * mySL * 4)); xbar_arm(xb2, (uint32_t)((dS + 1) * 4)); }
      const uint32_t recv_s = (uint32_t)__cvta_generic_to_shared(recv), gfull_s = (uint32_t)__cvta_generic_to_shared(gfull);
      for (int i = tid; i <= dS; i += NT) {
        const int owner = i / SL, off = i - owner * SL;
        st_async_f32(mapa_u32(recv_s + (uint32_t)(c.rank * SL + off) * 4u, (uint32_t)owner), gp[i], mapa_u32(xb1, (uint32_t)owner));
      }
      PROF(13);
      xbar_wait(xb1, par);
      for (int j = tid; j < mySL; j += NT) {
        float tv[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) tv[r] = r < c.G ? recv[r * SL + j] : 0.f;
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) sum += tv[r];
#pragma unroll
        for (int r = 0; r < 16; ++r)
          if (r < c.G) st_async_f32(mapa_u32(gfull_s + (uint32_t)(my0 + j) * 4u, (uint32_t)r), sum, mapa_u32(xb2, (uint32_t)r));
      }
      PROF(2);
      xbar_wait(xb2, par);
      if (has0) c.gg[i0] = gfull[i0];
      if (has1) c.gg[i1] = gfull[i1];
      if (tid == NT - 1) gp[dS + 1] = gfull[dS];
      __syncthreads();
    }
    PROF(14);
    // ---- prior, full gradient, and the sums the next B-step / handle_nans need -----------------------------------
    float gk[2] = {0.f, 0.f}, uk[2] = {0.f, 0.f};
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = k ? i1 : i0;
        if (k ? has1 : has0) {
          const float th = c.th[i], dlt = th - loc;
          float pg, pv;
          if (M.prior == MILE_PRIOR_NORMAL) { pv = (lognorm + dlt * dlt * inv_s2) * -0.5f; pg = -dlt * inv_s2; }
          else { pv = -lognorm - fabsf(dlt) * inv_scl; pg = -((dlt > 0.f) - (dlt < 0.f)) * inv_scl; }
          const float gi = gsum[i] + pg * P.prior_weight, ui = c.uu[i];
          c.gg[i] = gi; gk[k] = gi; uk[k] = ui;
          v[0] += pv * P.prior_weight; v[1] += gi * gi; v[2] += ui * gi; v[3] += isfinite(th) ? 0.f : 1.f;
        }
      }
      block_sum_bfly<4, NT>(v, c.red, c.phase);
      lp = v[0] + gp[ll_idx]; g2 = v[1]; ug = v[2]; nf = v[3];
    }
    PROF(11);
    if (h == 1) {
      // ---- end of MCLMC step s: last B and the partial refresh in one sweep (u, g still in registers) ------------
      const float er = P.refresh_mode ? 0.5f * eps : eps;
      if (warp == 0) {
        float ae, au;
        const float dk = esh_coeffs(d, eps, b1, g2, ug, ae, au);
        const float nu = isinf(Lc) ? 0.f : sqrtf((expf(2.f * er / Lc) - 1.f) / (float)d);
        if (lane == 0) { sc[4] = ae; sc[5] = au; sc[6] = dk; sc[7] = nu; }
      } else if (s + 1 < P.n_steps) {
        draw_noise(s + 1, tid - 32, NT - 32, 2, 3);
      }
      __syncthreads();
      const float ae = sc[4], au = sc[5], nu = sc[7];
      dK += sc[6];
      const float* zl = zs + (nslot - 1) * dS;
      float w[2] = {0.f, 0.f}, w2[2] = {0.f, 0.f};
      if (has0) { w[0] = ae * gk[0] + au * uk[0] + nu * zl[i0]; w2[0] += w[0] * w[0]; w2[1] += w[0] * gk[0]; }
      if (has1) { w[1] = ae * gk[1] + au * uk[1] + nu * zl[i1]; w2[0] += w[1] * w[1]; w2[1] += w[1] * gk[1]; }
      block_sum_bfly<2, NT>(w2, c.red, c.phase);
      float inv = rsqrtf(w2[0]);
      inv = inv * fmaf(-0.5f * w2[0] * inv, inv, 1.5f);
      if (has0) c.uu[i0] = w[0] * inv;
      if (has1) c.uu[i1] = w[1] * inv;
      ug = w2[1] * inv;
      float dE = dK - lp + lp_old;
      if (!tune) {
        if (P.info && c.rank == 0 && tid == 0) {
          float* o = P.info + ((long)s * P.C + ch) * 3;
          o[0] = lp; o[1] = dK; o[2] = dE;
        }
        const long idx = P.step_base + s;    // thinned sample capture (sampling.py:152-164)
        if (idx % P.thin == 0) {
          const long slot = idx / P.thin - P.sample_base;
          if (P.samples && c.rank == 0 && slot >= 0 && slot < P.n_slots) {
            if (has0) P.samples[(slot * P.C + ch) * d + i0] = c.th[i0];
            if (has1) P.samples[(slot * P.C + ch) * d + i1] = c.th[i1];
          }
        }
      } else {
        __syncthreads();                     // tune_epilogue re-reads u / g of other threads' making through shared memory
        eps = tune_epilogue<NT, false>(c, tr, eps, lp_old, nf, s, lp, dE, g2, ug);
      }
      PROF(12);
      // fused posterior-predictive fold at every kept position (whole block; needs the padded weight image)
      if (!tune && P.do_lppd && (P.step_base + s) % P.thin == 0) {
        __syncthreads();
        refresh_wp<NT>(c);
        __syncthreads();
        lppd_fold<NT>(c, ch);
      }
    }
  }
  // ---- epilogue: write the state back ----------------------------------------------------------------------------
  __syncthreads();
  if (c.rank == 0) {
    for (int i = tid; i < d; i += NT) {
      P.theta[(long)ch * d + i] = c.th[i];
      P.u[(long)ch * d + i] = c.uu[i];
      P.grad[(long)ch * d + i] = c.gg[i];
      if (tune) { P.avg_x[(long)ch * d + i] = c.avgx[i]; P.avg_x2[(long)ch * d + i] = c.avgx2[i]; }
    }
    if (tid == 0) {
      P.lp[ch] = lp;
      P.carry[2 * ch] = g2; P.carry[2 * ch + 1] = ug;
      if (tune) { P.t_time[ch] = tr.time; P.t_xavg[ch] = tr.xavg; P.t_epsmax[ch] = tr.epsmax; P.t_eps[ch] = eps; P.t_wtot[ch] = tr.wtot; }
    }
  }
  if (c.G > 1 && !P.sync_mode) cluster.sync();
}
