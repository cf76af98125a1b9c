// mile_fast.cuh -- warp-specialised layer pipeline for the narrow regression MLPs
// (hidden width 16, Gaussian head: the airfoil / bikesharing / protein configs of BASELINE.json).
//
// Why this shape (measured on B200, profiles/r1_lds_microbench.txt + profiles/r1a_*): a warp-wide
// LDS.128 costs ~3 SM cycles whatever its address pattern (register write-back of 512 B), so an
// FP32 tile GEMM fed from shared memory needs >= 4 FMAs per loaded float to be FMA-bound.  The
// generic 4x4-register-tile kernel gets 2.  Here every thread keeps one operand STATIONARY in
// registers for a whole gradient evaluation -- a [K][4] slab of the layer's weights (forward /
// backward) or of the weight-gradient accumulator (dW) -- and streams activation rows past it:
// 4 LDS.128 per 64 FMAs.  To keep those slabs resident across all row tiles, the 16 warps of the
// CTA are specialised by layer stage and the row tiles flow through them as a lock-step software
// pipeline (one bar.sync per tick):
//
//   NL=3 (2x16):  fwd0 | fwd1 | head(fwd2+loglik+bwd2+dW2) | bwd1 | dW0,dW1          depth 5
//   NL=4 (3x16):  fwd0 | fwd1 | fwd2 | head | bwd2 | bwd1 | dW0,dW1,dW2              depth 7
//
// Restates the same arithmetic as grad_eval (mile_kernel.cuh): probabilistic.py:92-138 through
// basic.py:41-61, differentiated by hand.
#pragma once
#include "mile_kernel.cuh"

template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
  if (ACT == MILE_ACT_RELU) return fmaxf(z, 0.f);
  if (ACT == MILE_ACT_SIGMOID) return 1.f / (1.f + expf(-z));
  if (ACT == MILE_ACT_TANH) return tanhf(z);
  if (ACT == MILE_ACT_LEAKY_RELU) return z >= 0.f ? z : 0.01f * z;
  return z;
}
// derivative expressed through the activation VALUE (so no act' buffer is needed)
template <int ACT>
__device__ __forceinline__ float act_bwd_from_value(float a) {
  if (ACT == MILE_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  if (ACT == MILE_ACT_SIGMOID) return a * (1.f - a);
  if (ACT == MILE_ACT_TANH) return 1.f - a * a;
  if (ACT == MILE_ACT_LEAKY_RELU) return a >= 0.f ? 1.f : 0.01f;
  return 1.f;
}

#ifdef MILE_PROFILE
// developer build (tools/phase_profile.py): cycles every warp of CTA 0 spends waiting at the per-tick barrier (slots 16 + warp)
__device__ __forceinline__ void bar_block() {
  const long long t0 = clock64();
  asm volatile("bar.sync 0;" ::: "memory");
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_prof[16 + (threadIdx.x >> 5)] += (unsigned long long)(clock64() - t0);
}
#else
__device__ __forceinline__ void bar_block() { asm volatile("bar.sync 0;" ::: "memory"); }
#endif

template <int NL, int FP, int ACT>
struct FastGE {
  static constexpr int NT = 512;
  static constexpr int H = 16;                    // hidden width (padded == actual)
  static constexpr int T = NL == 3 ? 64 : 32;     // rows per tile
  static constexpr int NH = NL - 1;               // hidden layers
  static constexpr int DEPTH = 2 * NL - 1;        // pipeline depth in ticks
  static constexpr int NBUF = 2 * NH;             // per slot: a_1..a_NH, D_0..D_{NH-1}
  static constexpr int SLOT_FLOATS = NBUF * T * H;
  static constexpr int RING_FLOATS = DEPTH * SLOT_FLOATS;

  // ---- warp -> stage tables (compile-time; warp w runs on SMSP w % 4, the tables balance the
  //      FMA work per SMSP) ----
  //   NL=3, T=64: fwd0 {0..3} fwd1 {4,5} head {6,7} bwd1 {8,9} dW0 {10..13} dW1 {14,15}
  //   NL=4, T=32: fwd0 {0} fwd1 {1,2} fwd2 {3,4} head {5} bwd2 {6,7} bwd1 {8,9} dW0 {10} dW1 {11,12} dW2 {13,14} idle {15}
  static constexpr int DW_STRIDE = (NL == 3 ? 32 : 16) * (H + 1) * H;   // scratch floats per dW stage (<= 32 / 16 row lanes)
  static constexpr int HEAD_STRIDE = 2 * H + 4;
  static __host__ __device__ constexpr int dw_row_lanes(int l) { return NL == 3 ? (l == 0 ? 32 : 16) : (l == 0 ? 8 : 16); }

  // buffers inside a ring slot
  static __device__ __forceinline__ float* abuf(float* ring, int slot, int l) {  // a_l, l = 1..NH
    return ring + slot * SLOT_FLOATS + (l - 1) * T * H;
  }
  static __device__ __forceinline__ float* dbuf(float* ring, int slot, int l) {  // D_l, l = 0..NH-1
    return ring + slot * SLOT_FLOATS + (NH + l) * T * H;
  }

  static __device__ __forceinline__ void prepare(Ctx&) {}

  // ---- stage bodies ---------------------------------------------------------------------------
  // forward GEMM stage: a_{l+1}[r][jt*4..] = act(b + a_l[r][:] W[:, jt*4..]);  stationary W[KP][4]
  template <int KP, int l, int W0, int NW, int TICK>
  static __device__ __forceinline__ void fwd_stage(Ctx& c, long r0, long nrows, int ntiles, float* ring) {
    const KParams& P = c.P;
    const DevModel& M = P.M;
    constexpr int RL = NW * 8;
    const int ls = threadIdx.x - W0 * 32, jt = ls & 3, rl = ls >> 2;
    float w[KP][4], b[4];
    {
      const float* W = c.wp + M.pw_off[l] + jt * 4;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(W + k * H);
        w[k][0] = v.x; w[k][1] = v.y; w[k][2] = v.z; w[k][3] = v.w;
      }
      const float4 bv = *reinterpret_cast<const float4*>(c.wp + M.pb_off[l] + jt * 4);
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
    }
    for (int tick = 0; tick < ntiles + DEPTH - 1; ++tick) {
      const int t = tick - TICK;
      if (t >= 0 && t < ntiles) {
        const int slot = t % DEPTH;
        const float* in = l == 0 ? c.xbuf + (long)t * T * FP : abuf(ring, slot, l);
        constexpr int sin_ = l == 0 ? FP : H;
        float* out = abuf(ring, slot, l + 1);
#pragma unroll
        for (int n = 0; n < T / RL; ++n) {
          const int r = rl + n * RL;
          float4 a[KP / 4];
          if (l == 0 && !P.resident) {   // X streamed from the padded HBM/L2 copy (row index clamped: finite data)
            const long gr = (long)t * T + r < nrows ? (long)t * T + r : nrows - 1;
            const float4* xp = reinterpret_cast<const float4*>(P.X + (r0 + gr) * FP);
#pragma unroll
            for (int k4 = 0; k4 < KP / 4; ++k4) a[k4] = __ldg(xp + k4);
          } else {
#pragma unroll
            for (int k4 = 0; k4 < KP / 4; ++k4) a[k4] = *reinterpret_cast<const float4*>(in + r * sin_ + k4 * 4);
          }
          float acc0 = b[0], acc1 = b[1], acc2 = b[2], acc3 = b[3];
#pragma unroll
          for (int k4 = 0; k4 < KP / 4; ++k4) {
            const float av[4] = {a[k4].x, a[k4].y, a[k4].z, a[k4].w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              acc0 = fmaf(av[kk], w[k4 * 4 + kk][0], acc0);
              acc1 = fmaf(av[kk], w[k4 * 4 + kk][1], acc1);
              acc2 = fmaf(av[kk], w[k4 * 4 + kk][2], acc2);
              acc3 = fmaf(av[kk], w[k4 * 4 + kk][3], acc3);
            }
          }
          *reinterpret_cast<float4*>(out + r * H + jt * 4) =
              make_float4(act_fwd<ACT>(acc0), act_fwd<ACT>(acc1), act_fwd<ACT>(acc2), act_fwd<ACT>(acc3));
        }
      }
      bar_block();
    }
  }

  // backward GEMM stage: D_{l-1}[r][it*4+ii] = act'(a_l[r][it*4+ii]) * sum_j D_l[r][j] W_l[it*4+ii][j]
  template <int l, int W0, int NW, int TICK>
  static __device__ __forceinline__ void bwd_stage(Ctx& c, int ntiles, float* ring) {
    const DevModel& M = c.P.M;
    constexpr int RL = NW * 8;
    const int ls = threadIdx.x - W0 * 32, itl = ls & 3, rl = ls >> 2;
    float w[4][H];
    {
      const float* W = c.wp + M.pw_off[l] + (itl * 4) * H;
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int j4 = 0; j4 < H / 4; ++j4) {
          const float4 v = *reinterpret_cast<const float4*>(W + ii * H + j4 * 4);
          w[ii][j4 * 4 + 0] = v.x; w[ii][j4 * 4 + 1] = v.y; w[ii][j4 * 4 + 2] = v.z; w[ii][j4 * 4 + 3] = v.w;
        }
    }
    for (int tick = 0; tick < ntiles + DEPTH - 1; ++tick) {
      const int t = tick - TICK;
      if (t >= 0 && t < ntiles) {
        const int slot = t % DEPTH;
        const float* D = dbuf(ring, slot, l);
        const float* A = abuf(ring, slot, l);
        float* Dp = dbuf(ring, slot, l - 1);
#pragma unroll
        for (int n = 0; n < T / RL; ++n) {
          const int r = rl + n * RL;
          float4 dd[H / 4];
#pragma unroll
          for (int j4 = 0; j4 < H / 4; ++j4) dd[j4] = *reinterpret_cast<const float4*>(D + r * H + j4 * 4);
          const float4 av = *reinterpret_cast<const float4*>(A + r * H + itl * 4);
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j4 = 0; j4 < H / 4; ++j4) {
            const float dv[4] = {dd[j4].x, dd[j4].y, dd[j4].z, dd[j4].w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) acc[ii] = fmaf(dv[jj], w[ii][j4 * 4 + jj], acc[ii]);
          }
          *reinterpret_cast<float4*>(Dp + r * H + itl * 4) =
              make_float4(acc[0] * act_bwd_from_value<ACT>(av.x), acc[1] * act_bwd_from_value<ACT>(av.y),
                          acc[2] * act_bwd_from_value<ACT>(av.z), acc[3] * act_bwd_from_value<ACT>(av.w));
        }
      }
      bar_block();
    }
  }

  // weight-gradient stage: acc[k][e] += a_l[r][k] * D_l[r][jt*4+e]; stationary accumulator [KP][4] (+ bias sums)
  template <int KP, int l, int W0, int NW, int TICK>
  static __device__ __forceinline__ void dw_stage(Ctx& c, long r0, long nrows, int ntiles, float* ring, float* scratch_out) {
    constexpr int RL = NW * 8;
    static_assert(RL == dw_row_lanes(l), "dW row-lane table out of sync");
    const int ls = threadIdx.x - W0 * 32, jt = ls & 3, rl = ls >> 2;
    float acc[KP][4], db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < KP; ++k) { acc[k][0] = 0.f; acc[k][1] = 0.f; acc[k][2] = 0.f; acc[k][3] = 0.f; }
    for (int tick = 0; tick < ntiles + DEPTH - 1; ++tick) {
      const int t = tick - TICK;
      if (t >= 0 && t < ntiles) {
        const int slot = t % DEPTH;
        const float* in = l == 0 ? c.xbuf + (long)t * T * FP : abuf(ring, slot, l);
        constexpr int sin_ = l == 0 ? FP : H;
        const float* D = dbuf(ring, slot, l);
#pragma unroll 2
        for (int n = 0; n < T / RL; ++n) {
          const int r = rl + n * RL;
          float4 a[KP / 4];
          if (l == 0 && !c.P.resident) {
            const long gr = (long)t * T + r < nrows ? (long)t * T + r : nrows - 1;
            const float4* xp = reinterpret_cast<const float4*>(c.P.X + (r0 + gr) * FP);
#pragma unroll
            for (int k4 = 0; k4 < KP / 4; ++k4) a[k4] = __ldg(xp + k4);
          } else {
#pragma unroll
            for (int k4 = 0; k4 < KP / 4; ++k4) a[k4] = *reinterpret_cast<const float4*>(in + r * sin_ + k4 * 4);
          }
          const float4 dd = *reinterpret_cast<const float4*>(D + r * H + jt * 4);
          db[0] += dd.x; db[1] += dd.y; db[2] += dd.z; db[3] += dd.w;
#pragma unroll
          for (int k4 = 0; k4 < KP / 4; ++k4) {
            const float av[4] = {a[k4].x, a[k4].y, a[k4].z, a[k4].w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              acc[k4 * 4 + kk][0] = fmaf(av[kk], dd.x, acc[k4 * 4 + kk][0]);
              acc[k4 * 4 + kk][1] = fmaf(av[kk], dd.y, acc[k4 * 4 + kk][1]);
              acc[k4 * 4 + kk][2] = fmaf(av[kk], dd.z, acc[k4 * 4 + kk][2]);
              acc[k4 * 4 + kk][3] = fmaf(av[kk], dd.w, acc[k4 * 4 + kk][3]);
            }
          }
        }
      }
      bar_block();
    }
    // partials -> scratch [rl][(KP+1)][16]  (row KP = bias sums); summed over rl by the whole block
    float* o = scratch_out + rl * ((KP + 1) * H) + jt * 4;
#pragma unroll
    for (int k = 0; k < KP; ++k) *reinterpret_cast<float4*>(o + k * H) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
    *reinterpret_cast<float4*>(o + KP * H) = make_float4(db[0], db[1], db[2], db[3]);
  }

  // head stage (thread per row): out = a W + b (K = 2), Gaussian log-likelihood, delta back through the
  // head, and the head's own weight gradient accumulated in registers.
  template <int W0, int TICK>
  static __device__ __forceinline__ float head_stage(Ctx& c, long r0, long nrows, int ntiles, float* ring,
                                                     float* scratch_out) {
    const KParams& P = c.P;
    const DevModel& M = P.M;
    const int l = NL - 1;
    const int ls = threadIdx.x - W0 * 32;   // 0..T-1 == row within the tile
    float w0[H], w1[H], g0[H], g1[H];
    {
      const float* W = c.wp + M.pw_off[l];   // [H][4] padded, columns 0,1 used
#pragma unroll
      for (int i = 0; i < H; ++i) { const float2 v = *reinterpret_cast<const float2*>(W + i * 4); w0[i] = v.x; w1[i] = v.y; g0[i] = 0.f; g1[i] = 0.f; }
    }
    const float b0 = c.wp[M.pb_off[l]], b1 = c.wp[M.pb_off[l] + 1];
    float gb0 = 0.f, gb1 = 0.f, ll_sum = 0.f;
    const float nb = M.n_batches;
    for (int tick = 0; tick < ntiles + DEPTH - 1; ++tick) {
      const int t = tick - TICK;
      if (t >= 0 && t < ntiles) {
        const int slot = t % DEPTH;
        const float* A = abuf(ring, slot, NH) + ls * H;
        float* D = dbuf(ring, slot, NH - 1) + ls * H;
        float a[H];
#pragma unroll
        for (int i4 = 0; i4 < H / 4; ++i4) {
          const float4 v = *reinterpret_cast<const float4*>(A + i4 * 4);
          a[i4 * 4] = v.x; a[i4 * 4 + 1] = v.y; a[i4 * 4 + 2] = v.z; a[i4 * 4 + 3] = v.w;
        }
        float mu = b0, s = b1;
#pragma unroll
        for (int i = 0; i < H; ++i) { mu = fmaf(a[i], w0[i], mu); s = fmaf(a[i], w1[i], s); }
        const long row = (long)t * T + ls;
        float dmu = 0.f, ds = 0.f;
        if (row < nrows) {
          const float yv = reinterpret_cast<const float*>(P.y)[r0 + row];
          // sigma = clip(exp(s), 1e-6, 1e6) (probabilistic.py:100): with sc = clip(s, ln 1e-6, ln 1e6) this is
          // log(2 pi sigma^2) = ln(2 pi) + 2 sc and 1 / sigma^2 = exp(-2 sc): one exp, no log, no divide on the
          // critical path of the pipeline's busiest stage (same values up to fp32 rounding).
          const float kLnClip = 13.815510557964274f;   // ln 1e6
          const float sc = fminf(fmaxf(s, -kLnClip), kLnClip);
          const float inside = (s > -kLnClip && s < kLnClip) ? 1.f : 0.f;
          const float inv_s2 = expf(-2.f * sc), res = yv - mu, q = res * res * inv_s2;
          float ll = (1.8378770664093453f + 2.f * sc + q) * -0.5f;
          dmu = res * inv_s2; ds = (q - 1.f) * inside;
          if (isnan(ll)) { ll = 0.f; dmu = 0.f; ds = 0.f; }
          ll_sum += ll;
          dmu *= nb; ds *= nb;
        }
        gb0 += dmu; gb1 += ds;
#pragma unroll
        for (int i4 = 0; i4 < H / 4; ++i4) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i4 * 4 + e;
            g0[i] = fmaf(a[i], dmu, g0[i]);
            g1[i] = fmaf(a[i], ds, g1[i]);
            o[e] = act_bwd_from_value<ACT>(a[i]) * fmaf(dmu, w0[i], ds * w1[i]);
          }
          *reinterpret_cast<float4*>(D + i4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      bar_block();
    }
    // partials -> scratch [lane][2H + 2]
    float* o = scratch_out + ls * HEAD_STRIDE;
#pragma unroll
    for (int i = 0; i < H; ++i) { o[2 * i] = g0[i]; o[2 * i + 1] = g1[i]; }
    o[2 * H] = gb0; o[2 * H + 1] = gb1;
    return ll_sum;
  }

  // ---- one gradient evaluation over this CTA's rows [r0, r1) ----------------------------------------
  static __device__ __forceinline__ void run(Ctx& c, long r0, long r1, float* gpart) {
    const KParams& P = c.P;
    const DevModel& M = P.M;
    float* ring = c.tile;
    const long nrows = r1 > r0 ? r1 - r0 : 0;
    const int ntiles = (int)((nrows + T - 1) / T);
    const int warp = threadIdx.x >> 5;
    // scratch of the final cross-lane reductions aliases the ring (dead once the pipeline has drained)
    float* scr = ring;
    float* scr_head = scr + NH * DW_STRIDE;
    static_assert(NH * DW_STRIDE + T * HEAD_STRIDE <= RING_FLOATS, "reduction scratch does not fit in the ring");
    float ll = 0.f;
    if constexpr (NL == 3) {
      // the stages that read the X tile get four warps, the hidden-layer GEMM stages two (per-warp barrier-wait timers:
      // with 2 / 4 / 2 / 4 / 2 / 2 the X stages were the critical roles and the four-warp stages idled ~40 % of a tick)
      if (warp < 4) fwd_stage<FP, 0, 0, 4, 0>(c, r0, nrows, ntiles, ring);
      else if (warp < 6) fwd_stage<H, 1, 4, 2, 1>(c, r0, nrows, ntiles, ring);
      else if (warp < 8) ll = head_stage<6, 2>(c, r0, nrows, ntiles, ring, scr_head);
      else if (warp < 10) bwd_stage<1, 8, 2, 3>(c, ntiles, ring);
      else if (warp < 14) dw_stage<FP, 0, 10, 4, 4>(c, r0, nrows, ntiles, ring, scr);
      else dw_stage<H, 1, 14, 2, 4>(c, r0, nrows, ntiles, ring, scr + DW_STRIDE);
    } else {
      if (warp < 1) fwd_stage<FP, 0, 0, 1, 0>(c, r0, nrows, ntiles, ring);
      else if (warp < 3) fwd_stage<H, 1, 1, 2, 1>(c, r0, nrows, ntiles, ring);
      else if (warp < 5) fwd_stage<H, 2, 3, 2, 2>(c, r0, nrows, ntiles, ring);
      else if (warp < 6) ll = head_stage<5, 3>(c, r0, nrows, ntiles, ring, scr_head);
      else if (warp < 8) bwd_stage<2, 6, 2, 4>(c, ntiles, ring);
      else if (warp < 10) bwd_stage<1, 8, 2, 5>(c, ntiles, ring);
      else if (warp < 11) dw_stage<FP, 0, 10, 1, 6>(c, r0, nrows, ntiles, ring, scr);
      else if (warp < 13) dw_stage<H, 1, 11, 2, 6>(c, r0, nrows, ntiles, ring, scr + DW_STRIDE);
      else if (warp < 15) dw_stage<H, 2, 13, 2, 6>(c, r0, nrows, ntiles, ring, scr + 2 * DW_STRIDE);
      else { for (int tick = 0; tick < ntiles + DEPTH - 1; ++tick) bar_block(); }
    }
    __syncthreads();
    // sum the dW partials over row lanes and scatter to the flat gradient layout
#pragma unroll
    for (int l = 0; l < NH; ++l) {
      const int KP = l == 0 ? FP : H, IN = M.dims[l];
      const int RL = dw_row_lanes(l), stride = (KP + 1) * H;
      const float* s0 = scr + l * DW_STRIDE;
      for (int o = threadIdx.x; o < (KP + 1) * H; o += NT) {
        float s = 0.f;
#pragma unroll 8
        for (int q = 0; q < RL; ++q) s += s0[q * stride + o];
        const int k = o / H, j = o % H;
        if (k < IN) gpart[M.kern_off[l] + k * H + j] = s;
        else if (k == KP) gpart[M.bias_off[l] + j] = s;
      }
    }
    {
      const int l = NL - 1;
      for (int o = threadIdx.x; o < 2 * H + 2; o += NT) {
        float s = 0.f;
#pragma unroll 8
        for (int q = 0; q < T; ++q) s += scr_head[q * HEAD_STRIDE + o];
        if (o < 2 * H) gpart[M.kern_off[l] + o] = s;     // kernel [H][2] row-major == interleaved (g0,g1)
        else gpart[M.bias_off[l] + (o - 2 * H)] = s;
      }
    }
    float v[1] = {ll};
    block_sum<1, NT>(v, c.red, c.phase);
    if (threadIdx.x == 0) gpart[P.dS] = v[0] * M.n_batches;   // probabilistic.py:136
    __syncthreads();   // the ring (== scratch) is reused by the next evaluation / lppd fold
  }
};
