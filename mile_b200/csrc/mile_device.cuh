// mile_device.cuh -- device-side building blocks of the fused MCLMC ensemble kernel.
//
// Reference arithmetic restated here (file:line under /root/reference, or blackjax 1.2.2):
//   FCN forward                      src/flax_building_blocks/basic.py:41-61
//   log-likelihood / log-prior       src/training/probabilistic.py:92-138, src/training/priors.py:101-128
//   ESH momentum update, McLachlan   blackjax/mcmc/integrators.py (isokinetic_mclachlan)
//   partial momentum refresh         blackjax/mcmc/integrators.py (partially_refresh_momentum)
//   step-size predictor, handle_nans src/training/warmup.py:276-352,468-483
//   pointwise LPPD                   src/inference/metrics.py:247-293
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "../../include/mile_b200.h"

namespace cg = cooperative_groups;

// Optional phase timers (build with -DMILE_PROFILE; tools/phase_profile.py): thread 0 of CTA 0
// accumulates clock64() deltas per phase into g_prof[].
#ifdef MILE_PROFILE
__device__ unsigned long long g_prof[32];
__device__ long long g_prof_t;
#define PROF_DECL do { if (threadIdx.x == 0 && blockIdx.x == 0) g_prof_t = clock64(); } while (0)
#define PROF(i)                                                                            \
  do {                                                                                     \
    if (threadIdx.x == 0 && blockIdx.x == 0) {                                             \
      long long n_ = clock64(); g_prof[i] += (unsigned long long)(n_ - g_prof_t); g_prof_t = n_; \
    }                                                                                      \
  } while (0)
#else
#define PROF_DECL
#define PROF(i)
#endif

#define MILE_THREADS 256            // block size of the generic kernel (the fast kernel uses 512)

// ------------------------------------------------------------------------------------
// Device-side model description (kernel parameter, lives in the constant bank).
// Activations and deltas are kept row-major in shared memory: buffer l holds
// [rows][stride_l] floats, stride_l = width padded to 4.  128-bit shared accesses are served
// per quarter-warp (8 lanes = 4 neuron tiles x 2 rows in the tile GEMMs), so dense rows of
// <= 16 floats never conflict (measured: tools/lds_microbench.cu, profiles/r1_*).
// ------------------------------------------------------------------------------------
struct DevModel {
  int F, NL, act, task, prior, d;
  float prior_loc, prior_scale, n_batches;
  int dims[MILE_MAX_LAYERS + 1];   // dims[0] = F, dims[l+1] = widths[l]
  int dimp[MILE_MAX_LAYERS + 1];   // padded to a multiple of 4
  int bias_off[MILE_MAX_LAYERS];   // flat-vector offsets
  int kern_off[MILE_MAX_LAYERS];
  int pb_off[MILE_MAX_LAYERS];     // padded shared-memory parameter image: bias[OUTP]
  int pw_off[MILE_MAX_LAYERS];     //                                       W[INP][OUTP]
  int pwt_off[MILE_MAX_LAYERS];    //                                       W^T[OUTP][INP] (layers >= 1, for the backward pass)
  int psize;                       // floats in the padded image
  int sA[MILE_MAX_LAYERS + 1];     // row stride of activation buffer l (l=0: X)
  int a_off[MILE_MAX_LAYERS + 1];  // float offset of activation buffer l inside a tile (l>=1)
  int d_off[MILE_MAX_LAYERS];      // float offset of delta buffer l (stride sA[l+1])
  int tile_floats;                 // floats per tile (excluding X when resident)
  int TR;                          // max rows per tile (multiple of 32)
};

// ------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum of NV values; every thread returns the same totals.
// scratch: [2][NV_MAX=4][NT/32] floats, phase toggles between the two halves so
// that a single __syncthreads per reduction suffices.
template <int NV, int NT, int BAR = 0>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* scratch, int& phase) {
  // BAR == 0: the whole block (NT threads) takes part and synchronises with bar 0;
  // BAR != 0: only the first NT threads of the block take part and use the named barrier BAR.
  constexpr int MILE_NWARPS = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* buf = scratch + phase * (4 * MILE_NWARPS);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float s = warp_sum(v[k]);
    if (lane == 0) buf[k * MILE_NWARPS + warp] = s;
  }
  if (BAR == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float s = 0.f;
    if (MILE_NWARPS % 4 == 0) {   // same summation order, a quarter of the shared-memory loads
#pragma unroll
      for (int w = 0; w < MILE_NWARPS; w += 4) {
        const float4 q = *reinterpret_cast<const float4*>(buf + k * MILE_NWARPS + w);
        s += q.x; s += q.y; s += q.z; s += q.w;
      }
    } else {
#pragma unroll
      for (int w = 0; w < MILE_NWARPS; ++w) s += buf[k * MILE_NWARPS + w];
    }
    v[k] = s;
  }
  phase ^= 1;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based so every CTA of a cluster draws the
// same noise for the same (seed, chain, step, element).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two standard normals for (seed, chain, step, stream, element pair) via Box-Muller: elements 2p and 2p + 1 of a
// noise vector are the cosine and sine branches of ONE Philox draw (counter word 0 = p).
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint32_t chain, uint64_t step, uint32_t stream, uint32_t pair,
                                               float& z0, float& z1) {
  uint32_t r[4];
  philox4x32_10(pair, (uint32_t)step, (uint32_t)(step >> 32), stream, (uint32_t)seed ^ (chain * 0x9E3779B9u),
                (uint32_t)(seed >> 32) + chain, r);
  const float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  const float rad = sqrtf(-2.0f * logf(u1));
  z0 = rad * c; z1 = rad * s;
}
// One standard normal for (seed, chain, step, stream, element).
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t chain, uint64_t step,
                                               uint32_t stream, uint32_t elem) {
  float z0, z1;
  philox_normal2(seed, chain, step, stream, elem >> 1, z0, z1);
  return (elem & 1u) ? z1 : z0;
}

// activation value and derivative (src/config/models/base.py:24-37; jax.nn semantics)
__device__ __forceinline__ void act_eval(int act, float z, float& a, float& da) {
  switch (act) {
    case MILE_ACT_RELU: a = fmaxf(z, 0.f); da = z > 0.f ? 1.f : 0.f; break;
    case MILE_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-z)); a = s; da = s * (1.f - s); } break;
    case MILE_ACT_TANH: { float t = tanhf(z); a = t; da = 1.f - t * t; } break;
    case MILE_ACT_GELU: {
      const float c = 0.7978845608028654f, k = 0.044715f;
      float t = tanhf(c * (z + k * z * z * z));
      a = 0.5f * z * (1.f + t);
      da = 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * c * (1.f + 3.f * k * z * z);
    } break;
    case MILE_ACT_LEAKY_RELU: a = z >= 0.f ? z : 0.01f * z; da = z >= 0.f ? 1.f : 0.01f; break;
    default: a = z; da = 1.f; break;
  }
}

#define MILE_LOG_2PI 1.8378770664093453f

// ------------------------------------------------------------------------------------
// Tile GEMMs through shared memory.  Thread tile = 4 rows x 4 neurons; rows of a thread are
// {q, q+Q, q+2Q, q+3Q} so that the 8 row-lanes of a warp read consecutive rows.
// ------------------------------------------------------------------------------------

// Z = A_in W + b ; hidden layers: A_out = act(Z), D_out = act'(Z) ; last layer: A_out = Z.
template <int NT>
__device__ __forceinline__ void fwd_layer(const DevModel& M, int l, const float* __restrict__ Wp,
                                          const float* __restrict__ Ain, int sin_,
                                          float* __restrict__ Aout, float* __restrict__ Dout,
                                          int sout, int Q, bool last) {
  const int IN = M.dims[l], INP = M.dimp[l], OUTP = M.dimp[l + 1], njt = OUTP >> 2;
  const float* __restrict__ W = Wp + M.pw_off[l];
  const float* __restrict__ B = Wp + M.pb_off[l];
  const int act = M.act;
  for (int it = threadIdx.x; it < Q * njt; it += NT) {
    const int jt = it % njt, q = it / njt;
    const float4 b4 = *reinterpret_cast<const float4*>(B + jt * 4);
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { acc[r][0] = b4.x; acc[r][1] = b4.y; acc[r][2] = b4.z; acc[r][3] = b4.w; }
    const float* ap = Ain + q * sin_;
    const float* wp = W + jt * 4;
    const int rstep = Q * sin_;
#pragma unroll 2
    for (int k = 0; k < INP; k += 4) {   // pads of A_in rows are finite, pads of W rows are zero
      float4 a[4], w[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(ap + r * rstep + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(wp + (k + kk) * OUTP);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float av[4] = {a[r].x, a[r].y, a[r].z, a[r].w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc[r][0] = fmaf(av[kk], w[kk].x, acc[r][0]);
          acc[r][1] = fmaf(av[kk], w[kk].y, acc[r][1]);
          acc[r][2] = fmaf(av[kk], w[kk].z, acc[r][2]);
          acc[r][3] = fmaf(av[kk], w[kk].w, acc[r][3]);
        }
      }
    }
    (void)IN;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = q + r * Q;
      float4 o, dd;
      if (!last) {
        act_eval(act, acc[r][0], o.x, dd.x);
        act_eval(act, acc[r][1], o.y, dd.y);
        act_eval(act, acc[r][2], o.z, dd.z);
        act_eval(act, acc[r][3], o.w, dd.w);
        *reinterpret_cast<float4*>(Dout + row * sout + jt * 4) = dd;
      } else {
        o = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      }
      *reinterpret_cast<float4*>(Aout + row * sout + jt * 4) = o;
    }
  }
}

// D_{l-1} <- D_{l-1}(=act') * (D_l W_l^T), using the transposed weight image W^T[OUTP][INP] so that
// the 4 input-neuron tiles of a quarter-warp read one contiguous 64 B segment (no bank conflicts).
template <int NT>
__device__ __forceinline__ void bwd_layer(const DevModel& M, int l, const float* __restrict__ Wp,
                                          const float* __restrict__ Dl, int sl,
                                          float* __restrict__ Dprev, int sp, int Q) {
  const int OUTP = M.dimp[l + 1], INP = M.dimp[l], nit = INP >> 2;
  const float* __restrict__ WT = Wp + M.pwt_off[l];
  for (int it = threadIdx.x; it < Q * nit; it += NT) {
    const int itl = it % nit, q = it / nit;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
    const float* dp = Dl + q * sl;
    const float* wp = WT + itl * 4;
    const int rstep = Q * sl;
#pragma unroll 2
    for (int j = 0; j < OUTP; j += 4) {
      float4 dd[4], w[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) dd[r] = *reinterpret_cast<const float4*>(dp + r * rstep + j);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) w[jj] = *reinterpret_cast<const float4*>(wp + (j + jj) * INP);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float dv[4] = {dd[r].x, dd[r].y, dd[r].z, dd[r].w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          acc[r][0] = fmaf(dv[jj], w[jj].x, acc[r][0]);
          acc[r][1] = fmaf(dv[jj], w[jj].y, acc[r][1]);
          acc[r][2] = fmaf(dv[jj], w[jj].z, acc[r][2]);
          acc[r][3] = fmaf(dv[jj], w[jj].w, acc[r][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4* p = reinterpret_cast<float4*>(Dprev + (q + r * Q) * sp + itl * 4);
      float4 g = *p;
      g.x *= acc[r][0]; g.y *= acc[r][1]; g.z *= acc[r][2]; g.w *= acc[r][3];
      *p = g;
    }
  }
}

// Per-row log-likelihood and d/d(out) (probabilistic.py:93-109).  Writes the output-layer
// delta (scaled by n_batches; zero for padded rows / NaN rows) and returns the thread's
// partial sum of log-likelihood terms.
template <int NT>
__device__ __forceinline__ float loglik_rows(const DevModel& M, const float* __restrict__ Out,
                                             float* __restrict__ Dout, int so, const void* __restrict__ y,
                                             long row0, int nvalid, int rows_pad, float sig_lo = 1e-6f,
                                             float* metric = nullptr) {
  // sig_lo: lower clip of sigma (1e-6 in the log-posterior, probabilistic.py:100; 1e-5 in the warm-start training loss,
  // where GaussianNLLLoss clips once more, src/inference/metrics.py:332).  metric (optional, per-thread accumulator):
  // sum of squared errors (regression) / number of correct argmax predictions (classification).
  const int K = M.dims[M.NL], KP = M.dimp[M.NL];
  float part = 0.f;
  for (int r = threadIdx.x; r < rows_pad; r += NT) {
    const float* o = Out + r * so;
    float* dd = Dout + r * so;
    if (r >= nvalid) {
      for (int k = 0; k < KP; ++k) dd[k] = 0.f;
      continue;
    }
    float ll;
    if (M.task == MILE_TASK_REGRESSION) {
      const float yv = reinterpret_cast<const float*>(y)[row0 + r];
      const float mu = o[0], s = o[1];
      const float e = expf(s);
      const float sigma = fminf(fmaxf(e, sig_lo), 1e6f);
      const float inside = (e > sig_lo && e < 1e6f) ? 1.f : 0.f;
      const float s2 = sigma * sigma;
      const float res = yv - mu;
      const float q = res * res / s2;
      ll = (logf(6.283185307179586f * s2) + q) / -2.f;
      float dmu = res / s2, ds = (q - 1.f) * inside;
      if (isnan(ll)) { ll = 0.f; dmu = 0.f; ds = 0.f; }
      if (metric) *metric += res * res;
      dd[0] = dmu * M.n_batches; dd[1] = ds * M.n_batches;
      for (int k = 2; k < KP; ++k) dd[k] = 0.f;
    } else {
      const int yi = reinterpret_cast<const int*>(y)[row0 + r];
      float m = o[0];
      for (int k = 1; k < K; ++k) m = fmaxf(m, o[k]);
      float se = 0.f;
      for (int k = 0; k < K; ++k) se += expf(o[k] - m);
      const float lse = m + logf(se);
      ll = o[yi] - lse;
      const bool bad = isnan(ll);
      if (metric) { int am = 0; for (int k = 1; k < K; ++k) if (o[k] > o[am]) am = k; *metric += am == yi ? 1.f : 0.f; }
      const float inv = 1.f / se;
      for (int k = 0; k < K; ++k) {
        float g = -expf(o[k] - m) * inv + (k == yi ? 1.f : 0.f);
        dd[k] = bad ? 0.f : g * M.n_batches;
      }
      for (int k = K; k < KP; ++k) dd[k] = 0.f;
      if (bad) ll = 0.f;
    }
    part += ll;
  }
  return part;
}

// Pointwise posterior-predictive log-density (metrics.py:278-293).
__device__ __forceinline__ float pointwise_lppd_row(const DevModel& M, const float* o, const void* y, long row) {
  if (M.task == MILE_TASK_REGRESSION) {
    const float yv = reinterpret_cast<const float*>(y)[row];
    const float sigma = fminf(fmaxf(expf(o[1]), 1e-6f), 1e6f);
    const float res = yv - o[0];
    return -(res * res) / (2.f * sigma * sigma) - logf(sigma) - 0.9189385332046727f;
  }
  const int K = M.dims[M.NL];
  const int yi = reinterpret_cast<const int*>(y)[row];
  float m = o[0];
  for (int k = 1; k < K; ++k) m = fmaxf(m, o[k]);
  float se = 0.f;
  for (int k = 0; k < K; ++k) se += expf(o[k] - m);
  return o[yi] - (m + logf(se));
}

// Weight-gradient accumulators: one 4x4 tile of dW (and 4 bias sums) per layer and thread,
// kept in registers across all row tiles of a gradient evaluation.
template <int NLMAX>
struct DwAcc {
  float w[NLMAX][16];
  float b[NLMAX][4];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int l = 0; l < NLMAX; ++l) {
#pragma unroll
      for (int e = 0; e < 16; ++e) w[l][e] = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) b[l][e] = 0.f;
    }
  }
};

struct DwRole { int ntile, nch, tile, chunk, itl, jt; bool active; };

template <int NT>
__device__ __forceinline__ DwRole dw_role(const DevModel& M, int l) {
  DwRole r;
  const int nit = M.dimp[l] >> 2, njt = M.dimp[l + 1] >> 2;
  r.ntile = nit * njt;
  r.nch = NT / r.ntile;          // host guarantees ntile <= NT
  if (r.nch > 1) r.nch &= ~1;              // even, so that chunk pairs cover 8-row groups
  r.tile = threadIdx.x % r.ntile;
  r.chunk = threadIdx.x / r.ntile;
  r.active = r.chunk < r.nch;
  r.jt = r.tile % njt;
  r.itl = r.tile / njt;
  return r;
}

// dW_l += A_{l}^T D_l over the rows of the current tile (rows_pad = 4Q rows, invalid rows have D = 0).
template <int NLMAX, int NT>
__device__ __forceinline__ void dw_accumulate(const DevModel& M, int l, DwAcc<NLMAX>& acc, int la,
                                              const float* __restrict__ A, int sa,
                                              const float* __restrict__ D, int sd, int rows_pad) {
  const DwRole R = dw_role<NT>(M, l);
  if (!R.active) return;
  const float* ap = A + R.itl * 4;
  const float* dp = D + R.jt * 4;
  // chunk pair p = chunk/2 handles 8-row groups p, p+npairs, ...; parity picks rows +0..3 / +4..7
  const int npairs = R.nch > 1 ? (R.nch >> 1) : 1;
  const int rows_per_iter = R.nch > 1 ? 8 : 4;
  int base = (R.nch > 1) ? ((R.chunk >> 1) * 8 + (R.chunk & 1) * 4) : 0;
  const int stride_rows = npairs * rows_per_iter;
  for (; base < rows_pad; base += stride_rows) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int row = base + t;
      const float4 a = *reinterpret_cast<const float4*>(ap + row * sa);
      const float4 dd = *reinterpret_cast<const float4*>(dp + row * sd);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc.w[la][i * 4 + 0] = fmaf(av[i], dd.x, acc.w[la][i * 4 + 0]);
        acc.w[la][i * 4 + 1] = fmaf(av[i], dd.y, acc.w[la][i * 4 + 1]);
        acc.w[la][i * 4 + 2] = fmaf(av[i], dd.z, acc.w[la][i * 4 + 2]);
        acc.w[la][i * 4 + 3] = fmaf(av[i], dd.w, acc.w[la][i * 4 + 3]);
      }
      if (R.itl == 0) {
        acc.b[la][0] += dd.x; acc.b[la][1] += dd.y; acc.b[la][2] += dd.z; acc.b[la][3] += dd.w;
      }
    }
  }
}
