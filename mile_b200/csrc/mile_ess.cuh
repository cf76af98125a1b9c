// mile_ess.cuh -- effective sample size of the phase-3 positions on the device (src/training/warmup.py:442-463 ->
// blackjax.diagnostics.effective_sample_size, one chain per call as the reference uses it: flat_samples[None, ...]).
//
// The reference transforms every series with an FFT.  On a B200 the direct autocovariance is cheaper than staging an FFT:
// Geyer's initial-positive-sequence rule only ever reads the lags up to the first non-positive pair sum, so the lags are
// produced lazily (a few hundred per series for a tuned MCLMC chain instead of all n), each as one warp-wide dot product
// over the series held in shared memory.  Worst case (a series that never turns negative) is n^2 / 2 FMAs = 12.5 M for
// n = 5000: still microseconds per series.
//   1. ess_transpose_kernel: positions [n][C][d] (with optional parameter / sample index lists, the reference's
//      "> 2000 parameters: random subset", "> 10000 samples: linspace" rules) -> series-major [C * d_sel][n_sel], coalesced
//      on both sides through a 32 x 33 shared tile;
//   2. ess_series_kernel: one CTA per series: mean, centring, lazy autocovariance, the initial positive / initial monotone
//      sequence estimators and tau, exactly in the order of the reference's formulas (fp32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct EssParams {
  const float* pos;      // [n_total][C][d]
  float* series;         // [C * d_sel][n]   (scratch)
  const int* pidx;       // [d_sel] or null (identity)
  const int* sidx;       // [n] or null (identity)
  float* ess;            // [C][d_sel]
  int n, C, d, d_sel;
};

__global__ void ess_transpose_kernel(const EssParams E) {
  __shared__ float tile[32][33];
  const long n_series = (long)E.C * E.d_sel;
  const long s0 = (long)blockIdx.x * 32;   // series block
  const int i0 = blockIdx.y * 32;          // sample block
  // read: 32 consecutive series (adjacent parameters of one chain mostly) for each of 32 samples
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r;
    const long s = s0 + threadIdx.x;
    float v = 0.f;
    if (i < E.n && s < n_series) {
      const int c = (int)(s / E.d_sel), j = (int)(s % E.d_sel);
      const int p = E.pidx ? E.pidx[j] : j;
      const long row = E.sidx ? E.sidx[i] : i;
      v = E.pos[(row * E.C + c) * E.d + p];
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long s = s0 + r;
    const int i = i0 + threadIdx.x;
    if (s < n_series && i < E.n) E.series[s * E.n + i] = tile[threadIdx.x][r];
  }
}

#define ESS_THREADS 256
#define ESS_LAGS_PER_ROUND 32   // 8 warps x 4 lags

__global__ void __launch_bounds__(ESS_THREADS) ess_series_kernel(const EssParams E) {
  extern __shared__ float sm[];
  const int n = E.n, n_even = n - (n & 1), T = n_even / 2;
  float* xs = sm;              // [n] centred series
  float* ac = sm + n;          // [n_even] autocovariance (filled lazily)
  __shared__ float red[ESS_THREADS / 32];
  __shared__ int stop_pair;    // first pair index with a non-positive sum, or -1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* src = E.series + (long)blockIdx.x * n;
  float s = 0.f;
  for (int i = tid; i < n; i += ESS_THREADS) { const float v = src[i]; xs[i] = v; s += v; }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  if (tid == 0) stop_pair = -1;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < ESS_THREADS / 32; ++w) mean += red[w];
  mean /= (float)n;
  float q = 0.f;
  for (int i = tid; i < n; i += ESS_THREADS) { const float v = xs[i] - mean; xs[i] = v; q += v * v; }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __syncthreads();
  if (lane == 0) red[warp] = q;
  __syncthreads();
  float ac0 = 0.f;
  for (int w = 0; w < ESS_THREADS / 32; ++w) ac0 += red[w];
  ac0 /= (float)n;
  const float fn = (float)n;
  const float var0 = ac0 * fn / (fn - 1.0f);          // mean_var0
  const float wvar = var0 * (fn - 1.0f) / fn;         // weighted_var (one chain)
  if (tid == 0 && n_even > 0) ac[0] = ac0;
  // ---- lazy autocovariance: lags in rounds of 32 until the first non-positive pair sum rho[2k] + rho[2k+1] ----------
  int have = 1;   // lags [0, have) are in ac
  int first_bad = -1;
  while (have < n_even) {
    const int base = have;
#pragma unroll 1
    for (int u = 0; u < ESS_LAGS_PER_ROUND / (ESS_THREADS / 32); ++u) {
      const int t = base + u * (ESS_THREADS / 32) + warp;
      if (t < n_even) {
        float a = 0.f;
        const int m = n - t;
        for (int i = lane; i < m; i += 32) a = fmaf(xs[i], xs[i + t], a);
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) ac[t] = a / fn;
      }
    }
    const int upto = min(n_even, base + ESS_LAGS_PER_ROUND);
    __syncthreads();
    if (tid == 0) {
      // complete pairs inside [first unchecked pair, upto / 2)
      for (int k = base / 2; 2 * k + 1 < upto; ++k) {   // pairs below base / 2 were checked in the earlier rounds
        const float re = k == 0 ? 1.f : 1.f - (var0 - ac[2 * k]) / wvar;
        const float ro = 1.f - (var0 - ac[2 * k + 1]) / wvar;
        if (!(re + ro > 0.f)) { stop_pair = k; break; }
      }
    }
    __syncthreads();
    have = upto;
    first_bad = stop_pair;
    // the even term one pair past the last positive pair is still read ("improve estimation" step): it belongs to the
    // pair that ended the sequence, so it is already there
    if (first_bad >= 0) break;
  }
  if (tid != 0) return;
  // ---- Geyer initial positive sequence + initial monotone sequence (thread 0; a few hundred terms) --------------------
  // mask[k] = all pairs 0..k positive; max_t = last such k (0 when even pair 0 fails); sel = max_t + 1
  const int n_pos = first_bad >= 0 ? first_bad : T;          // pairs 0 .. n_pos-1 are positive
  const int max_t = n_pos > 0 ? n_pos - 1 : 0;
  const int sel = max_t + 1;
  auto rho = [&](int t) -> float { return t == 0 ? 1.f : 1.f - (var0 - ac[t]) / wvar; };
  const int last_k = sel < T ? sel : T - 1;                  // JAX clamps the out-of-bounds gather
  float run_min = 0.f, total = 0.f, last_even_f = 0.f;
  for (int k = 0; k <= last_k; ++k) {
    float re, ro;
    if (k < n_pos) { re = rho(2 * k); ro = rho(2 * k + 1); }
    else if (k == sel) { const float r = rho(2 * k); re = r > 0.f ? r : 0.f; ro = 0.f; }   // only reached when sel < T
    else { re = 0.f; ro = 0.f; }
    const float sum = re + ro;
    if (k == 0) run_min = sum; else run_min = fminf(run_min, sum);
    if (sum > run_min) { re = run_min / 2.f; ro = run_min / 2.f; }
    total += re + ro;
    if (k == last_k) last_even_f = re;
  }
  float tau = -1.f + 2.f * total - last_even_f;
  tau = fmaxf(tau, 1.f / log10f(fn));
  E.ess[blockIdx.x] = fn / tau;
}

// ---- all chains pooled per parameter (the report's ESS: src/inference/metrics.py -> blackjax / numpyro effective_sample_size on
// [chains, samples, dim]): autocovariance averaged over the chains, between-chain variance of the chain means in the
// normalisation, ess_raw = chains x samples.  One CTA per parameter; the C centred series of the parameter sit in shared
// memory ([C][n] + [n_even] floats: up to ~4500 samples for 12 chains; longer runs use the torch form in diagnostics.py).
__global__ void __launch_bounds__(ESS_THREADS) ess_pooled_kernel(const EssParams E) {
  extern __shared__ float sm[];
  const int n = E.n, C = E.C, n_even = n - (n & 1), T = n_even / 2;
  float* xs = sm;                    // [C][n] centred series
  float* ac = sm + (size_t)C * n;    // [n_even] mean autocovariance (filled lazily)
  __shared__ float red[ESS_THREADS / 32];
  __shared__ float cmean[64];
  __shared__ int stop_pair;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x;          // selected parameter
  float q = 0.f;
  for (int c = 0; c < C; ++c) {
    const float* src = E.series + ((long)c * E.d_sel + j) * n;
    float* x = xs + (size_t)c * n;
    float s = 0.f;
    for (int i = tid; i < n; i += ESS_THREADS) { const float v = src[i]; x[i] = v; s += v; }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();                 // (red free again)
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float mean = 0.f;
    for (int w = 0; w < ESS_THREADS / 32; ++w) mean += red[w];
    mean /= (float)n;
    if (tid == 0) cmean[c] = mean;
    for (int i = tid; i < n; i += ESS_THREADS) { const float v = x[i] - mean; x[i] = v; q += v * v; }
  }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __syncthreads();
  if (lane == 0) red[warp] = q;
  if (tid == 0) stop_pair = -1;
  __syncthreads();
  const float fn = (float)n, fc = (float)C;
  float ac0 = 0.f;
  for (int w = 0; w < ESS_THREADS / 32; ++w) ac0 += red[w];
  ac0 = ac0 / fn / fc;                                   // mean over chains of acov_c[0]
  float vb = 0.f;                                        // variance of the chain means (ddof = 1)
  if (C > 1) {
    float m = 0.f;
    for (int c = 0; c < C; ++c) m += cmean[c];
    m /= fc;
    for (int c = 0; c < C; ++c) { const float dlt = cmean[c] - m; vb += dlt * dlt; }
    vb /= (fc - 1.f);
  }
  const float var0 = ac0 * fn / (fn - 1.0f);
  const float wvar = var0 * (fn - 1.0f) / fn + vb;
  if (tid == 0 && n_even > 0) ac[0] = ac0;
  int have = 1, first_bad = -1;
  while (have < n_even) {
    const int base = have;
#pragma unroll 1
    for (int u = 0; u < ESS_LAGS_PER_ROUND / (ESS_THREADS / 32); ++u) {
      const int t = base + u * (ESS_THREADS / 32) + warp;
      if (t < n_even) {
        float a = 0.f;
        const int m = n - t;
        for (int c = 0; c < C; ++c) {
          const float* x = xs + (size_t)c * n;
          for (int i = lane; i < m; i += 32) a = fmaf(x[i], x[i + t], a);
        }
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) ac[t] = a / fn / fc;
      }
    }
    const int upto = min(n_even, base + ESS_LAGS_PER_ROUND);
    __syncthreads();
    if (tid == 0) {
      for (int k = base / 2; 2 * k + 1 < upto; ++k) {
        const float re = k == 0 ? 1.f : 1.f - (var0 - ac[2 * k]) / wvar;
        const float ro = 1.f - (var0 - ac[2 * k + 1]) / wvar;
        if (!(re + ro > 0.f)) { stop_pair = k; break; }
      }
    }
    __syncthreads();
    have = upto;
    first_bad = stop_pair;
    if (first_bad >= 0) break;
  }
  if (tid != 0) return;
  const int n_pos = first_bad >= 0 ? first_bad : T;
  const int max_t = n_pos > 0 ? n_pos - 1 : 0;
  const int sel = max_t + 1;
  auto rho = [&](int t) -> float { return t == 0 ? 1.f : 1.f - (var0 - ac[t]) / wvar; };
  const int last_k = sel < T ? sel : T - 1;
  float run_min = 0.f, total = 0.f, last_even_f = 0.f;
  for (int k = 0; k <= last_k; ++k) {
    float re, ro;
    if (k < n_pos) { re = rho(2 * k); ro = rho(2 * k + 1); }
    else if (k == sel) { const float r = rho(2 * k); re = r > 0.f ? r : 0.f; ro = 0.f; }
    else { re = 0.f; ro = 0.f; }
    const float sum = re + ro;
    if (k == 0) run_min = sum; else run_min = fminf(run_min, sum);
    if (sum > run_min) { re = run_min / 2.f; ro = run_min / 2.f; }
    total += re + ro;
    if (k == last_k) last_even_f = re;
  }
  const float ess_raw = fn * fc;
  float tau = -1.f + 2.f * total - last_even_f;
  tau = fmaxf(tau, 1.f / log10f(ess_raw));
  E.ess[j] = ess_raw / tau;
}
