/*
 * mile_oracle.c -- plain-C restatement of the MILE MCLMC sampling hot path (CPU, fp32).
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (same role and same caveat as oracle/mile_oracle.py:
 * PARITY UNPINNED -- the reference ships no golden vectors and its sampler arithmetic lives in the
 * un-vendored blackjax==1.2.2; see the header of mile_oracle.py).  Nothing under mile_b200/ links or
 * loads this file.  It exists because the numpy oracle is too slow to serve as the "reference CPU
 * implementation on all host threads" leg of bench.py: here one chain runs per OpenMP thread,
 * mirroring the reference's one-virtual-XLA-device-per-chain pmap (train.py:16,
 * src/training/sampling.py:181-184).
 *
 * Follows, function by function:
 *   mo_logpost_value_and_grad   src/training/probabilistic.py:92-138, src/training/priors.py:101-128,
 *                               src/flax_building_blocks/basic.py:41-61 (hand-differentiated)
 *   mo_esh_update / mo_step     blackjax 1.2.2 mcmc/integrators.py (isokinetic_mclachlan,
 *                               esh_dynamics_momentum_update_one_step, partially_refresh_momentum),
 *                               mcmc/mclmc.py (kernel) -- SURVEY.md Appendix A
 *   mo_run_sampling             src/training/sampling.py:134-177 (scan of sampler.step)
 * Validated against oracle/mile_oracle.py in tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MO_MAX_LAYERS 12
#define MO_MAX_WIDTH 512

typedef struct {
  int32_t n_features, n_layers;
  int32_t widths[MO_MAX_LAYERS];
  int32_t bias_off[MO_MAX_LAYERS], kernel_off[MO_MAX_LAYERS];
  int32_t activation; /* 0 identity 1 relu 2 sigmoid 3 tanh 4 gelu 5 leaky_relu */
  int32_t task;       /* 0 regression 1 classification */
  int32_t prior;      /* 0 normal 1 laplace */
  float prior_loc, prior_scale, n_batches;
} mo_spec;

static int mo_nparams(const mo_spec* s) {
  int d = 0, in = s->n_features;
  for (int l = 0; l < s->n_layers; ++l) { d += s->widths[l] * (in + 1); in = s->widths[l]; }
  return d;
}

static inline void act_eval(int act, float z, float* a, float* da) {
  switch (act) {
    case 1: *a = z > 0.f ? z : 0.f; *da = z > 0.f ? 1.f : 0.f; break;
    case 2: { float s = 1.f / (1.f + expf(-z)); *a = s; *da = s * (1.f - s); } break;
    case 3: { float t = tanhf(z); *a = t; *da = 1.f - t * t; } break;
    case 4: {
      const float c = 0.7978845608028654f, k = 0.044715f;
      float t = tanhf(c * (z + k * z * z * z));
      *a = 0.5f * z * (1.f + t);
      *da = 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * c * (1.f + 3.f * k * z * z);
    } break;
    case 5: *a = z >= 0.f ? z : 0.01f * z; *da = z >= 0.f ? 1.f : 0.01f; break;
    default: *a = z; *da = 1.f; break;
  }
}

/* Likelihood part of value_and_grad over the rows [n0, n1): grad [d] overwritten with the partial gradient, returns the
 * partial sum of log-likelihood terms. */
static float mo_loglik_rows(const mo_spec* s, const float* theta, const float* X, const void* y, int64_t n0, int64_t n1,
                            float* grad) {
  const int NL = s->n_layers, d = mo_nparams(s);
  int dims[MO_MAX_LAYERS + 1];
  dims[0] = s->n_features;
  for (int l = 0; l < NL; ++l) dims[l + 1] = s->widths[l];
  memset(grad, 0, sizeof(float) * (size_t)d);
  /* per-row activations: a[l] (inputs of layer l), da[l] (act' at layer l output), delta */
  float a[MO_MAX_LAYERS + 1][MO_MAX_WIDTH], da[MO_MAX_LAYERS][MO_MAX_WIDTH], dl[2][MO_MAX_WIDTH];
  float ll_sum = 0.f;
  const int K = dims[NL];
  for (int64_t n = n0; n < n1; ++n) {
    for (int i = 0; i < dims[0]; ++i) a[0][i] = X[n * dims[0] + i];
    for (int l = 0; l < NL; ++l) {
      const int IN = dims[l], OUT = dims[l + 1];
      const float* W = theta + s->kernel_off[l];
      const float* B = theta + s->bias_off[l];
      float z[MO_MAX_WIDTH];
      for (int j = 0; j < OUT; ++j) z[j] = B[j];
      for (int i = 0; i < IN; ++i) {
        const float ai = a[l][i];
        const float* restrict w = W + (size_t)i * OUT;
#pragma omp simd
        for (int j = 0; j < OUT; ++j) z[j] += ai * w[j];
      }
      if (l < NL - 1) for (int j = 0; j < OUT; ++j) act_eval(s->activation, z[j], &a[l + 1][j], &da[l][j]);
      else for (int j = 0; j < OUT; ++j) a[l + 1][j] = z[j];
    }
    const float* out = a[NL];
    float* dout = dl[(NL - 1) & 1];
    float ll;
    if (s->task == 0) {
      const float yv = ((const float*)y)[n], mu = out[0], sg = out[1];
      const float e = expf(sg);
      const float sigma = fminf(fmaxf(e, 1e-6f), 1e6f);
      const float inside = (e > 1e-6f && e < 1e6f) ? 1.f : 0.f;
      const float s2 = sigma * sigma, r = yv - mu, q = r * r / s2;
      ll = (logf(6.283185307179586f * s2) + q) / -2.f;
      dout[0] = r / s2; dout[1] = (q - 1.f) * inside;
      for (int k = 2; k < K; ++k) dout[k] = 0.f;
    } else {
      const int yi = ((const int32_t*)y)[n];
      float m = out[0];
      for (int k = 1; k < K; ++k) m = fmaxf(m, out[k]);
      float se = 0.f;
      for (int k = 0; k < K; ++k) se += expf(out[k] - m);
      ll = out[yi] - (m + logf(se));
      for (int k = 0; k < K; ++k) dout[k] = -expf(out[k] - m) / se + (k == yi ? 1.f : 0.f);
    }
    if (isnan(ll)) { ll = 0.f; for (int k = 0; k < K; ++k) dout[k] = 0.f; } /* jnp.nansum */
    ll_sum += ll;
    for (int k = 0; k < K; ++k) dout[k] *= s->n_batches;
    for (int l = NL - 1; l >= 0; --l) {
      const int IN = dims[l], OUT = dims[l + 1];
      const float* W = theta + s->kernel_off[l];
      float* gW = grad + s->kernel_off[l];
      float* gB = grad + s->bias_off[l];
      const float* dcur = dl[l & 1];
      float* dprev = dl[(l + 1) & 1];
      for (int j = 0; j < OUT; ++j) gB[j] += dcur[j];
      for (int i = 0; i < IN; ++i) {
        const float ai = a[l][i];
        const float* restrict w = W + (size_t)i * OUT;
        float* restrict gw = gW + (size_t)i * OUT;
        float acc = 0.f;
#pragma omp simd
        for (int j = 0; j < OUT; ++j) gw[j] += ai * dcur[j];
        if (l > 0) {
#pragma omp simd reduction(+ : acc)
          for (int j = 0; j < OUT; ++j) acc += w[j] * dcur[j];
          dprev[i] = acc * da[l - 1][i];
        }
      }
    }
  }
  return ll_sum;
}

/* Threads that split the rows of ONE chain's evaluation (bench: host cores / chains when the host has more cores than
 * chains; 1 = the plain serial evaluation). */
static int g_row_threads = 1;
void mo_set_row_threads(int n) { g_row_threads = n < 1 ? 1 : n; }

/* value_and_grad of log_unnormalized_posterior for one chain; grad [d] overwritten. */
float mo_logpost_value_and_grad(const mo_spec* s, const float* theta, const float* X, const void* y, int64_t N,
                                float* grad) {
  const int d = mo_nparams(s);
  float ll_sum = 0.f;
  int T = g_row_threads;
  if (T > 1 && N >= 64 * (int64_t)T) {
#ifdef _OPENMP
    float* part = (float*)malloc(sizeof(float) * (size_t)d * (size_t)T);
    float lls[64];
    if (T > 64) T = 64;
#pragma omp parallel num_threads(T)
    {
      const int t = omp_get_thread_num(), nt = omp_get_num_threads();
      const int64_t per = (N + nt - 1) / nt, a0 = per * t < N ? per * t : N, a1 = a0 + per < N ? a0 + per : N;
      lls[t] = mo_loglik_rows(s, theta, X, y, a0, a1, part + (size_t)t * d);
#pragma omp barrier
#pragma omp for schedule(static)
      for (int i = 0; i < d; ++i) {
        float acc = 0.f;
        for (int q = 0; q < nt; ++q) acc += part[(size_t)q * d + i];
        grad[i] = acc;
      }
#pragma omp single
      { for (int q = 0; q < nt; ++q) ll_sum += lls[q]; }
    }
    free(part);
#else
    ll_sum = mo_loglik_rows(s, theta, X, y, 0, N, grad);
#endif
  } else {
    ll_sum = mo_loglik_rows(s, theta, X, y, 0, N, grad);
  }
  /* prior (priors.py:101-128) */
  float pv = 0.f;
  const float loc = s->prior_loc, sc = s->prior_scale, s2 = sc * sc;
  for (int i = 0; i < d; ++i) {
    const float dlt = theta[i] - loc;
    if (s->prior == 0) { pv += (logf(6.283185307179586f * s2) + dlt * dlt / s2) / -2.f; grad[i] += -dlt / s2; }
    else { pv += -logf(2.f * sc) - fabsf(dlt) / sc; grad[i] += -((dlt > 0.f) - (dlt < 0.f)) / sc; }
  }
  return pv + ll_sum * s->n_batches;
}

/* ---- Philox4x32-10 + Box-Muller: same counter layout as mile_b200/csrc/mile_device.cuh ---------- */
static inline void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* o) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
float mo_philox_normal(uint64_t seed, uint32_t chain, uint64_t step, uint32_t stream, uint32_t elem) {
  /* elements 2p and 2p + 1 are the cosine and sine branches of one draw (counter word 0 = p), as on the device */
  uint32_t r[4];
  philox(elem >> 1, (uint32_t)step, (uint32_t)(step >> 32), stream, (uint32_t)seed ^ (chain * 0x9E3779B9u),
         (uint32_t)(seed >> 32) + chain, r);
  const float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * logf(u1)), ang = 6.283185307179586f * u2;
  return (elem & 1u) ? rad * sinf(ang) : rad * cosf(ang);
}

/* B-step (literal blackjax formula, fp32). Returns the kinetic-energy change of this sub-step. */
static float mo_esh_update(float* u, const float* g, int d, float eps, float coef) {
  float g2 = 0.f;
  for (int i = 0; i < d; ++i) g2 += g[i] * g[i];
  const float gn = sqrtf(g2);
  const float ginv = gn > 1e-13f ? 1.f / gn : 1.f;
  float p = 0.f;
  for (int i = 0; i < d; ++i) p += u[i] * (g[i] * ginv);
  const float delta = eps * coef * gn / (float)(d - 1);
  const float zeta = expf(-delta);
  float rn2 = 0.f;
  for (int i = 0; i < d; ++i) {
    const float raw = (g[i] * ginv) * (1.f - zeta) * (1.f + zeta + p * (1.f - zeta)) + 2.f * zeta * u[i];
    u[i] = raw; rn2 += raw * raw;
  }
  const float rn = sqrtf(rn2), rinv = rn > 1e-13f ? 1.f / rn : 1.f;
  for (int i = 0; i < d; ++i) u[i] *= rinv;
  return (delta - 0.6931471805599453f + logf(1.f + p + (1.f - p) * zeta * zeta)) * (float)(d - 1);
}

/* One MCLMC kernel step for one chain; z [d] normal draws.  info = (logdensity, dK, dE). */
void mo_step(const mo_spec* s, const float* X, const void* y, int64_t N, float* theta, float* u, float* lp, float* g,
             float eps, float L, const float* z, float* info) {
  const int d = mo_nparams(s);
  const float b1 = 0.1931833275037836f, b2 = 1.f - 2.f * 0.1931833275037836f;
  const float lp_old = *lp;
  float dK = mo_esh_update(u, g, d, eps, b1);
  for (int i = 0; i < d; ++i) theta[i] += eps * 0.5f * u[i];
  *lp = mo_logpost_value_and_grad(s, theta, X, y, N, g);
  dK += mo_esh_update(u, g, d, eps, b2);
  for (int i = 0; i < d; ++i) theta[i] += eps * 0.5f * u[i];
  *lp = mo_logpost_value_and_grad(s, theta, X, y, N, g);
  dK += mo_esh_update(u, g, d, eps, b1);
  if (!isinf(L)) {
    const float nu = sqrtf((expf(2.f * eps / L) - 1.f) / (float)d);
    float n2 = 0.f;
    for (int i = 0; i < d; ++i) { u[i] += nu * z[i]; n2 += u[i] * u[i]; }
    const float inv = 1.f / sqrtf(n2);
    for (int i = 0; i < d; ++i) u[i] *= inv;
  }
  if (info) { info[0] = *lp; info[1] = dK; info[2] = dK - *lp + lp_old; }
}

/* All chains: value_and_grad (one chain per OpenMP thread). */
void mo_logpost_batch(const mo_spec* s, const float* thetas, int C, const float* X, const void* y, int64_t N, float* lps,
                      float* grads, int threads) {
  const int d = mo_nparams(s);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int c = 0; c < C; ++c) lps[c] = mo_logpost_value_and_grad(s, thetas + (size_t)c * d, X, y, N, grads + (size_t)c * d);
}

/* mclmc.init for all chains: lp, grad, unit momentum from z0 (or Philox when z0 == NULL). */
void mo_init(const mo_spec* s, const float* X, const void* y, int64_t N, int C, const float* theta, float* u, float* lp,
             float* g, const float* z0, uint64_t seed, int threads) {
  const int d = mo_nparams(s);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int c = 0; c < C; ++c) {
    lp[c] = mo_logpost_value_and_grad(s, theta + (size_t)c * d, X, y, N, g + (size_t)c * d);
    float n2 = 0.f;
    for (int i = 0; i < d; ++i) {
      const float zz = z0 ? z0[(size_t)c * d + i] : mo_philox_normal(seed, (uint32_t)c, 0xFFFFFFFFFFFFFFFFull, 0u, (uint32_t)i);
      u[(size_t)c * d + i] = zz; n2 += zz * zz;
    }
    const float inv = 1.f / sqrtf(n2);
    for (int i = 0; i < d; ++i) u[(size_t)c * d + i] *= inv;
  }
}

/* scan(sampler.step) for all chains (sampling.py:134-177): z [n_steps,C,d] or NULL (Philox).
 * samples [n_kept,C,d] (kept when (step_base+i) % thin == 0) or NULL; info [n_steps,C,3] or NULL. */
void mo_run_sampling(const mo_spec* s, const float* X, const void* y, int64_t N, int C, float* theta, float* u, float* lp,
                     float* g, const float* eps, const float* L, int n_steps, int64_t step_base, int thin, const float* z,
                     uint64_t seed, float* samples, float* info, int threads) {
  const int d = mo_nparams(s);
  const int64_t first = (step_base + thin - 1) / thin;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int c = 0; c < C; ++c) {
    float* zbuf = (float*)malloc(sizeof(float) * (size_t)d);
    float* th = theta + (size_t)c * d; float* uu = u + (size_t)c * d; float* gg = g + (size_t)c * d;
    for (int i = 0; i < n_steps; ++i) {
      const float* zz;
      if (z) zz = z + ((size_t)i * C + c) * d;
      else { for (int k = 0; k < d; ++k) zbuf[k] = mo_philox_normal(seed, (uint32_t)c, (uint64_t)(step_base + i), 1u, (uint32_t)k); zz = zbuf; }
      mo_step(s, X, y, N, th, uu, lp + c, gg, eps[c], L[c], zz, info ? info + ((size_t)i * C + c) * 3 : NULL);
      const int64_t idx = step_base + i;
      if (samples && idx % thin == 0) memcpy(samples + ((size_t)(idx / thin - first) * C + c) * d, th, sizeof(float) * (size_t)d);
    }
    free(zbuf);
  }
}

int mo_max_threads(void) {
#ifdef _OPENMP
  omp_set_max_active_levels(2);   /* chains outside, rows of one chain inside (mo_set_row_threads) */
  return omp_get_max_threads();
#else
  return 1;
#endif
}
