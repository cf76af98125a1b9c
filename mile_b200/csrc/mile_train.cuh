// mile_train.cuh -- deep-ensemble warm-start training on the device: the phase that produces
// the warm-start members the MCLMC chains start from (src/training/trainer.py:330-538: train_warmstart / train_de_member,
// single_step_regr / single_step_class l.662-760, predict_* / compute_metrics_* l.763-868).
//
// One CTA per ensemble member, persistent for one EPOCH: parameters and the AdamW moments live in shared memory, every
// minibatch is gathered from the resident training matrix by row index, pushed through the same forward / likelihood /
// backward tile code as the sampler's generic evaluator (mile_kernel.cuh: grad_eval), and the optimizer update
// (optax.adamw = scale_by_adam -> add_decayed_weights -> scale by -learning_rate; optax.adam; optax.sgd) is fused behind
// it: zero launches per minibatch, one per epoch.  Loss = mean over the batch of GaussianNLLLoss (sigma clipped to
// [1e-5, 1e6]: trainer.py:706-710 then src/inference/metrics.py:332) or of the softmax cross-entropy.
// The validation / test passes (predict_regr / predict_class) are mile_metrics_kernel: mean loss and RMSE / accuracy of
// every member over a whole split.
#pragma once
#include "mile_kernel.cuh"

enum { MILE_OPT_ADAMW = 0, MILE_OPT_ADAM = 1, MILE_OPT_SGD = 2 };

struct TrainParams {
  KParams K;                       // model, training matrix (K.X, K.y, K.N), shared-memory carve-up, K.theta = parameters [C, d]
  const int* batch_idx;            // [n_batches, B] row indices into the training split (same batches for every member,
                                   // like the reference's loader: one permutation key for all devices, tabular.py:186-190)
  int n_batches, B;
  int opt_kind; float lr, b1, b2, eps, wd;
  const unsigned char* stopped;    // [C] early-stopping flags (trainer.py:432-437: a stopped member skips its steps) or NULL
  float* metrics;                  // [n_batches, C, 2] = (mean batch loss, RMSE | accuracy) BEFORE the update; NaN when stopped
  float* m; float* v; int* t;      // AdamW state [C, d], [C, d], [C]
  int off_y;                       // float offset of the gathered labels in shared memory
};

template <int NLMAX>
__global__ void __launch_bounds__(MILE_THREADS, 1) mile_train_kernel(const __grid_constant__ TrainParams T) {
  constexpr int NT = MILE_THREADS;
  extern __shared__ __align__(16) float smem[];
  const KParams& P = T.K;
  const DevModel& M = P.M;
  Ctx c(P);
  c.G = 1; c.rank = 0; c.chain = blockIdx.x; c.phase = 0; c.lead = threadIdx.x == 0;
  c.wp = smem + P.off_wp; c.th = smem + P.off_th; c.uu = smem + P.off_u; c.gg = smem + P.off_g;
  c.thb = smem + P.off_thb; c.ub = smem + P.off_ub; c.gb = smem + P.off_gb; c.gpart = smem + P.off_gpart;
  c.avgx = smem + P.off_avgx; c.avgx2 = smem + P.off_avgx2; c.pmap = reinterpret_cast<int*>(smem + P.off_pmap);
  c.red = smem + P.off_red; c.red2 = c.red + 128; c.phase2 = 0; c.tile = smem + P.off_tile;
  c.xstream = c.tile + M.tile_floats;
  c.xbuf = c.xstream;                      // the gathered minibatch [B][sA[0]]
  c.aux = smem;
  float* ybuf = smem + T.off_y;
  float metric = 0.f;
  c.y_override = ybuf; c.force_resident = 1; c.sig_lo = 1e-5f; c.metric = &metric;
  const int d = M.d, ch = c.chain, tid = threadIdx.x, B = T.B, sx = M.sA[0];
  const bool regr = M.task == MILE_TASK_REGRESSION;
  if (T.stopped && T.stopped[ch]) {         // _fallback of single_step_*: state untouched, metrics NaN
    if (T.metrics)
      for (int b = tid; b < T.n_batches; b += NT) { T.metrics[((long)b * P.C + ch) * 2] = nanf(""); T.metrics[((long)b * P.C + ch) * 2 + 1] = nanf(""); }
    return;
  }
  for (int i = tid; i < M.psize; i += NT) c.wp[i] = 0.f;
  build_pmap<NT>(M, c.pmap, P.dS);
  float* mm = c.uu; float* vv = c.gg;       // first / second moment live where the sampler keeps u / g
  for (int i = tid; i < d; i += NT) { c.th[i] = P.theta[(long)ch * d + i]; mm[i] = T.m[(long)ch * d + i]; vv[i] = T.v[(long)ch * d + i]; }
  int t = T.t[ch];
  __syncthreads();
  refresh_wp<NT>(c);
  const int Q = (((B + 3) >> 2) + 7) & ~7, rows_pad = Q * 4;
  float* gp = c.gpart;
  const float invB = 1.f / (float)B;
#pragma unroll 1
  for (int b = 0; b < T.n_batches; ++b) {
    // gather the minibatch (rows beyond B are zero; their deltas are zeroed by loglik_rows)
    const int* idx = T.batch_idx + (long)b * B;
    for (int e = tid; e < rows_pad * (sx >> 2); e += NT) {
      const int r = e / (sx >> 2), q = e % (sx >> 2);
      float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < B) v4 = __ldg(reinterpret_cast<const float4*>(P.X + (long)idx[r] * sx) + q);
      reinterpret_cast<float4*>(c.xbuf)[e] = v4;
    }
    for (int r = tid; r < B; r += NT) ybuf[r] = reinterpret_cast<const float*>(P.y)[idx[r]];   // fp32 or int32 bit pattern
    metric = 0.f;
    __syncthreads();
    grad_eval<NLMAX, NT>(c, 0, B, gp);       // gp = d(sum log-lik)/d theta, gp[dS] = sum log-lik (no prior)
    float mv[1] = {metric};
    block_sum<1, NT>(mv, c.red, c.phase);
    __syncthreads();
    const float loss = -gp[P.dS] * invB;
    if (T.metrics && tid == 0) {
      float* o = T.metrics + ((long)b * P.C + ch) * 2;
      o[0] = loss; o[1] = regr ? sqrtf(mv[0] * invB) : mv[0] * invB;
    }
    // ---- optimizer update on the mean-loss gradient g = -(1/B) d(sum ll)/d theta ---------------------------------
    ++t;
    const float bc1 = 1.f - powf(T.b1, (float)t), bc2 = 1.f - powf(T.b2, (float)t);
    for (int i = tid; i < d; i += NT) {
      const float g = -gp[i] * invB;
      float th = c.th[i];
      if (T.opt_kind == MILE_OPT_SGD) {
        th -= T.lr * g;
      } else {
        const float m1 = T.b1 * mm[i] + (1.f - T.b1) * g;
        const float v1 = T.b2 * vv[i] + (1.f - T.b2) * g * g;
        mm[i] = m1; vv[i] = v1;
        float upd = (m1 / bc1) / (sqrtf(v1 / bc2) + T.eps);
        if (T.opt_kind == MILE_OPT_ADAMW) upd += T.wd * th;
        th -= T.lr * upd;
      }
      c.th[i] = th;
      store_param(c, i, th);
    }
    __syncthreads();
  }
  for (int i = tid; i < d; i += NT) { P.theta[(long)ch * d + i] = c.th[i]; T.m[(long)ch * d + i] = mm[i]; T.v[(long)ch * d + i] = vv[i]; }
  if (tid == 0) T.t[ch] = t;
}

// predict_regr / predict_class (trainer.py:763-868): per member mean loss and RMSE | accuracy over a whole split.
struct MetricsParams {
  KParams K;               // K.theta_in = parameters [n, d]; K.which selects the split (0 train, 1 test)
  float* out;              // [n, 2]
};

template <int NLMAX>
__global__ void __launch_bounds__(MILE_THREADS, 1) mile_metrics_kernel(const __grid_constant__ MetricsParams T) {
  constexpr int NT = MILE_THREADS;
  extern __shared__ __align__(16) float smem[];
  const KParams& P = T.K;
  const DevModel& M = P.M;
  Ctx c(P);
  c.G = 1; c.rank = 0; c.chain = blockIdx.x; c.phase = 0; c.lead = threadIdx.x == 0;
  c.wp = smem + P.off_wp; c.th = smem + P.off_th; c.pmap = reinterpret_cast<int*>(smem + P.off_pmap);
  c.red = smem + P.off_red; c.tile = smem + P.off_tile; c.xstream = c.tile + M.tile_floats; c.xbuf = c.xstream;
  const int d = M.d, ch = c.chain, tid = threadIdx.x;
  for (int i = tid; i < M.psize; i += NT) c.wp[i] = 0.f;
  build_pmap<NT>(M, c.pmap, P.dS);
  for (int i = tid; i < d; i += NT) c.th[i] = P.theta_in[(long)ch * d + i];
  __syncthreads();
  refresh_wp<NT>(c);
  __syncthreads();
  const float* Xs = P.which ? P.Xt : P.X;
  const void* ys = P.which ? P.yt : P.y;
  const long Nr = P.which ? P.Nt : P.N;
  const int K = M.dims[M.NL];
  float v[2] = {0.f, 0.f};
  for (long row0 = 0; row0 < Nr; row0 += M.TR) {
    const int nvalid = (int)((Nr - row0) < M.TR ? (Nr - row0) : M.TR);
    const int Q = (((nvalid + 3) >> 2) + 7) & ~7;
    load_x_tile<NT>(c.xstream, Xs, row0, nvalid, Q * 4, M.sA[0]);
    __syncthreads();
    const float* out = forward_tile<NT>(c, c.xstream, Q);
    for (int r = tid; r < nvalid; r += NT) {
      const float* o = out + r * M.sA[M.NL];
      if (M.task == MILE_TASK_REGRESSION) {
        const float yv = reinterpret_cast<const float*>(ys)[row0 + r];
        const float sigma = fminf(fmaxf(expf(o[1]), 1e-5f), 1e6f), res = yv - o[0];
        v[0] += 0.5f * logf(6.283185307179586f * sigma * sigma) + res * res / (2.f * sigma * sigma);
        v[1] += res * res;
      } else {
        const int yi = reinterpret_cast<const int*>(ys)[row0 + r];
        float mx = o[0]; int am = 0;
        for (int k = 1; k < K; ++k) if (o[k] > mx) { mx = o[k]; am = k; }
        float se = 0.f;
        for (int k = 0; k < K; ++k) se += expf(o[k] - mx);
        v[0] += (mx + logf(se)) - o[yi];
        v[1] += am == yi ? 1.f : 0.f;
      }
    }
    __syncthreads();
  }
  block_sum<2, NT>(v, c.red, c.phase);
  if (tid == 0) {
    const float inv = 1.f / (float)Nr;
    T.out[ch * 2] = v[0] * inv;
    T.out[ch * 2 + 1] = M.task == MILE_TASK_REGRESSION ? sqrtf(v[1] * inv) : v[1] * inv;
  }
}
