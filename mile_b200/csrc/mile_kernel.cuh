// mile_kernel.cuh -- the persistent cluster-per-chain MCLMC kernel.
//
// One thread-block cluster of G CTAs owns one chain for the whole launch:
//   * theta / momentum / gradient live replicated in every CTA's shared memory,
//   * the training rows are split across the G CTAs (X slice resident in shared memory
//     when it fits, otherwise streamed tile by tile from the padded HBM/L2 copy),
//   * every gradient evaluation ends in a DSMEM all-reduce of the G partial gradients
//     (one cluster barrier, fixed summation order => all CTAs hold bit-identical state),
//   * B/A/B/A/B + partial refresh + energy bookkeeping + tuning statistics + thinned
//     sample capture + test-set logsumexp run in the same kernel: zero launches per step.
//
// Reference call stack replaced: src/training/sampling.py:134-177 (scan of sampler.step),
// src/training/warmup.py:276-352 (tuning step), blackjax 1.2.2 mclmc kernel.
#pragma once
#include "mile_device.cuh"
#include <float.h>

enum { MODE_EVAL = 0, MODE_INIT = 1, MODE_SAMPLE = 2, MODE_TUNE = 3, MODE_LPPD = 4, MODE_PREDICT = 5 };

// NUTS branch (mile_nuts.cuh): adaptation state and per-CTA scratch; everything else comes from KParams
struct NutsParams {
  float* scratch;                 // [CTAs][(10 + 2 D) * dS]
  float *imm, *w_mean, *w_m2;     // [C][d] inverse mass matrix (diagonal), Welford mean / m2 of the open window
  float* da;                      // [C][8] log_x, log_x_avg, step, avg_error, mu, Welford count, step_size, -
  const unsigned char* schedule;  // [n_steps] bit 0 = slow stage, bit 1 = window end; null = no adaptation (sampling)
  const float* uni;               // [n_steps][C][uni_len] host-supplied uniforms or null (Philox)
  float* info;                    // [n_steps][C][8] or null
  int uni_len, max_doublings;
  float divergence_threshold, target_accept;
  int smem_off;                   // float offset of the scratch inside dynamic shared memory, or -1 (global scratch)
  int push;                       // 1 = DSMEM push exchange where the plan has its buffers (tensor-evaluator plans)
};

struct KParams {
  DevModel M;
  // data: padded row-major copies [N][sA[0]]
  const float* X; const void* y; long N;
  const float* Xt; const void* yt; long Nt;
  int C, G, resident, rows_res;   // rows_res: rows of the resident X slice buffer (multiple of TR)
  // chain state [C,d],[C,d],[C,d],[C]
  float* theta; float* u; float* grad; float* lp;
  int mode, n_steps, thin, refresh_mode, do_lppd;
  long step_base, sample_base, n_slots;
  const float* eps; const float* L;  // [C]
  const float* z; unsigned long long seed;
  float* samples; float* info;
  float* lppd_m; float* lppd_s;
  float* carry; int carry_valid;   // [C,2] (sum g^2, u.g) carried across launches so chunking is bit-exact
  // tuning state [C] / [C,d]
  float *t_time, *t_xavg, *t_epsmax, *t_eps, *t_L, *t_wtot, *avg_x, *avg_x2, *tune_info;
  int tune1, tune2; float ev_start, ev_end, trust, neff;
  // eval / init / lppd / predict io
  const float* theta_in; float* lp_out; float* grad_out; float* pred_out; int n_eval, which;
  // global-memory exchange (sync_mode 1): any number of CTAs per chain, cooperative launch, partial gradients and an
  // arrival counter in HBM/L2 instead of DSMEM + cluster barrier (lets <= 12 chains use all 148 SMs)
  // Exchange words are {value, epoch flag} pairs written with one 8-byte store each (the NCCL "LL" idea): a reader
  // spins on the data word itself, so there is no fence, no atomic counter and no separate barrier on the critical path.
  int sync_mode; float2* xchg; unsigned int xbase;   // xchg [C][2][G][dS+4] (value, flag); flag = xbase + eval + 1
  float2* xchg2;         // [C][2][dS+4] summed slices of the reduce-scatter form of the exchange (mile_mma_step_kernel)
  int out_stride;        // row stride of grad_out / lp_out (d, or d+1 for the packed [C,d+1] all-reduce buffer)
  float prior_weight;    // 1, or 1/world when the rows are sharded across ranks (the sum over ranks counts the prior once)
  // shared-memory carve-up (float offsets)
  int dS, off_wp, off_th, off_u, off_g, off_thb, off_ub, off_gb, off_gpart, off_avgx, off_avgx2,
      off_pmap, off_red, off_tile, off_x, off_aux, off_z, off_gs;
  // global id of local chain 0: the Philox streams are keyed by (seed, chain_base + chain, step, element), so the ranks
  // of a chain-partitioned ensemble draw independent noise (the reference splits one key per chain, sampling.py:181-184)
  int chain_base;
  // multi-rank exchange (rows sharded across the GPUs of one box, mile_shard_*): the step loop stays ONE persistent kernel
  // per rank.  After the local reduce-scatter every CTA pushes its summed slice as flagged 8-byte words into the
  // `mr_sums` region of EVERY rank (peer memory mapped with CUDA IPC, stores travel over NVLink), and all CTAs then poll
  // only their own GPU's copy: one NVLink store hop per evaluation, no NCCL call, no kernel boundary.
  // diagonal preconditioning (warmup.py:391-393, blackjax `sqrt_diag_cov`): [C][d] or null.  The B-steps see the scaled
  // gradient m .* g, the A-steps move by eps * m .* u.  Served by the generic step loop and the integrator kernel.
  const float* sdc;
  // partition sampling (src/training/partition_sampling.py, trainer.py:613-659): pmask [d] = 1 for sampled parameters, 0
  // for frozen ones (shared by all chains), d_eff = number of sampled parameters (the dimension of the MCLMC dynamics:
  // ESH normalisation, refresh, tuning).  Frozen parameters keep their value, see no prior and get zero gradient,
  // momentum and noise.  Null = everything is sampled (d_eff = d).  Generic step loop only.
  const float* pmask; int d_eff;
  int mr_world, mr_rank; unsigned int mr_base;   // flag = mr_base + eval + 1 (advanced identically on every rank)
  float2* mr_sums[8];    // rank r's region [C][2 parities][world][dS+4] (own region for r == mr_rank)
  NutsParams nuts;
};

struct Ctx {
  const KParams& P;
  float *wp, *th, *uu, *gg, *thb, *ub, *gb, *gpart, *avgx, *avgx2, *red, *tile, *xbuf, *xstream, *aux;
  int* pmap;
  int phase;   // block_sum double-buffer phase (whole block)
  float* red2; int phase2;   // scratch / phase of the integrator thread group (named barrier 1)
  int rank, G, chain;
  // element-split mode (ES, wide / large-d integrator): the d elements are strided over the CTAs of a cluster
  int e0, estride; float* csum; int csum_phase;
  // warm-start training overrides (mile_train.cuh): labels of the gathered minibatch, X always in c.xbuf, sigma clip, metric
  const void* y_override = nullptr; int force_resident = 0; float sig_lo = 1e-6f; float* metric = nullptr;
  const float* sdc = nullptr;   // this chain's sqrt_diag_cov [d] (global memory) or null = identity
  const float* pmask = nullptr; // partition mask [d] or null; deff = dimension of the sampled sub-space
  int deff = 0;
  bool lead;   // the one thread that reports per-chain scalars (thread 0 of the block; lane 0 of the integrator warp in warp mode)
  __device__ Ctx(const KParams& p) : P(p) {}
};

// element i of the preconditioner (1 when none is set)
__device__ __forceinline__ float sdc_at(const Ctx& c, int i) { return c.sdc ? __ldg(c.sdc + i) : 1.f; }
// 1 for a sampled parameter, 0 for a frozen one (partition sampling)
__device__ __forceinline__ float pm_at(const Ctx& c, int i) { return c.pmask ? __ldg(c.pmask + i) : 1.f; }
__device__ __forceinline__ int dim_eff(const Ctx& c) { return c.pmask ? c.deff : c.P.M.d; }

// flat element i -> position in the padded parameter image (pmap) and in the transposed image
// (pmap + dS; biases and layer 0, which need no transpose, point at their pmap slot again)
template <int NT>
__device__ __forceinline__ void build_pmap(const DevModel& M, int* pmap, int dS) {
  for (int l = 0; l < M.NL; ++l) {
    const int IN = M.dims[l], OUT = M.dims[l + 1], OUTP = M.dimp[l + 1], INP = M.dimp[l];
    for (int j = threadIdx.x; j < OUT; j += NT) {
      pmap[M.bias_off[l] + j] = M.pb_off[l] + j;
      pmap[dS + M.bias_off[l] + j] = M.pb_off[l] + j;
    }
    for (int e = threadIdx.x; e < IN * OUT; e += NT) {
      const int i = e / OUT, j = e % OUT;
      pmap[M.kern_off[l] + e] = M.pw_off[l] + i * OUTP + j;
      pmap[dS + M.kern_off[l] + e] = l >= 1 ? M.pwt_off[l] + j * INP + i : M.pw_off[l] + i * OUTP + j;
    }
  }
}

__device__ __forceinline__ void store_param(Ctx& c, int i, float t) {
  c.wp[c.pmap[i]] = t;
  c.wp[c.pmap[c.P.dS + i]] = t;
}

template <int NT>
__device__ __forceinline__ void refresh_wp(Ctx& c) {
  const int d = c.P.M.d;
  for (int i = threadIdx.x; i < d; i += NT) store_param(c, i, c.th[i]);
}

// Copy rows [row0, row0+rows_pad) of the padded global matrix into a shared tile, zero beyond nvalid.
template <int NT>
__device__ __forceinline__ void load_x_tile(float* dst, const float* __restrict__ src, long row0, int nvalid,
                                            int rows_pad, int sx) {
  const float4* s4 = reinterpret_cast<const float4*>(src + row0 * sx);
  float4* d4 = reinterpret_cast<float4*>(dst);
  const int nv4 = nvalid * (sx >> 2), np4 = rows_pad * (sx >> 2);
  for (int i = threadIdx.x; i < np4; i += NT)
    d4[i] = i < nv4 ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Forward pass of one row tile; returns pointer to the output buffer (stride sA[NL]).
template <int NT>
__device__ __forceinline__ const float* forward_tile(Ctx& c, const float* Xt, int Q) {
  const DevModel& M = c.P.M;
  const float* in = Xt;
  int sin_ = M.sA[0];
  for (int l = 0; l < M.NL; ++l) {
    float* out = c.tile + M.a_off[l + 1];
    fwd_layer<NT>(M, l, c.wp, in, sin_, out, c.tile + M.d_off[l], M.sA[l + 1], Q, l == M.NL - 1);
    __syncthreads();
    in = out; sin_ = M.sA[l + 1];
  }
  return in;
}

// Full-batch value_and_grad restricted to this CTA's rows [r0, r1): partial gradient (flat
// layout, likelihood part only) and partial log-likelihood into gpart[0..dS] (ll at [dS]).
template <int NLMAX, int NT>
__device__ __forceinline__ void grad_eval(Ctx& c, long r0, long r1, float* gpart) {
  const KParams& P = c.P;
  const DevModel& M = P.M;
  DwAcc<NLMAX> acc;
  acc.zero();
  float llpart = 0.f;
  PROF_DECL;
  const int TR = M.TR;
  const long nrows = r1 > r0 ? r1 - r0 : 0;
  const int ntiles = (int)((nrows + TR - 1) / TR);
  for (int t = 0; t < ntiles; ++t) {
    const long row0 = r0 + (long)t * TR;
    const int nvalid = (int)((r1 - row0) < TR ? (r1 - row0) : TR);
    const int Q = (((nvalid + 3) >> 2) + 7) & ~7;
    const int rows_pad = Q * 4;
    const float* Xt;
    if (P.resident || c.force_resident) {
      Xt = c.xbuf + (long)t * TR * M.sA[0];
    } else {
      load_x_tile<NT>(c.xbuf, P.X, row0, nvalid, rows_pad, M.sA[0]);
      __syncthreads();
      Xt = c.xbuf;
    }
    PROF(0);
    const float* out = forward_tile<NT>(c, Xt, Q);
    PROF(1);
    llpart += loglik_rows<NT>(M, out, c.tile + M.d_off[M.NL - 1], M.sA[M.NL], c.y_override ? c.y_override : P.y, row0, nvalid,
                              rows_pad, c.sig_lo, c.metric);
    __syncthreads();
    PROF(2);
    for (int l = M.NL - 1; l >= 1; --l) {
      bwd_layer<NT>(M, l, c.wp, c.tile + M.d_off[l], M.sA[l + 1], c.tile + M.d_off[l - 1], M.sA[l], Q);
      __syncthreads();
    }
    PROF(3);
#pragma unroll
    for (int l = 0; l < NLMAX; ++l)
      if (l < M.NL)
        dw_accumulate<NLMAX, NT>(M, l, acc, l, l == 0 ? Xt : c.tile + M.a_off[l], M.sA[l],
                             c.tile + M.d_off[l], M.sA[l + 1], rows_pad);
    __syncthreads();
    PROF(4);
  }
  // cross-chunk reduction of the per-thread 4x4 tiles (scratch aliases the tile buffers)
  float* scr = c.tile;
  float* scrb = c.tile + NT * 16;
#pragma unroll
  for (int l = 0; l < NLMAX; ++l) {
    if (l < M.NL) {
      const DwRole R = dw_role<NT>(M, l);
      if (R.active) {
        float4* s4 = reinterpret_cast<float4*>(scr + (R.chunk * R.ntile + R.tile) * 16);
        s4[0] = make_float4(acc.w[l][0], acc.w[l][1], acc.w[l][2], acc.w[l][3]);
        s4[1] = make_float4(acc.w[l][4], acc.w[l][5], acc.w[l][6], acc.w[l][7]);
        s4[2] = make_float4(acc.w[l][8], acc.w[l][9], acc.w[l][10], acc.w[l][11]);
        s4[3] = make_float4(acc.w[l][12], acc.w[l][13], acc.w[l][14], acc.w[l][15]);
        if (R.itl == 0)
          *reinterpret_cast<float4*>(scrb + (R.chunk * R.ntile + R.tile) * 4) =
              make_float4(acc.b[l][0], acc.b[l][1], acc.b[l][2], acc.b[l][3]);
      }
      __syncthreads();
      const int IN = M.dims[l], OUT = M.dims[l + 1], njt = M.dimp[l + 1] >> 2;
      for (int o = threadIdx.x; o < R.ntile * 16; o += NT) {
        const int tile = o >> 4, e = o & 15;
        float s = 0.f;
        for (int ch = 0; ch < R.nch; ++ch) s += scr[(ch * R.ntile + tile) * 16 + e];
        const int i = (tile / njt) * 4 + (e >> 2), j = (tile % njt) * 4 + (e & 3);
        if (i < IN && j < OUT) gpart[M.kern_off[l] + i * OUT + j] = s;
      }
      for (int o = threadIdx.x; o < njt * 4; o += NT) {
        const int tile = o >> 2, e = o & 3;   // itl == 0 -> tile == jt
        float s = 0.f;
        for (int ch = 0; ch < R.nch; ++ch) s += scrb[(ch * R.ntile + tile) * 4 + e];
        if (o < OUT) gpart[M.bias_off[l] + o] = s;
      }
      __syncthreads();
    }
  }
  float v[1] = {llpart};
  block_sum<1, NT>(v, c.red, c.phase);
  if (threadIdx.x == 0) gpart[P.dS] = v[0] * M.n_batches;   // probabilistic.py:136
  PROF(5);
}

#define MILE_RED(c, BAR) ((BAR) ? (c).red2 : (c).red)
#define MILE_PH(c, BAR) ((BAR) ? (c).phase2 : (c).phase)

// DSMEM all-reduce of the cluster's partial gradients + prior: afterwards gg = full gradient of
// the log-posterior (bit-identical in every CTA), returns the log-posterior value.  Also returns
// sum g^2, sum u.g and the number of non-finite entries of theta for the next B-step / handle_nans.
// 8-byte flagged word of the global exchange (sync_mode 1)
__device__ __forceinline__ void ll_store(float2* p, float v, unsigned int flag) {
  asm volatile("st.volatile.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(flag) : "memory");
}
// sum over the G ranks' flagged words at the same offset; spins (bounded) until every word carries `flag`
__device__ __forceinline__ float ll_sum(const float2* base, int stride, int G, unsigned int flag) {
  uint32_t v[16], f[16];
  long spin = 0;
  bool ok;
  do {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r < G) asm volatile("ld.volatile.global.v2.b32 {%0, %1}, [%2];" : "=r"(v[r]), "=r"(f[r]) : "l"(base + (long)r * stride) : "memory");
      else { v[r] = 0u; f[r] = flag; }
    }
    ok = true;
#pragma unroll
    for (int r = 0; r < 16; ++r) ok = ok && (f[r] == flag);
    if (!ok && ++spin > (1L << 24)) __trap();
  } while (!ok);
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 16; ++r) s += __uint_as_float(v[r]);
  return s;
}
// two offsets at once (a thread that owns two elements keeps both sets of loads in flight: one L2 round trip, not two)
__device__ __forceinline__ void ll_sum_pair(const float2* base0, const float2* base1, int stride, int G, unsigned int flag, float& s0, float& s1) {
  uint32_t v[16], f[16], w[16], h[16];
  long spin = 0;
  bool ok;
  do {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r < G) {
        asm volatile("ld.volatile.global.v2.b32 {%0, %1}, [%2];" : "=r"(v[r]), "=r"(f[r]) : "l"(base0 + (long)r * stride) : "memory");
        asm volatile("ld.volatile.global.v2.b32 {%0, %1}, [%2];" : "=r"(w[r]), "=r"(h[r]) : "l"(base1 + (long)r * stride) : "memory");
      } else { v[r] = 0u; f[r] = flag; w[r] = 0u; h[r] = flag; }
    }
    ok = true;
#pragma unroll
    for (int r = 0; r < 16; ++r) ok = ok && (f[r] == flag) && (h[r] == flag);
    if (!ok && ++spin > (1L << 24)) __trap();
  } while (!ok);
  s0 = 0.f; s1 = 0.f;
#pragma unroll
  for (int r = 0; r < 16; ++r) { s0 += __uint_as_float(v[r]); s1 += __uint_as_float(w[r]); }
}

template <int NT, int BAR>
__device__ __forceinline__ float cluster_reduce_grad(Ctx& c, float* gpart, const float2* gslab, unsigned int flag, float& g2, float& ug,
                                                     float& nonfinite, const bool use_ll, const int ng) {
  // use_ll: the partials arrive as `ng` flagged words per element at gslab (+ r * (dS + 4)): the G CTAs of the chain
  // (sync_mode 1), or the `world` ranks' summed slices of the multi-rank exchange
  // (DSMEM mode: the caller has already executed the cluster / block barrier that publishes every CTA's gpart;
  //  global mode: gslab = this chain's parity slab of flagged words, waited on element by element)
  const KParams& P = c.P;
  const DevModel& M = P.M;
  cg::cluster_group cluster = cg::this_cluster();
  // rank loops are unrolled to the maximum cluster size with predication so that all remote (DSMEM)
  // loads of an element are in flight together (~200 cycles each when serialised)
  const float* rp[16];
  const int stride_g = P.dS + 4;
#pragma unroll
  for (int r = 0; r < 16; ++r) rp[r] = (!use_ll && c.G > 1 && r < c.G) ? cluster.map_shared_rank(gpart, r) : gpart;
  float ll = 0.f;
  if (use_ll) {
    // only one thread fetches the log-likelihood partials (it folds them into its v[0] below): a second serial L2
    // round trip for every thread would sit on the critical path
  } else {
    float t[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) t[r] = r < c.G ? rp[r][P.dS] : 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) ll += t[r];
  }
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  const float loc = M.prior_loc, sc = M.prior_scale, s2 = sc * sc;
  const float lognorm = M.prior == MILE_PRIOR_NORMAL ? logf(6.283185307179586f * s2) : logf(2.f * sc);
  auto consume = [&](int i, float s) {
    const float th = c.th[i];
    float pg, pv;
    if (M.prior == MILE_PRIOR_NORMAL) {
      const float dlt = th - loc;
      pv = (lognorm + dlt * dlt / s2) / -2.f;
      pg = -dlt / s2;
    } else {
      const float dlt = th - loc;
      pv = -lognorm - fabsf(dlt) / sc;
      pg = -((dlt > 0.f) - (dlt < 0.f)) / sc;
    }
    const float mk = pm_at(c, i);
    const float g = (s + pg * P.prior_weight) * mk;
    c.gg[i] = g;
    const float gs = g * sdc_at(c, i);
    v[0] += pv * P.prior_weight * mk; v[1] += gs * gs; v[2] += c.uu[i] * gs; v[3] += isfinite(th) ? 0.f : 1.f;
  };
  if (use_ll) {
    for (int i = threadIdx.x; i < M.d; i += 2 * NT) {
      const int i2 = i + NT;
      if (i2 < M.d) {
        float s0, s1;
        ll_sum_pair(gslab + i, gslab + i2, stride_g, ng, flag, s0, s1);
        consume(i, s0); consume(i2, s1);
      } else {
        consume(i, ll_sum(gslab + i, stride_g, ng, flag));
      }
    }
  } else {
    for (int i = threadIdx.x; i < M.d; i += NT) {
      float t[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) t[r] = r < c.G ? rp[r][i] : 0.f;
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) s += t[r];
      consume(i, s);
    }
  }
  PROF(14);
  if (use_ll && threadIdx.x == NT - 1) v[0] += ll_sum(gslab + P.dS, stride_g, ng, flag);
  block_sum<4, NT, BAR>(v, MILE_RED(c, BAR), MILE_PH(c, BAR));
  g2 = v[1]; ug = v[2]; nonfinite = v[3];
  return v[0] + ll;
}


// ES (element split): element loops start at c.e0 with stride c.estride, and every block_sum is followed by a
// DSMEM sum over the cluster's CTAs in rank order (double-buffered slots: one cluster barrier per reduction).
#define MILE_I0(c, ES) ((ES) ? (c).e0 : (int)threadIdx.x)
#define MILE_IS(c, ES, NT) ((ES) ? (c).estride : (NT))
template <int NV>
__device__ __forceinline__ void cluster_sum(Ctx& c, float (&v)[NV]) {
  cg::cluster_group cl = cg::this_cluster();
  float* slot = c.csum + (c.csum_phase & 1) * 8;
  c.csum_phase++;
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) slot[k] = v[k];
  cl.sync();
  float t[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) t[k] = 0.f;
  const int G = (int)cl.num_blocks();
  for (int r = 0; r < G; ++r) {
    const float* rs = cl.map_shared_rank(slot, r);
#pragma unroll
    for (int k = 0; k < NV; ++k) t[k] += rs[k];
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = t[k];
}
#define MILE_WARP_MODE 99   // BAR value: the caller is ONE warp (elements strided by lane through c.e0 / c.estride, ES = true);
                            // reductions are xor-shuffles only: no shared memory, no barrier
template <int NV, int NT, int BAR, bool ES>
__device__ __forceinline__ void all_sum(Ctx& c, float (&v)[NV]) {
  if (BAR == MILE_WARP_MODE) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    return;
  }
  block_sum<NV, NT, BAR == MILE_WARP_MODE ? 0 : BAR>(v, MILE_RED(c, BAR), MILE_PH(c, BAR));
  if (ES) cluster_sum<NV>(c, v);
}

// ESH momentum update B(coef) (blackjax esh_dynamics_momentum_update_one_step, sqrt_diag_cov = 1).
// delta-small-safe forms: 1-zeta = -expm1(-delta), log(1+p+(1-p)zeta^2) - ln2 = log1p(-(1-p)(1-zeta^2)/2).
// scalar part: u' = ae * g + au * u; returns the kinetic-energy change of this sub-step
__device__ __forceinline__ float esh_coeffs(int d, float eps, float coef, float g2, float ug, float& ae, float& au) {
  // The scalar chain below sits on the critical path of every half step (one thread-serial dependency chain per
  // evaluation), so the IEEE sqrt / divide / expm1f / log1pf library sequences are replaced on their common range by
  // MUFU.RSQ + one Newton step and by short Horner series (all <= 1e-9 relative on the stated range); outside it the
  // library forms are used.
  float gn, ginv;
  if (g2 > 1e-26f && g2 < 1e37f) {
    float y = rsqrtf(g2);
    y = y * fmaf(-0.5f * g2 * y, y, 1.5f);
    ginv = y; gn = g2 * y;
  } else {
    gn = sqrtf(g2);
    ginv = gn > 1e-13f ? 1.f / gn : 1.f;
  }
  const float p = ug * ginv;
  const float delta = eps * coef * gn / (float)(d - 1);
  float omz;                             // 1 - zeta = 1 - exp(-delta) without cancellation
  if (delta >= 0.f && delta < 0.125f) {
    float q = fmaf(delta, -1.f / 7.f, 1.f);
    q = fmaf(delta * (-1.f / 6.f), q, 1.f);
    q = fmaf(delta * (-1.f / 5.f), q, 1.f);
    q = fmaf(delta * (-1.f / 4.f), q, 1.f);
    q = fmaf(delta * (-1.f / 3.f), q, 1.f);
    q = fmaf(delta * (-1.f / 2.f), q, 1.f);
    omz = delta * q;
  } else {
    omz = -expm1f(-delta);
  }
  const float zeta = 1.f - omz;
  const float ce = omz * (1.f + zeta + p * omz);
  const float cu = 2.f * zeta;
  // |raw|^2 in closed form: raw = ce * e + cu * u with |e| = |u| = 1 and u.e = p  (saves one block reduction per
  // B-step; |u| is re-normalised numerically by the partial refresh once per step, so rounding cannot accumulate)
  const float rn2 = ce * ce + cu * cu + 2.f * ce * cu * p;
  const float rinv = rn2 > 1e-26f ? rsqrtf(rn2) : 1.f;
  ae = ce * ginv * rinv; au = cu * rinv;
  const float omz2 = omz * (1.f + zeta);   // 1 - zeta^2
  const float yy = -0.5f * (1.f - p) * omz2;
  float l1p;
  if (fabsf(yy) < 0.0625f) {             // log1p(y) = y (1 - y/2 + y^2/3 - ... + y^6/7)
    float q = fmaf(yy, 1.f / 7.f, -1.f / 6.f);
    q = fmaf(yy, q, 1.f / 5.f);
    q = fmaf(yy, q, -1.f / 4.f);
    q = fmaf(yy, q, 1.f / 3.f);
    q = fmaf(yy, q, -1.f / 2.f);
    q = fmaf(yy, q, 1.f);
    l1p = yy * q;
  } else {
    l1p = log1pf(yy);
  }
  return (delta + l1p) * (float)(d - 1);
}
template <int NT, int BAR = 0, bool ES = false>
__device__ __forceinline__ float esh_update(Ctx& c, float eps, float coef, float g2, float ug) {
  const int d = c.P.M.d;
  float ae, au;
  const float dk = esh_coeffs(dim_eff(c), eps, coef, g2, ug, ae, au);
  // (restrict + unroll: in the element-split / global-memory form every access is an L2 round trip; the loads of four
  //  elements must be in flight together)
  float* __restrict__ uu = c.uu;
  const float* __restrict__ gg = c.gg;
#pragma unroll 4
  for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) uu[i] = ae * (gg[i] * sdc_at(c, i)) + au * uu[i];
  return dk;
}

// A(coef): theta += eps*coef*u and refresh the padded weight image (the caller synchronises the block
// before the next gradient evaluation reads the image).
template <int NT>
__device__ __forceinline__ void position_update(Ctx& c, float eps, float coef) {
  const int d = c.P.M.d;
  const float s = eps * coef;
  for (int i = threadIdx.x; i < d; i += NT) {
    const float t = c.th[i] + s * sdc_at(c, i) * c.uu[i];
    c.th[i] = t;
    store_param(c, i, t);
  }
}

__device__ __forceinline__ float noise_at(const KParams& P, int chain, long step_local, int slot, int nslot, int i) {
  if (P.z) return P.z[(((long)step_local * nslot + slot) * P.C + chain) * P.M.d + i];
  return philox_normal(P.seed, (uint32_t)(P.chain_base + chain), (uint64_t)(P.step_base + step_local), (uint32_t)slot + 1u, (uint32_t)i);
}

// partially_refresh_momentum: u <- normalise(u + nu z); also returns u.g for the next B-step.
template <int NT, int BAR = 0, bool ES = false>
__device__ __forceinline__ void refresh_momentum(Ctx& c, float eps, float L, long step_local, int slot, int nslot,
                                                 float& ug_out, const float* zsm = nullptr) {
  const KParams& P = c.P;
  const int d = P.M.d;
  if (isinf(L)) {   // no refresh: still re-normalise numerically (the B-steps use the closed-form norm)
    float v[2] = {0.f, 0.f};
    for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) { v[0] += c.uu[i] * c.uu[i]; v[1] += c.uu[i] * (c.gg[i] * sdc_at(c, i)); }
    all_sum<2, NT, BAR, ES>(c, v);
    const float inv = 1.f / sqrtf(v[0]);
    for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) c.uu[i] *= inv;
    ug_out = v[1] * inv;
    return;
  }
  const float nu = sqrtf((expf(2.f * eps / L) - 1.f) / (float)dim_eff(c));
  float v[2] = {0.f, 0.f};
  float* __restrict__ uu = c.uu;
  const float* __restrict__ gg = c.gg;
#pragma unroll 4
  for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) {
    const float w = uu[i] + nu * pm_at(c, i) * (zsm ? zsm[i] : noise_at(P, c.chain, step_local, slot, nslot, i));
    uu[i] = w;
    v[0] += w * w; v[1] += w * (gg[i] * sdc_at(c, i));
  }
  all_sum<2, NT, BAR, ES>(c, v);
  const float inv = 1.f / sqrtf(v[0]);
#pragma unroll 4
  for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) uu[i] *= inv;
  ug_out = v[1] * inv;
}

__device__ __forceinline__ float nan_to_num(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0.f ? FLT_MAX : -FLT_MAX;
  return x;
}

// Per-chain scalars of the adaptive state (warmup.py:358-363) carried in registers (uniform across the block).
struct TuneRegs { float time, xavg, epsmax, wtot; };

// End of a tuning iteration (warmup.py:293-350): handle_nans, energy-variance step-size predictor, streaming
// average of (x, x^2).  Returns the new step size; lp / dE / g2 / ug are updated in place.
template <int NT, bool HAS_WP, int BAR = 0, bool ES = false>
__device__ __forceinline__ float tune_epilogue(Ctx& c, TuneRegs& t, float eps, float lp_old, float nf, long s_local,
                                               float& lp, float& dE, float& g2, float& ug) {
  const KParams& P = c.P;
  const int d = P.M.d, tid = threadIdx.x, ch = c.chain;
  float& t_time = t.time; float& t_xavg = t.xavg; float& t_epsmax = t.epsmax; float& t_wtot = t.wtot;
  const long s = s_local;
      // handle_nans (warmup.py:468-483)
      const bool success = nf == 0.f;
      if (!success) {
        for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) { c.th[i] = c.thb[i]; c.uu[i] = c.ub[i]; c.gg[i] = c.gb[i]; if (HAS_WP) store_param(c, i, c.thb[i]); }
        lp = lp_old;
        t_epsmax = eps * 0.8f;
        dE = 0.f;
        float v[2] = {0.f, 0.f};
        for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) { const float gs = c.gg[i] * sdc_at(c, i); v[0] += gs * gs; v[1] += c.uu[i] * gs; }
        all_sum<2, NT, BAR, ES>(c, v);
        g2 = v[0]; ug = v[1];
      } else {
        bool changed = false;
        for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) {
          const float u0 = c.uu[i], g0 = c.gg[i];
          const float u1 = nan_to_num(u0), g1 = nan_to_num(g0);
          if (u1 != u0 || g1 != g0 || isnan(u0) || isnan(g0)) { c.uu[i] = u1; c.gg[i] = g1; changed = true; }
        }
        lp = nan_to_num(lp);
        t_epsmax = nan_to_num(t_epsmax);
        dE = nan_to_num(dE);
        float chv[1] = {changed ? 1.f : 0.f};
        all_sum<1, NT, BAR, ES>(c, chv);
        if (chv[0] != 0.f) {
          float v[2] = {0.f, 0.f};
          for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) { const float gs = c.gg[i] * sdc_at(c, i); v[0] += gs * gs; v[1] += c.uu[i] * gs; }
          all_sum<2, NT, BAR, ES>(c, v);
          g2 = v[0]; ug = v[1];
        }
      }
      // step-size predictor (warmup.py:302-322)
      const long it = P.step_base + s;
      const float total = (float)(P.tune1 + P.tune2 + 1);
      float target;
      if (P.ev_start > 2.0f) {
        const float ex = expf(-(float)it / (total / 4.f));
        target = P.ev_start * ex + P.ev_end * (1.f - ex);
      } else {
        const float progress = fminf((float)it / total, 1.f);
        target = P.ev_start - (P.ev_start - P.ev_end) * progress;
      }
      const float decay = (P.neff - 1.f) / (P.neff + 1.f);
      const float xi = dE * dE / ((float)dim_eff(c) * target) + 1e-8f;
      const float lx = logf(xi) / (6.f * P.trust);
      const float wgt = expf(-0.5f * lx * lx);
      t_xavg = decay * t_xavg + wgt * (xi / powf(eps, 6.f));
      t_time = decay * t_time + wgt;
      float eps_new = powf(t_xavg / t_time, -1.f / 6.f);
      eps_new = (eps_new < t_epsmax ? eps_new : 0.f) + (eps_new > t_epsmax ? t_epsmax : 0.f);
      // streaming average of (x, x^2) in phase 2 (warmup.py:341-348); phase 1 keeps it at 0
      if (it >= P.tune1) {
        const float w = (success ? 1.f : 0.f) * eps_new;
        const float denom = t_wtot + w;
        for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) {
          const float x = c.th[i];
          c.avgx[i] = (t_wtot * c.avgx[i] + w * x) / denom;
          c.avgx2[i] = (t_wtot * c.avgx2[i] + w * (x * x)) / denom;
        }
        t_wtot += w;
      }
      if (P.tune_info && c.rank == 0 && c.lead) {
        float* o = P.tune_info + ((long)s * P.C + ch) * 4;
        o[0] = dE; o[1] = eps_new; o[2] = t_epsmax; o[3] = success ? 1.f : 0.f;
      }
  return eps_new;
}

// Online logsumexp over the test split for the current theta (weights already in c.wp).
template <int NT>
__device__ __forceinline__ void lppd_fold(Ctx& c, int chain) {
  const KParams& P = c.P;
  const DevModel& M = P.M;
  const int TR = M.TR;
  const long per = (P.Nt + c.G - 1) / c.G;
  const long r0 = per * c.rank, r1 = (r0 + per) < P.Nt ? (r0 + per) : P.Nt;
  for (long row0 = r0; row0 < r1; row0 += TR) {
    const int nvalid = (int)((r1 - row0) < TR ? (r1 - row0) : TR);
    const int Q = (((nvalid + 3) >> 2) + 7) & ~7;
    float* xt = c.xstream;
    load_x_tile<NT>(xt, P.Xt, row0, nvalid, Q * 4, M.sA[0]);
    __syncthreads();
    const float* out = forward_tile<NT>(c, xt, Q);
    for (int r = threadIdx.x; r < nvalid; r += NT) {
      const float lp = pointwise_lppd_row(M, out + r * M.sA[M.NL], P.yt, row0 + r);
      const long idx = (long)chain * P.Nt + row0 + r;
      const float m = P.lppd_m[idx], s = P.lppd_s[idx];
      const float nm = fmaxf(m, lp);
      const float safe = isfinite(nm) ? nm : 0.f;
      P.lppd_m[idx] = nm;
      P.lppd_s[idx] = s * expf((isfinite(m) ? m : -INFINITY) - safe) + expf(lp - safe);
    }
    __syncthreads();
  }
}

// Gradient-evaluation policy of the generic kernel: 4x4 register tiles, runtime layer shapes.
template <int NLMAX>
struct GenericGE {
  static constexpr int NT = 256;
  static __device__ __forceinline__ void prepare(Ctx&) {}
  static __device__ __forceinline__ void run(Ctx& c, long r0, long r1, float* gp) { grad_eval<NLMAX, NT>(c, r0, r1, gp); }
};

template <class GE>
__global__ void __launch_bounds__(GE::NT, 1) mile_mclmc_kernel(const __grid_constant__ KParams P) {
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
  c.xstream = c.tile + M.tile_floats;              // one streamed X tile [TR][sA[0]]
  c.xbuf = P.resident ? smem + P.off_x : c.xstream;  // resident slice or the streamed tile
  c.aux = smem + P.off_aux;
  const int d = M.d, ch = c.chain;
  const int tid = threadIdx.x;
  c.sdc = P.sdc ? P.sdc + (long)ch * d : nullptr;
  c.pmask = P.pmask; c.deff = P.d_eff;

  // ---- prologue: parameter image, state, resident X slice --------------------------------
  for (int i = tid; i < M.psize; i += NT) c.wp[i] = 0.f;
  build_pmap<NT>(M, c.pmap, P.dS);
  GE::prepare(c);
  const bool from_input = (P.mode == MODE_EVAL || P.mode == MODE_INIT || P.mode == MODE_LPPD || P.mode == MODE_PREDICT);
  const float* th_src = from_input ? P.theta_in + (long)ch * d : P.theta + (long)ch * d;
  for (int i = tid; i < d; i += NT) {
    c.th[i] = th_src[i];
    if (!from_input) { c.uu[i] = P.u[(long)ch * d + i]; c.gg[i] = P.grad[(long)ch * d + i]; }
    else { c.uu[i] = 0.f; c.gg[i] = 0.f; }
  }
  __syncthreads();
  refresh_wp<NT>(c);
  const long per = (P.N + c.G - 1) / c.G;
  const long r0 = per * c.rank < P.N ? per * c.rank : P.N;
  const long r1 = (r0 + per) < P.N ? (r0 + per) : P.N;
  const bool needs_train = (P.mode == MODE_EVAL || P.mode == MODE_INIT || P.mode == MODE_SAMPLE || P.mode == MODE_TUNE);
  if (P.resident && needs_train) {
    const int sx = M.sA[0];
    const long nv4 = (r1 - r0) * (sx >> 2), np4 = (long)P.rows_res * (sx >> 2);
    const float4* s4 = reinterpret_cast<const float4*>(P.X + r0 * sx);
    float4* d4 = reinterpret_cast<float4*>(c.xbuf);
    for (long i = tid; i < np4; i += NT) d4[i] = i < nv4 ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  if (P.mode == MODE_PREDICT) {
    // forward only: out [n, Nrows, K]
    const float* Xsrc = P.which ? P.Xt : P.X;
    const long Nr = P.which ? P.Nt : P.N;
    const int K = M.dims[M.NL];
    const long pr = (Nr + c.G - 1) / c.G;
    const long a0 = pr * c.rank, a1 = (a0 + pr) < Nr ? (a0 + pr) : Nr;
    for (long row0 = a0; row0 < a1; row0 += M.TR) {
      const int nvalid = (int)((a1 - row0) < M.TR ? (a1 - row0) : M.TR);
      const int Q = (((nvalid + 3) >> 2) + 7) & ~7;
      float* xt = c.xstream;
      load_x_tile<NT>(xt, Xsrc, row0, nvalid, Q * 4, M.sA[0]);
      __syncthreads();
      const float* out = forward_tile<NT>(c, xt, Q);
      for (int e = tid; e < nvalid * K; e += NT)
        P.pred_out[((long)ch * Nr + row0 + e / K) * K + e % K] = out[(e / K) * M.sA[M.NL] + e % K];
      __syncthreads();
    }
    if (c.G > 1 && !P.sync_mode) cluster.sync();
    return;
  }
  if (P.mode == MODE_LPPD) {
    lppd_fold<NT>(c, ch);
    if (c.G > 1 && !P.sync_mode) cluster.sync();
    return;
  }

  // ---- one loop over gradient evaluations: 1 for EVAL / INIT, 2 per MCLMC step otherwise.  A single
  //      call site keeps GE::run inlined exactly once (shared address space stays visible to ptxas).
  //      The integrator phases (B / A / refresh / reductions / tuning) touch only d ~ 500-2000 floats: they run
  //      on the first NI = 128 threads (one warp per scheduler) with named barrier 1, so the scalar math is
  //      issued by 4 warps instead of all of them; everybody else waits at the block barrier.
  constexpr int NI = NT, IB = 0;   // measured: the phases are latency-bound, so all threads (shorter per-thread chains) beat a 128-thread group
  const bool integ = tid < NI;
  const bool stepping = P.mode == MODE_SAMPLE || P.mode == MODE_TUNE;
  const bool tune = P.mode == MODE_TUNE;
  float g2 = 0.f, ug = 0.f, nf = 0.f;
  float lp = 0.f, eps = 0.f, Lc = 0.f;
  float t_time = 0.f, t_xavg = 0.f, t_epsmax = INFINITY, t_wtot = 0.f;
  const float b1 = 0.1931833275037836f, b2 = 1.f - 2.f * 0.1931833275037836f;
  const int nslot = P.refresh_mode ? 2 : 1;
  if (stepping && integ) {
    lp = P.lp[ch];
    eps = tune ? P.t_eps[ch] : P.eps[ch];
    Lc = tune ? P.t_L[ch] : P.L[ch];
    if (tune) {
      t_time = P.t_time[ch]; t_xavg = P.t_xavg[ch]; t_epsmax = P.t_epsmax[ch]; t_wtot = P.t_wtot[ch];
      for (int i = tid; i < d; i += NI) { c.avgx[i] = P.avg_x[(long)ch * d + i]; c.avgx2[i] = P.avg_x2[(long)ch * d + i]; }
    }
    if (P.carry_valid) {
      g2 = P.carry[2 * ch]; ug = P.carry[2 * ch + 1];
    } else {  // cached gradient: sum g^2 and u.g for the first B-step
      float v[2] = {0.f, 0.f};
      for (int i = tid; i < d; i += NI) { const float gs = c.gg[i] * sdc_at(c, i); v[0] += gs * gs; v[1] += c.uu[i] * gs; }
      block_sum<2, NI, IB>(v, c.red2, c.phase2);
      g2 = v[0]; ug = v[1];
    }
  }
  const int n_evals = stepping ? 2 * P.n_steps : 1;
  float lp_old = 0.f, dK = 0.f;
  PROF_DECL;
#pragma unroll 1
  for (int e = 0; e < n_evals; ++e) {
    const int h = e & 1, s = e >> 1;   // half-step h of MCLMC step s
    if (stepping && integ) {
      if (h == 0) {
        lp_old = lp; dK = 0.f;
        if (tune) for (int i = tid; i < d; i += NI) { c.thb[i] = c.th[i]; c.ub[i] = c.uu[i]; c.gb[i] = c.gg[i]; }
        if (P.refresh_mode) refresh_momentum<NI, IB>(c, 0.5f * eps, Lc, s, 0, nslot, ug);
      }
      // B(b1) A(1/2) grad | B(1-2 b1) A(1/2) grad
      dK += esh_update<NI, IB>(c, eps, h == 0 ? b1 : b2, g2, ug);
      PROF(8);
      position_update<NI>(c, eps, 0.5f);
      PROF(9);
    }
    __syncthreads();   // weight image (and, after a handle_nans restore, theta) visible to the whole block
    float* gp = c.gpart + (e & 1) * (P.dS + 4);
    GE::run(c, r0, r1, gp);
    PROF(10);
    const float2* gslab = nullptr;
    unsigned int xflag = P.xbase + (unsigned int)e + 1u;
    bool use_ll = c.G > 1 && P.sync_mode;
    int ng = c.G;
    if (c.G > 1 && P.sync_mode) {
      // global exchange: publish this CTA's partial as flagged 8-byte words; readers wait on the words themselves.
      // Two parity slabs suffice: a CTA can only publish eval e+2 after it has read every rank's eval e+1, which
      // each rank publishes after it finished reading eval e.
      float2* slab = P.xchg + ((long)ch * 2 + (e & 1)) * c.G * (P.dS + 4);
      float2* mine = slab + c.rank * (P.dS + 4);
      __syncthreads();
      for (int i = tid; i <= P.dS; i += NT) ll_store(mine + i, gp[i], xflag);
      PROF(13);
      gslab = slab;
    } else if (c.G > 1) {
      cluster.sync();      // publishes every CTA's partial gradient (DSMEM)
    } else {
      __syncthreads();
    }
    if (P.mr_world > 1) {
      // multi-rank exchange: local reduce-scatter (CTA j sums slice j of the G local partials in rank order), push of the
      // summed slice into every rank's region over NVLink, then everybody polls its own GPU's copy of the world's slices.
      // Every rank adds the same `world` values in the same order, so the replicas of a chain stay bit-identical.
      // Two parities suffice by the same argument as for the local slabs: a CTA pushes eval e+2 only after it has read
      // every rank's eval e+1, which each CTA of each rank pushes after it finished reading eval e.
      const int W = P.mr_world, stride_g = P.dS + 4;
      const int SL = (P.dS + 1 + c.G - 1) / c.G;
      const unsigned int rflag = P.mr_base + (unsigned int)e + 1u;
      const long roff = (((long)ch * 2 + (e & 1)) * W + P.mr_rank) * stride_g;
      for (int j = tid; j < SL; j += NT) {
        const int idx = c.rank * SL + j;
        if (idx <= P.dS) {
          const float sl = c.G > 1 ? ll_sum(gslab + idx, stride_g, c.G, xflag) : gp[idx];
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if (r < W) ll_store(P.mr_sums[r] + roff + idx, sl, rflag);
        }
      }
      gslab = P.mr_sums[P.mr_rank] + ((long)ch * 2 + (e & 1)) * W * stride_g;
      xflag = rflag; use_ll = true; ng = W;
    }
    if (integ) {
      const float lp_new = cluster_reduce_grad<NI, IB>(c, gp, gslab, xflag, g2, ug, nf, use_ll, ng);
      PROF(11);
      if (!stepping) {
        if (P.mode == MODE_EVAL) {
          if (c.rank == 0) {
            for (int i = tid; i < d; i += NI) P.grad_out[(long)ch * P.out_stride + i] = c.gg[i];
            if (tid == 0) P.lp_out[(long)ch * (P.out_stride == d ? 1 : P.out_stride)] = lp_new;
          }
        } else {
          // blackjax.mcmc.mclmc.init: generate_unit_vector u = z / |z|
          float v[1] = {0.f};
          for (int i = tid; i < d; i += NI) {
            const float zz = pm_at(c, i) * (P.z ? P.z[(long)ch * d + i] : philox_normal(P.seed, (uint32_t)(P.chain_base + ch), 0xFFFFFFFFFFFFFFFFull, 0u, (uint32_t)i));
            c.uu[i] = zz; v[0] += zz * zz;
          }
          block_sum<1, NI, IB>(v, c.red2, c.phase2);
          const float inv = 1.f / sqrtf(v[0]);
          if (c.rank == 0) {
            for (int i = tid; i < d; i += NI) {
              P.theta[(long)ch * d + i] = c.th[i];
              P.u[(long)ch * d + i] = c.uu[i] * inv;
              P.grad[(long)ch * d + i] = c.gg[i];
            }
            if (tid == 0) P.lp[ch] = lp_new;
          }
        }
      } else {
        lp = lp_new;
        if (h == 1) {
          // ---- end of MCLMC step s: last B, partial refresh, energy bookkeeping ----------------------
          dK += esh_update<NI, IB>(c, eps, b1, g2, ug);
          PROF(8);
          refresh_momentum<NI, IB>(c, P.refresh_mode ? 0.5f * eps : eps, Lc, s, nslot - 1, nslot, ug);
          PROF(12);
          float dE = dK - lp + lp_old;
          if (!tune) {
            if (P.info && c.rank == 0 && tid == 0) {
              float* o = P.info + ((long)s * P.C + ch) * 3;
              o[0] = lp; o[1] = dK; o[2] = dE;
            }
            // thinned sample capture (sampling.py:152-164)
            const long idx = P.step_base + s;
            if (idx % P.thin == 0) {
              const long slot = idx / P.thin - P.sample_base;
              if (P.samples && c.rank == 0 && slot >= 0 && slot < P.n_slots)
                for (int i = tid; i < d; i += NI) P.samples[(slot * P.C + ch) * d + i] = c.th[i];
            }
          } else {
            TuneRegs tr{t_time, t_xavg, t_epsmax, t_wtot};
            const float eps_new = tune_epilogue<NI, true, IB>(c, tr, eps, lp_old, nf, s, lp, dE, g2, ug);
            t_time = tr.time; t_xavg = tr.xavg; t_epsmax = tr.epsmax; t_wtot = tr.wtot;
            eps = eps_new;
          }
        }
      }
    }
    if (!stepping) break;
    // fused posterior-predictive fold at every kept position (whole block; the weight image still holds theta)
    if (h == 1 && !tune && P.do_lppd && (P.step_base + s) % P.thin == 0) {
      __syncthreads();
      lppd_fold<NT>(c, ch);
    }
  }
  // ---- epilogue: write state back ------------------------------------------------------------
  if (stepping && integ && c.rank == 0) {
    for (int i = tid; i < d; i += NI) {
      P.theta[(long)ch * d + i] = c.th[i];
      P.u[(long)ch * d + i] = c.uu[i];
      P.grad[(long)ch * d + i] = c.gg[i];
      if (tune) { P.avg_x[(long)ch * d + i] = c.avgx[i]; P.avg_x2[(long)ch * d + i] = c.avgx2[i]; }
    }
    if (tid == 0) {
      P.lp[ch] = lp;
      P.carry[2 * ch] = g2; P.carry[2 * ch + 1] = ug;
      if (tune) { P.t_time[ch] = t_time; P.t_xavg[ch] = t_xavg; P.t_epsmax[ch] = t_epsmax; P.t_eps[ch] = eps; P.t_wtot[ch] = t_wtot; }
    }
  }
  if (c.G > 1 && !P.sync_mode) cluster.sync();
}
