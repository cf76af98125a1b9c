// mile_nuts.cuh -- the NUTS branch of the sampling seam as ONE persistent kernel per launch.
//
// Reference call stack replaced: src/training/sampling.py:70-81,107-210 (scan of sampler.step with sampler = blackjax.nuts)
// and src/training/warmup.py:27-152 (`custom_window_adaptation`: nuts transition -> adapt_step per warmup step), with
// blackjax 1.2.2's iterative NUTS (mcmc/nuts.py, mcmc/trajectory.py, mcmc/termination.py, mcmc/proposal.py,
// mcmc/metrics.py diagonal Gaussian-Euclidean, velocity Verlet) and window adaptation (dual averaging + Welford).
//
// Same ownership as the MCLMC kernel (mile_kernel.cuh): a cluster of G CTAs owns one chain for the whole launch, the rows
// are split across the CTAs, every gradient evaluation ends in the same exchange, and all CTAs hold bit-identical
// state -- so the data-dependent control flow of NUTS (trajectory length, U-turn, divergence, progressive sampling) is
// taken identically by every CTA of a chain without any extra communication.  Chains diverge freely from each other.
//
// Working set per CTA:  shared memory  th / uu / gg   the state the integrator advances (position, momentum, gradient)
//                                      avgx           inverse mass matrix (diagonal)
//                                      avgx2, ub      momentum sums of the sub-trajectory / of the whole trajectory
//                                      thb, gb        proposal of the sub-trajectory (position, gradient)
//                       shared memory when it fits, else global (L2):  left / right end states, the transition's proposal, the U-turn checkpoints
//                                      [2][D][d], the Welford moments: element i is always touched by the same thread, so
//                                      these need no barriers.
#pragma once
#include "mile_kernel.cuh"

enum { MODE_NUTS = 6 };
enum { NUTS_INFO = 8 };   // floats per (transition, chain): num_integration_steps, acceptance_rate, num_trajectory_expansions,
                          // is_divergent, energy, is_turning, logdensity, step_size used

__device__ __forceinline__ float nuts_logaddexp(float a, float b) {
  if (a == b) return a + 0.6931471805599453f;
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(-fabsf(a - b)));
}

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t chain, uint64_t step, uint32_t stream, uint32_t idx) {
  uint32_t r[4];
  philox4x32_10(idx >> 2, (uint32_t)step, (uint32_t)(step >> 32), stream, (uint32_t)seed ^ (chain * 0x9E3779B9u),
                (uint32_t)(seed >> 32) + chain, r);
  const uint32_t w = (idx & 2u) ? ((idx & 1u) ? r[3] : r[2]) : ((idx & 1u) ? r[1] : r[0]);
  return (float)(w >> 8) * (1.0f / 16777216.0f);   // [0, 1)
}

// The gradient exchange of one leapfrog fused with everything that follows it element by element: sum of the G partials in rank
// order + prior (as cluster_reduce_grad, mile_kernel.cuh), second half of the momentum update p += h g, kinetic energy,
// momentum sum of the sub-trajectory, the checkpoint of an even leaf or the innermost U-turn check of an odd one -- one sweep
// over d and ONE block reduction per leapfrog.  Returns the log-density; out = {p . M^-1 p, U-turn dots (left, right)}.
template <int NT>
__device__ __forceinline__ float nuts_reduce_half_step(Ctx& c, float* gpart, const float2* gslab, unsigned int flag, const bool use_ll,
                                                       const float h, const float* imm, float* ssum, float* cp, float* cs,
                                                       const bool first, const bool even, float (&out)[3],
                                                       const float* gfull = nullptr) {
  // gfull != null: the push exchange (mile_mma.cuh) has already delivered the rank-ordered sums [0, dS] to this CTA
  const KParams& P = c.P;
  const DevModel& M = P.M;
  cg::cluster_group cluster = cg::this_cluster();
  const float* rp[16];
  const int stride_g = P.dS + 4;
#pragma unroll
  for (int r = 0; r < 16; ++r) rp[r] = (!use_ll && c.G > 1 && r < c.G) ? cluster.map_shared_rank(gpart, r) : gpart;
  float ll = 0.f;
  if (gfull) {
    ll = gfull[P.dS];
  } else if (!use_ll) {
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
    const float dlt = c.th[i] - loc;
    float pg, pv;
    if (M.prior == MILE_PRIOR_NORMAL) { pv = (lognorm + dlt * dlt / s2) / -2.f; pg = -dlt / s2; }
    else { pv = -lognorm - fabsf(dlt) / sc; pg = -((dlt > 0.f) - (dlt < 0.f)) / sc; }
    const float mk = pm_at(c, i);
    const float g = (s + pg * P.prior_weight) * mk;
    c.gg[i] = g;
    v[0] += pv * P.prior_weight * mk;
    const float p = c.uu[i] + h * g;
    c.uu[i] = p;
    const float im = imm[i];
    v[1] += im * p * p;
    const float sm = first ? p : ssum[i] + p;
    ssum[i] = sm;
    if (even) { cp[i] = p; cs[i] = sm; }
    else {
      const float pl = cp[i];
      const float rho = (sm - cs[i] + pl) - (p + pl) / 2.f;
      v[2] += im * pl * rho; v[3] += im * p * rho;
    }
  };
  if (gfull) {
    for (int i = threadIdx.x; i < M.d; i += NT) consume(i, gfull[i]);
  } else if (use_ll) {
    for (int i = threadIdx.x; i < M.d; i += 2 * NT) {
      const int i2 = i + NT;
      if (i2 < M.d) {
        float s0, s1;
        ll_sum_pair(gslab + i, gslab + i2, stride_g, c.G, flag, s0, s1);
        consume(i, s0); consume(i2, s1);
      } else {
        consume(i, ll_sum(gslab + i, stride_g, c.G, flag));
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
  if (use_ll && threadIdx.x == NT - 1) v[0] += ll_sum(gslab + P.dS, stride_g, c.G, flag);
  block_sum<4, NT, 0>(v, c.red, c.phase);
  out[0] = v[1]; out[1] = v[2]; out[2] = v[3];
  return v[0] + ll;
}

template <class GE>
__global__ void __launch_bounds__(GE::NT, 1) mile_nuts_kernel(const __grid_constant__ KParams P) {
  constexpr int NT = GE::NT;
  extern __shared__ __align__(16) float smem[];
  const DevModel& M = P.M;
  const NutsParams& Q = P.nuts;
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
  const int d = M.d, ch = c.chain, tid = threadIdx.x, D = Q.max_doublings;
  c.pmask = P.pmask; c.deff = P.d_eff;   // partition sampling: frozen parameters keep their value (zero momentum and gradient)
  float* imm = c.avgx; float* ssum = c.avgx2; float* psum = c.ub; float* sp_th = c.thb; float* sp_g = c.gb;
  // this CTA's global scratch: (10 + 2 D) vectors of dS floats
  // (in shared memory behind the plan's carve-up when it fits -- the narrow MLPs -- else in this CTA's slice of the L2-resident scratch)
  float* S = Q.smem_off >= 0 ? smem + Q.smem_off : Q.scratch + (size_t)blockIdx.x * (size_t)(10 + 2 * D) * P.dS;
  float *L_th = S, *L_p = S + P.dS, *L_g = S + 2 * P.dS, *R_th = S + 3 * P.dS, *R_p = S + 4 * P.dS, *R_g = S + 5 * P.dS;
  float *P_th = S + 6 * P.dS, *P_g = S + 7 * P.dS, *w_mean = S + 8 * P.dS, *w_m2 = S + 9 * P.dS;
  float* ck_p = S + 10 * (size_t)P.dS; float* ck_s = ck_p + (size_t)D * P.dS;

  // ---- prologue ---------------------------------------------------------------------------------------------------
  for (int i = tid; i < M.psize; i += NT) c.wp[i] = 0.f;
  build_pmap<NT>(M, c.pmap, P.dS);
  GE::prepare(c);
  for (int i = tid; i < d; i += NT) {
    c.th[i] = P.theta[(long)ch * d + i]; c.gg[i] = P.grad[(long)ch * d + i]; c.uu[i] = 0.f;
    imm[i] = Q.imm[(long)ch * d + i];
    if (Q.schedule) { w_mean[i] = Q.w_mean[(long)ch * d + i]; w_m2[i] = Q.w_m2[(long)ch * d + i]; }
  }
  __syncthreads();
  refresh_wp<NT>(c);
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
  __syncthreads();
  // DSMEM push exchange of the tensor evaluator's plans (mile_mma.cuh: reduce-scatter + all-gather with st.async + mbarrier
  // transaction counts, no cluster.sync): recv[G][SL] + gfull[dS + 4] in the plan's gslice area, two mbarriers behind the
  // reduction scratch.  Other plans keep the pull form (cluster.sync + remote loads) or the flagged words through L2.
  const bool push = c.G > 1 && !P.sync_mode && P.off_gs > 0 && Q.push;
  float* recv = smem + P.off_gs;
  float* gfull = recv + (P.dS + 16);
  const uint32_t xb1 = (uint32_t)__cvta_generic_to_shared(c.red + 160), xb2 = xb1 + 8;
  if (push) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xb1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xb2));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();      // every CTA's barriers exist before the first remote store
  }
  float lp = P.lp[ch];
  // adaptation state (window_adaptation.base): dual averaging + Welford count + the step size in use
  float* da = Q.da + (long)ch * 8;
  float da_logx = da[0], da_logx_avg = da[1], da_step = da[2], da_err = da[3], da_mu = da[4], w_count = da[5];
  float eps = da[6];
  unsigned int e = 0;   // gradient evaluations of this launch (exchange parity / flags)

#pragma unroll 1
  for (int s = 0; s < P.n_steps; ++s) {
    const long tstep = P.step_base + s;
    const float* zrow = P.z ? P.z + ((long)s * P.C + ch) * d : nullptr;
    const float* urow = Q.uni ? Q.uni + ((long)s * P.C + ch) * Q.uni_len : nullptr;
    auto uni_at = [&](int idx) -> float {
      return urow ? urow[idx] : philox_uniform(P.seed, (uint32_t)(P.chain_base + ch), (uint64_t)tstep, 17u, (uint32_t)idx);
    };
    // warm-up positions on request (warmup.py:102-109 saves the position BEFORE the transition of warm-up step n)
    if (Q.schedule && P.samples && c.rank == 0)
      for (int i = tid; i < d; i += NT) P.samples[((long)s * P.C + ch) * d + i] = c.th[i];
    // ---- momentum_generator + initial proposal / trajectory (nuts.py kernel, iterative_nuts_proposal.propose) ------
    float v1[1] = {0.f};
    for (int i = tid; i < d; i += NT) {
      const float zz = zrow ? zrow[i] : philox_normal(P.seed, (uint32_t)(P.chain_base + ch), (uint64_t)tstep, 16u, (uint32_t)i);
      const float p = pm_at(c, i) * zz / sqrtf(imm[i]);
      c.uu[i] = p; psum[i] = p;
      const float th = c.th[i], g = c.gg[i];
      L_th[i] = th; L_p[i] = p; L_g[i] = g; R_th[i] = th; R_p[i] = p; R_g[i] = g; P_th[i] = th; P_g[i] = g;
      v1[0] += imm[i] * p * p;
    }
    block_sum<1, NT, 0>(v1, c.red, c.phase);
    const float e0 = -lp + 0.5f * v1[0];
    float L_lp = lp, R_lp = lp;
    float pr_lp = lp, pr_e = e0, pr_w = 0.f, pr_sl = -INFINITY;
    int cur_side = 0;           // which end state th/uu/gg currently hold: 0 both (start), +1 right, -1 left
    int n_states = 0, n_leap = 0, step = 0;
    bool diverging = false, turning = false;
#pragma unroll 1
    while (step < D && !diverging && !turning) {
      const int dir = uni_at(step) < 0.5f ? 1 : -1;
      if (cur_side != 0 && cur_side != dir) {
        const float *sth = dir > 0 ? R_th : L_th, *sp = dir > 0 ? R_p : L_p, *sg = dir > 0 ? R_g : L_g;
        for (int i = tid; i < d; i += NT) { c.th[i] = sth[i]; c.uu[i] = sp[i]; c.gg[i] = sg[i]; }
        lp = dir > 0 ? R_lp : L_lp;
      }
      const float h = 0.5f * (float)dir * eps, fs = (float)dir * eps;
      // ---- trajectory.dynamic_progressive_integration -------------------------------------------------------------
      float sp_lp = 0.f, sp_e = 0.f, sp_w = 0.f, sp_sl = 0.f;
      bool s_div = false, s_term = false;
      int k = 0, s_n = 0;
      const int kmax = 1 << step;
#pragma unroll 1
      while (k < kmax && !s_term && !s_div) {
        // velocity Verlet: p += h g; theta += eps M^-1 p; (gradient); p += h g'
        for (int i = tid; i < d; i += NT) {
          const float p = c.uu[i] + h * c.gg[i];
          c.uu[i] = p;
          const float t = c.th[i] + fs * (imm[i] * p);
          c.th[i] = t;
          store_param(c, i, t);
        }
        __syncthreads();
        float* gp = c.gpart + (e & 1u) * (P.dS + 4);
        GE::run(c, r0, r1, gp);
        const float2* gslab = nullptr;
        const unsigned int xflag = P.xbase + e + 1u;
        const bool use_ll = c.G > 1 && P.sync_mode;
        if (use_ll) {
          float2* slab = P.xchg + ((long)ch * 2 + (e & 1u)) * c.G * (P.dS + 4);
          float2* mine = slab + c.rank * (P.dS + 4);
          __syncthreads();
          for (int i = tid; i <= P.dS; i += NT) ll_store(mine + i, gp[i], xflag);
          gslab = slab;
        } else if (push) {
          // hop 1: element i of this CTA's partial -> the owner's recv[rank][i - owner * SL]; the owner adds the G partials in
          // rank order; hop 2: each sum -> every CTA's gfull[i].  (Buffer reuse across evaluations: argument in mile_mma.cuh.)
          const int dS = P.dS;
          const int SL = (dS + 1 + c.G - 1) / c.G;
          const int my0 = c.rank * SL;
          const int mySL = my0 > dS ? 0 : (my0 + SL > dS + 1 ? dS + 1 - my0 : SL);
          const uint32_t par = e & 1u;
          __syncthreads();                     // this CTA's partial is complete
          if (tid == 0) { xbar_arm(xb1, (uint32_t)(c.G * mySL * 4)); xbar_arm(xb2, (uint32_t)((dS + 1) * 4)); }
          const uint32_t recv_s = (uint32_t)__cvta_generic_to_shared(recv), gfull_s = (uint32_t)__cvta_generic_to_shared(gfull);
          for (int i = tid; i <= dS; i += NT) {
            const int owner = i / SL, off = i - owner * SL;
            st_async_f32(mapa_u32(recv_s + (uint32_t)(c.rank * SL + off) * 4u, (uint32_t)owner), gp[i], mapa_u32(xb1, (uint32_t)owner));
          }
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
          xbar_wait(xb2, par);
        } else if (c.G > 1) {
          cluster.sync();
        } else {
          __syncthreads();
        }
        // termination.iterative_uturn_numpyro: checkpoint on even leaves, check the open sub-trees on odd ones
        const int idx_max = __popc(k >> 1);
        const int idx_min = idx_max - (__ffs(~k) - 1) + 1;
        const bool even = (k & 1) == 0;
        float* cp = ck_p + (size_t)idx_max * P.dS; float* cs = ck_s + (size_t)idx_max * P.dS;
        float v[3];
        lp = nuts_reduce_half_step<NT>(c, gp, gslab, xflag, use_ll, h, imm, ssum, cp, cs, k == 0, even, v, push ? gfull : nullptr);
        ++e;
        // proposal.update: weight = initial energy - new energy (NaN -> -inf), divergence beyond the threshold
        const float new_e = -lp + 0.5f * v[0];
        float delta = e0 - new_e;
        if (isnan(delta)) delta = -INFINITY;
        s_div = fabsf(delta) > Q.divergence_threshold;
        const float n_w = delta, n_sl = fminf(delta, 0.f);
        bool take;
        if (k == 0) {
          take = true; sp_w = n_w; sp_sl = n_sl;
        } else {   // proposal.progressive_uniform_sampling
          const float p_acc = 1.f / (1.f + expf(-(n_w - sp_w)));
          take = uni_at(2 * D + n_leap) < p_acc;
          sp_w = nuts_logaddexp(sp_w, n_w); sp_sl = nuts_logaddexp(sp_sl, n_sl);
        }
        if (take) {
          sp_lp = lp; sp_e = new_e;
          for (int i = tid; i < d; i += NT) { sp_th[i] = c.th[i]; sp_g[i] = c.gg[i]; }
        }
        if (!even) {
          s_term = (v[1] <= 0.f) || (v[2] <= 0.f);
#pragma unroll 1
          for (int j = idx_max - 1; j >= idx_min && !s_term; --j) {
            const float* qp = ck_p + (size_t)j * P.dS; const float* qs = ck_s + (size_t)j * P.dS;
            float w[2] = {0.f, 0.f};
            for (int i = tid; i < d; i += NT) {
              const float pl = qp[i], pr = c.uu[i];
              const float rho = (ssum[i] - qs[i] + pl) - (pr + pl) / 2.f;
              w[0] += imm[i] * pl * rho; w[1] += imm[i] * pr * rho;
            }
            block_sum<2, NT, 0>(w, c.red, c.phase);
            s_term = (w[0] <= 0.f) || (w[1] <= 0.f);
          }
        }
        ++k; ++n_leap; ++s_n;
      }
      // ---- trajectory.dynamic_multiplicative_expansion: merge, biased progressive sampling, whole-trajectory U-turn ---
      {
        float *eth = dir > 0 ? R_th : L_th, *ep = dir > 0 ? R_p : L_p, *eg = dir > 0 ? R_g : L_g;
        for (int i = tid; i < d; i += NT) { eth[i] = c.th[i]; ep[i] = c.uu[i]; eg[i] = c.gg[i]; }
        if (dir > 0) R_lp = lp; else L_lp = lp;
        cur_side = dir;
      }
      n_states += s_n;
      bool take = false;
      if (s_div || s_term) {
        pr_sl = nuts_logaddexp(pr_sl, sp_sl);
      } else {   // proposal.progressive_biased_sampling
        const float p_acc = fminf(1.f, expf(sp_w - pr_w));
        take = uni_at(D + step) < p_acc;
        pr_w = nuts_logaddexp(pr_w, sp_w); pr_sl = nuts_logaddexp(pr_sl, sp_sl);
        if (take) { pr_lp = sp_lp; pr_e = sp_e; }
      }
      float w[2] = {0.f, 0.f};
      for (int i = tid; i < d; i += NT) {
        const float sm = psum[i] + ssum[i];
        psum[i] = sm;
        if (take) { P_th[i] = sp_th[i]; P_g[i] = sp_g[i]; }
        const float pl = L_p[i], pr = R_p[i];
        const float rho = sm - (pr + pl) / 2.f;
        w[0] += imm[i] * pl * rho; w[1] += imm[i] * pr * rho;
      }
      block_sum<2, NT, 0>(w, c.red, c.phase);
      diverging = s_div;
      turning = s_term || (w[0] <= 0.f) || (w[1] <= 0.f);
      ++step;
    }
    // ---- the transition's result becomes the chain state ------------------------------------------------------------
    for (int i = tid; i < d; i += NT) { c.th[i] = P_th[i]; c.gg[i] = P_g[i]; }
    lp = pr_lp;
    const float acc = expf(pr_sl) / (float)(n_states > 0 ? n_states : 1);
    if (Q.info && c.rank == 0 && tid == 0) {
      float* o = Q.info + ((long)s * P.C + ch) * NUTS_INFO;
      o[0] = (float)n_states; o[1] = acc; o[2] = (float)step; o[3] = diverging ? 1.f : 0.f; o[4] = pr_e;
      o[5] = turning ? 1.f : 0.f; o[6] = lp; o[7] = eps;
    }
    if (Q.schedule) {
      // ---- window_adaptation.base.update (warmup.py:96-101): Welford in the slow stage, dual averaging always ---------
      const int sc = Q.schedule[s];
      const bool slow = sc & 1, window_end = sc & 2;
      if (slow) {
        w_count += 1.f;
        for (int i = tid; i < d; i += NT) {
          const float x = c.th[i], mu = w_mean[i];
          const float dl = x - mu, mn = mu + dl / w_count;
          w_mean[i] = mn; w_m2[i] += dl * (x - mn);
        }
      }
      {
        const float grad = Q.target_accept - acc;
        const float reg = da_step + 10.f, eta = powf(da_step, -0.75f);
        da_err = (1.f - 1.f / reg) * da_err + grad / reg;
        da_logx = da_mu - (sqrtf(da_step) / 0.05f) * da_err;
        da_logx_avg = eta * da_logx + (1.f - eta) * da_logx_avg;
        da_step += 1.f;
        eps = expf(da_logx);
      }
      if (window_end) {   // slow_final: metric of the finished window, dual averaging restarted at the averaged step size
        const float cnt = w_count;
        for (int i = tid; i < d; i += NT) {
          const float cov = w_m2[i] / (cnt - 1.f);
          imm[i] = (cnt / (cnt + 5.f)) * cov + 1e-3f * (5.f / (cnt + 5.f));
          w_mean[i] = 0.f; w_m2[i] = 0.f;
        }
        w_count = 0.f;
        const float x0 = expf(da_logx_avg);
        da_logx = logf(x0); da_logx_avg = 0.f; da_step = 1.f; da_err = 0.f; da_mu = logf(10.f * x0);
        eps = expf(da_logx);
      }
    } else {
      // thinned sample capture (sampling.py:107-177: the position after the transition)
      if (tstep % P.thin == 0) {
        const long slot = tstep / P.thin - P.sample_base;
        if (P.samples && c.rank == 0 && slot >= 0 && slot < P.n_slots)
          for (int i = tid; i < d; i += NT) P.samples[(slot * P.C + ch) * d + i] = c.th[i];
        if (P.do_lppd) {
          __syncthreads();
          refresh_wp<NT>(c);
          __syncthreads();
          lppd_fold<NT>(c, ch);
        }
      }
    }
    __syncthreads();
  }
  // ---- epilogue ---------------------------------------------------------------------------------------------------
  if (c.rank == 0) {
    for (int i = tid; i < d; i += NT) {
      P.theta[(long)ch * d + i] = c.th[i]; P.grad[(long)ch * d + i] = c.gg[i];
      if (Q.schedule) {
        Q.imm[(long)ch * d + i] = imm[i]; Q.w_mean[(long)ch * d + i] = w_mean[i]; Q.w_m2[(long)ch * d + i] = w_m2[i];
      }
    }
    if (tid == 0) {
      P.lp[ch] = lp;
      if (Q.schedule) {
        da[0] = da_logx; da[1] = da_logx_avg; da[2] = da_step; da[3] = da_err; da[4] = da_mu; da[5] = w_count; da[6] = eps;
      }
    }
  }
  if (c.G > 1 && !P.sync_mode) cluster.sync();
}
