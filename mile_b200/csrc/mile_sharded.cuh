// mile_sharded.cuh -- integrator-only kernel for the DATA-SHARDED variant (covertype):
// the training rows are split across ranks, every rank holds all chains, and each gradient evaluation is
//     local value_and_grad (MODE_EVAL of mile_mclmc_kernel, prior weighted 1/world)
//  -> ncclAllReduce(sum) of the packed [C, d+1] buffer (gradient | log-density)
//  -> this kernel: the B / A / refresh / bookkeeping stage that consumes it.
// Every rank applies the identical update to its replica of the chain state (same inputs, same order), so the
// replicas stay bit-identical without any further exchange.  One CTA per chain.
#pragma once
#include "mile_kernel.cuh"

enum { SH_BEGIN = 0, SH_MID = 1, SH_END = 2 };

struct ShardParams {
  KParams K;              // model, chain state pointers, eps/L, z/seed, thinning, samples, info, tuning fields
  const float* gl;        // [C, d+1] all-reduced (gradient | log-density) -- consumed by SH_MID / SH_END
  float* scal;            // [C, 4]  lp_old, dK carried between the three launches of a step
  float* thb; float* ub; float* gb;   // [C,d] state backups for handle_nans (tune mode)
  int stage, tune;
  long s_local;           // index of the step inside this call (noise / info / tune_info addressing)
  // Peer-memory all-reduce (world > 1, CUDA IPC over NVLink): instead of an ncclAllReduce between the gradient kernel and
  // this one, every rank's gradient kernel leaves its partial [C, d+1] in its own exchange buffer, and this kernel --
  // the consumer -- publishes "my partial of evaluation `epoch` is complete", waits for the same flag of every peer and
  // sums the partials straight out of peer memory in rank order (identical on every rank, so the replicas stay
  // bit-identical).  Compute step and collective in one kernel: no NCCL launch, no extra pass over the buffer.
  int p2p, world, rank;
  unsigned int epoch;
  const float* peer_data[8];        // rank r's partial of this evaluation (peer pointer; own buffer for r == rank)
  unsigned int* peer_flag[8];       // rank r's completion flag for this parity
};

__device__ __forceinline__ float ld_peer(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");   // peer memory must not be served by L1
  return v;
}
// sum of the world's partials at packed index k (rank order; all loads in flight)
__device__ __forceinline__ float p2p_sum(const ShardParams& S, long k) {
  float t[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) t[r] = r < S.world ? ld_peer(S.peer_data[r] + k) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) s += t[r];
  return s;
}

// ES = element split: a cluster of CTAs per chain, the d elements strided over all of its threads and every reduction
// completed over DSMEM -- the large-d (wide) path, where one CTA per chain would be a latency-bound 200k-element loop.
template <int NT, bool ES = false>
__global__ void __launch_bounds__(NT, 1) mile_integrator_kernel(const __grid_constant__ ShardParams S) {
  // works IN PLACE on the global chain state (every element is owned by one thread), so it has
  // no shared-memory size limit and also serves the wide / large-d path (mile_wide.cuh)
  __shared__ __align__(16) float red[2 * 4 * (NT / 32) + 64];
  __shared__ float csum[16];
  const KParams& P = S.K;
  const int d = P.M.d, tid = threadIdx.x;
  int ch = blockIdx.x, crank = 0, csize = 1;
  if (ES) {
    cg::cluster_group cl = cg::this_cluster();
    csize = (int)cl.num_blocks(); crank = (int)cl.block_rank(); ch = blockIdx.x / csize;
  }
  Ctx c(P);
  c.G = 1; c.rank = crank; c.chain = ch; c.phase = 0; c.phase2 = 0; c.lead = tid == 0;
  c.e0 = crank * NT + tid; c.estride = csize * NT; c.csum = csum; c.csum_phase = 0;
  const bool writer = tid == 0 && crank == 0;
  c.th = P.theta + (long)ch * d; c.uu = P.u + (long)ch * d; c.gg = P.grad + (long)ch * d;
  c.thb = S.thb + (long)ch * d; c.ub = S.ub + (long)ch * d; c.gb = S.gb + (long)ch * d;
  c.avgx = P.avg_x + (long)ch * d; c.avgx2 = P.avg_x2 + (long)ch * d; c.red = red; c.red2 = red + 2 * 4 * (NT / 32);
  c.wp = nullptr; c.pmap = nullptr; c.gpart = nullptr; c.tile = nullptr; c.xbuf = nullptr; c.xstream = nullptr;
  c.sdc = P.sdc ? P.sdc + (long)ch * d : nullptr;
  const bool fresh = S.stage != SH_BEGIN;      // a newly all-reduced gradient arrives with MID / END
  const bool p2p = S.p2p && fresh;
  if (p2p) {
    // (the gradient kernel of this rank finished before this kernel started: stream order)
    if (blockIdx.x == 0 && tid == 0) {
      __threadfence_system();
      asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(S.peer_flag[S.rank]), "r"(S.epoch) : "memory");
    }
    if (tid < S.world && tid != S.rank) {
      unsigned int f;
      long spin = 0;
      do {
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(f) : "l"(S.peer_flag[tid]) : "memory");
        if ((int)(f - S.epoch) < 0 && ++spin > (1L << 26)) __trap();     // a dead peer must not hang the GPU
      } while ((int)(f - S.epoch) < 0);
      __threadfence_system();
    }
    __syncthreads();
  }
  float lp = p2p ? p2p_sum(S, (long)ch * (d + 1) + d) : (fresh ? S.gl[(long)ch * (d + 1) + d] : P.lp[ch]);
  if (p2p) {
    // sum the ranks' partial gradients out of peer memory into the chain's gradient array, four elements x world loads in
    // flight per thread (an NVLink round trip is ~2 us: one element at a time would serialise d / NT of them)
    const int IS = MILE_IS(c, ES, NT);
    for (int i = MILE_I0(c, ES); i < d; i += 4 * IS) {
      float t[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int r = 0; r < 8; ++r)
          t[k][r] = (i + k * IS < d && r < S.world) ? ld_peer(S.peer_data[r] + (long)ch * (d + 1) + i + k * IS) : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i + k * IS < d) {
          float sum = 0.f;
#pragma unroll
          for (int r = 0; r < 8; ++r) sum += t[k][r];
          c.gg[i + k * IS] = sum;
        }
    }
  }
  // (the copy and the sums touch the same elements in the same thread: no barrier needed in between)
  float v[3] = {0.f, 0.f, 0.f};
  {
    float* __restrict__ gg = c.gg;
    const float* __restrict__ uu = c.uu;
    const float* __restrict__ th = c.th;
    const float* __restrict__ gl = S.gl + (long)ch * (d + 1);
    const bool copy = fresh && !p2p;
#pragma unroll 4
    for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) {
      const float gi = copy ? gl[i] : gg[i];   // (p2p: this thread just wrote gg[i])
      if (copy) gg[i] = gi;
      const float gs = gi * sdc_at(c, i);
      v[0] += gs * gs; v[1] += uu[i] * gs; v[2] += isfinite(th[i]) ? 0.f : 1.f;
    }
  }
  all_sum<3, NT, 0, ES>(c, v);
  float g2 = v[0], ug = v[1];
  const float nf = v[2];
  const bool tune = S.tune != 0;
  float eps = tune ? P.t_eps[ch] : P.eps[ch];
  const float Lc = tune ? P.t_L[ch] : P.L[ch];
  const float b1 = 0.1931833275037836f, b2 = 1.f - 2.f * 0.1931833275037836f;
  const int nslot = P.refresh_mode ? 2 : 1;
  float lp_old = S.scal[ch * 4 + 0], dK = S.scal[ch * 4 + 1];
  if (S.stage == SH_BEGIN) {
    lp_old = lp; dK = 0.f;
    if (tune) for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) { c.thb[i] = c.th[i]; c.ub[i] = c.uu[i]; c.gb[i] = c.gg[i]; }
    if (P.refresh_mode) refresh_momentum<NT, 0, ES>(c, 0.5f * eps, Lc, S.s_local, 0, nslot, ug);
  }
  if (S.stage != SH_END) {
    dK += esh_update<NT, 0, ES>(c, eps, S.stage == SH_BEGIN ? b1 : b2, g2, ug);
    const float st = eps * 0.5f;
    float* __restrict__ th = c.th;
    const float* __restrict__ uu = c.uu;
#pragma unroll 4
    for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) th[i] += st * sdc_at(c, i) * uu[i];
  } else {
    dK += esh_update<NT, 0, ES>(c, eps, b1, g2, ug);
    refresh_momentum<NT, 0, ES>(c, P.refresh_mode ? 0.5f * eps : eps, Lc, S.s_local, nslot - 1, nslot, ug);
    float dE = dK - lp + lp_old;
    if (!tune) {
      if (P.info && writer) { float* o = P.info + ((long)S.s_local * P.C + ch) * 3; o[0] = lp; o[1] = dK; o[2] = dE; }
      const long idx = P.step_base + S.s_local;
      if (idx % P.thin == 0) {
        const long slot = idx / P.thin - P.sample_base;
        if (P.samples && slot >= 0 && slot < P.n_slots)
          for (int i = MILE_I0(c, ES); i < d; i += MILE_IS(c, ES, NT)) P.samples[(slot * P.C + ch) * d + i] = c.th[i];
      }
    } else {
      TuneRegs tr{P.t_time[ch], P.t_xavg[ch], P.t_epsmax[ch], P.t_wtot[ch]};
      eps = tune_epilogue<NT, false, 0, ES>(c, tr, eps, lp_old, nf, S.s_local, lp, dE, g2, ug);
      if (writer) { P.t_time[ch] = tr.time; P.t_xavg[ch] = tr.xavg; P.t_epsmax[ch] = tr.epsmax; P.t_wtot[ch] = tr.wtot; P.t_eps[ch] = eps; }
    }
  }
  if (writer) { P.lp[ch] = lp; S.scal[ch * 4 + 0] = lp_old; S.scal[ch * 4 + 1] = dK; }
  if (ES) cg::this_cluster().sync();   // no CTA may exit while a peer still reads its reduction slots
}

// blackjax generate_unit_vector for all chains: u = z / |z| (z host-supplied or the same Philox draw MODE_INIT uses)
template <int NT>
__global__ void __launch_bounds__(NT) mile_unit_momentum_kernel(float* __restrict__ u, const float* __restrict__ z0,
                                                                unsigned long long seed, int d, int chain_base) {
  __shared__ __align__(16) float red[64];
  int phase = 0;
  const int ch = blockIdx.x;
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < d; i += NT) {
    const float zz = z0 ? z0[(long)ch * d + i] : philox_normal(seed, (uint32_t)(chain_base + ch), 0xFFFFFFFFFFFFFFFFull, 0u, (uint32_t)i);
    u[(long)ch * d + i] = zz; v[0] += zz * zz;
  }
  block_sum<1, NT>(v, red, phase);
  const float inv = 1.f / sqrtf(v[0]);
  for (int i = threadIdx.x; i < d; i += NT) u[(long)ch * d + i] *= inv;
}
