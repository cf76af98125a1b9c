// mile_api.cu -- host side of the C ABI declared in include/mile_b200.h.
// Owns the device buffers of one ensemble wave, plans the shared-memory carve-up and the
// cluster shape, and launches the persistent kernel of mile_kernel.cuh.  Links cudart only.
#include "mile_mma.cuh"
#include "mile_sharded.cuh"
#include "mile_wide.cuh"
#include "mile_train.cuh"
#include "mile_nuts.cuh"
#include "mile_ess.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <tuple>
#include <vector>

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                  std::to_string(__LINE__) + ")");                                            \
  } while (0)

static const size_t kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

// ---- data-sharded variant: NCCL resolved at run time (dlopen) so the library links cudart only ----
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.ok) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);   // the copy torch already loaded, else the system one
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(std::string("cannot load libnccl: ") + dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) return fail("libnccl lacks required symbols");
  g_nccl.ok = true;
  return 0;
}

struct mile_ctx {
  mile_model_desc desc;
  DevModel M;
  int C = 0, device = 0, d = 0;
  // options
  int opt_cluster = 0, opt_tile_rows = 0, opt_refresh = 0, opt_resident = -1, opt_fast = 2, opt_tensor = 2, opt_chain_base = 0, opt_steploop = 1, opt_kslices = 0;   // fast: 0 generic tiles, 1 FFMA layer pipeline (mile_fast.cuh), 2 register-chained 3xTF32 MMA evaluator (mile_mma.cuh);   // tensor: 0 SIMT, 1 tcgen05 (staged), 2 (default) tcgen05 TMA-fed, one CTA per tile, 3 = 2 + CTA pairs (cta_group::2) for the K-major GEMMs (measured 5 % slower, profiles/r3a_*)
  // data
  float* X = nullptr; void* y = nullptr; long N = 0;
  float* Xt = nullptr; void* yt = nullptr; long Nt = 0;
  // state
  float *theta = nullptr, *u = nullptr, *grad = nullptr, *lp = nullptr;
  float *t_time = nullptr, *t_xavg = nullptr, *t_epsmax = nullptr, *t_eps = nullptr, *t_L = nullptr,
        *t_wtot = nullptr, *avg_x = nullptr, *avg_x2 = nullptr;
  float *lppd_m = nullptr, *lppd_s = nullptr; long lppd_count = 0;
  float* carry = nullptr; int carry_valid = 0;
  float *tr_m = nullptr, *tr_v = nullptr; int* tr_t = nullptr;   // warm-start training: AdamW moments [C,d] and step counts [C]
  float2* xchg = nullptr; unsigned int xepoch = 0; size_t xchg_bytes = 0; int n_sms = 148, opt_sync = -1;
  // data-sharded variant (rows split across ranks, NCCL all-reduce per gradient evaluation)
  void* nccl_comm = nullptr; int world = 1, rank = 0; int shard_lppd = 0;
  // peer-memory all-reduce (mile_sharded.cuh): own exchange region [2 parities][xr_n floats] + 2 flags, peers' regions via CUDA IPC
  float* xr = nullptr; size_t xr_n = 0; unsigned int xr_epoch = 0; int p2p = 0; float* xr_peer[8] = {nullptr};
  // multi-rank exchange of the fused step loop (KParams::mr_*): float offset of the [C][2][world][dS+4] float2 area inside
  // the exchange region, flag epoch (advanced identically on every rank), option switch (mile_set_option "shard_fused")
  size_t mr_off = 0; unsigned int mr_epoch = 0; int opt_shard_fused = 1;
  float *gl = nullptr, *scal = nullptr, *thb = nullptr, *ub = nullptr, *gb = nullptr;
  // wide / large-d path (mile_wide.cuh): HBM-resident activations, chain-batched GEMMs
  int wide = 0; long w_rows = 0; int w_chains = 0, w_kslices = 1, w_nblk = 0;
  float *w_act = nullptr, *w_delta[2] = {nullptr, nullptr}, *w_part = nullptr, *w_llpart = nullptr, *w_ones = nullptr;
  float* w_gl = nullptr;   // packed [n, d+1] output of a stand-alone value_and_grad call
  float* wp_act = nullptr; size_t wp_act_floats = 0;   // activations of a forward-only pass over a split (mile_predict / LPPD)
  float* wp_out = nullptr; size_t wp_out_floats = 0;   // its [n, N, K] outputs when the caller wants them folded (LPPD)
  // tcgen05 v2: tf32 remainders of activations / deltas / weights + cached TMA tensor maps
  long w_part_per_chain = 0;
  // NUTS branch (mile_nuts.cuh): inverse mass matrix / Welford moments [C,d], dual-averaging state [C,8], per-CTA scratch
  float *nuts_imm = nullptr, *nuts_mean = nullptr, *nuts_m2 = nullptr, *nuts_da = nullptr, *nuts_scratch = nullptr;
  size_t nuts_scratch_floats = 0; int opt_nuts_smem = 1, opt_nuts_push = 1, nuts_max_doublings = 10; float nuts_div = 1000.f, nuts_target = 0.8f;
  float* pmask = nullptr; int pmask_on = 0, d_eff = 0;   // partition sampling: 1 = sampled / 0 = frozen per parameter
  float* sdc = nullptr; int sdc_on = 0;      // diagonal preconditioner [C][d] (warmup.py:391-393); used when sdc_on
  int integ_cluster = 8;      // cluster size of the large-d integrator kernel (16 = non-portable size: measured slower, 69 vs 55 us)
  float* w_fin = nullptr;     // finalize scratch: [chains][32 CTAs][2] partial scalars, then [chains] arrival tickets
  float* w_arena = nullptr; size_t w_arena_floats = 0;   // partial sums of one evaluation, summed by wide_finalize_kernel (WideJobs)
  int opt_head_fused = 1;                                 // wide path: fused output-layer kernel (0 = separate streaming kernels)
  float *w_wpk = nullptr, *w_wpk_lo = nullptr, *w_wpkT = nullptr,
        *w_wpkT_lo = nullptr;
  long w_n8 = 0, w_wstride = 0; long w_woff[MILE_MAX_LAYERS] = {0};
  std::map<std::tuple<const void*, long, long, long, long, int, int, int>, CUtensorMap> tmaps;
  // staging for the *_host entry points
  std::vector<std::pair<void*, size_t>> scratch;  // slot -> (ptr, bytes)
  cudaStream_t own_stream = nullptr;
  long launches = 0;
};

static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static int stride_for(int w) { return round_up(w, 4); }  // dense rows: see DevModel comment

static void build_model(mile_ctx* c) {
  const mile_model_desc& D = c->desc;
  DevModel& M = c->M;
  memset(&M, 0, sizeof(M));
  M.F = D.n_features; M.NL = D.n_layers; M.act = D.activation; M.task = D.task; M.prior = D.prior;
  M.prior_loc = D.prior_loc; M.prior_scale = D.prior_scale; M.n_batches = D.n_batches;
  M.dims[0] = D.n_features;
  for (int l = 0; l < M.NL; ++l) M.dims[l + 1] = D.widths[l];
  int d = 0, ps = 0;
  for (int l = 0; l <= M.NL; ++l) { M.dimp[l] = round_up(M.dims[l], 4); M.sA[l] = stride_for(M.dims[l]); }
  for (int l = 0; l < M.NL; ++l) {
    M.bias_off[l] = D.bias_off[l]; M.kern_off[l] = D.kernel_off[l];
    d += M.dims[l + 1] + M.dims[l] * M.dims[l + 1];
    M.pb_off[l] = ps; ps += M.dimp[l + 1];
    M.pw_off[l] = ps; ps += M.dimp[l] * M.dimp[l + 1];
    if (l >= 1) { M.pwt_off[l] = ps; ps += M.dimp[l] * M.dimp[l + 1]; } else M.pwt_off[l] = M.pw_off[l];
  }
  M.d = d; M.psize = ps;
  c->d = d;
}

struct Plan {
  int G, TR, resident, rows_res, fast = 0, fast_fp = 0, sync_mode = 0;
  size_t smem;
  KParams kp;  // offsets + model filled in
};

static int make_plan_impl(const mile_ctx* c, int n_chains, long nrows_for_split, bool want_resident, Plan& pl, int force_g,
                          int force_sync) {
  DevModel M = c->M;
  int G = force_g > 0 ? force_g : c->opt_cluster, sync_mode = 0;
  if (G <= 0) {
    if (n_chains * 8 <= 128) G = 8; else if (n_chains * 4 <= 148) G = 4; else if (n_chains * 2 <= 148) G = 2; else G = 1;
    while (G > 1 && nrows_for_split / G < 64) G >>= 1;
    // few chains: more than 8 CTAs per chain only fit with the global-memory exchange (cooperative launch)
    const int gmax = c->n_sms / n_chains;
    if (c->opt_sync != 0 && G == 8 && gmax > 9 && nrows_for_split / gmax >= 64) { G = gmax > 16 ? 16 : gmax; sync_mode = 1; }
  } else if (c->opt_sync == 1 || (c->opt_sync != 0 && force_sync != 0 && ((G > 8) || (G & (G - 1))))) {
    sync_mode = 1;      // (sync_mode 0 asked for explicitly: a thread-block cluster of any size <= 16, non-portable above 8)
  }
  if (force_sync == 1) sync_mode = 1;    // (multi-rank step loop: the exchange is flagged words only, every CTA co-resident)
  if (G < 1 || G > 16) return fail("cluster_size must be in [1, 16]");
  if (sync_mode && (long)G * n_chains > c->n_sms) return fail("cluster_size x chains exceeds the SM count (cooperative launch)");
  if (G == 1) sync_mode = 0;
  const long rows_cta = (nrows_for_split + G - 1) / G;
  const int dS = round_up(M.d, 4);
  int S1 = 0;
  for (int l = 1; l <= M.NL; ++l) S1 += M.sA[l];
  const size_t fixed = (size_t)M.psize + 8 * (size_t)dS + 2 * (size_t)(dS + 4) + 2 * dS + 192;
  // ---- register-chained tensor evaluator (mile_mma.cuh): hidden width 16 + Gaussian head, relu, F <= 16 ----
  {
    bool ok = want_resident && c->opt_fast >= 2 && M.task == MILE_TASK_REGRESSION && M.dims[M.NL] == 2 &&
              (M.NL == 3 || M.NL == 4) && M.act == MILE_ACT_RELU && M.dims[0] <= 16;
    for (int l = 1; ok && l < M.NL; ++l) ok = M.dims[l] == 16;
    if (ok) {
      const int FPm = M.dims[0] <= 8 ? 8 : 16;
      int tile_f = 0, aux_i = 0;
      if (M.NL == 3 && FPm == 8) { tile_f = MmaGE<3, 8, 512>::TILE_FLOATS; aux_i = MmaGE<3, 8, 512>::AUX_INTS; }
      else if (M.NL == 3) { tile_f = MmaGE<3, 16, 512>::TILE_FLOATS; aux_i = MmaGE<3, 16, 512>::AUX_INTS; }
      else if (FPm == 8) { tile_f = MmaGE<4, 8, 512>::TILE_FLOATS; aux_i = MmaGE<4, 8, 512>::AUX_INTS; }
      else { tile_f = MmaGE<4, 16, 512>::TILE_FLOATS; aux_i = MmaGE<4, 16, 512>::AUX_INTS; }
      const int TRg = 64;  // tile of the generic forward used by the fused lppd fold / predict
      const size_t gen = (size_t)2 * TRg * S1;
      const size_t tile = (size_t)tile_f > gen ? (size_t)tile_f : gen;
      // resident slice: 16-row tiles plus one zero guard tile (k-steps may read up to 12 floats past a row's stride)
      const int rows_res = (int)((rows_cta + 15) / 16) * 16 + 16;
      const size_t base_need = (fixed + tile + (size_t)TRg * M.sA[0] + round_up(aux_i, 4) + 6 * (size_t)dS + 32) * 4;
      const int res = (c->opt_resident != 0 && base_need + (size_t)rows_res * M.sA[0] * 4 <= kSmemLimit) ? 1 : 0;
      if (base_need <= kSmemLimit) {
        M.TR = TRg; M.tile_floats = (int)tile;
        int off = 0;
        for (int l = 1; l <= M.NL; ++l) { M.a_off[l] = off; off += TRg * M.sA[l]; }
        for (int l = 0; l < M.NL; ++l) { M.d_off[l] = off; off += TRg * M.sA[l + 1]; }
        KParams& k = pl.kp;
        memset(&k, 0, sizeof(k));
        k.M = M; k.dS = dS;
        int o = 0;
        k.off_wp = o; o += round_up(M.psize, 4);
        k.off_th = o; o += dS; k.off_u = o; o += dS; k.off_g = o; o += dS;
        k.off_thb = o; o += dS; k.off_ub = o; o += dS; k.off_gb = o; o += dS;
        k.off_avgx = o; o += dS; k.off_avgx2 = o; o += dS;
        k.off_gpart = o; o += 2 * (dS + 4);
        k.off_pmap = o; o += 2 * dS;
        k.off_red = o; o += 192;
        k.off_aux = o; o += round_up(aux_i, 4);
        k.off_z = o; o += 4 * dS;            // refresh noise of two steps x two slots
        k.off_gs = o; o += 2 * (dS + 16);    // DSMEM push exchange: recv[G][SL] slots, then the gathered sums
        k.off_tile = o; o += (int)tile + TRg * M.sA[0];
        k.off_x = o; if (res) o += rows_res * M.sA[0];
        pl.G = G; pl.TR = 16; pl.resident = res; pl.rows_res = rows_res; pl.fast = 2; pl.fast_fp = FPm; pl.sync_mode = sync_mode;
        pl.smem = (size_t)o * 4;
        k.G = G; k.resident = res; k.rows_res = rows_res; k.C = n_chains;
        return 0;
      }
    }
  }
  // ---- fast path: warp-specialised pipeline (mile_fast.cuh) for hidden width 16 + Gaussian head ----
  {
    bool ok = want_resident && c->opt_fast >= 1 && M.task == MILE_TASK_REGRESSION &&
              M.dims[M.NL] == 2 && (M.NL == 3 || M.NL == 4) && M.act == MILE_ACT_RELU &&
              (M.dimp[0] == 8 || M.dimp[0] == 12);
    for (int l = 1; ok && l < M.NL; ++l) ok = M.dims[l] == 16;
    if (ok) {
      const int T = M.NL == 3 ? 64 : 32, depth = 2 * M.NL - 1, nbuf = 2 * (M.NL - 1);
      const int TRg = 64;  // tile of the generic forward used by the fused lppd fold inside the fast kernel
      size_t ring = (size_t)depth * nbuf * T * 16, gen = (size_t)2 * TRg * S1;
      size_t tile = ring > gen ? ring : gen;
      const int rows_res = (int)((rows_cta + T - 1) / T) * T;
      const size_t base_need = (fixed + tile + (size_t)TRg * M.sA[0]) * 4;
      const int res = (c->opt_resident != 0 && base_need + (size_t)rows_res * M.sA[0] * 4 <= kSmemLimit) ? 1 : 0;
      if (base_need <= kSmemLimit) {
        M.TR = TRg; M.tile_floats = (int)tile;
        int off = 0;
        for (int l = 1; l <= M.NL; ++l) { M.a_off[l] = off; off += TRg * M.sA[l]; }
        for (int l = 0; l < M.NL; ++l) { M.d_off[l] = off; off += TRg * M.sA[l + 1]; }
        KParams& k = pl.kp;
        memset(&k, 0, sizeof(k));
        k.M = M; k.dS = dS;
        int o = 0;
        k.off_wp = o; o += round_up(M.psize, 4);
        k.off_th = o; o += dS; k.off_u = o; o += dS; k.off_g = o; o += dS;
        k.off_thb = o; o += dS; k.off_ub = o; o += dS; k.off_gb = o; o += dS;
        k.off_avgx = o; o += dS; k.off_avgx2 = o; o += dS;
        k.off_gpart = o; o += 2 * (dS + 4);
        k.off_pmap = o; o += 2 * dS;
        k.off_red = o; o += 192;
        k.off_tile = o; o += (int)tile + TRg * M.sA[0];
        k.off_x = o; if (res) o += rows_res * M.sA[0];
        pl.G = G; pl.TR = T; pl.resident = res; pl.rows_res = rows_res; pl.fast = 1; pl.fast_fp = M.dimp[0]; pl.sync_mode = sync_mode;
        pl.smem = (size_t)o * 4;
        k.G = G; k.resident = res; k.rows_res = rows_res; k.C = n_chains;
        return 0;
      }
    }
  }
  int TR = c->opt_tile_rows > 0 ? round_up(c->opt_tile_rows, 32) : 256;
  const int want = round_up((int)(rows_cta < 32 ? 32 : (rows_cta > 256 ? 256 : rows_cta)), 32);
  if (c->opt_tile_rows <= 0 && TR > want) TR = want;
  for (;; TR -= 32) {
    if (TR < 32) return fail("model too large for the CUDA-core path (shared memory): use the wide path");
    size_t tile = (size_t)2 * TR * S1;
    if (tile < MILE_THREADS * 20) tile = MILE_THREADS * 20;
    const size_t need = (fixed + tile + (size_t)TR * M.sA[0]) * 4;
    if (need <= kSmemLimit) break;
    if (c->opt_tile_rows > 0) return fail("tile_rows does not fit in shared memory");
  }
  size_t tile = (size_t)2 * TR * S1;
  if (tile < MILE_THREADS * 20) tile = MILE_THREADS * 20;
  M.TR = TR; M.tile_floats = (int)tile;
  int off = 0;
  for (int l = 1; l <= M.NL; ++l) { M.a_off[l] = off; off += TR * M.sA[l]; }
  for (int l = 0; l < M.NL; ++l) { M.d_off[l] = off; off += TR * M.sA[l + 1]; }
  const int rows_res = (int)((rows_cta + TR - 1) / TR) * TR;
  size_t base = fixed + tile + (size_t)TR * M.sA[0];
  int resident = 0;
  if (want_resident && c->opt_resident != 0 && (base + (size_t)rows_res * M.sA[0]) * 4 <= kSmemLimit) resident = 1;
  KParams& k = pl.kp;
  memset(&k, 0, sizeof(k));
  k.M = M; k.dS = dS;
  int o = 0;
  k.off_wp = o; o += round_up(M.psize, 4);
  k.off_th = o; o += dS; k.off_u = o; o += dS; k.off_g = o; o += dS;
  k.off_thb = o; o += dS; k.off_ub = o; o += dS; k.off_gb = o; o += dS;
  k.off_avgx = o; o += dS; k.off_avgx2 = o; o += dS;
  k.off_gpart = o; o += 2 * (dS + 4);
  k.off_pmap = o; o += 2 * dS;
  k.off_red = o; o += 192;
  k.off_tile = o; o += (int)tile + TR * M.sA[0];
  k.off_x = o; if (resident) o += rows_res * M.sA[0];
  pl.G = G; pl.TR = TR; pl.resident = resident; pl.rows_res = rows_res; pl.sync_mode = sync_mode;
  pl.smem = (size_t)o * 4;
  if (pl.smem > kSmemLimit) return fail("internal: shared-memory plan exceeds the limit");
  k.G = G; k.resident = resident; k.rows_res = rows_res; k.C = n_chains;
  for (int l = 0; l < M.NL; ++l)
    if ((M.dimp[l] / 4) * (M.dimp[l + 1] / 4) > MILE_THREADS)
      return fail("layer too wide for the CUDA-core path (more than 256 4x4 weight tiles): use the wide path");
  return 0;
}

template <class GE, int V2 = 0>   // V2: 0 generic step loop, 1 restructured step loop of the tensor evaluator, 2 NUTS
static int launch_t(const Plan& pl, int n_chains, cudaStream_t st) {
  void (*kern)(const KParams) = mile_mclmc_kernel<GE>;
  if constexpr (V2 == 1) kern = mile_mma_step_kernel<GE>;   // restructured step loop of the tensor evaluator (mile_mma.cuh)
  if constexpr (V2 == 2) kern = mile_nuts_kernel<GE>;       // NUTS transitions (mile_nuts.cuh)
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  if (pl.G > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(n_chains * pl.G), 1, 1);
  cfg.blockDim = dim3(GE::NT, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pl.sync_mode) {   // global-memory exchange: every CTA must be co-resident -> cooperative launch, no cluster
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
  } else {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  }
  cfg.attrs = attr; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, pl.kp));
  return 0;
}

// Can `n_chains` clusters of `G` CTAs of the tensor evaluator's step kernel be resident at the same time?  (cached)
static bool mma_clusters_fit(const mile_ctx* c, const Plan& pl, int n_chains) {
  static std::map<std::tuple<int, int, int, size_t, int, int>, bool> cache;
  const auto key = std::make_tuple(c->device, c->M.NL, pl.fast_fp, pl.smem, pl.G, n_chains);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  void (*kern)(const KParams) = nullptr;
  const int NL = c->M.NL;
  if (NL == 3 && pl.fast_fp == 8) kern = mile_mma_step_kernel<MmaGE<3, 8, 512>>;
  else if (NL == 3) kern = mile_mma_step_kernel<MmaGE<3, 16, 512>>;
  else if (pl.fast_fp == 8) kern = mile_mma_step_kernel<MmaGE<4, 8, 512>>;
  else kern = mile_mma_step_kernel<MmaGE<4, 16, 512>>;
  bool ok = false;
  if (cudaSetDevice(c->device) == cudaSuccess &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit) == cudaSuccess &&
      (pl.G <= 8 || cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess)) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(n_chains * pl.G), 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = pl.smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess) ok = n >= n_chains;
    else cudaGetLastError();
  }
  cache[key] = ok;
  return ok;
}

// Plan of a launch.  Automatic choice for FEW chains on the tensor evaluator: the L2 flagged-word exchange lets a chain use
// up to 16 CTAs, but when the row slice of a CTA is at most one wave of 16-row tiles either way (latency-bound regime:
// airfoil) a thread-block cluster with the DSMEM push exchange is faster even with fewer CTAs (9 x 12 chains: 834 k
// chain-steps/s against 769 k for 12 x 12 through L2, profiles/r3h_*).  The largest cluster size (<= 16, non-portable
// above 8) whose clusters are all co-resident is taken.
static int make_plan(const mile_ctx* c, int n_chains, long nrows_for_split, bool want_resident, Plan& pl, int force_g = 0,
                     bool force_sync = false) {
  const int rc = make_plan_impl(c, n_chains, nrows_for_split, want_resident, pl, force_g, force_sync ? 1 : -1);
  if (rc || force_g > 0 || force_sync || c->opt_cluster > 0 || c->opt_sync >= 0 || pl.fast != 2 || !pl.sync_mode) return rc;
  const std::string keep_err = g_err;
  for (int gc = pl.G < 16 ? pl.G : 16; gc >= 9; --gc) {
    const long rows_cta = (nrows_for_split + gc - 1) / gc;
    if ((rows_cta + 15) / 16 > 8) break;             // more than two tiles per scheduler: the extra CTAs of the L2 form pay
    Plan alt;
    if (make_plan_impl(c, n_chains, nrows_for_split, want_resident, alt, gc, 0) || alt.fast != 2 || alt.sync_mode) continue;
    if (mma_clusters_fit(c, alt, n_chains)) { pl = alt; break; }
  }
  g_err = keep_err;
  return 0;
}

static int launch(mile_ctx* c, Plan& pl, int n_chains, cudaStream_t st) {
  CK(cudaSetDevice(c->device));
  pl.kp.sync_mode = pl.sync_mode;
  if (pl.sync_mode) {
    const size_t nA = (size_t)n_chains * 2 * pl.G * (pl.kp.dS + 4), nB = (size_t)n_chains * 2 * (pl.kp.dS + 4);
    const size_t need = (nA + nB) * sizeof(float2);   // partials of every rank + the summed slices (reduce-scatter form)
    unsigned int adv = 2u * (unsigned int)(pl.kp.n_steps > 0 ? pl.kp.n_steps : 0) + 2u;
    if (pl.kp.mode == MODE_NUTS) {   // at most 2^D gradient evaluations per transition
      const unsigned long long ev = (unsigned long long)(pl.kp.n_steps > 0 ? pl.kp.n_steps : 0) << pl.kp.nuts.max_doublings;
      if (ev > 0x40000000ull) return fail("too many NUTS transitions in one launch for the exchange flags: use smaller chunks");
      adv = (unsigned int)ev + 2u;
    }
    if (need > c->xchg_bytes || c->xepoch > 0xF0000000u - adv) {   // (re)allocate, or restart the flag epoch before it wraps
      if (need > c->xchg_bytes) {
        if (c->xchg) cudaFree(c->xchg);
        CK(cudaMalloc(&c->xchg, need));
        c->xchg_bytes = need;
      }
      CK(cudaMemsetAsync(c->xchg, 0, c->xchg_bytes, st));
      c->xepoch = 0;
    }
    pl.kp.xchg = c->xchg; pl.kp.xbase = c->xepoch;   // flags of this launch: xbase+1 .. xbase+n_evals (never 0, never reused)
    pl.kp.xchg2 = c->xchg + nA;
    c->xepoch += adv;
  }
  const int NL = c->M.NL;
  int rc;
  if (pl.kp.mode == MODE_NUTS) {   // NUTS: tensor evaluator or the generic tiles for the gradients (plans with fast = 1 are not made)
    if (pl.fast == 2) {
      if (NL == 3 && pl.fast_fp == 8) rc = launch_t<MmaGE<3, 8, 512>, 2>(pl, n_chains, st);
      else if (NL == 3) rc = launch_t<MmaGE<3, 16, 512>, 2>(pl, n_chains, st);
      else if (pl.fast_fp == 8) rc = launch_t<MmaGE<4, 8, 512>, 2>(pl, n_chains, st);
      else rc = launch_t<MmaGE<4, 16, 512>, 2>(pl, n_chains, st);
    }
    else if (pl.fast) return fail("internal: NUTS has no FFMA-pipeline plan");
    else if (NL <= 2) rc = launch_t<GenericGE<2>, 2>(pl, n_chains, st);
    else if (NL <= 4) rc = launch_t<GenericGE<4>, 2>(pl, n_chains, st);
    else if (NL <= 8) rc = launch_t<GenericGE<8>, 2>(pl, n_chains, st);
    else rc = launch_t<GenericGE<12>, 2>(pl, n_chains, st);
    if (rc == 0) c->launches++;
    return rc;
  }
  // (the preconditioned dynamics live in the generic step loop; the tensor evaluator still computes the gradients)
  if (pl.fast == 2 && (pl.kp.mode == MODE_SAMPLE || pl.kp.mode == MODE_TUNE) && c->opt_steploop != 0 && !pl.kp.sdc && !pl.kp.pmask) {
    if (NL == 3 && pl.fast_fp == 8) rc = launch_t<MmaGE<3, 8, 512>, 1>(pl, n_chains, st);
    else if (NL == 3) rc = launch_t<MmaGE<3, 16, 512>, 1>(pl, n_chains, st);
    else if (pl.fast_fp == 8) rc = launch_t<MmaGE<4, 8, 512>, 1>(pl, n_chains, st);
    else rc = launch_t<MmaGE<4, 16, 512>, 1>(pl, n_chains, st);
  }
  else if (pl.fast == 2) {
    if (NL == 3 && pl.fast_fp == 8) rc = launch_t<MmaGE<3, 8, 512>>(pl, n_chains, st);
    else if (NL == 3) rc = launch_t<MmaGE<3, 16, 512>>(pl, n_chains, st);
    else if (pl.fast_fp == 8) rc = launch_t<MmaGE<4, 8, 512>>(pl, n_chains, st);
    else rc = launch_t<MmaGE<4, 16, 512>>(pl, n_chains, st);
  }
  else if (pl.fast) {
    if (NL == 3 && pl.fast_fp == 8) rc = launch_t<FastGE<3, 8, MILE_ACT_RELU>>(pl, n_chains, st);
    else if (NL == 3) rc = launch_t<FastGE<3, 12, MILE_ACT_RELU>>(pl, n_chains, st);
    else if (pl.fast_fp == 8) rc = launch_t<FastGE<4, 8, MILE_ACT_RELU>>(pl, n_chains, st);
    else rc = launch_t<FastGE<4, 12, MILE_ACT_RELU>>(pl, n_chains, st);
  }
  else if (NL <= 2) rc = launch_t<GenericGE<2>>(pl, n_chains, st);
  else if (NL <= 3) rc = launch_t<GenericGE<3>>(pl, n_chains, st);
  else if (NL <= 4) rc = launch_t<GenericGE<4>>(pl, n_chains, st);
  else if (NL <= 6) rc = launch_t<GenericGE<6>>(pl, n_chains, st);
  else if (NL <= 8) rc = launch_t<GenericGE<8>>(pl, n_chains, st);
  else rc = launch_t<GenericGE<12>>(pl, n_chains, st);
  if (rc == 0) c->launches++;
  return rc;
}

static void fill_common(mile_ctx* c, KParams& k) {
  k.X = c->X; k.y = c->y; k.N = c->N; k.Xt = c->Xt; k.yt = c->yt; k.Nt = c->Nt;
  k.theta = c->theta; k.u = c->u; k.grad = c->grad; k.lp = c->lp;
  k.t_time = c->t_time; k.t_xavg = c->t_xavg; k.t_epsmax = c->t_epsmax; k.t_eps = c->t_eps; k.t_L = c->t_L;
  k.t_wtot = c->t_wtot; k.avg_x = c->avg_x; k.avg_x2 = c->avg_x2;
  k.lppd_m = c->lppd_m; k.lppd_s = c->lppd_s;
  k.carry = c->carry; k.carry_valid = c->carry_valid;
  k.refresh_mode = c->opt_refresh; k.thin = 1;
  k.out_stride = c->d; k.prior_weight = 1.f; k.chain_base = c->opt_chain_base;
  k.sdc = c->sdc_on ? c->sdc : nullptr;
  k.pmask = c->pmask_on ? c->pmask : nullptr; k.d_eff = c->pmask_on ? c->d_eff : c->d;
}

// ---- small utility kernels ------------------------------------------------------------------
__global__ void pad_rows_kernel(const float* __restrict__ X, float* __restrict__ Xp, long N, int F, int sx) {
  const long total = N * sx;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / sx; const int f = (int)(i % sx);
    Xp[i] = f < F ? X[r * F + f] : 0.f;
  }
}
__global__ void fill_kernel(float* p, long n, float v) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void tune_L_kernel(const float* __restrict__ ax, const float* __restrict__ ax2, float* L, int d) {
  // L = sqrt(sum(E[x^2] - E[x]^2))  (warmup.py:387-390); one CTA per chain
  __shared__ float red[64];
  int phase = 0;
  const int c = blockIdx.x;
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < d; i += MILE_THREADS) {
    const float m = ax[(long)c * d + i];
    v[0] += ax2[(long)c * d + i] - m * m;
  }
  block_sum<1, MILE_THREADS>(v, red, phase);
  if (threadIdx.x == 0) L[c] = sqrtf(v[0]);
}

__global__ void precond_from_moments_kernel(const float* __restrict__ ax, const float* __restrict__ ax2, float* __restrict__ sdc,
                                            float* __restrict__ L, int d, int d_eff) {
  // sqrt_diag_cov = sqrt(E[x^2] - E[x]^2), L = sqrt(d)  (warmup.py:388-394); a variance that rounds below zero gives 0
  const int c = blockIdx.x;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float m = ax[(long)c * d + i];
    sdc[(long)c * d + i] = sqrtf(fmaxf(ax2[(long)c * d + i] - m * m, 0.f));
  }
  if (threadIdx.x == 0) L[c] = sqrtf((float)d_eff);
}

static void* scratch(mile_ctx* c, int slot, size_t bytes) {
  if ((int)c->scratch.size() <= slot) c->scratch.resize(slot + 1, {nullptr, 0});
  auto& s = c->scratch[slot];
  if (s.second < bytes) {
    if (s.first) cudaFree(s.first);
    s.first = nullptr; s.second = 0;
    if (cudaMalloc(&s.first, bytes) != cudaSuccess) return nullptr;
    s.second = bytes;
  }
  return s.first;
}

// give a large staging buffer back (the captured positions of phase 3 can be tens of GB for a wide model)
static void scratch_release(mile_ctx* c, int slot, size_t keep_below) {
  if ((int)c->scratch.size() <= slot) return;
  auto& s = c->scratch[slot];
  if (s.first && s.second > keep_below) { cudaFree(s.first); s.first = nullptr; s.second = 0; }
}

template <int NLMAX>
static int launch_train(const TrainParams& T, int n_chains, size_t smem, cudaStream_t st) {
  CK(cudaFuncSetAttribute(mile_train_kernel<NLMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  mile_train_kernel<NLMAX><<<n_chains, MILE_THREADS, smem, st>>>(T);
  CK(cudaGetLastError());
  return 0;
}
template <int NLMAX>
static int launch_metrics(const MetricsParams& T, int n, size_t smem, cudaStream_t st) {
  CK(cudaFuncSetAttribute(mile_metrics_kernel<NLMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  mile_metrics_kernel<NLMAX><<<n, MILE_THREADS, smem, st>>>(T);
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

static int wide_forward(mile_ctx* c, const float* theta, int n, const float* Xs, long N, long N8, float* actbuf, cudaStream_t st, int n_layers = -1);
static int wide_predict(mile_ctx* c, const float* theta, int n, int which, float* out, cudaStream_t st);
static int wide_lppd_fold(mile_ctx* c, const float* theta, int n, cudaStream_t st);
static int wide_alloc(mile_ctx* c, int n_chains);
static int wide_eval(mile_ctx* c, const float* theta, int n, float* gl, float prior_weight, cudaStream_t st);
static int shard_fused_plan(mile_ctx* c, Plan& pl);

const char* mile_last_error(void) { return g_err.c_str(); }
int mile_version(void) { return 100; }

int mile_create(const mile_model_desc* desc, int32_t n_chains, int32_t device, mile_ctx** out) {
  if (!desc || !out) return fail("null argument");
  if (desc->n_layers < 1 || desc->n_layers > MILE_MAX_LAYERS) return fail("n_layers out of range");
  if (n_chains < 1) return fail("n_chains must be >= 1");
  if (desc->task == MILE_TASK_REGRESSION && desc->widths[desc->n_layers - 1] < 2)
    return fail("regression needs an output width >= 2 (mean, log-sigma)");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("no such CUDA device (this library has no CPU fallback)");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 9) return fail("mile_b200 needs thread-block clusters (built for sm_100a)");
  mile_ctx* c = new mile_ctx();
  c->n_sms = prop.multiProcessorCount;
  c->desc = *desc; c->C = n_chains; c->device = device;
  build_model(c);
  if (c->d < 2) { delete c; return fail("The target distribution must have more than 1 dimension for MCLMC."); }
  const size_t Cd = (size_t)n_chains * c->d * 4, Cb = (size_t)n_chains * 4;
  auto alloc_all = [&]() -> int {
    CK(cudaMalloc(&c->theta, Cd)); CK(cudaMalloc(&c->u, Cd)); CK(cudaMalloc(&c->grad, Cd)); CK(cudaMalloc(&c->lp, Cb));
    CK(cudaMalloc(&c->avg_x, Cd)); CK(cudaMalloc(&c->avg_x2, Cd));
    CK(cudaMalloc(&c->t_time, Cb)); CK(cudaMalloc(&c->t_xavg, Cb)); CK(cudaMalloc(&c->t_epsmax, Cb));
    CK(cudaMalloc(&c->t_eps, Cb)); CK(cudaMalloc(&c->t_L, Cb)); CK(cudaMalloc(&c->t_wtot, Cb));
    CK(cudaMalloc(&c->carry, 2 * Cb));
    CK(cudaMemset(c->theta, 0, Cd)); CK(cudaMemset(c->u, 0, Cd)); CK(cudaMemset(c->grad, 0, Cd)); CK(cudaMemset(c->lp, 0, Cb));
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    return 0;
  };
  if (alloc_all()) { const std::string e = g_err; mile_destroy(c); g_err = e; return -1; }
  {  // shapes the shared-memory kernels cannot hold run on the HBM-resident layer-by-layer path
    Plan probe;
    if (make_plan(c, n_chains, 4096, true, probe) != 0) {
      if (desc->activation == MILE_ACT_GELU) { mile_destroy(c); return fail("gelu is not supported on the wide path"); }
      c->wide = 1; g_err.clear();
    }
  }
  *out = c;
  return 0;
}

void mile_destroy(mile_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  void* ptrs[] = {c->X, c->y, c->Xt, c->yt, c->theta, c->u, c->grad, c->lp, c->t_time, c->t_xavg, c->t_epsmax,
                  c->t_eps, c->t_L, c->t_wtot, c->avg_x, c->avg_x2, c->lppd_m, c->lppd_s, c->carry,
                  c->gl, c->scal, c->thb, c->ub, c->gb, c->tr_m, c->tr_v, (float*)c->tr_t, (float*)c->xchg, c->w_act, c->w_delta[0], c->w_delta[1], c->w_part, c->w_llpart,
                  c->w_ones, c->w_gl, c->wp_act, c->wp_out, c->w_wpk, c->w_wpk_lo, c->w_wpkT, c->w_wpkT_lo, c->w_arena, c->w_fin, c->sdc, c->pmask,
                  c->nuts_imm, c->nuts_mean, c->nuts_m2, c->nuts_da, c->nuts_scratch};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& s : c->scratch) if (s.first) cudaFree(s.first);
  for (int r = 0; r < 8; ++r) if (c->xr_peer[r] && c->xr_peer[r] != c->xr) cudaIpcCloseMemHandle(c->xr_peer[r]);
  if (c->xr) cudaFree(c->xr);
  if (c->nccl_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int32_t mile_n_params(const mile_ctx* c) { return c ? c->d : -1; }

int mile_set_option(mile_ctx* c, const char* key, int64_t v) {
  if (!c || !key) return fail("null argument");
  if (!strcmp(key, "cluster_size")) c->opt_cluster = (int)v;
  else if (!strcmp(key, "tile_rows")) c->opt_tile_rows = (int)v;
  else if (!strcmp(key, "refresh_mode")) c->opt_refresh = (int)v;
  else if (!strcmp(key, "resident")) c->opt_resident = (int)v;
  else if (!strcmp(key, "fast")) c->opt_fast = (int)v;
  else if (!strcmp(key, "nuts_smem")) c->opt_nuts_smem = (int)v;
  else if (!strcmp(key, "nuts_push")) c->opt_nuts_push = (int)v;
  else if (!strcmp(key, "tensor")) c->opt_tensor = (int)v;
  else if (!strcmp(key, "sync_mode")) c->opt_sync = (int)v;
  else if (!strcmp(key, "chain_base")) c->opt_chain_base = (int)v;
  else if (!strcmp(key, "kslices")) { c->opt_kslices = (int)v; c->w_rows = -1; }   // wide path: split-K slices of the dW GEMMs (0 = auto)
  else if (!strcmp(key, "p2p")) { if (v == 0) c->p2p = 0; else if (!c->xr_peer[c->rank == 0 ? 1 : 0]) return fail("p2p: the peer mapping has not been opened"); else c->p2p = 1; }
  else if (!strcmp(key, "head_fused")) c->opt_head_fused = (int)v;   // wide path: fused output-layer kernel on / off
  else if (!strcmp(key, "shard_fused")) c->opt_shard_fused = (int)v;   // 1: multi-rank step loop as ONE persistent kernel per rank (needs the peer mapping), 0: one launch per phase
  else if (!strcmp(key, "steploop")) c->opt_steploop = (int)v;   // 1: integrator-warp step loop of the tensor evaluator, 0: generic loop
  else return fail(std::string("unknown option ") + key);
  return 0;
}

int64_t mile_get_option(const mile_ctx* c, const char* key) {
  if (!c || !key) return -1;
  if (!strcmp(key, "cluster_size") || !strcmp(key, "tile_rows") || !strcmp(key, "resident") || !strcmp(key, "smem_bytes") ||
      !strcmp(key, "fast") || !strcmp(key, "sync_mode")) {
    Plan pl;
    if (make_plan(c, c->C, c->N > 0 ? c->N : 1, true, pl)) return -1;
    if (!strcmp(key, "cluster_size")) return pl.G;
    if (!strcmp(key, "tile_rows")) return pl.TR;
    if (!strcmp(key, "resident")) return pl.resident;
    if (!strcmp(key, "fast")) return pl.fast;
    if (!strcmp(key, "sync_mode")) return pl.sync_mode;
    return (int64_t)pl.smem;
  }
  if (!strcmp(key, "refresh_mode")) return c->opt_refresh;
  if (!strcmp(key, "chain_base")) return c->opt_chain_base;
  if (!strcmp(key, "wide")) return c->wide;
  if (!strcmp(key, "p2p")) return c->p2p;
  if (!strcmp(key, "d_eff")) return c->pmask_on ? c->d_eff : c->d;
  if (!strcmp(key, "shard_fused")) {   // is the multi-rank step loop the fused persistent kernel?
    Plan pl;
    return shard_fused_plan(const_cast<mile_ctx*>(c), pl) == 0 ? 1 : 0;
  }
  if (!strcmp(key, "tensor")) return c->opt_tensor;
  if (!strcmp(key, "row_stride")) return c->M.sA[0];
  return -1;
}

static int set_split(mile_ctx* c, const float* X_dev, const void* y_dev, long N, cudaStream_t st, float** Xd,
                     void** yd, long* Nd) {
  if (N < 0) return fail("n_rows < 0");
  CK(cudaSetDevice(c->device));
  const int sx = c->M.sA[0];
  if (!(*Xd && *yd && *Nd == N && N > 0)) {      // a split of the same size reuses its buffers (cudaFree / cudaMalloc of tens
    if (*Xd) { CK(cudaFree(*Xd)); *Xd = nullptr; }   // of MB cost 20-600 ms per call: tools/time_e2e_phases.py)
    if (*yd) { CK(cudaFree(*yd)); *yd = nullptr; }
    *Nd = N;
    if (N == 0) return 0;
    // 16 zero floats of slack: the k-steps of the tensor evaluator may read up to 12 floats past the last row's stride
    CK(cudaMalloc(Xd, ((size_t)N * sx + 16) * 4));
    CK(cudaMalloc(yd, (size_t)N * 4));
  }
  CK(cudaMemsetAsync(*Xd + (size_t)N * sx, 0, 16 * 4, st));
  pad_rows_kernel<<<296, 256, 0, st>>>(X_dev, *Xd, N, c->M.F, sx);
  CK(cudaGetLastError());
  c->launches++;
  CK(cudaMemcpyAsync(*yd, y_dev, (size_t)N * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int lppd_alloc(mile_ctx* c, cudaStream_t st) {
  if (c->lppd_m) { CK(cudaFree(c->lppd_m)); c->lppd_m = nullptr; }
  if (c->lppd_s) { CK(cudaFree(c->lppd_s)); c->lppd_s = nullptr; }
  if (c->Nt == 0) return 0;
  const long n = (long)c->C * c->Nt;
  CK(cudaMalloc(&c->lppd_m, n * 4)); CK(cudaMalloc(&c->lppd_s, n * 4));
  fill_kernel<<<148, 256, 0, st>>>(c->lppd_m, n, -INFINITY);
  fill_kernel<<<148, 256, 0, st>>>(c->lppd_s, n, 0.f);
  CK(cudaGetLastError());
  c->launches += 2; c->lppd_count = 0;
  return 0;
}

int mile_set_data(mile_ctx* c, const float* X_dev, const void* y_dev, int64_t N, void* stream) {
  if (!c) return fail("null ctx");
  return set_split(c, X_dev, y_dev, (long)N, (cudaStream_t)stream, &c->X, &c->y, &c->N);
}
int mile_set_test(mile_ctx* c, const float* X_dev, const void* y_dev, int64_t N, void* stream) {
  if (!c) return fail("null ctx");
  if (set_split(c, X_dev, y_dev, (long)N, (cudaStream_t)stream, &c->Xt, &c->yt, &c->Nt)) return -1;
  // the [C, Nt] logsumexp state is allocated on first use (lppd_reset / accumulate / lppd=1 sampling): mile_predict
  // contexts with thousands of (chain, sample) rows never need it
  if (c->lppd_m) { CK(cudaFree(c->lppd_m)); c->lppd_m = nullptr; }
  if (c->lppd_s) { CK(cudaFree(c->lppd_s)); c->lppd_s = nullptr; }
  c->lppd_count = 0;
  return 0;
}

static int set_split_host(mile_ctx* c, const float* X, const void* y, long N, bool test) {
  CK(cudaSetDevice(c->device));
  float* Xd = (float*)scratch(c, 0, (size_t)(N > 0 ? N : 1) * c->M.F * 4);
  void* yd = scratch(c, 1, (size_t)(N > 0 ? N : 1) * 4);
  if (!Xd || !yd) return fail("cudaMalloc failed (staging)");
  CK(cudaMemcpyAsync(Xd, X, (size_t)N * c->M.F * 4, cudaMemcpyHostToDevice, c->own_stream));
  CK(cudaMemcpyAsync(yd, y, (size_t)N * 4, cudaMemcpyHostToDevice, c->own_stream));
  int rc = test ? mile_set_test(c, Xd, yd, N, c->own_stream) : mile_set_data(c, Xd, yd, N, c->own_stream);
  CK(cudaStreamSynchronize(c->own_stream));
  return rc;
}
int mile_set_data_host(mile_ctx* c, const float* X, const void* y, int64_t N) {
  if (!c) return fail("null ctx");
  return set_split_host(c, X, y, (long)N, false);
}
int mile_set_test_host(mile_ctx* c, const float* X, const void* y, int64_t N) {
  if (!c) return fail("null ctx");
  return set_split_host(c, X, y, (long)N, true);
}

int mile_logpost_value_and_grad(mile_ctx* c, const float* theta_dev, int32_t n, float* lp_dev, float* grad_dev,
                                void* stream) {
  if (!c) return fail("null ctx");
  if (!c->X) return fail("mile_set_data has not been called");
  if (n < 1) return fail("n must be >= 1");
  if (c->wide) {
    cudaStream_t st = (cudaStream_t)stream;
    if (wide_eval(c, theta_dev, n, nullptr, 1.f, st)) return -1;
    const size_t rowb = (size_t)(c->d + 1) * 4;
    CK(cudaMemcpy2DAsync(grad_dev, (size_t)c->d * 4, c->w_gl, rowb, (size_t)c->d * 4, n, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpy2DAsync(lp_dev, 4, c->w_gl + c->d, rowb, 4, n, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  Plan pl;
  if (make_plan(c, n, c->N, true, pl)) return -1;
  fill_common(c, pl.kp);
  pl.kp.C = n; pl.kp.mode = MODE_EVAL; pl.kp.theta_in = theta_dev; pl.kp.lp_out = lp_dev; pl.kp.grad_out = grad_dev;
  pl.kp.n_eval = n;
  return launch(c, pl, n, (cudaStream_t)stream);
}

int mile_logpost_value_and_grad_host(mile_ctx* c, const float* theta, int32_t n, float* lp, float* grad) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  const size_t nd = (size_t)n * c->d * 4;
  float* th = (float*)scratch(c, 2, nd); float* g = (float*)scratch(c, 3, nd); float* l = (float*)scratch(c, 4, (size_t)n * 4);
  if (!th || !g || !l) return fail("cudaMalloc failed (staging)");
  CK(cudaMemcpyAsync(th, theta, nd, cudaMemcpyHostToDevice, c->own_stream));
  if (mile_logpost_value_and_grad(c, th, n, l, g, c->own_stream)) return -1;
  CK(cudaMemcpyAsync(lp, l, (size_t)n * 4, cudaMemcpyDeviceToHost, c->own_stream));
  CK(cudaMemcpyAsync(grad, g, nd, cudaMemcpyDeviceToHost, c->own_stream));
  CK(cudaStreamSynchronize(c->own_stream));
  return 0;
}

int mile_shard_init(mile_ctx* c, const void* unique_id128, int32_t rank, int32_t world);
int mile_shard_mclmc_init(mile_ctx* c, const float* theta0_dev, const float* z0_dev, uint64_t seed, void* stream);
int mile_shard_mclmc_sample(mile_ctx* c, int32_t n_steps, int64_t step_base, int32_t n_thinning, int64_t sample_base,
                            const float* step_size_dev, const float* L_dev, const float* z_dev, uint64_t seed,
                            float* samples_dev, int64_t n_slots, float* info_dev, void* stream);
int mile_shard_mclmc_tune(mile_ctx* c, int32_t n_steps, int64_t step_base, const mile_tune_cfg* cfg, const float* z_dev,
                          uint64_t seed, float* tune_info_dev, void* stream);

int mile_mclmc_init(mile_ctx* c, const float* theta0_dev, const float* z0_dev, uint64_t seed, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->X) return fail("mile_set_data has not been called");
  if (c->wide) {   // HBM-resident layer-by-layer path: same launch structure as the data-sharded variant, world = 1
    if (!c->gl && mile_shard_init(c, nullptr, 0, 1)) return -1;
    return mile_shard_mclmc_init(c, theta0_dev, z0_dev, seed, stream);
  }
  Plan pl;
  if (make_plan(c, c->C, c->N, true, pl)) return -1;
  fill_common(c, pl.kp);
  pl.kp.mode = MODE_INIT; pl.kp.theta_in = theta0_dev; pl.kp.z = z0_dev; pl.kp.seed = seed;
  c->carry_valid = 0;
  return launch(c, pl, c->C, (cudaStream_t)stream);
}

int mile_mclmc_init_host(mile_ctx* c, const float* theta0, const float* z0, uint64_t seed) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  const size_t Cd = (size_t)c->C * c->d * 4;
  float* th = (float*)scratch(c, 2, Cd); float* z = z0 ? (float*)scratch(c, 3, Cd) : nullptr;
  if (!th || (z0 && !z)) return fail("cudaMalloc failed (staging)");
  CK(cudaMemcpyAsync(th, theta0, Cd, cudaMemcpyHostToDevice, c->own_stream));
  if (z0) CK(cudaMemcpyAsync(z, z0, Cd, cudaMemcpyHostToDevice, c->own_stream));
  if (mile_mclmc_init(c, th, z, seed, c->own_stream)) return -1;
  CK(cudaStreamSynchronize(c->own_stream));
  return 0;
}

int mile_set_state_host(mile_ctx* c, const float* theta, const float* u, const float* lp, const float* grad) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  const size_t Cd = (size_t)c->C * c->d * 4;
  c->carry_valid = 0;
  if (theta) CK(cudaMemcpy(c->theta, theta, Cd, cudaMemcpyHostToDevice));
  if (u) CK(cudaMemcpy(c->u, u, Cd, cudaMemcpyHostToDevice));
  if (grad) CK(cudaMemcpy(c->grad, grad, Cd, cudaMemcpyHostToDevice));
  if (lp) CK(cudaMemcpy(c->lp, lp, (size_t)c->C * 4, cudaMemcpyHostToDevice));
  return 0;
}
int mile_get_state_host(mile_ctx* c, float* theta, float* u, float* lp, float* grad) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  const size_t Cd = (size_t)c->C * c->d * 4;
  if (theta) CK(cudaMemcpy(theta, c->theta, Cd, cudaMemcpyDeviceToHost));
  if (u) CK(cudaMemcpy(u, c->u, Cd, cudaMemcpyDeviceToHost));
  if (grad) CK(cudaMemcpy(grad, c->grad, Cd, cudaMemcpyDeviceToHost));
  if (lp) CK(cudaMemcpy(lp, c->lp, (size_t)c->C * 4, cudaMemcpyDeviceToHost));
  return 0;
}
int mile_get_state(mile_ctx* c, float* theta, float* u, float* lp, float* grad, void* stream) {
  if (!c) return fail("null ctx");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Cd = (size_t)c->C * c->d * 4;
  if (theta) CK(cudaMemcpyAsync(theta, c->theta, Cd, cudaMemcpyDeviceToDevice, st));
  if (u) CK(cudaMemcpyAsync(u, c->u, Cd, cudaMemcpyDeviceToDevice, st));
  if (grad) CK(cudaMemcpyAsync(grad, c->grad, Cd, cudaMemcpyDeviceToDevice, st));
  if (lp) CK(cudaMemcpyAsync(lp, c->lp, (size_t)c->C * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int mile_mclmc_sample(mile_ctx* c, int32_t n_steps, int64_t step_base, int32_t n_thinning, int64_t sample_base,
                      const float* step_size_dev, const float* L_dev, const float* z_dev, uint64_t seed,
                      float* samples_dev, int64_t n_slots, float* info_dev, int32_t lppd, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->X) return fail("mile_set_data has not been called");
  if (n_steps < 0 || n_thinning < 1) return fail("n_steps must be >= 0 and n_thinning >= 1");
  if (lppd && !c->Xt) return fail("lppd requested but mile_set_test has not been called");
  if (!step_size_dev || !L_dev) return fail("step_size / L are required");
  if (n_steps == 0) return 0;
  if (lppd && !c->wide && !c->lppd_m && lppd_alloc(c, (cudaStream_t)stream)) return -1;
  if (c->wide) {
    if (!c->gl && mile_shard_init(c, nullptr, 0, 1)) return -1;
    if (lppd && !c->lppd_m && lppd_alloc(c, (cudaStream_t)stream)) return -1;
    c->shard_lppd = lppd ? 1 : 0;      // shard_run folds every kept position through the wide forward pass
    const int rc = mile_shard_mclmc_sample(c, n_steps, step_base, n_thinning, sample_base, step_size_dev, L_dev, z_dev, seed,
                                           samples_dev, n_slots, info_dev, stream);
    c->shard_lppd = 0;
    if (rc) return rc;
    if (lppd) {
      const long first = (step_base + n_thinning - 1) / n_thinning, last = (step_base + n_steps - 1) / n_thinning;
      c->lppd_count += (last >= first) ? (last - first + 1) : 0;
    }
    return 0;
  }
  Plan pl;
  if (make_plan(c, c->C, c->N, true, pl)) return -1;
  fill_common(c, pl.kp);
  KParams& k = pl.kp;
  k.mode = MODE_SAMPLE; k.n_steps = n_steps; k.step_base = step_base; k.thin = n_thinning; k.sample_base = sample_base;
  k.n_slots = n_slots; k.eps = step_size_dev; k.L = L_dev; k.z = z_dev; k.seed = seed; k.samples = samples_dev;
  k.info = info_dev; k.do_lppd = lppd;
  if (launch(c, pl, c->C, (cudaStream_t)stream)) return -1;
  c->carry_valid = 1;
  if (lppd) {
    // kept positions in [step_base, step_base+n_steps)
    const long first = (step_base + n_thinning - 1) / n_thinning, last = (step_base + n_steps - 1) / n_thinning;
    c->lppd_count += (last >= first) ? (last - first + 1) : 0;
  }
  return 0;
}

int mile_mclmc_sample_host(mile_ctx* c, int32_t n_steps, int64_t step_base, int32_t n_thinning,
                           const float* step_size, const float* L, const float* z, uint64_t seed, float* samples,
                           int64_t n_slots, float* info, int32_t lppd) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->own_stream;
  const size_t Cb = (size_t)c->C * 4, Cd = (size_t)c->C * c->d * 4;
  const int nslot = c->opt_refresh ? 2 : 1;
  float* e = (float*)scratch(c, 5, Cb); float* l = (float*)scratch(c, 6, Cb);
  float* zd = z ? (float*)scratch(c, 7, (size_t)n_steps * nslot * Cd) : nullptr;
  float* sd = samples ? (float*)scratch(c, 8, (size_t)n_slots * Cd) : nullptr;
  float* id = info ? (float*)scratch(c, 9, (size_t)n_steps * c->C * 3 * 4) : nullptr;
  if (!e || !l || (z && !zd) || (samples && !sd) || (info && !id)) return fail("cudaMalloc failed (staging)");
  CK(cudaMemcpyAsync(e, step_size, Cb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(l, L, Cb, cudaMemcpyHostToDevice, st));
  if (z) CK(cudaMemcpyAsync(zd, z, (size_t)n_steps * nslot * Cd, cudaMemcpyHostToDevice, st));
  const int64_t sample_base = (step_base + n_thinning - 1) / n_thinning;
  if (mile_mclmc_sample(c, n_steps, step_base, n_thinning, sample_base, e, l, zd, seed, sd, n_slots, id, lppd, st)) return -1;
  if (samples) CK(cudaMemcpyAsync(samples, sd, (size_t)n_slots * Cd, cudaMemcpyDeviceToHost, st));
  if (info) CK(cudaMemcpyAsync(info, id, (size_t)n_steps * c->C * 3 * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int mile_tune_reset(mile_ctx* c, float step_size_init, void* stream) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long C = c->C, Cd = (long)c->C * c->d;
  fill_kernel<<<8, 256, 0, st>>>(c->t_time, C, 0.f);
  fill_kernel<<<8, 256, 0, st>>>(c->t_xavg, C, 0.f);
  fill_kernel<<<8, 256, 0, st>>>(c->t_epsmax, C, INFINITY);
  fill_kernel<<<8, 256, 0, st>>>(c->t_eps, C, step_size_init);
  fill_kernel<<<8, 256, 0, st>>>(c->t_L, C, fmaxf(sqrtf((float)(c->pmask_on ? c->d_eff : c->d)), 15.0f));
  fill_kernel<<<8, 256, 0, st>>>(c->t_wtot, C, 0.f);
  fill_kernel<<<148, 256, 0, st>>>(c->avg_x, Cd, 0.f);
  fill_kernel<<<148, 256, 0, st>>>(c->avg_x2, Cd, 0.f);
  CK(cudaGetLastError());
  c->launches += 8;
  return 0;
}

int mile_mclmc_tune(mile_ctx* c, int32_t n_steps, int64_t step_base, const mile_tune_cfg* cfg, const float* z_dev,
                    uint64_t seed, float* tune_info_dev, void* stream) {
  if (!c || !cfg) return fail("null argument");
  if (!c->X) return fail("mile_set_data has not been called");
  if (n_steps <= 0) return 0;
  if (c->wide) {
    if (!c->gl && mile_shard_init(c, nullptr, 0, 1)) return -1;
    return mile_shard_mclmc_tune(c, n_steps, step_base, cfg, z_dev, seed, tune_info_dev, stream);
  }
  Plan pl;
  if (make_plan(c, c->C, c->N, true, pl)) return -1;
  fill_common(c, pl.kp);
  KParams& k = pl.kp;
  k.mode = MODE_TUNE; k.n_steps = n_steps; k.step_base = step_base; k.z = z_dev; k.seed = seed; k.tune_info = tune_info_dev;
  k.tune1 = cfg->tune1_steps; k.tune2 = cfg->tune2_steps; k.ev_start = cfg->desired_energy_var_start;
  k.ev_end = cfg->desired_energy_var_end; k.trust = cfg->trust_in_estimate; k.neff = cfg->num_effective_samples;
  if (launch(c, pl, c->C, (cudaStream_t)stream)) return -1;
  c->carry_valid = 1;
  return 0;
}

int mile_mclmc_tune_host(mile_ctx* c, int32_t n_steps, int64_t step_base, const mile_tune_cfg* cfg, const float* z,
                         uint64_t seed, float* tune_info) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->own_stream;
  const size_t Cd = (size_t)c->C * c->d * 4;
  const int nslot = c->opt_refresh ? 2 : 1;
  float* zd = z ? (float*)scratch(c, 7, (size_t)n_steps * nslot * Cd) : nullptr;
  float* id = tune_info ? (float*)scratch(c, 9, (size_t)n_steps * c->C * 4 * 4) : nullptr;
  if ((z && !zd) || (tune_info && !id)) return fail("cudaMalloc failed (staging)");
  if (z) CK(cudaMemcpyAsync(zd, z, (size_t)n_steps * nslot * Cd, cudaMemcpyHostToDevice, st));
  if (mile_mclmc_tune(c, n_steps, step_base, cfg, zd, seed, id, st)) return -1;
  if (tune_info) CK(cudaMemcpyAsync(tune_info, id, (size_t)n_steps * c->C * 4 * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int mile_tune_finish_phase2(mile_ctx* c, void* stream) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  tune_L_kernel<<<c->C, MILE_THREADS, 0, (cudaStream_t)stream>>>(c->avg_x, c->avg_x2, c->t_L, c->d);
  CK(cudaGetLastError());
  c->launches++;
  return 0;
}

// ---- diagonal preconditioning -------------------------------------------------------------------------------------
static int sdc_alloc(mile_ctx* c) {
  if (!c->sdc) CK(cudaMalloc(&c->sdc, (size_t)c->C * c->d * 4));
  return 0;
}
int mile_precondition_from_moments(mile_ctx* c, void* stream) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  if (sdc_alloc(c)) return -1;
  precond_from_moments_kernel<<<c->C, 256, 0, (cudaStream_t)stream>>>(c->avg_x, c->avg_x2, c->sdc, c->t_L, c->d, c->pmask_on ? c->d_eff : c->d);
  CK(cudaGetLastError());
  c->launches++;
  c->sdc_on = 1; c->carry_valid = 0;     // the carried (sum g^2, u.g) refer to the unscaled gradient
  return 0;
}
int mile_set_sqrt_diag_cov_host(mile_ctx* c, const float* sdc) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  c->carry_valid = 0;
  if (!sdc) { c->sdc_on = 0; return 0; }
  if (sdc_alloc(c)) return -1;
  CK(cudaMemcpy(c->sdc, sdc, (size_t)c->C * c->d * 4, cudaMemcpyHostToDevice));
  c->sdc_on = 1;
  return 0;
}
int mile_get_sqrt_diag_cov_host(mile_ctx* c, float* out) {
  if (!c || !out) return fail("null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  const size_t n = (size_t)c->C * c->d;
  if (c->sdc_on) CK(cudaMemcpy(out, c->sdc, n * 4, cudaMemcpyDeviceToHost));
  else for (size_t i = 0; i < n; ++i) out[i] = 1.f;
  return 0;
}

// ---- partition sampling ---------------------------------------------------------------------------------------------
int mile_set_frozen_mask_host(mile_ctx* c, const uint8_t* frozen) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  c->carry_valid = 0;
  if (!frozen) { c->pmask_on = 0; return 0; }
  if (c->wide) return fail("partition sampling is served by the shared-memory kernels only (this model runs on the wide path)");
  std::vector<float> m((size_t)c->d);
  int n = 0;
  for (int i = 0; i < c->d; ++i) { m[i] = frozen[i] ? 0.f : 1.f; n += frozen[i] ? 0 : 1; }
  if (n < 2) return fail("partition sampling needs at least two sampled parameters");
  if (!c->pmask) CK(cudaMalloc(&c->pmask, (size_t)c->d * 4));
  CK(cudaMemcpy(c->pmask, m.data(), (size_t)c->d * 4, cudaMemcpyHostToDevice));
  c->pmask_on = 1; c->d_eff = n;
  return 0;
}

int mile_get_tuning_host(mile_ctx* c, float* step_size, float* L, float* step_size_max, float* mean_x, float* mean_x2) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  const size_t Cb = (size_t)c->C * 4, Cd = (size_t)c->C * c->d * 4;
  if (step_size) CK(cudaMemcpy(step_size, c->t_eps, Cb, cudaMemcpyDeviceToHost));
  if (L) CK(cudaMemcpy(L, c->t_L, Cb, cudaMemcpyDeviceToHost));
  if (step_size_max) CK(cudaMemcpy(step_size_max, c->t_epsmax, Cb, cudaMemcpyDeviceToHost));
  if (mean_x) CK(cudaMemcpy(mean_x, c->avg_x, Cd, cudaMemcpyDeviceToHost));
  if (mean_x2) CK(cudaMemcpy(mean_x2, c->avg_x2, Cd, cudaMemcpyDeviceToHost));
  return 0;
}
int mile_set_tuning_host(mile_ctx* c, const float* step_size, const float* L) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  const size_t Cb = (size_t)c->C * 4;
  if (step_size) CK(cudaMemcpy(c->t_eps, step_size, Cb, cudaMemcpyHostToDevice));
  if (L) CK(cudaMemcpy(c->t_L, L, Cb, cudaMemcpyHostToDevice));
  return 0;
}
int mile_tuning_ptrs(mile_ctx* c, float** step_size_dev, float** L_dev) {
  if (!c) return fail("null ctx");
  if (step_size_dev) *step_size_dev = c->t_eps;
  if (L_dev) *L_dev = c->t_L;
  return 0;
}

int mile_lppd_reset(mile_ctx* c, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->Xt) return fail("mile_set_test has not been called");
  return lppd_alloc(c, (cudaStream_t)stream);
}

int mile_lppd_accumulate(mile_ctx* c, const float* theta_dev, int32_t n, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->Xt) return fail("mile_set_test has not been called");
  if (n < 1 || n > c->C) return fail("n must be in [1, n_chains]");
  if (!c->lppd_m && lppd_alloc(c, (cudaStream_t)stream)) return -1;
  if (c->wide) {
    if (wide_lppd_fold(c, theta_dev, n, (cudaStream_t)stream)) return -1;
    c->lppd_count += 1;
    return 0;
  }
  Plan pl;
  if (make_plan(c, c->C, c->Nt, false, pl)) return -1;
  fill_common(c, pl.kp);
  pl.kp.mode = MODE_LPPD; pl.kp.theta_in = theta_dev;
  if (launch(c, pl, n, (cudaStream_t)stream)) return -1;
  c->lppd_count += 1;
  return 0;
}

int mile_lppd_state_host(mile_ctx* c, float* m, float* s, int64_t* count) {
  if (!c) return fail("null ctx");
  if (!c->Xt) return fail("mile_set_test has not been called");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  const size_t n = (size_t)c->C * c->Nt * 4;
  if (!c->lppd_m && lppd_alloc(c, nullptr)) return -1;
  CK(cudaDeviceSynchronize());
  if (m) CK(cudaMemcpy(m, c->lppd_m, n, cudaMemcpyDeviceToHost));
  if (s) CK(cudaMemcpy(s, c->lppd_s, n, cudaMemcpyDeviceToHost));
  if (count) *count = c->lppd_count;
  return 0;
}

int mile_predict(mile_ctx* c, const float* theta_dev, int32_t n, int32_t which, float* out_dev, void* stream) {
  if (!c) return fail("null ctx");
  if (which ? !c->Xt : !c->X) return fail("requested split has not been set");
  if (n < 1) return fail("n must be >= 1");
  if (c->wide) return wide_predict(c, theta_dev, n, which, out_dev, (cudaStream_t)stream);
  Plan pl;
  if (make_plan(c, n, which ? c->Nt : c->N, false, pl)) return -1;
  fill_common(c, pl.kp);
  pl.kp.C = n; pl.kp.mode = MODE_PREDICT; pl.kp.theta_in = theta_dev; pl.kp.pred_out = out_dev; pl.kp.which = which;
  return launch(c, pl, n, (cudaStream_t)stream);
}


// ---- deep-ensemble warm-start training (mile_train.cuh) ---------------------------------------------------------
int mile_train_init(mile_ctx* c, const float* theta0_dev, void* stream) {
  if (!c || !theta0_dev) return fail("null argument");
  if (c->wide) return fail("warm-start training is implemented for the shared-memory models only");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Cd = (size_t)c->C * c->d * 4;
  if (!c->tr_m) { CK(cudaMalloc(&c->tr_m, Cd)); CK(cudaMalloc(&c->tr_v, Cd)); CK(cudaMalloc(&c->tr_t, (size_t)c->C * 4)); }
  CK(cudaMemcpyAsync(c->theta, theta0_dev, Cd, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(c->tr_m, 0, Cd, st)); CK(cudaMemsetAsync(c->tr_v, 0, Cd, st)); CK(cudaMemsetAsync(c->tr_t, 0, (size_t)c->C * 4, st));
  c->carry_valid = 0;
  return 0;
}

int mile_train_epoch(mile_ctx* c, const int32_t* batch_idx_dev, int32_t n_batches, int32_t batch_size, const mile_opt_cfg* opt,
                     const uint8_t* stopped_dev, float* metrics_dev, void* stream) {
  if (!c || !batch_idx_dev || !opt) return fail("null argument");
  if (!c->X) return fail("mile_set_data has not been called");
  if (!c->tr_m) return fail("mile_train_init has not been called");
  if (n_batches < 0 || batch_size < 1 || batch_size > 256) return fail("batch_size must be in [1, 256] and n_batches >= 0");
  if (opt->kind < 0 || opt->kind > 2) return fail("unknown optimizer kind");
  if (n_batches == 0) return 0;
  CK(cudaSetDevice(c->device));
  const int keep_fast = c->opt_fast, keep_tr = c->opt_tile_rows;
  c->opt_fast = 0; c->opt_tile_rows = 0;       // the generic tile evaluator serves any FCN shape
  Plan pl;
  const int rc = make_plan(c, c->C, batch_size, false, pl, 1);
  c->opt_fast = keep_fast; c->opt_tile_rows = keep_tr;
  if (rc) return -1;
  fill_common(c, pl.kp);
  TrainParams T;
  memset(&T, 0, sizeof(T));
  T.K = pl.kp; T.K.G = 1; T.K.resident = 0; T.K.C = c->C;
  T.batch_idx = batch_idx_dev; T.n_batches = n_batches; T.B = batch_size;
  T.opt_kind = opt->kind; T.lr = opt->learning_rate; T.b1 = opt->b1; T.b2 = opt->b2; T.eps = opt->eps; T.wd = opt->weight_decay;
  T.stopped = stopped_dev; T.metrics = metrics_dev; T.m = c->tr_m; T.v = c->tr_v; T.t = c->tr_t;
  T.off_y = (int)(pl.smem / 4);
  const size_t smem = pl.smem + 256 * 4;
  if (smem > kSmemLimit) return fail("model too large for the training kernel (shared memory)");
  const int NL = c->M.NL;
  cudaStream_t st = (cudaStream_t)stream;
  int r2;
  if (NL <= 2) r2 = launch_train<2>(T, c->C, smem, st);
  else if (NL <= 3) r2 = launch_train<3>(T, c->C, smem, st);
  else if (NL <= 4) r2 = launch_train<4>(T, c->C, smem, st);
  else if (NL <= 6) r2 = launch_train<6>(T, c->C, smem, st);
  else if (NL <= 8) r2 = launch_train<8>(T, c->C, smem, st);
  else r2 = launch_train<12>(T, c->C, smem, st);
  if (r2 == 0) c->launches++;
  c->carry_valid = 0;
  return r2;
}

int mile_eval_metrics(mile_ctx* c, const float* theta_dev, int32_t n, int32_t which, float* out_dev, void* stream) {
  if (!c || !out_dev) return fail("null argument");
  if (which ? !c->Xt : !c->X) return fail("requested split has not been set");
  if (c->wide) return fail("mile_eval_metrics is implemented for the shared-memory models only");
  if (!theta_dev) { theta_dev = c->theta; n = c->C; }
  if (n < 1) return fail("n must be >= 1");
  CK(cudaSetDevice(c->device));
  const int keep_fast = c->opt_fast;
  c->opt_fast = 0;
  Plan pl;
  const int rc = make_plan(c, n, which ? c->Nt : c->N, false, pl, 1);
  c->opt_fast = keep_fast;
  if (rc) return -1;
  fill_common(c, pl.kp);
  MetricsParams T;
  memset(&T, 0, sizeof(T));
  T.K = pl.kp; T.K.C = n; T.K.theta_in = theta_dev; T.K.which = which; T.out = out_dev;
  const int NL = c->M.NL;
  cudaStream_t st = (cudaStream_t)stream;
  int r2;
  if (NL <= 2) r2 = launch_metrics<2>(T, n, pl.smem, st);
  else if (NL <= 3) r2 = launch_metrics<3>(T, n, pl.smem, st);
  else if (NL <= 4) r2 = launch_metrics<4>(T, n, pl.smem, st);
  else if (NL <= 6) r2 = launch_metrics<6>(T, n, pl.smem, st);
  else if (NL <= 8) r2 = launch_metrics<8>(T, n, pl.smem, st);
  else r2 = launch_metrics<12>(T, n, pl.smem, st);
  if (r2 == 0) c->launches++;
  return r2;
}

int mile_train_get_state(mile_ctx* c, float* theta_dev, float* m_dev, float* v_dev, int32_t* t_dev, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->tr_m) return fail("mile_train_init has not been called");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Cd = (size_t)c->C * c->d * 4;
  if (theta_dev) CK(cudaMemcpyAsync(theta_dev, c->theta, Cd, cudaMemcpyDeviceToDevice, st));
  if (m_dev) CK(cudaMemcpyAsync(m_dev, c->tr_m, Cd, cudaMemcpyDeviceToDevice, st));
  if (v_dev) CK(cudaMemcpyAsync(v_dev, c->tr_v, Cd, cudaMemcpyDeviceToDevice, st));
  if (t_dev) CK(cudaMemcpyAsync(t_dev, c->tr_t, (size_t)c->C * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

#ifdef MILE_PROFILE
int mile_debug_read_profile(unsigned long long* out32, int reset) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out32, g_prof, sizeof(unsigned long long) * 32));
  if (reset) { unsigned long long z[32] = {0}; CK(cudaMemcpyToSymbol(g_prof, z, sizeof(z))); }
  return 0;
}
#endif


#define NCK(call)                                                                                         \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess) return fail(std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
  } while (0)

int mile_nccl_unique_id(void* out128) {
  if (!out128) return fail("null argument");
  if (nccl_load()) return -1;
  ncclUniqueId id;
  NCK(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out128, &id, 128);
  return 0;
}

int mile_shard_init(mile_ctx* c, const void* unique_id128, int32_t rank, int32_t world) {
  if (!c) return fail("null ctx");
  if (world < 1 || rank < 0 || rank >= world) return fail("bad rank / world");
  CK(cudaSetDevice(c->device));
  c->world = world; c->rank = rank;
  if (world > 1) {
    if (!unique_id128) return fail("unique id required for world > 1");
    if (nccl_load()) return -1;
    ncclUniqueId id;
    memcpy(&id, unique_id128, 128);
    ncclComm_t comm;
    NCK(g_nccl.CommInitRank(&comm, world, id, rank));
    c->nccl_comm = comm;
  }
  const size_t Cd = (size_t)c->C * c->d * 4;
  if (!c->gl) {
    CK(cudaMalloc(&c->gl, (size_t)c->C * (c->d + 1) * 4)); CK(cudaMalloc(&c->scal, (size_t)c->C * 16));
    CK(cudaMalloc(&c->thb, Cd)); CK(cudaMalloc(&c->ub, Cd)); CK(cudaMalloc(&c->gb, Cd));
    CK(cudaMemset(c->scal, 0, (size_t)c->C * 16));
  }
  return 0;
}


// ---- wide path orchestration ----------------------------------------------------------------------------
static int wide_alloc(mile_ctx* c, int n_chains) {
  if (c->w_act && c->w_rows == c->N && c->w_chains >= n_chains) return 0;
  float** ptrs[] = {&c->w_act, &c->w_delta[0], &c->w_delta[1], &c->w_part, &c->w_llpart, &c->w_ones, &c->w_gl,
                    &c->w_wpk, &c->w_wpk_lo, &c->w_wpkT, &c->w_wpkT_lo};
  for (float** p : ptrs) if (*p) { cudaFree(*p); *p = nullptr; }
  c->tmaps.clear();
  const DevModel& M = c->M;
  const long N = c->N, N8 = (N + 7) / 8 * 8;   // rows padded to the 8-row core matrices of the TMA views; pad rows stay 0
  long act_per_row = 0; int maxw = 0; long maxio = 0, wsum = 0;
  for (int l = 1; l <= M.NL; ++l) { act_per_row += M.dims[l]; if (M.dims[l] > maxw) maxw = M.dims[l]; }
  for (int l = 0; l < M.NL; ++l) {
    long io = (long)M.dims[l] * M.dims[l + 1]; if (io > maxio) maxio = io;
    c->w_woff[l] = wsum; wsum += (long)((M.dims[l] + 7) / 8 * 8) * ((M.dims[l + 1] + 7) / 8 * 8);
  }
  c->w_wstride = wsum; c->w_n8 = N8;
  {
    // split-K factor of the dW GEMMs (K = rows): the smallest one whose tile count fills whole waves of the persistent
    // tcgen05 kernel (tiles = m-blocks x n-blocks x chains x slices on n_sms CTAs), at least 256 rows per slice
    const long mt = (maxw + 127) / 128, nt = (maxw + 255) / 256;
    int best = 1; double best_eff = 0.0;
    for (int ks = 1; ks <= 64 && N / ks >= 256; ++ks) {
      const long tiles = mt * nt * n_chains * ks, waves = (tiles + c->n_sms - 1) / c->n_sms;
      const double eff = (double)tiles / (double)(waves * c->n_sms);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = ks; }
      if (eff >= 0.95) { best = ks; break; }   // (measured: one full wave of 9 slices beats four waves of 37: fewer epilogues, 4x smaller slice reduction)
    }
    c->w_kslices = c->opt_kslices > 0 ? c->opt_kslices : best;
  }
  c->w_nblk = (int)((N + 255) / 256);
  // (+64: a context that only predicts has no training rows yet; the cache check above needs non-null buffers)
  const size_t actb = (size_t)n_chains * N8 * act_per_row * 4 + 64, delb = (size_t)n_chains * N8 * maxw * 4 + 64;
  CK(cudaMalloc(&c->w_act, actb));
  CK(cudaMalloc(&c->w_delta[0], delb)); CK(cudaMalloc(&c->w_delta[1], delb));
  CK(cudaMalloc(&c->w_wpk, (size_t)n_chains * wsum * 4)); CK(cudaMalloc(&c->w_wpk_lo, (size_t)n_chains * wsum * 4));
  CK(cudaMalloc(&c->w_wpkT, (size_t)n_chains * wsum * 4)); CK(cudaMalloc(&c->w_wpkT_lo, (size_t)n_chains * wsum * 4));
  CK(cudaMemset(c->w_wpkT, 0, (size_t)n_chains * wsum * 4)); CK(cudaMemset(c->w_wpkT_lo, 0, (size_t)n_chains * wsum * 4));
  CK(cudaMemset(c->w_act, 0, actb));
  CK(cudaMemset(c->w_delta[0], 0, delb)); CK(cudaMemset(c->w_delta[1], 0, delb));
  CK(cudaMemset(c->w_wpk, 0, (size_t)n_chains * wsum * 4)); CK(cudaMemset(c->w_wpk_lo, 0, (size_t)n_chains * wsum * 4));
  c->w_part_per_chain = (long)c->w_kslices * maxio;
  CK(cudaMalloc(&c->w_part, 64));      // (legacy scratch; the partial sums of an evaluation live in the arena below)
  {
    // arena of partial sums, per chain and layer: split-K slices (GEMM layers) or 148 row slices (skinny layers) of dW,
    // and row-group column sums (<= 4 per 128-row block, or 148 row slices) for the bias; + the fused head's partials
    size_t per_chain = 0;
    const long nmb = (N + 127) / 128;
    for (int l = 0; l < M.NL; ++l) {
      const long IN = M.dims[l], OUT = M.dims[l + 1];
      const long ns_w = (IN <= WS_KMAX || OUT <= WS_KMAX) ? 148 : c->w_kslices;
      per_chain += (size_t)ns_w * IN * OUT + 32;
      per_chain += (size_t)(4 * nmb > 148 ? 4 * nmb : 148) * OUT + 32;
    }
    per_chain += (size_t)4 * 148 * ((size_t)maxw * 9 + 16) + 128;      // fused head: <= 592 CTAs x (IN*K + IN + K + 1), K <= 8
    if (c->w_arena) { cudaFree(c->w_arena); c->w_arena = nullptr; }
    if (c->w_fin) { cudaFree(c->w_fin); c->w_fin = nullptr; }
    CK(cudaMalloc(&c->w_fin, (size_t)n_chains * (64 + 1) * 4 + 64));
    CK(cudaMemset(c->w_fin, 0, (size_t)n_chains * (64 + 1) * 4 + 64));
    c->w_arena_floats = per_chain * n_chains + 1024;
    CK(cudaMalloc(&c->w_arena, c->w_arena_floats * 4));
  }
  CK(cudaMalloc(&c->w_llpart, (size_t)n_chains * c->w_nblk * 4 + 64));
  CK(cudaMalloc(&c->w_ones, (size_t)N * 4 + 64));
  CK(cudaMalloc(&c->w_gl, (size_t)n_chains * (c->d + 1) * 4));
  fill_kernel<<<148, 256>>>(c->w_ones, N, 1.f);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  c->w_rows = N; c->w_chains = n_chains;
  return 0;
}

// ---- TMA tensor maps for the tcgen05 v2 core: 5-D core-matrix view of a row-major [R x Ccols] fp32 matrix ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int tmap_get(mile_ctx* c, const float* ptr, long R, long Ccols, long ld, long bstride, int nbatch, int box_cols, int box_rows,
                    CUtensorMap* out, int atom32 = 0) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
      return fail("cuTensorMapEncodeTiled is not available");
    g_encode = (EncodeTiledFn)fn;
  }
  auto key = std::make_tuple((const void*)ptr, R, Ccols, ld, bstride, nbatch, box_cols, box_rows + 100000 * atom32);
  auto it = c->tmaps.find(key);
  if (it != c->tmaps.end()) { *out = it->second; return 0; }
  alignas(64) CUtensorMap tm;
  // row-major [nbatch][R][Ccols] fp32, K (= columns) contiguous; box = box_rows x 32 floats (one 128-byte swizzle atom per row)
  cuuint64_t dims[3] = {(cuuint64_t)Ccols, (cuuint64_t)R, (cuuint64_t)(nbatch > 0 ? nbatch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(bstride > 0 ? bstride : R * ld) * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  c->tmaps[key] = tm;
  *out = tm;
  return 0;
}

static bool tc2_operand_ok(const float* p, long ld, long bstride) {
  return p && (ld & 3) == 0 && (bstride & 3) == 0 && ((uintptr_t)p & 15) == 0;
}

// returns 1 if launched on the v2 core, 0 if not eligible, -1 on error
static int wide_gemm_tc2(mile_ctx* c, const GemmArgs& g, cudaStream_t st) {
  const bool a_mn = g.sam == 1 && g.sak != 1, b_mn = g.sbn == 1 && g.sbk != 1;
  const bool a_k = g.sak == 1, b_k = g.sbk == 1;
  const bool mn = a_mn && b_mn;                  // dW = a^T delta: both operands MN-major, no packed remainders
  if (!(a_k && b_k) && !mn) return 0;            // mixed orientations: handled by the v1 core (transposes while staging)
  const long a_ld = a_mn ? g.sak : g.sam, b_ld = b_mn ? g.sbk : g.sbn;
  if (!tc2_operand_ok(g.A, a_ld, g.a_batch) || !tc2_operand_ok(g.B, b_ld, g.b_batch)) return 0;
  if (!mn && !tc2_operand_ok(g.B_lo, b_ld, g.b_batch)) return 0;
  if (mn ? ((g.M & 3) || (g.N & 3) || g.epi != 0) : (g.K & 3) != 0) return 0;
  // K-major products on CTA pairs (tcgen05.mma.cta_group::2, wide_gemm_tc2x_kernel) unless tensor == 2 asks for one CTA per tile
  const bool pairs = !mn && c->opt_tensor >= 3 && g.kslices == 1 && g.M > T2_BM;
  Tc2Args t;
  memset(&t, 0, sizeof(t));
  if (mn) {   // global [K rows][MN cols]: boxes of 32 MN-floats x 32 K-rows
    if (tmap_get(c, g.A, g.K, g.M, a_ld, g.a_batch, g.nbatch, 32, T2_BK, &t.a_hi, 1)) return -1;
    if (tmap_get(c, g.B, g.K, g.N, b_ld, g.b_batch, g.nbatch, 32, T2_BK, &t.b_hi, 1)) return -1;
  } else {
    const int brows = pairs ? 128 : T2_BN;      // CTA pairs: every CTA stages its half of the weight tile
    if (tmap_get(c, g.A, g.M, g.K, a_ld, g.a_batch, g.nbatch, T2_BK, T2_BM, &t.a_hi)) return -1;
    if (tmap_get(c, g.B, g.N, g.K, b_ld, g.b_batch, g.nbatch, T2_BK, brows, &t.b_hi)) return -1;
    if (tmap_get(c, g.B_lo, g.N, g.K, b_ld, g.b_batch, g.nbatch, T2_BK, brows, &t.b_lo)) return -1;
  }
  t.M = g.M; t.N = g.N; t.K = g.K; t.kslices = g.kslices; t.nbatch = g.nbatch;
  t.C = g.C; t.c_batch = g.c_batch; t.c_slice = g.c_slice; t.ldc = g.ldc; t.epi = g.epi; t.act = g.act;
  t.bias = g.bias; t.bias_batch = g.bias_batch; t.aux = g.aux; t.aux_batch = g.aux_batch; t.ldaux = g.ldaux;
  t.csum = g.csum;
  const int ntiles = ((g.M + T2_BM - 1) / T2_BM) * ((g.N + T2_BN - 1) / T2_BN) * g.nbatch * g.kslices;
  const int grid = ntiles < c->n_sms ? ntiles : c->n_sms;      // persistent: one CTA per SM, tiles strided over the grid
  const bool relu = g.act == MILE_ACT_RELU;
  if (pairs) {
    const int ntiles2 = ((g.M + 2 * T2_BM - 1) / (2 * T2_BM)) * ((g.N + T2_BN - 1) / T2_BN) * g.nbatch;
    const int npairs = ntiles2 < c->n_sms / 2 ? ntiles2 : c->n_sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * npairs), 1, 1); cfg.blockDim = dim3(T2_THREADS, 1, 1);
    cfg.dynamicSmemBytes = X2_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
#define X2_LAUNCH(...)                                                                                                      \
    do {                                                                                                                    \
      CK(cudaFuncSetAttribute(wide_gemm_tc2x_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, X2_SMEM_BYTES)); \
      CK(cudaLaunchKernelEx(&cfg, wide_gemm_tc2x_kernel<__VA_ARGS__>, t));                                                  \
    } while (0)
    if (g.epi == 0) X2_LAUNCH(0, false);
    else if (g.epi == 2) X2_LAUNCH(2, false);
    else if (g.epi == 1) { if (relu) X2_LAUNCH(1, true); else X2_LAUNCH(1, false); }
    else { if (relu) X2_LAUNCH(3, true); else X2_LAUNCH(3, false); }
#undef X2_LAUNCH
    CK(cudaGetLastError());
    c->launches++;
    return 1;
  }
#define T2_LAUNCH(...)                                                                                                    \
  do {                                                                                                                    \
    CK(cudaFuncSetAttribute(wide_gemm_tc2_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES)); \
    wide_gemm_tc2_kernel<__VA_ARGS__><<<grid, T2_THREADS, T2_SMEM_BYTES, st>>>(t);                                        \
  } while (0)
  if (mn) T2_LAUNCH(0, false, true);
  else if (g.epi == 0) T2_LAUNCH(0, false);
  else if (g.epi == 2) T2_LAUNCH(2, false);
  else if (g.epi == 1) { if (relu) T2_LAUNCH(1, true); else T2_LAUNCH(1, false); }
  else { if (relu) T2_LAUNCH(3, true); else T2_LAUNCH(3, false); }
#undef T2_LAUNCH
  CK(cudaGetLastError());
  c->launches++;
  return 1;
}

static int wide_gemm(mile_ctx* c, const GemmArgs& g, cudaStream_t st) {
  if (c->opt_tensor >= 2 && g.K >= 32 && g.N >= 64 && g.M >= 64) {   // TMA-fed warp-specialised tcgen05 core
    const int r = wide_gemm_tc2(c, g, st);
    if (r != 0) return r < 0 ? -1 : 0;
  }
  if (c->opt_tensor && g.K >= 32 && g.N >= 64 && g.M >= 64) {   // large contraction: tcgen05 / TMEM core (3xTF32)
    CK(cudaFuncSetAttribute(wide_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    dim3 grid((g.N + TC_BN - 1) / TC_BN, (g.M + TC_BM - 1) / TC_BM, g.nbatch * g.kslices);
    wide_gemm_tc_kernel<<<grid, 256, TC_SMEM_BYTES, st>>>(g);
    CK(cudaGetLastError());
    c->launches++;
    return 0;
  }
  dim3 grid((g.N + WG_BN - 1) / WG_BN, (g.M + WG_BM - 1) / WG_BM, g.nbatch * g.kslices);
  wide_gemm_kernel<<<grid, 256, 0, st>>>(g);
  CK(cudaGetLastError());
  c->launches++;
  return 0;
}

// skinny GEMM (K <= 16 or N <= 8): streaming kernels, no tile padding
static int wide_skinny(mile_ctx* c, const GemmArgs& g, cudaStream_t st) {
  if (g.K <= WS_KMAX && g.epi == 1 && g.sak == 1 && (g.N & 3) == 0 && (g.ldc & 3) == 0 && (g.c_batch & 3) == 0 &&
      ((uintptr_t)g.C & 15) == 0) {            // forward of the first layer: dedicated register-stationary kernel
    const dim3 grid((g.M + 127) / 128, (g.N + 255) / 256, g.nbatch);
    const int KQ = (g.K + 3) / 4;
    const bool relu = g.act == MILE_ACT_RELU;
#define FIRST_LAUNCH(KQv) do { if (relu) wide_first_kernel<KQv, true><<<grid, 256, 0, st>>>(g); else wide_first_kernel<KQv, false><<<grid, 256, 0, st>>>(g); } while (0)
    if (KQ == 1) FIRST_LAUNCH(1); else if (KQ == 2) FIRST_LAUNCH(2); else if (KQ == 3) FIRST_LAUNCH(3); else FIRST_LAUNCH(4);
#undef FIRST_LAUNCH
  }
  else if (g.K <= WS_KMAX) wide_smallk_kernel<<<dim3((g.M + 63) / 64, (g.N + 255) / 256, g.nbatch), 256, 0, st>>>(g);
  else wide_smalln_kernel<<<dim3((g.M + 63) / 64, 1, g.nbatch), 256, 0, st>>>(g);
  CK(cudaGetLastError());
  c->launches++;
  return 0;
}

// ---- partial-sum arena + reduction jobs of one evaluation (summed by wide_finalize_kernel) -----------------------
struct WideEvalState { WideJobs J; size_t used; };
static float* arena_take(mile_ctx* c, WideEvalState& E, size_t floats) {
  floats = (floats + 31) / 32 * 32;
  if (E.used + floats > c->w_arena_floats) return nullptr;
  float* p = c->w_arena + E.used;
  E.used += floats;
  return p;
}
static int add_job(WideEvalState& E, const float* src, long src_batch, long slice, int nslices, int dst_off, int len) {
  if (E.J.n >= WJ_MAX) return fail("wide path: too many reduction jobs");
  const int j = E.J.n++;
  E.J.src[j] = src; E.J.src_batch[j] = src_batch; E.J.slice[j] = slice; E.J.nslices[j] = nslices; E.J.dst_off[j] = dst_off; E.J.len[j] = len;
  return 0;
}

// partials of  sum_r Wd(r, t) S(r, q)  (S == nullptr: column sums) over row slices; the sum over the slices is a finalize job
static int wide_rowreduce(mile_ctx* c, WideEvalState& E, const float* Wd, long wd_batch, long wd_ld, int WD, const float* S, long s_batch,
                          long s_ld, int s, int small_is_row, long rows, int n, int dst_off, long n_out, cudaStream_t st) {
  RowReduceArgs a;
  a.Wd = Wd; a.wd_batch = wd_batch; a.wd_ld = wd_ld; a.WD = WD; a.S = S; a.s_batch = s_batch; a.s_ld = s_ld; a.s = s;
  a.small_is_row = small_is_row; a.rows = rows;
  // dW of a layer with few inputs (S = its input rows): float4-streaming kernel with S staged in shared memory
  const bool dw_small = S && small_is_row && (WD & 3) == 0 && (wd_ld & 3) == 0 && (wd_batch & 3) == 0 && ((uintptr_t)Wd & 15) == 0;
  long nsl = dw_small ? (4 * c->n_sms + n - 1) / n : 148;   // with 8 chains: 592 / 1184 CTAs
  if (nsl > 148) nsl = 148;
  if (nsl > rows) nsl = rows;
  float* part = arena_take(c, E, (size_t)n * nsl * n_out);
  if (!part) return fail("wide path: partial arena too small");
  a.nslices = (int)nsl; a.part = part; a.p_slice = n_out; a.p_batch = nsl * n_out;
  const dim3 grid((unsigned)nsl, (WD + 255) / 256, n);
  if (dw_small) {
    const int KQ = (s + 3) / 4;
    if (KQ == 1) wide_dw_small_kernel<1><<<grid, 256, 0, st>>>(a);
    else if (KQ == 2) wide_dw_small_kernel<2><<<grid, 256, 0, st>>>(a);
    else if (KQ == 3) wide_dw_small_kernel<3><<<grid, 256, 0, st>>>(a);
    else wide_dw_small_kernel<4><<<grid, 256, 0, st>>>(a);
  } else wide_rowreduce_kernel<<<grid, 256, 0, st>>>(a);
  CK(cudaGetLastError());
  c->launches++;
  return add_job(E, part, a.p_batch, a.p_slice, (int)nsl, dst_off, (int)n_out);
}

// forward pass of n chains over the rows of one split: activations a_1..a_NL into actbuf ([n][N8][dims[l]] each, a_NL last)
static int wide_forward(mile_ctx* c, const float* theta, int n, const float* Xs, long N, long N8, float* actbuf, cudaStream_t st, int n_layers) {
  const DevModel& M = c->M;
  const int d = c->d, NL = M.NL;
  if (n_layers < 0) n_layers = NL;
  std::vector<long> aoff(NL + 2, 0);
  for (int l = 1; l <= NL; ++l) aoff[l + 1] = aoff[l] + (long)n * N8 * M.dims[l];
  auto act = [&](int l) { return actbuf + aoff[l]; };
  const bool tc2 = c->opt_tensor >= 2;
  if (tc2) {                                  // aligned packed copies of the weights + their tf32 remainders (GEMM layers only)
    PackArgs pa; memset(&pa, 0, sizeof(pa));
    for (int l = 0; l < NL && pa.n_layers < 13; ++l) {
      // layers whose forward AND backward products are skinny stream their weights from theta (first layer with few
      // features: no input delta is ever formed; output layer with few outputs)
      if ((l == 0 && M.dims[0] <= WS_KMAX) || (M.dims[l + 1] <= WS_NMAX && M.dims[l] <= 1024)) continue;
      const int k = pa.n_layers++;
      pa.kern_off[k] = M.kern_off[l]; pa.IN[k] = M.dims[l]; pa.OUT[k] = M.dims[l + 1]; pa.pack_off[k] = c->w_woff[l];
    }
    if (pa.n_layers > 0) {
      int max_tiles = 1;
      for (int k = 0; k < pa.n_layers; ++k) {
        const int tiles = ((pa.IN[k] + 31) / 32) * ((pa.OUT[k] + 31) / 32);
        if (tiles > max_tiles) max_tiles = tiles;
      }
      wide_pack_weights_kernel<<<dim3(max_tiles, n, pa.n_layers), 256, 0, st>>>(theta, d, pa, c->w_wpk, c->w_wpk_lo, c->w_wpkT, c->w_wpkT_lo,
                                                                               c->w_wstride);
      CK(cudaGetLastError());
      c->launches++;
    }
  }
  for (int l = 0; l < n_layers; ++l) {        // forward
    GemmArgs g; memset(&g, 0, sizeof(g));
    const int IN = M.dims[l], OUT = M.dims[l + 1];
    if (l == 0) { g.A = Xs; g.a_batch = 0; g.sam = M.sA[0]; g.sak = 1; }
    else { g.A = act(l); g.a_batch = N8 * IN; g.sam = IN; g.sak = 1; }
    if (tc2) {   // W^T [OUT x IN]: K-major B for the TMA-fed core (MN-major tf32 operands are not used)
      g.B = c->w_wpkT + c->w_woff[l]; g.B_lo = c->w_wpkT_lo + c->w_woff[l]; g.b_batch = c->w_wstride; g.sbk = 1; g.sbn = IN;
    } else { g.B = theta + M.kern_off[l]; g.b_batch = d; g.sbk = OUT; g.sbn = 1; }
    g.C = act(l + 1);
    g.c_batch = N8 * OUT; g.ldc = OUT; g.M = (int)N; g.N = OUT; g.K = IN; g.kslices = 1;
    g.epi = l < NL - 1 ? 1 : 2; g.bias = theta + M.bias_off[l]; g.bias_batch = d; g.act = M.act; g.nbatch = n;
    if (IN <= WS_KMAX || (OUT <= WS_NMAX && IN <= 1024)) {   // skinny layer: stream it from theta directly
      g.B = theta + M.kern_off[l]; g.B_lo = nullptr; g.b_batch = d; g.sbk = OUT; g.sbn = 1;
      if (wide_skinny(c, g, st)) return -1;
    } else if (wide_gemm(c, g, st)) return -1;
  }
  return 0;
}

// forward-only pass of n chains over a split (which: 0 train, 1 test) -> out [n][N][K]
static int wide_predict(mile_ctx* c, const float* theta, int n, int which, float* out, cudaStream_t st) {
  if (wide_alloc(c, n > c->C ? n : c->C)) return -1;        // packed-weight buffers and their offsets
  const DevModel& M = c->M;
  const float* Xs = which ? c->Xt : c->X;
  const long N = which ? c->Nt : c->N, N8 = (N + 7) / 8 * 8;
  long per_row = 0, before_last = 0;
  for (int l = 1; l <= M.NL; ++l) { if (l == M.NL) before_last = per_row; per_row += M.dims[l]; }
  const size_t need = (size_t)n * N8 * per_row;
  if (need > c->wp_act_floats) {
    if (c->wp_act) { CK(cudaFree(c->wp_act)); c->wp_act = nullptr; c->wp_act_floats = 0; }
    c->tmaps.clear();
    CK(cudaMalloc(&c->wp_act, need * 4));
    c->wp_act_floats = need;
  }
  CK(cudaMemsetAsync(c->wp_act, 0, need * 4, st));            // pad rows (N..N8) must stay finite for the tensor tiles
  if (wide_forward(c, theta, n, Xs, N, N8, c->wp_act, st)) return -1;
  const int K = M.dims[M.NL];
  const float* last = c->wp_act + (size_t)n * N8 * before_last;
  CK(cudaMemcpy2DAsync(out, (size_t)N * K * 4, last, (size_t)N8 * K * 4, (size_t)N * K * 4, n, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// fold theta [n,d] (row c = chain c) into the per-chain online logsumexp state through the wide forward pass
static int wide_lppd_fold(mile_ctx* c, const float* theta, int n, cudaStream_t st) {
  const size_t need = (size_t)n * c->Nt * c->M.dims[c->M.NL];
  if (need > c->wp_out_floats) {
    if (c->wp_out) { CK(cudaFree(c->wp_out)); c->wp_out = nullptr; c->wp_out_floats = 0; }
    CK(cudaMalloc(&c->wp_out, need * 4));
    c->wp_out_floats = need;
  }
  if (wide_predict(c, theta, n, 1, c->wp_out, st)) return -1;
  wide_lppd_fold_kernel<<<296, 256, 0, st>>>(c->M, c->wp_out, c->yt, c->lppd_m, c->lppd_s, n, c->Nt);
  CK(cudaGetLastError());
  c->launches++;
  return 0;
}

// value_and_grad of n chains (theta [n,d]) over the local rows into the packed buffer gl [n, d+1]
static int wide_eval(mile_ctx* c, const float* theta, int n, float* gl, float prior_weight, cudaStream_t st) {
  if (wide_alloc(c, n > c->C ? n : c->C)) return -1;
  if (!gl) gl = c->w_gl;
  const DevModel& M = c->M;
  const long N = c->N, N8 = c->w_n8;          // buffers hold N8 rows per chain (pad rows stay zero)
  const int d = c->d, NL = M.NL;
  std::vector<long> aoff(NL + 2, 0);          // activation buffers a_1..a_NL, each [n][N8][dims[l]]
  for (int l = 1; l <= NL; ++l) aoff[l + 1] = aoff[l] + (long)n * N8 * M.dims[l];
  auto act = [&](int l) { return c->w_act + aoff[l]; };
  const bool tc2 = c->opt_tensor >= 2;
  WideEvalState E; memset(&E, 0, sizeof(E));
  // fused output layer: last hidden width 128 or 256, <= 8 outputs, activation whose derivative follows from its value
  const int HIN = NL >= 2 ? M.dims[NL - 1] : 0, HK = M.dims[NL];
  const bool head_fused = c->opt_head_fused && NL >= 2 && (HIN == 128 || HIN == 256) && HK <= 8;
  if (wide_forward(c, theta, n, c->X, N, N8, c->w_act, st, head_fused ? NL - 1 : NL)) return -1;
  const float* llpart = c->w_llpart; int n_llpart = c->w_nblk;
  bool have_colsum = false;                   // the column sums (bias gradient) of the current delta are already a job
  if (head_fused) {
    int nblk = (4 * c->n_sms + n - 1) / n;
    if (nblk > (N + 31) / 32) nblk = (int)((N + 31) / 32);
    const int KP = HK <= 2 ? 2 : 8;
    HeadArgs h; memset(&h, 0, sizeof(h));
    h.a_last = act(NL - 1); h.a_batch = N8 * HIN; h.D = c->w_delta[(NL - 2) & 1]; h.d_batch = N8 * HIN;
    h.theta = theta; h.d = d; h.kern_off = M.kern_off[NL - 1]; h.bias_off = M.bias_off[NL - 1];
    h.y = c->y; h.N = N; h.rows_per_cta = (int)((N + nblk - 1) / nblk);
    h.pW = arena_take(c, E, (size_t)n * nblk * HIN * HK); h.pb = arena_take(c, E, (size_t)n * nblk * HK);
    h.pcol = arena_take(c, E, (size_t)n * nblk * HIN); h.pll = arena_take(c, E, (size_t)n * nblk);
    if (!h.pW || !h.pb || !h.pcol || !h.pll) return fail("wide path: partial arena too small");
    const size_t smem = ((size_t)HIN * KP + KP + 8 * ((size_t)HIN * KP + HIN + KP + 1)) * 4;
#define HEAD_LAUNCH(NV, KPv)                                                                                              \
    do {                                                                                                                  \
      CK(cudaFuncSetAttribute(wide_head_kernel<NV, KPv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
      wide_head_kernel<NV, KPv><<<dim3(nblk, n), 256, smem, st>>>(M, h);                                                  \
    } while (0)
    if (HIN == 128) { if (KP == 2) HEAD_LAUNCH(1, 2); else HEAD_LAUNCH(1, 8); }
    else { if (KP == 2) HEAD_LAUNCH(2, 2); else HEAD_LAUNCH(2, 8); }
#undef HEAD_LAUNCH
    CK(cudaGetLastError());
    c->launches++;
    if (add_job(E, h.pW, (long)nblk * HIN * HK, (long)HIN * HK, nblk, M.kern_off[NL - 1], HIN * HK)) return -1;
    if (add_job(E, h.pb, (long)nblk * HK, HK, nblk, M.bias_off[NL - 1], HK)) return -1;
    if (add_job(E, h.pcol, (long)nblk * HIN, HIN, nblk, M.bias_off[NL - 2], HIN)) return -1;
    llpart = h.pll; n_llpart = nblk;
    have_colsum = true;
  } else {
    wide_loglik_kernel<<<dim3(c->w_nblk, n), 256, 0, st>>>(M, act(NL), c->w_delta[(NL - 1) & 1], c->y, N, N8, c->w_llpart);
    CK(cudaGetLastError());
    c->launches++;
  }
  for (int l = head_fused ? NL - 2 : NL - 1; l >= 0; --l) {         // dW_l, db_l, then the delta of the layer below
    const int IN = M.dims[l], OUT = M.dims[l + 1];
    float* D = c->w_delta[l & 1];
    const float* Aop = l == 0 ? c->X : act(l);
    const long a_bs = l == 0 ? 0 : N8 * IN, a_ld = l == 0 ? M.sA[0] : IN;
    if (IN <= WS_KMAX) {          // dW[i][j] = sum_r A(r,i) D(r,j), few rows i: thread = column j of D
      if (wide_rowreduce(c, E, D, N8 * OUT, OUT, OUT, Aop, a_bs, a_ld, IN, 1, N, n, M.kern_off[l], (long)IN * OUT, st)) return -1;
    } else if (OUT <= WS_KMAX) {  // few columns j: thread = column i of A
      if (wide_rowreduce(c, E, Aop, a_bs, a_ld, IN, D, N8 * OUT, OUT, OUT, 0, N, n, M.kern_off[l], (long)IN * OUT, st)) return -1;
    } else {
      GemmArgs g; memset(&g, 0, sizeof(g));
      g.A = Aop; g.a_batch = a_bs; g.sam = 1; g.sak = a_ld;
      g.B = D; g.b_batch = N8 * OUT; g.sbk = OUT; g.sbn = 1;
      g.M = IN; g.N = OUT; g.K = (int)N8; g.kslices = c->w_kslices; g.nbatch = n; g.epi = 0;
      float* part = arena_take(c, E, (size_t)n * c->w_kslices * IN * OUT);
      if (!part) return fail("wide path: partial arena too small");
      g.C = part; g.c_slice = (long)IN * OUT; g.c_batch = (long)c->w_kslices * IN * OUT; g.ldc = OUT;
      if (wide_gemm(c, g, st)) return -1;
      if (add_job(E, part, g.c_batch, g.c_slice, c->w_kslices, M.kern_off[l], IN * OUT)) return -1;
    }
    // bias gradient = column sums of D (already a job when the kernel that produced D summed its columns)
    if (!have_colsum &&
        wide_rowreduce(c, E, D, N8 * OUT, OUT, OUT, nullptr, 0, 0, 1, 1, N, n, M.bias_off[l], (long)OUT, st)) return -1;
    have_colsum = false;
    if (l > 0) {
      GemmArgs w; memset(&w, 0, sizeof(w));
      w.A = D; w.a_batch = N8 * OUT; w.sam = OUT; w.sak = 1;
      if (tc2) { w.B = c->w_wpk + c->w_woff[l]; w.B_lo = c->w_wpk_lo + c->w_woff[l]; w.b_batch = c->w_wstride; }
      else { w.B = theta + M.kern_off[l]; w.b_batch = d; }
      w.sbk = 1; w.sbn = OUT;                                                   // B(k=j, n=i) = W[i][j]
      w.C = c->w_delta[(l - 1) & 1];
      w.c_batch = N8 * IN; w.ldc = IN; w.M = (int)N; w.N = IN; w.K = OUT; w.kslices = 1;
      w.epi = 3; w.aux = act(l); w.aux_batch = N8 * IN; w.ldaux = IN; w.act = M.act; w.nbatch = n;
      if (OUT <= WS_KMAX) {
        w.B = theta + M.kern_off[l]; w.B_lo = nullptr; w.b_batch = d;
        if (wide_skinny(c, w, st)) return -1;
      } else {
        // TMA core with a full-width epilogue: it also leaves the column sums of the new delta (bias gradient of layer l-1)
        const long nmb = (N + T2_BM - 1) / T2_BM;
        const bool cs_ok = tc2 && OUT >= 32 && IN >= 64 && N >= 64 && (IN % T2_BN) == 0 && (OUT & 3) == 0 &&
                           tc2_operand_ok(w.A, OUT, w.a_batch) && tc2_operand_ok(w.B, OUT, w.b_batch) && tc2_operand_ok(w.B_lo, OUT, w.b_batch) &&
                           tc2_operand_ok(w.C, IN, w.c_batch) && tc2_operand_ok(w.aux, IN, w.aux_batch);
        if (cs_ok) {
          w.csum = arena_take(c, E, (size_t)n * nmb * 4 * IN);
          if (!w.csum) return fail("wide path: partial arena too small");
        }
        if (wide_gemm(c, w, st)) return -1;
        if (cs_ok) {
          if (add_job(E, w.csum, nmb * 4 * IN, IN, (int)(nmb * 4), M.bias_off[l - 1], IN)) return -1;
          have_colsum = true;
        }
      }
    }
  }
  {
    int NB = c->n_sms / n;       // CTAs per chain: the pass over the partials (89 MB for 4x256 x 8 chains) needs every SM
    if (NB < 1) NB = 1;
    if (NB > 32) NB = 32;
    const int wc = c->w_chains > n ? c->w_chains : n;
    wide_finalize_kernel<<<dim3(NB, n), 1024, 0, st>>>(M, theta, gl, llpart, n_llpart, prior_weight, c->w_fin,
                                                       reinterpret_cast<unsigned int*>(c->w_fin + (size_t)wc * 64), E.J);
  }
  CK(cudaGetLastError());
  c->launches++;
  return 0;
}

// local value_and_grad of the rank's row shard into the packed buffer, then all-reduce over the ranks: through NCCL, or
// (use_p2p, step loop with an open peer mapping) left in this rank's exchange region for the integrator kernel to sum out
// of peer memory
static float* xr_slot(mile_ctx* c, float* region, unsigned int epoch) { return region + (size_t)(epoch & 1u) * c->xr_n; }
static unsigned int* xr_flag(mile_ctx* c, float* region, unsigned int epoch) {
  return reinterpret_cast<unsigned int*>(region + 2 * c->xr_n) + (epoch & 1u);
}
static int shard_eval(mile_ctx* c, cudaStream_t st, bool use_p2p = false) {
  const bool p2p = use_p2p && c->p2p && c->world > 1;
  float* target = c->gl;
  if (p2p) { c->xr_epoch++; target = xr_slot(c, c->xr, c->xr_epoch); }
  if (c->wide) {
    if (wide_eval(c, c->theta, c->C, target, 1.f / (float)c->world, st)) return -1;
  } else {
    Plan pl;
    if (make_plan(c, c->C, c->N, true, pl)) return -1;
    fill_common(c, pl.kp);
    pl.kp.mode = MODE_EVAL; pl.kp.theta_in = c->theta; pl.kp.grad_out = target; pl.kp.lp_out = target + c->d;
    pl.kp.out_stride = c->d + 1; pl.kp.prior_weight = 1.f / (float)c->world; pl.kp.n_eval = c->C;
    if (launch(c, pl, c->C, st)) return -1;
  }
  if (c->world > 1 && !p2p)
    NCK(g_nccl.AllReduce(c->gl, c->gl, (size_t)c->C * (c->d + 1), ncclFloat, ncclSum, (ncclComm_t)c->nccl_comm, st));
  return 0;
}

static int shard_integ(mile_ctx* c, ShardParams& S, int stage, long s_local, cudaStream_t st, bool use_p2p = false) {
  S.stage = stage; S.s_local = s_local;
  S.p2p = (use_p2p && c->p2p && c->world > 1) ? 1 : 0; S.world = c->world; S.rank = c->rank; S.epoch = c->xr_epoch;
  if (S.p2p)
    for (int r = 0; r < c->world; ++r) {
      float* region = r == c->rank ? c->xr : c->xr_peer[r];
      S.peer_data[r] = xr_slot(c, region, c->xr_epoch); S.peer_flag[r] = xr_flag(c, region, c->xr_epoch);
    }
  if (c->d > 8192) {   // large d: a cluster of 8 CTAs per chain, elements strided over its 8192 threads (DSMEM reductions)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    // 16-CTA clusters (non-portable size) when all of them fit at once: twice the threads per chain for a pass that is
    // bound by the latency of its element loops and cluster reductions; 8 otherwise (or when the device refuses 16)
    int cs = (c->integ_cluster == 16 && (long)c->C * 16 <= c->n_sms) ? 16 : 8;
    for (;;) {
      cfg.gridDim = dim3((unsigned)c->C * cs, 1, 1); cfg.blockDim = dim3(1024, 1, 1); cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      if (cs == 16) cudaFuncSetAttribute(mile_integrator_kernel<1024, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      const cudaError_t e = cudaLaunchKernelEx(&cfg, mile_integrator_kernel<1024, true>, S);
      if (e == cudaSuccess) break;
      if (cs == 16) { cudaGetLastError(); cs = 8; c->integ_cluster = 8; continue; }
      CK(e);
    }
  } else mile_integrator_kernel<256><<<c->C, 256, 0, st>>>(S);
  CK(cudaGetLastError());
  c->launches++;
  return 0;
}

static int shard_params(mile_ctx* c, ShardParams& S) {
  memset(&S, 0, sizeof(S));
  S.K.M = c->M; S.K.dS = round_up(c->d, 4); S.K.C = c->C;
  fill_common(c, S.K);
  S.gl = c->gl; S.scal = c->scal; S.thb = c->thb; S.ub = c->ub; S.gb = c->gb;
  return 0;
}

// ---- peer-memory all-reduce over NVLink (CUDA IPC) --------------------------------------------------------------
int mile_shard_p2p_handle(mile_ctx* c, void* out64) {
  if (!c || !out64) return fail("null argument");
  if (!c->gl) return fail("mile_shard_init has not been called");
  CK(cudaSetDevice(c->device));
  if (!c->xr) {
    c->xr_n = ((size_t)c->C * (c->d + 1) + 31) / 32 * 32;
    c->mr_off = 2 * c->xr_n + 32;
    // + the flagged-word area of the fused multi-rank step loop: [C][2][world][dS+4] float2
    const size_t mr_floats = c->wide ? 0 : (size_t)c->C * 2 * c->world * (round_up(c->d, 4) + 4) * 2;
    CK(cudaMalloc(&c->xr, (c->mr_off + mr_floats) * 4));
    CK(cudaMemset(c->xr, 0, (c->mr_off + mr_floats) * 4));
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->xr));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  memcpy(out64, &h, 64);
  return 0;
}
int mile_shard_p2p_open(mile_ctx* c, const void* handles64) {
  if (!c || !handles64) return fail("null argument");
  if (!c->xr) return fail("mile_shard_p2p_handle has not been called");
  if (c->world < 2 || c->world > 8) return fail("the peer-memory all-reduce serves 2..8 ranks of one box");
  CK(cudaSetDevice(c->device));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->xr_peer[r] = c->xr; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles64 + 64 * r, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->xr_peer[r] = (float*)p;
  }
  c->p2p = 1;
  return 0;
}

int mile_shard_mclmc_init(mile_ctx* c, const float* theta0_dev, const float* z0_dev, uint64_t seed, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->X) return fail("mile_set_data has not been called");
  if (!c->gl) return fail("mile_shard_init has not been called");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Cd = (size_t)c->C * c->d * 4;
  CK(cudaMemcpyAsync(c->theta, theta0_dev, Cd, cudaMemcpyDeviceToDevice, st));
  if (shard_eval(c, st)) return -1;
  mile_unit_momentum_kernel<256><<<c->C, 256, 0, st>>>(c->u, z0_dev, seed, c->d, c->opt_chain_base);
  CK(cudaGetLastError());
  c->launches++;
  CK(cudaMemcpy2DAsync(c->grad, (size_t)c->d * 4, c->gl, (size_t)(c->d + 1) * 4, (size_t)c->d * 4, c->C, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpy2DAsync(c->lp, 4, c->gl + c->d, (size_t)(c->d + 1) * 4, 4, c->C, cudaMemcpyDeviceToDevice, st));
  c->carry_valid = 0;
  return 0;
}

static int shard_run(mile_ctx* c, ShardParams& S, int n_steps, cudaStream_t st) {
  for (int s = 0; s < n_steps; ++s) {
    if (shard_integ(c, S, SH_BEGIN, s, st)) return -1;
    if (shard_eval(c, st, true)) return -1;
    if (shard_integ(c, S, SH_MID, s, st, true)) return -1;
    if (shard_eval(c, st, true)) return -1;
    if (shard_integ(c, S, SH_END, s, st, true)) return -1;
    if (c->shard_lppd && c->wide && !S.tune && (S.K.step_base + s) % S.K.thin == 0 && wide_lppd_fold(c, c->theta, c->C, st)) return -1;
  }
  c->carry_valid = 0;
  return 0;
}

// Fused multi-rank step loop: when the peer mapping is open and the model runs on the shared-memory kernels, the whole
// call is ONE launch of mile_mclmc_kernel per rank (rows of this rank, flagged-word exchange over NVLink inside the
// kernel, KParams::mr_*).  Returns 1 when the plan does not qualify (the caller then runs the launch-per-phase loop).
static int shard_fused_plan(mile_ctx* c, Plan& pl) {
  if (!(c->p2p && c->world > 1 && !c->wide && c->opt_shard_fused && c->mr_off)) return 1;
  const int save_fast = c->opt_fast;
  c->opt_fast = save_fast < 1 ? save_fast : 1;       // the exchange lives in mile_mclmc_kernel (generic / FastGE evaluators)
  const int rc = make_plan(c, c->C, c->N, true, pl, 0, true);
  c->opt_fast = save_fast;
  if (rc) return 1;
  if ((long)pl.G * c->C > c->n_sms) return 1;      // every CTA of every rank must be resident: the ranks wait on each other
  if (pl.G == 1) pl.sync_mode = 0;
  return 0;
}
static int shard_fused_launch(mile_ctx* c, Plan& pl, cudaStream_t st) {
  KParams& k = pl.kp;
  k.mr_world = c->world; k.mr_rank = c->rank; k.mr_base = c->mr_epoch;
  for (int r = 0; r < c->world; ++r) k.mr_sums[r] = reinterpret_cast<float2*>((r == c->rank ? c->xr : c->xr_peer[r]) + c->mr_off);
  c->mr_epoch += 2u * (unsigned int)k.n_steps + 2u;     // same sequence of calls on every rank -> same epochs
  if (launch(c, pl, c->C, st)) return -1;
  c->carry_valid = 1;
  return 0;
}

int mile_shard_mclmc_sample(mile_ctx* c, int32_t n_steps, int64_t step_base, int32_t n_thinning, int64_t sample_base,
                            const float* step_size_dev, const float* L_dev, const float* z_dev, uint64_t seed,
                            float* samples_dev, int64_t n_slots, float* info_dev, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->gl) return fail("mile_shard_init has not been called");
  if (!step_size_dev || !L_dev) return fail("step_size / L are required");
  if (n_steps <= 0) return 0;
  {
    Plan pl;
    if (shard_fused_plan(c, pl) == 0) {
      fill_common(c, pl.kp);
      KParams& k = pl.kp;
      k.mode = MODE_SAMPLE; k.n_steps = n_steps; k.step_base = step_base; k.thin = n_thinning; k.sample_base = sample_base;
      k.n_slots = n_slots; k.eps = step_size_dev; k.L = L_dev; k.z = z_dev; k.seed = seed; k.samples = samples_dev;
      k.info = info_dev; k.do_lppd = 0;
      return shard_fused_launch(c, pl, (cudaStream_t)stream);
    }
  }
  ShardParams S;
  shard_params(c, S);
  KParams& k = S.K;
  k.n_steps = n_steps; k.step_base = step_base; k.thin = n_thinning; k.sample_base = sample_base; k.n_slots = n_slots;
  k.eps = step_size_dev; k.L = L_dev; k.z = z_dev; k.seed = seed; k.samples = samples_dev; k.info = info_dev;
  S.tune = 0;
  return shard_run(c, S, n_steps, (cudaStream_t)stream);
}

int mile_shard_mclmc_tune(mile_ctx* c, int32_t n_steps, int64_t step_base, const mile_tune_cfg* cfg, const float* z_dev,
                          uint64_t seed, float* tune_info_dev, void* stream) {
  if (!c || !cfg) return fail("null argument");
  if (!c->gl) return fail("mile_shard_init has not been called");
  if (n_steps <= 0) return 0;
  {
    Plan pl;
    if (shard_fused_plan(c, pl) == 0) {
      fill_common(c, pl.kp);
      KParams& k = pl.kp;
      k.mode = MODE_TUNE; k.n_steps = n_steps; k.step_base = step_base; k.z = z_dev; k.seed = seed; k.tune_info = tune_info_dev;
      k.tune1 = cfg->tune1_steps; k.tune2 = cfg->tune2_steps; k.ev_start = cfg->desired_energy_var_start;
      k.ev_end = cfg->desired_energy_var_end; k.trust = cfg->trust_in_estimate; k.neff = cfg->num_effective_samples;
      return shard_fused_launch(c, pl, (cudaStream_t)stream);
    }
  }
  ShardParams S;
  shard_params(c, S);
  KParams& k = S.K;
  k.n_steps = n_steps; k.step_base = step_base; k.z = z_dev; k.seed = seed; k.tune_info = tune_info_dev;
  k.tune1 = cfg->tune1_steps; k.tune2 = cfg->tune2_steps; k.ev_start = cfg->desired_energy_var_start;
  k.ev_end = cfg->desired_energy_var_end; k.trust = cfg->trust_in_estimate; k.neff = cfg->num_effective_samples;
  S.tune = 1;
  return shard_run(c, S, n_steps, (cudaStream_t)stream);
}


#ifdef MILE_PROFILE
// Developer hook (only in the -DMILE_PROFILE build of tools/phase_profile.py; not part of the shipped ABI):
// C[M,N] = A(M,K) * B(K,N) with chosen operand orientations through the wide-path GEMM cores
// (core: 0 SIMT, 1 tcgen05 v1, 2 tcgen05 v2 TMA).  a_mn: A stored [K x M] (m contiguous) instead of [M x K];
// b_mn: B stored [K x N] (n contiguous) instead of [N x K].  Host pointers.
int mile_debug_wide_gemm(int32_t device, int32_t core, int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                         const float* A_host, const float* B_host, float* C_host) {
  CK(cudaSetDevice(device));
  mile_ctx ctx;   // only the tensor-map cache / launch counter are used
  ctx.opt_tensor = core;
  float *A, *Alo, *B, *Blo, *Cd;
  const long K8 = (K + 7) / 8 * 8;
  const size_t ab = (size_t)(a_mn ? K8 * M : (long)((M + 7) / 8 * 8) * K) * 4, bb = (size_t)(b_mn ? K8 * N : (long)((N + 7) / 8 * 8) * K) * 4;
  CK(cudaMalloc(&A, ab)); CK(cudaMalloc(&Alo, ab)); CK(cudaMalloc(&B, bb)); CK(cudaMalloc(&Blo, bb)); CK(cudaMalloc(&Cd, (size_t)M * N * 4));
  CK(cudaMemset(A, 0, ab)); CK(cudaMemset(Alo, 0, ab)); CK(cudaMemset(B, 0, bb)); CK(cudaMemset(Blo, 0, bb));
  std::vector<float> lo;
  auto up = [&](float* hi_d, float* lo_d, const float* h, size_t n) -> int {
    lo.resize(n);
    for (size_t i = 0; i < n; ++i) { uint32_t u; memcpy(&u, &h[i], 4); u &= 0xFFFFE000u; float t; memcpy(&t, &u, 4); lo[i] = h[i] - t; }
    CK(cudaMemcpy(hi_d, h, n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(lo_d, lo.data(), n * 4, cudaMemcpyHostToDevice));
    return 0;
  };
  if (up(A, Alo, A_host, (size_t)M * K) || up(B, Blo, B_host, (size_t)N * K)) return -1;
  GemmArgs g; memset(&g, 0, sizeof(g));
  g.A = A; g.A_lo = Alo; g.B = B; g.B_lo = Blo; g.C = Cd; g.ldc = N; g.M = M; g.N = N; g.K = K; g.kslices = 1; g.nbatch = 1;
  if (a_mn) { g.sam = 1; g.sak = M; } else { g.sam = K; g.sak = 1; }
  if (b_mn) { g.sbk = N; g.sbn = 1; } else { g.sbk = 1; g.sbn = K; }
  if (wide_gemm(&ctx, g, 0)) return -1;
  CK(cudaDeviceSynchronize());
  if (getenv("MILE_DEBUG_TIMING")) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    for (int r = 0; r < 10; ++r) if (wide_gemm(&ctx, g, 0)) return -1;
    cudaEventRecord(e1, 0);
    CK(cudaEventSynchronize(e1));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[mile debug] gemm core %d M %d N %d K %d: %.1f us per launch\n", core, M, N, K, ms * 100.f);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  CK(cudaMemcpy(C_host, Cd, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  cudaFree(A); cudaFree(Alo); cudaFree(B); cudaFree(Blo); cudaFree(Cd);
  return 0;
}
#endif  // MILE_PROFILE

// ---- NUTS branch of the sampling seam (mile_nuts.cuh) ---------------------------------------------------------------
__global__ void nuts_adapt_init_kernel(float* imm, float* mean, float* m2, float* da, int C, int d, float eps0) {
  const long n = (long)C * d;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    imm[i] = 1.f; mean[i] = 0.f; m2[i] = 0.f;
  }
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float* a = da + (long)c * 8;   // window_adaptation.base.init: dual averaging at log(eps0), mu = log(10 eps0)
    a[0] = logf(eps0); a[1] = 0.f; a[2] = 1.f; a[3] = 0.f; a[4] = logf(10.f * eps0); a[5] = 0.f; a[6] = eps0; a[7] = 0.f;
  }
}
__global__ void nuts_adapt_final_kernel(float* da, int C) {   // adapt_final: step_size = exp(log_step_size_avg)
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) da[(long)c * 8 + 6] = expf(da[(long)c * 8 + 1]);
}

static int nuts_alloc(mile_ctx* c) {
  if (c->nuts_imm) return 0;
  const size_t Cd = (size_t)c->C * c->d * 4;
  CK(cudaMalloc(&c->nuts_imm, Cd)); CK(cudaMalloc(&c->nuts_mean, Cd)); CK(cudaMalloc(&c->nuts_m2, Cd));
  CK(cudaMalloc(&c->nuts_da, (size_t)c->C * 8 * 4));
  return 0;
}

int mile_nuts_init(mile_ctx* c, const float* theta0_dev, const mile_nuts_cfg* cfg, void* stream) {
  if (!c || !cfg) return fail("null ctx / cfg");
  if (!c->X) return fail("mile_set_data has not been called");
  if (c->wide) return fail("NUTS is not available on the wide path");
  if (c->sdc_on) return fail("NUTS with an MCLMC preconditioner (sqrt_diag_cov) set is not supported: clear it first");
  if (cfg->max_num_doublings < 1 || cfg->max_num_doublings > 12) return fail("max_num_doublings must be in [1, 12]");
  if (!(cfg->initial_step_size > 0.f)) return fail("initial_step_size must be positive");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (nuts_alloc(c)) return -1;
  c->nuts_max_doublings = cfg->max_num_doublings; c->nuts_div = cfg->divergence_threshold; c->nuts_target = cfg->target_acceptance_rate;
  // hmc.init: position, logdensity, gradient
  if (mile_logpost_value_and_grad(c, theta0_dev, c->C, c->lp, c->grad, stream)) return -1;
  CK(cudaMemcpyAsync(c->theta, theta0_dev, (size_t)c->C * c->d * 4, cudaMemcpyDeviceToDevice, st));
  nuts_adapt_init_kernel<<<64, 256, 0, st>>>(c->nuts_imm, c->nuts_mean, c->nuts_m2, c->nuts_da, c->C, c->d, cfg->initial_step_size);
  CK(cudaGetLastError());
  c->carry_valid = 0;
  return 0;
}

static int nuts_launch(mile_ctx* c, int n_steps, long step_base, const unsigned char* schedule_dev, int thin, long sample_base,
                       const float* z_dev, const float* uni_dev, uint64_t seed, float* samples_dev, long n_slots,
                       float* info_dev, int lppd, cudaStream_t st) {
  if (!c->nuts_imm) return fail("mile_nuts_init has not been called");
  if (n_steps == 0) return 0;
  Plan pl;
  const int keep_fast = c->opt_fast;
  if (c->opt_fast == 1) c->opt_fast = 0;   // (no NUTS instantiation of the FFMA layer pipeline)
  const int prc = make_plan(c, c->C, c->N, true, pl);
  c->opt_fast = keep_fast;
  if (prc) return -1;
  const int D = c->nuts_max_doublings;
  const size_t need = (size_t)c->C * pl.G * (size_t)(10 + 2 * D) * pl.kp.dS;
  if (need > c->nuts_scratch_floats) {
    if (c->nuts_scratch) cudaFree(c->nuts_scratch);
    c->nuts_scratch = nullptr; c->nuts_scratch_floats = 0;
    CK(cudaMalloc(&c->nuts_scratch, need * 4));
    c->nuts_scratch_floats = need;
  }
  fill_common(c, pl.kp);
  KParams& k = pl.kp;
  k.mode = MODE_NUTS; k.n_steps = n_steps; k.step_base = step_base; k.thin = thin; k.sample_base = sample_base;
  k.n_slots = n_slots; k.z = z_dev; k.seed = seed; k.samples = samples_dev; k.do_lppd = lppd;
  k.sdc = nullptr;
  NutsParams& q = k.nuts;
  q.scratch = c->nuts_scratch; q.imm = c->nuts_imm; q.w_mean = c->nuts_mean; q.w_m2 = c->nuts_m2; q.da = c->nuts_da;
  q.schedule = schedule_dev; q.uni = uni_dev; q.info = info_dev;
  q.uni_len = 2 * D + (1 << D); q.max_doublings = D; q.divergence_threshold = c->nuts_div; q.target_accept = c->nuts_target;
  // end states, proposal, checkpoints and Welford moments in shared memory when they fit behind the plan (1 CTA per SM either way)
  const size_t extra = (size_t)(10 + 2 * D) * pl.kp.dS * 4;
  q.smem_off = -1; q.push = c->opt_nuts_push;
  if (c->opt_nuts_smem && pl.smem + extra <= kSmemLimit) { q.smem_off = (int)(pl.smem / 4); pl.smem += extra; }
  return launch(c, pl, c->C, st);
}

int mile_nuts_warmup(mile_ctx* c, int32_t n_steps, int64_t step_base, const uint8_t* schedule_dev, const float* z_dev,
                     const float* uni_dev, uint64_t seed, float* positions_dev, float* info_dev, void* stream) {
  if (!c) return fail("null ctx");
  if (n_steps < 0) return fail("n_steps must be >= 0");
  if (!schedule_dev && n_steps > 0) return fail("the adaptation schedule is required");
  CK(cudaSetDevice(c->device));
  return nuts_launch(c, n_steps, step_base, schedule_dev, 1, 0, z_dev, uni_dev, seed, positions_dev, n_steps, info_dev, 0, (cudaStream_t)stream);
}

int mile_nuts_finish_warmup(mile_ctx* c, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->nuts_imm) return fail("mile_nuts_init has not been called");
  CK(cudaSetDevice(c->device));
  nuts_adapt_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c->nuts_da, c->C);
  CK(cudaGetLastError());
  return 0;
}

int mile_nuts_sample(mile_ctx* c, int32_t n_steps, int64_t step_base, int32_t n_thinning, int64_t sample_base, const float* z_dev,
                     const float* uni_dev, uint64_t seed, float* samples_dev, int64_t n_slots, float* info_dev, int32_t lppd,
                     void* stream) {
  if (!c) return fail("null ctx");
  if (n_steps < 0 || n_thinning < 1) return fail("n_steps must be >= 0 and n_thinning >= 1");
  if (lppd && !c->Xt) return fail("lppd requested but mile_set_test has not been called");
  CK(cudaSetDevice(c->device));
  if (lppd && !c->lppd_m && lppd_alloc(c, (cudaStream_t)stream)) return -1;
  if (nuts_launch(c, n_steps, step_base, nullptr, n_thinning, sample_base, z_dev, uni_dev, seed, samples_dev, n_slots, info_dev,
                  lppd, (cudaStream_t)stream)) return -1;
  if (lppd && n_steps > 0) {
    const long first = (step_base + n_thinning - 1) / n_thinning, last = (step_base + n_steps - 1) / n_thinning;
    c->lppd_count += (last >= first) ? (last - first + 1) : 0;
  }
  return 0;
}

int mile_nuts_init_host(mile_ctx* c, const float* theta0, const mile_nuts_cfg* cfg) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  const size_t Cd = (size_t)c->C * c->d * 4;
  float* th = (float*)scratch(c, 2, Cd);
  if (!th) return fail("cudaMalloc failed (staging)");
  CK(cudaMemcpyAsync(th, theta0, Cd, cudaMemcpyHostToDevice, c->own_stream));
  if (mile_nuts_init(c, th, cfg, c->own_stream)) return -1;
  CK(cudaStreamSynchronize(c->own_stream));
  return 0;
}

// schedule [n_steps] (warm-up) or NULL (sampling); z [n_steps,C,d] / uni [n_steps,C,2D+2^D] or NULL (in-kernel Philox)
int mile_nuts_run_host(mile_ctx* c, int32_t n_steps, int64_t step_base, const uint8_t* schedule, int32_t n_thinning,
                       const float* z, const float* uni, uint64_t seed, float* samples, int64_t n_slots, float* info,
                       int32_t lppd) {
  if (!c) return fail("null ctx");
  if (n_steps < 0) return fail("n_steps must be >= 0");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->own_stream;
  const size_t Cd = (size_t)c->C * c->d * 4;
  const size_t ulen = (size_t)(2 * c->nuts_max_doublings + (1 << c->nuts_max_doublings));
  const size_t ub = (size_t)n_steps * c->C * ulen * 4, ib = (size_t)n_steps * c->C * NUTS_INFO * 4;
  unsigned char* sd = schedule ? (unsigned char*)scratch(c, 5, (size_t)n_steps + 4) : nullptr;
  float* zd = z ? (float*)scratch(c, 7, (size_t)n_steps * Cd) : nullptr;
  float* ud = uni ? (float*)scratch(c, 6, ub) : nullptr;
  float* smp = samples ? (float*)scratch(c, 8, (size_t)n_slots * Cd) : nullptr;
  float* id = info ? (float*)scratch(c, 9, ib) : nullptr;
  if ((schedule && !sd) || (z && !zd) || (uni && !ud) || (samples && !smp) || (info && !id)) return fail("cudaMalloc failed (staging)");
  if (schedule) CK(cudaMemcpyAsync(sd, schedule, (size_t)n_steps, cudaMemcpyHostToDevice, st));
  if (z) CK(cudaMemcpyAsync(zd, z, (size_t)n_steps * Cd, cudaMemcpyHostToDevice, st));
  if (uni) CK(cudaMemcpyAsync(ud, uni, ub, cudaMemcpyHostToDevice, st));
  if (schedule) {
    if (samples && n_slots != n_steps) return fail("warm-up positions: n_slots must equal n_steps");
    if (mile_nuts_warmup(c, n_steps, step_base, sd, zd, ud, seed, smp, id, st)) return -1;
    if (samples) CK(cudaMemcpyAsync(samples, smp, (size_t)n_slots * Cd, cudaMemcpyDeviceToHost, st));
  } else {
    const int64_t sample_base = (step_base + n_thinning - 1) / n_thinning;
    if (mile_nuts_sample(c, n_steps, step_base, n_thinning, sample_base, zd, ud, seed, smp, n_slots, id, lppd, st)) return -1;
    if (samples) CK(cudaMemcpyAsync(samples, smp, (size_t)n_slots * Cd, cudaMemcpyDeviceToHost, st));
  }
  if (info) CK(cudaMemcpyAsync(info, id, ib, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int mile_nuts_get_params_host(mile_ctx* c, float* step_size, float* inverse_mass_matrix) {
  if (!c) return fail("null ctx");
  if (!c->nuts_imm) return fail("mile_nuts_init has not been called");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  if (step_size) CK(cudaMemcpy2D(step_size, 4, c->nuts_da + 6, 32, 4, c->C, cudaMemcpyDeviceToHost));
  if (inverse_mass_matrix) CK(cudaMemcpy(inverse_mass_matrix, c->nuts_imm, (size_t)c->C * c->d * 4, cudaMemcpyDeviceToHost));
  return 0;
}
int mile_nuts_set_params_host(mile_ctx* c, const float* step_size, const float* inverse_mass_matrix) {
  if (!c) return fail("null ctx");
  if (!c->nuts_imm) return fail("mile_nuts_init has not been called");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  if (step_size) CK(cudaMemcpy2D(c->nuts_da + 6, 32, step_size, 4, 4, c->C, cudaMemcpyHostToDevice));
  if (inverse_mass_matrix) CK(cudaMemcpy(c->nuts_imm, inverse_mass_matrix, (size_t)c->C * c->d * 4, cudaMemcpyHostToDevice));
  return 0;
}

// ---- phase 3 of the warmup on the device (warmup.py:408-465): capture + effective sample size -----------------------
static int ess_run(mile_ctx* c, const float* pos_dev, int n_total, const int32_t* param_idx, int n_sel, const int32_t* sample_idx,
                   int n_samples_sel, float* ess_host, cudaStream_t st, bool pooled = false) {
  const int d_sel = param_idx ? n_sel : c->d;
  const int n = sample_idx ? n_samples_sel : n_total;
  if (n < 4) return fail("effective sample size needs at least 4 samples");
  if (d_sel < 1) return fail("no parameters selected");
  const long n_series = (long)c->C * d_sel;
  int* pidx_d = nullptr; int* sidx_d = nullptr;
  if (param_idx) {
    pidx_d = (int*)scratch(c, 10, (size_t)d_sel * 4);
    if (!pidx_d) return fail("cudaMalloc failed (ess)");
    for (int j = 0; j < d_sel; ++j) if (param_idx[j] < 0 || param_idx[j] >= c->d) return fail("parameter index out of range");
    CK(cudaMemcpyAsync(pidx_d, param_idx, (size_t)d_sel * 4, cudaMemcpyHostToDevice, st));
  }
  if (sample_idx) {
    sidx_d = (int*)scratch(c, 11, (size_t)n * 4);
    if (!sidx_d) return fail("cudaMalloc failed (ess)");
    for (int i = 0; i < n; ++i) if (sample_idx[i] < 0 || sample_idx[i] >= n_total) return fail("sample index out of range");
    CK(cudaMemcpyAsync(sidx_d, sample_idx, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  }
  float* series = (float*)scratch(c, 12, (size_t)n_series * n * 4);
  float* ess_d = (float*)scratch(c, 13, (size_t)n_series * 4);
  if (!series || !ess_d) return fail("cudaMalloc failed (ess)");
  EssParams E;
  E.pos = pos_dev; E.series = series; E.pidx = pidx_d; E.sidx = sidx_d; E.ess = ess_d; E.n = n; E.C = c->C; E.d = c->d; E.d_sel = d_sel;
  const size_t smem = ((size_t)n + (size_t)(n - (n & 1))) * 4;
  const size_t smem_max = kSmemLimit - 1024;   // (the kernel also holds a few static words)
  if (smem > smem_max) return fail("series too long for the on-device effective sample size (more than ~28000 samples): thin them first");
  CK(cudaFuncSetAttribute(ess_series_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  dim3 tg((unsigned)((n_series + 31) / 32), (unsigned)((n + 31) / 32), 1);
  if (tg.y > 65535) return fail("too many samples for the transpose grid");
  ess_transpose_kernel<<<tg, dim3(32, 8, 1), 0, st>>>(E);
  CK(cudaGetLastError());
  size_t n_out = (size_t)n_series;
  if (pooled) {     // all chains of a parameter in one CTA: [C][n] + [n_even] floats of shared memory
    if (c->C > 64) return fail("pooled effective sample size: at most 64 chains");
    const size_t smem_p = ((size_t)c->C * n + (size_t)(n - (n & 1))) * 4;
    if (smem_p > smem_max) return fail("pooled effective sample size: chains x samples do not fit in shared memory");
    CK(cudaFuncSetAttribute(ess_pooled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    ess_pooled_kernel<<<(unsigned)d_sel, ESS_THREADS, smem_p, st>>>(E);
    n_out = (size_t)d_sel;
  } else {
    ess_series_kernel<<<(unsigned)n_series, ESS_THREADS, smem, st>>>(E);
  }
  CK(cudaGetLastError());
  c->launches += 2;
  CK(cudaMemcpyAsync(ess_host, ess_d, n_out * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int mile_ess_positions(mile_ctx* c, const float* pos_dev, int32_t n, const int32_t* param_idx, int32_t n_params_sel,
                       const int32_t* sample_idx, int32_t n_samples_sel, float* ess_host, void* stream) {
  if (!c || !pos_dev || !ess_host) return fail("null argument");
  CK(cudaSetDevice(c->device));
  return ess_run(c, pos_dev, n, param_idx, n_params_sel, sample_idx, n_samples_sel, ess_host, (cudaStream_t)stream);
}

int mile_ess_positions_host(mile_ctx* c, const float* pos, int32_t n, const int32_t* param_idx, int32_t n_params_sel,
                            const int32_t* sample_idx, int32_t n_samples_sel, float* ess_host) {
  if (!c || !pos || !ess_host) return fail("null argument");
  CK(cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * c->C * c->d * 4;
  float* pd = (float*)scratch(c, 14, bytes);
  if (!pd) return fail("cudaMalloc failed (ess positions)");
  CK(cudaMemcpyAsync(pd, pos, bytes, cudaMemcpyHostToDevice, c->own_stream));
  return ess_run(c, pd, n, param_idx, n_params_sel, sample_idx, n_samples_sel, ess_host, c->own_stream);
}

int mile_ess_pooled_host(mile_ctx* c, const float* pos, int32_t n, const int32_t* param_idx, int32_t n_params_sel,
                         const int32_t* sample_idx, int32_t n_samples_sel, float* ess_host) {
  if (!c || !pos || !ess_host) return fail("null argument");
  CK(cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * c->C * c->d * 4;
  float* pd = (float*)scratch(c, 14, bytes);
  if (!pd) return fail("cudaMalloc failed (ess positions)");
  CK(cudaMemcpyAsync(pd, pos, bytes, cudaMemcpyHostToDevice, c->own_stream));
  return ess_run(c, pd, n, param_idx, n_params_sel, sample_idx, n_samples_sel, ess_host, c->own_stream, true);
}

int mile_mclmc_phase3_ess(mile_ctx* c, int32_t n_steps, const float* step_size_host, const float* L_host, uint64_t seed,
                          const int32_t* param_idx, int32_t n_params_sel, const int32_t* sample_idx, int32_t n_samples_sel,
                          float* ess_host) {
  if (!c || !step_size_host || !L_host || !ess_host) return fail("null argument");
  if (n_steps < 4) return fail("phase 3 needs at least 4 steps");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->own_stream;
  const size_t Cb = (size_t)c->C * 4;
  float* e = (float*)scratch(c, 5, Cb); float* l = (float*)scratch(c, 6, Cb);
  float* pos = (float*)scratch(c, 14, (size_t)n_steps * c->C * c->d * 4);
  if (!e || !l || !pos) return fail("cudaMalloc failed (phase 3 positions)");
  CK(cudaMemcpyAsync(e, step_size_host, Cb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(l, L_host, Cb, cudaMemcpyHostToDevice, st));
  const int chunk = 10000;
  for (int done = 0; done < n_steps; done += chunk) {     // HOT LOOP B: every position captured in HBM
    const int n = n_steps - done < chunk ? n_steps - done : chunk;
    if (mile_mclmc_sample(c, n, done, 1, 0, e, l, nullptr, seed, pos, n_steps, nullptr, 0, st)) return -1;
  }
  const int rc = ess_run(c, pos, n_steps, param_idx, n_params_sel, sample_idx, n_samples_sel, ess_host, st);
  scratch_release(c, 14, (size_t)256 << 20);     // positions and series-major copy: kept only when small
  scratch_release(c, 12, (size_t)256 << 20);
  return rc;
}

int64_t mile_launch_count(const mile_ctx* c) { return c ? c->launches : -1; }
int mile_synchronize(mile_ctx* c) {
  if (!c) return fail("null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  return 0;
}

}  // extern "C"
