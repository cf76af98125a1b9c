/*
 * mile_b200.h -- C ABI of the B200-native MCLMC ensemble sampling path for MILE.
 *
 * The reference (zhiyuan-yang/MILE) is 100% Python on JAX/BlackJAX and has no FFI of
 * its own; the seams this library sits behind are the Python callables listed in
 * SURVEY.md section 8(b).  Each entry point below cites the reference interface
 * (file:line under /root/reference) whose work it performs.  Host code (the Python
 * shim in mile_b200/, or any other binding) talks to the library only through this
 * header: plain pointers and sizes, opaque context handle, int status return
 * (0 = OK, negative = error, text via mile_last_error()).  No exceptions cross the
 * ABI.  One context per GPU; a context is not thread-safe.
 *
 * Pointer naming: *_dev = device pointer, *_host = host pointer.  All float data is
 * fp32 (the reference computes in fp32: src/flax_building_blocks/basic.py:29).
 *
 * Flat parameter layout (theta, u, grad, noise): jax.flatten_util.ravel_pytree order,
 * i.e. for every layer: bias(out) then kernel(in,out) row-major; the offsets are
 * given explicitly in mile_model_desc so that any leaf order ('layer10' < 'layer2')
 * works.
 */
#ifndef MILE_B200_H
#define MILE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MILE_MAX_LAYERS 12

/* src/config/models/base.py:24-37 (flax.linen.<name>) */
enum mile_activation {
  MILE_ACT_IDENTITY = 0,
  MILE_ACT_RELU = 1,
  MILE_ACT_SIGMOID = 2,
  MILE_ACT_TANH = 3,
  MILE_ACT_GELU = 4, /* tanh approximation, flax default */
  MILE_ACT_LEAKY_RELU = 5
};
/* src/config/data.py Task */
enum mile_task { MILE_TASK_REGRESSION = 0, MILE_TASK_CLASSIFICATION = 1 };
/* src/training/priors.py:12-56 */
enum mile_prior { MILE_PRIOR_NORMAL = 0, MILE_PRIOR_LAPLACE = 1 };

/* What `partial(prob_model.log_unnormalized_posterior, x=train_x, y=train_y)` closes
 * over (src/training/trainer.py:576-580): FCN shape (src/models/tabular/fcn.py:11-28),
 * task, prior (src/training/probabilistic.py:20-47). */
typedef struct mile_model_desc {
  int32_t n_features;
  int32_t n_layers;                  /* len(hidden_structure), output layer included */
  int32_t widths[MILE_MAX_LAYERS];   /* hidden_structure */
  int32_t bias_off[MILE_MAX_LAYERS]; /* offset of layer l bias in the flat vector */
  int32_t kernel_off[MILE_MAX_LAYERS];
  int32_t activation;                /* enum mile_activation */
  int32_t task;                      /* enum mile_task */
  int32_t prior;                     /* enum mile_prior */
  float prior_loc;
  float prior_scale;
  float n_batches;                   /* probabilistic.py:136, == 1 for full batch */
} mile_model_desc;

/* Arguments of custom_mclmc_warmup / mclmc_find_L_and_step_size
 * (src/training/warmup.py:155-228,486-495). */
typedef struct mile_tune_cfg {
  int32_t tune1_steps;
  int32_t tune2_steps;
  float desired_energy_var_start;
  float desired_energy_var_end;
  float trust_in_estimate;
  float num_effective_samples;
} mile_tune_cfg;

/* Arguments of the NUTS branch: blackjax.nuts defaults (max_num_doublings 10, divergence_threshold 1000) and
 * custom_window_adaptation's (src/training/warmup.py:27-36: initial_step_size 1.0, target_acceptance_rate 0.80). */
typedef struct mile_nuts_cfg {
  int32_t max_num_doublings;
  float divergence_threshold;
  float target_acceptance_rate;
  float initial_step_size;
} mile_nuts_cfg;

typedef struct mile_ctx mile_ctx;

const char* mile_last_error(void);
int mile_version(void);

/* ---- lifetime ------------------------------------------------------------------ */
/* One context = one ensemble wave on one GPU (reference: one `inference_loop` call,
 * src/training/sampling.py:32-40, over `len(step_ids)` chains). */
int mile_create(const mile_model_desc* desc, int32_t n_chains, int32_t device, mile_ctx** out);
void mile_destroy(mile_ctx* ctx);
int32_t mile_n_params(const mile_ctx* ctx);
/* Execution knobs: "cluster_size" CTAs per chain of the persistent kernel (0 = auto), "refresh_mode" (0 = single
 * post-step refresh, blackjax 1.2.2; 1 = half-step refreshes around the integrator, later blackjax
 * `with_isokinetic_maruyama`), "chain_base" (global id of this context's chain 0: the in-kernel Philox streams are keyed by
 * chain_base + local chain, so the ranks of a chain-partitioned ensemble draw independent noise, like the reference's
 * jax.random.split per device, sampling.py:181-184), "fast" (narrow-MLP evaluator: 2 = 3xTF32 register MMA, 1 = FFMA layer
 * pipeline, 0 = generic tiles), "tensor" (wide path GEMM core), "sync_mode", "resident", "tile_rows", "steploop". */
int mile_set_option(mile_ctx* ctx, const char* key, int64_t value);
int64_t mile_get_option(const mile_ctx* ctx, const char* key);

/* ---- data (closure constants of the log-posterior, trainer.py:576-580) ---------- */
/* X [N,F] row-major fp32; y fp32[N] (regression) or int32[N] (classification).
 * Copies into the library's padded HBM layout; the caller keeps ownership. */
int mile_set_data(mile_ctx* ctx, const float* X_dev, const void* y_dev, int64_t n_rows, void* stream);
int mile_set_data_host(mile_ctx* ctx, const float* X_host, const void* y_host, int64_t n_rows);
/* Test split used by the fused posterior-predictive LPPD (src/inference/evaluation.py:378-400). */
int mile_set_test(mile_ctx* ctx, const float* X_dev, const void* y_dev, int64_t n_rows, void* stream);
int mile_set_test_host(mile_ctx* ctx, const float* X_host, const void* y_host, int64_t n_rows);

/* ---- a1: value_and_grad of log_unnormalized_posterior --------------------------- */
/* src/training/probabilistic.py:115-138 differentiated as blackjax does
 * (jax.value_and_grad).  theta [n,d] -> lp [n], grad [n,d]; n <= n_chains. */
int mile_logpost_value_and_grad(mile_ctx* ctx, const float* theta_dev, int32_t n,
                                float* lp_dev, float* grad_dev, void* stream);
int mile_logpost_value_and_grad_host(mile_ctx* ctx, const float* theta_host, int32_t n,
                                     float* lp_host, float* grad_host);

/* ---- a14: blackjax.mcmc.mclmc.init (call site warmup.py:539-541) ---------------- */
/* theta0 [C,d]; z0 [C,d] normal draws for the initial unit momentum, or NULL to draw
 * them on the device from `seed` (Philox4x32-10). */
int mile_mclmc_init(mile_ctx* ctx, const float* theta0_dev, const float* z0_dev, uint64_t seed,
                    void* stream);
int mile_mclmc_init_host(mile_ctx* ctx, const float* theta0_host, const float* z0_host, uint64_t seed);
/* Overwrite / read the full chain state (position, momentum, logdensity, logdensity_grad)
 * = blackjax IntegratorState, [C,d],[C,d],[C],[C,d].  NULL pointers are skipped. */
int mile_set_state_host(mile_ctx* ctx, const float* theta, const float* u, const float* lp,
                        const float* grad);
int mile_get_state_host(mile_ctx* ctx, float* theta, float* u, float* lp, float* grad);
int mile_get_state(mile_ctx* ctx, float* theta_dev, float* u_dev, float* lp_dev, float* grad_dev,
                   void* stream);

/* ---- a6/a14: sampler.step scan (sampling.py:134-177; blackjax mclmc kernel) ------ */
/* Runs n_steps MCLMC steps for every chain with fixed per-chain step_size / L.
 *   z_dev        [n_steps,C,d] host-supplied normal draws (parity mode) or NULL
 *                (Philox from `seed`; draws depend on (seed, chain, step_base+i)).
 *   step_base    index of the first step (the `idx` of sampling.py:150); position i is
 *                kept when (step_base+i) % n_thinning == 0 (sampling.py:162), written to
 *                samples_dev[((step_base+i)/n_thinning - sample_base) , c, :].
 *   samples_dev  [n_slots,C,d] or NULL.   info_dev [n_steps,C,3] = MCLMCInfo
 *                (logdensity, kinetic_change, energy_change) or NULL.
 *   lppd         non-zero: fold every kept position into the online test-set
 *                logsumexp state (needs mile_set_test).
 */
int mile_mclmc_sample(mile_ctx* ctx, int32_t n_steps, int64_t step_base, int32_t n_thinning,
                      int64_t sample_base, const float* step_size_dev, const float* L_dev,
                      const float* z_dev, uint64_t seed, float* samples_dev, int64_t n_slots,
                      float* info_dev, int32_t lppd, void* stream);
/* Same with host buffers: uploads step_size/L/z, downloads samples/info (the e2e path). */
int mile_mclmc_sample_host(mile_ctx* ctx, int32_t n_steps, int64_t step_base, int32_t n_thinning,
                           const float* step_size_host, const float* L_host, const float* z_host,
                           uint64_t seed, float* samples_host, int64_t n_slots, float* info_host,
                           int32_t lppd);

/* ---- a9-a11: make_L_step_size_adaptation (warmup.py:231-405) -------------------- */
/* Resets the adaptive state: L = max(sqrt(d),15), step_size = step_size_init,
 * (time, x_average, step_size_max) = (0,0,inf), streaming averages 0 (warmup.py:204-209,358-363). */
int mile_tune_reset(mile_ctx* ctx, float step_size_init, void* stream);
/* n_steps iterations of HOT LOOP A starting at iteration `step_base` (phase 1 while
 * step < tune1_steps, then phase 2): kernel step, handle_nans, step-size predictor,
 * streaming mean of (x, x^2).  tune_info_dev [n_steps,C,4] = (energy_change after
 * handle_nans, step_size after update, step_size_max, success) or NULL. */
int mile_mclmc_tune(mile_ctx* ctx, int32_t n_steps, int64_t step_base, const mile_tune_cfg* cfg,
                    const float* z_dev, uint64_t seed, float* tune_info_dev, void* stream);
int mile_mclmc_tune_host(mile_ctx* ctx, int32_t n_steps, int64_t step_base, const mile_tune_cfg* cfg,
                         const float* z_host, uint64_t seed, float* tune_info_host);
/* L = sqrt(sum(E[x^2]-E[x]^2)) (warmup.py:383-390); writes it into the tuning state. */
int mile_tune_finish_phase2(mile_ctx* ctx, void* stream);
/* step_size [C], L [C], step_size_max [C], mean_x [C,d], mean_x2 [C,d]; NULLs skipped. */
int mile_get_tuning_host(mile_ctx* ctx, float* step_size, float* L, float* step_size_max,
                         float* mean_x, float* mean_x2);
int mile_set_tuning_host(mile_ctx* ctx, const float* step_size, const float* L);
/* Device views of the tuned parameters (valid until mile_destroy). */
int mile_tuning_ptrs(mile_ctx* ctx, float** step_size_dev, float** L_dev);

/* ---- a15: posterior-predictive LPPD (metrics.py:247-312, evaluation.py:378-400) --- */
int mile_lppd_reset(mile_ctx* ctx, void* stream);
/* Fold theta [n,d] (n <= C; row c belongs to chain c) into the online state. */
int mile_lppd_accumulate(mile_ctx* ctx, const float* theta_dev, int32_t n, void* stream);
/* Running max m [C,Nt], scaled sum s [C,Nt], number of folded samples per chain. */
int mile_lppd_state_host(mile_ctx* ctx, float* m, float* s, int64_t* count);
/* Forward pass only: theta [n,d] on the test (which=1) or train (which=0) split -> out [n,N,K]. */
int mile_predict(mile_ctx* ctx, const float* theta_dev, int32_t n, int32_t which, float* out_dev,
                 void* stream);

/* ---- deep-ensemble warm-start training ---------------------------------------------------- */
/* src/config/warmstart.py:17-41 (OptimizerConfig -> optax.adamw / adam / sgd with its `parameters`). */
enum mile_optimizer { MILE_OPT_KIND_ADAMW = 0, MILE_OPT_KIND_ADAM = 1, MILE_OPT_KIND_SGD = 2 };
typedef struct mile_opt_cfg {
  int32_t kind;          /* enum mile_optimizer */
  float learning_rate;
  float b1, b2, eps;     /* optax defaults 0.9, 0.999, 1e-8 */
  float weight_decay;    /* optax.adamw default 1e-4 */
} mile_opt_cfg;
/* get_initial_state (src/training/trainer.py:870-892) with host-initialised parameters: theta0 [C,d] (one row per ensemble
 * member), optimizer moments and step counts zeroed. */
int mile_train_init(mile_ctx* ctx, const float* theta0_dev, void* stream);
/* One epoch of train_de_member's inner loop (trainer.py:441-460) for all members in ONE launch: for every minibatch
 * (rows batch_idx[b, 0..B) of the training split set with mile_set_data; the same batches for every member) the mean
 * Gaussian-NLL / cross-entropy loss, its gradient and the optimizer update (single_step_regr / single_step_class,
 * trainer.py:662-760).  stopped_dev [C] (or NULL): members whose early-stopping flag is set skip the epoch.
 * metrics_dev [n_batches, C, 2] (or NULL) = (loss, RMSE | accuracy) of every step before its update, NaN when stopped. */
int mile_train_epoch(mile_ctx* ctx, const int32_t* batch_idx_dev, int32_t n_batches, int32_t batch_size,
                     const mile_opt_cfg* opt, const uint8_t* stopped_dev, float* metrics_dev, void* stream);
/* predict_regr / predict_class (trainer.py:763-868): mean loss and RMSE | accuracy of theta [n,d] (NULL: the training
 * state of all members) over the train (which = 0) or the mile_set_test (which = 1) split -> out [n,2]. */
int mile_eval_metrics(mile_ctx* ctx, const float* theta_dev, int32_t n, int32_t which, float* out_dev, void* stream);
/* Parameters, AdamW moments [C,d] and step counts [C] (device buffers; NULLs skipped). */
int mile_train_get_state(mile_ctx* ctx, float* theta_dev, float* m_dev, float* v_dev, int32_t* t_dev, void* stream);

/* ---- data-sharded variant (covertype, rows split across the GPUs of one box) -------- */
/* ---- sample files: src/training/callbacks.py:17-44 (`save_position`: one compressed npz per chain and kept position).
 * Writes n_files archives with the same n_members members each: member m of file i = header bytes (the .npy header of
 * its shape) followed by member_floats[m] fp32 values taken from data[i][offset of m] (members back to back in leaf
 * order, i.e. a flat position vector when the leaves are in ravel order).  Deflated by n_threads host threads; plain zip
 * archives that np.load / load_samples_from_dir (utils.py:131-161) read unchanged.  Host memory only; 0 on success. */
int mile_write_npz_batch(const char* const* paths, int32_t n_files, const char* const* member_names,
                         const uint8_t* const* member_headers, const int32_t* header_lens, const int64_t* member_floats,
                         int32_t n_members, const float* data /* [n_files][sum member_floats] host */, int32_t n_threads);

/* ---- partition sampling: src/training/partition_sampling.py:32-330 with trainer.py:613-659
 * (`log_unnormalized_posterior_partition`): only the first and the last layer are sampled, the hidden layers stay at their
 * warm-start values.  frozen[i] != 0 freezes flat parameter i for ALL chains: it keeps its value, contributes no prior
 * term, and gets zero gradient / momentum / noise; the dimension of the dynamics (ESH normalisation d - 1, refresh,
 * tuning, L_0) becomes the number of sampled parameters.  NULL clears the mask.  Generic step loop; rejected on the wide path. */
int mile_set_frozen_mask_host(mile_ctx* ctx, const uint8_t* frozen /* [d] host, or NULL */);

/* ---- diagonal preconditioning: src/training/warmup.py:385-401 (`diagonal_preconditioning=True`, the reference's default,
 * off in every MCLMC YAML) and blackjax's `sqrt_diag_cov` argument of mclmc.build_kernel / isokinetic_mclachlan.
 * With a preconditioner m set, the B-steps see the scaled gradient m .* g and the A-steps move by eps * m .* u in every
 * following mile_mclmc_tune / mile_mclmc_sample / mile_shard_* call (generic step loop; gradients still come from the
 * fastest evaluator).  mile_precondition_from_moments: m = sqrt(E[x^2] - E[x]^2) from the streaming moments of tuning
 * phase 2 and L = sqrt(d) (warmup.py:388-394).  NULL clears it.  The reference DROPS m after warmup (sampling.py:291:
 * only step_size and L are returned), so its sampling phase runs unpreconditioned; the mirror does the same. */
int mile_precondition_from_moments(mile_ctx* ctx, void* stream);
int mile_set_sqrt_diag_cov_host(mile_ctx* ctx, const float* sqrt_diag_cov /* [C,d] host, or NULL = identity */);
int mile_get_sqrt_diag_cov_host(mile_ctx* ctx, float* out /* [C,d] host */);

/* Every rank holds ALL chains and 1/world of the training rows (mile_set_data with the local shard).  Each
 * gradient evaluation = local value_and_grad (prior weighted 1/world) + ncclAllReduce(sum) of the packed
 * [C, d+1] (gradient | log-density) buffer + an integrator-only kernel; all ranks apply identical updates, so
 * the chain state stays replicated bit-identically.  The reference has no counterpart (it replicates the data on
 * every virtual device, src/training/sampling.py:181-184); the arithmetic per step is the same as
 * mile_mclmc_sample / mile_mclmc_tune.  NCCL is resolved with dlopen at run time. */
int mile_nccl_unique_id(void* out128);   /* rank 0: 128-byte ncclUniqueId to broadcast to the other ranks */
int mile_shard_init(mile_ctx* ctx, const void* unique_id128, int32_t rank, int32_t world);
/* Optional peer-memory all-reduce (one box, 2..8 ranks, NVLink): every rank exports the CUDA-IPC handle of its exchange
 * region (64 bytes), the host all-gathers them, and after mile_shard_p2p_open the step loop's integrator kernel sums the
 * ranks' partial [C, d+1] buffers straight out of peer memory (flag + data, rank order) instead of waiting for an
 * ncclAllReduce: compute step and collective in one kernel.  mile_get_option("p2p") reports whether it is active. */
int mile_shard_p2p_handle(mile_ctx* ctx, void* out64);
int mile_shard_p2p_open(mile_ctx* ctx, const void* handles64 /* [world][64] */);
int mile_shard_mclmc_init(mile_ctx* ctx, const float* theta0_dev, const float* z0_dev, uint64_t seed, void* stream);
int mile_shard_mclmc_sample(mile_ctx* ctx, int32_t n_steps, int64_t step_base, int32_t n_thinning,
                            int64_t sample_base, const float* step_size_dev, const float* L_dev,
                            const float* z_dev, uint64_t seed, float* samples_dev, int64_t n_slots,
                            float* info_dev, void* stream);
int mile_shard_mclmc_tune(mile_ctx* ctx, int32_t n_steps, int64_t step_base, const mile_tune_cfg* cfg,
                          const float* z_dev, uint64_t seed, float* tune_info_dev, void* stream);

/* ---- a12: phase 3 of the warmup on the device: src/training/warmup.py:408-465 (`make_adaptation_L`): n_steps sampling
 * steps with every position kept in HBM, then blackjax.diagnostics.effective_sample_size of every (chain, parameter)
 * series (one chain per series, as the reference calls it: `flat_samples[None, ...]`), by hand-written kernels: a tiled
 * transpose to series-major order and one CTA per series (lazy direct autocovariance in shared memory + Geyer's initial
 * positive / monotone sequence).  param_idx [n_params_sel] / sample_idx [n_samples_sel] (host, or NULL = all) are the
 * reference's subsampling rules (> 2000 parameters: random subset; > 10000 samples: linspace).  ess_host [C, n_sel].
 * The caller finishes with L = 0.4 * step_size * mean(n_steps / ess) (warmup.py:461-463). */
int mile_mclmc_phase3_ess(mile_ctx* ctx, int32_t n_steps, const float* step_size_host, const float* L_host, uint64_t seed,
                          const int32_t* param_idx, int32_t n_params_sel, const int32_t* sample_idx,
                          int32_t n_samples_sel, float* ess_host);
/* the same estimator on positions the caller holds: pos [n, C, d] on the device / on the host */
int mile_ess_positions(mile_ctx* ctx, const float* pos_dev, int32_t n, const int32_t* param_idx, int32_t n_params_sel,
                       const int32_t* sample_idx, int32_t n_samples_sel, float* ess_host, void* stream);
int mile_ess_positions_host(mile_ctx* ctx, const float* pos, int32_t n, const int32_t* param_idx, int32_t n_params_sel,
                            const int32_t* sample_idx, int32_t n_samples_sel, float* ess_host);
/* all chains pooled per parameter -- the report's ESS (src/inference/metrics.py:354-425 -> effective_sample_size on
 * [chains, samples, dim]: autocovariance averaged over the chains, between-chain variance of the chain means in the
 * normalisation, chains x samples in the numerator).  pos [n, C, d] on the host; ess_host [n_selected parameters]. */
int mile_ess_pooled_host(mile_ctx* ctx, const float* pos, int32_t n, const int32_t* param_idx, int32_t n_params_sel,
                         const int32_t* sample_idx, int32_t n_samples_sel, float* ess_host);

/* ---- NUTS branch of the sampling seam: src/training/sampling.py:70-81,107-210 (sampler = blackjax.nuts) and
 * src/training/warmup.py:27-152 (`custom_window_adaptation`), called from `warmup_nuts` (sampling.py:220-262).  One
 * persistent kernel per call runs whole transitions (momentum draw, trajectory doubling with the iterative U-turn
 * checkpoints, progressive sampling, divergence test) for every chain; the warm-up call also applies `adapt_step`
 * (dual averaging of the step size; Welford mass matrix in the slow windows) after every transition.
 * schedule [n_steps] bytes: bit 0 = slow stage, bit 1 = end of a slow window (window_adaptation.build_schedule).
 * z [n_steps,C,d] standard normals and uni [n_steps,C,2 D + 2^D] uniforms (directions | merge acceptances | per-leapfrog
 * acceptances) replace the in-kernel Philox streams when given (trajectory-level tests against the oracle).
 * info [n_steps,C,8] = num_integration_steps, acceptance_rate, num_trajectory_expansions, is_divergent, energy, is_turning
 * (the NUTSInfo fields sampling.py:200-210 keeps), logdensity, step size used. */
int mile_nuts_init(mile_ctx* ctx, const float* theta0_dev, const mile_nuts_cfg* cfg, void* stream);
int mile_nuts_init_host(mile_ctx* ctx, const float* theta0, const mile_nuts_cfg* cfg);
/* positions_dev [n_steps,C,d] or NULL: the position BEFORE each warm-up transition (what warmup.py:102-109 writes under
 * saving_path when keep_warmup is set) */
int mile_nuts_warmup(mile_ctx* ctx, int32_t n_steps, int64_t step_base, const uint8_t* schedule_dev, const float* z_dev,
                     const float* uni_dev, uint64_t seed, float* positions_dev, float* info_dev, void* stream);
/* adapt_final: step_size = exp(averaged log step size) */
int mile_nuts_finish_warmup(mile_ctx* ctx, void* stream);
int mile_nuts_sample(mile_ctx* ctx, int32_t n_steps, int64_t step_base, int32_t n_thinning, int64_t sample_base,
                     const float* z_dev, const float* uni_dev, uint64_t seed, float* samples_dev, int64_t n_slots,
                     float* info_dev, int32_t lppd, void* stream);
/* host-buffer form of both: schedule != NULL = warm-up transitions (samples = the pre-transition positions, n_slots =
 * n_steps), NULL = sampling transitions */
int mile_nuts_run_host(mile_ctx* ctx, int32_t n_steps, int64_t step_base, const uint8_t* schedule, int32_t n_thinning,
                       const float* z, const float* uni, uint64_t seed, float* samples, int64_t n_slots, float* info,
                       int32_t lppd);
/* step_size [C], inverse_mass_matrix [C,d] (the `parameters` dict warmup_nuts returns); NULLs skipped */
int mile_nuts_get_params_host(mile_ctx* ctx, float* step_size, float* inverse_mass_matrix);
int mile_nuts_set_params_host(mile_ctx* ctx, const float* step_size, const float* inverse_mass_matrix);

/* ---- bookkeeping ----------------------------------------------------------------- */
/* Number of kernels this library has launched since mile_create (bench `gpu_launches`). */
int64_t mile_launch_count(const mile_ctx* ctx);
int mile_synchronize(mile_ctx* ctx);
/* Measured FP32 CUDA-core peak in TFLOP/s (roofline denominator for the narrow-MLP configs,
 * SURVEY.md section 8d): variant 0 = scalar FFMA, 1 = packed fma.rn.f32x2 (FFMA2), 2 = the register-operand tensor
 * instruction of the narrow-MLP evaluator, mma.sync.m16n8k8 tf32 (TFLOP/s of tf32 products; the evaluator issues three per
 * fp32 product). */
int mile_measure_fp32_peak(int32_t device, int32_t variant, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* MILE_B200_H */
