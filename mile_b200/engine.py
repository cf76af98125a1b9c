"""Thin object wrapper around the C ABI: one `Ensemble` = one `mile_ctx` = one ensemble wave
on one GPU.  Host arrays are numpy; device arrays are torch CUDA tensors (torch is only the
allocator / stream owner here, all compute happens in libmile_b200.so)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass(frozen=True)
class FCNSpec:
    """What `partial(prob_model.log_unnormalized_posterior, x=..., y=...)` closes over
    (src/training/trainer.py:576-580): FCN shape (src/models/tabular/fcn.py:11-28, the
    hidden_structure INCLUDES the output width), task, prior."""

    n_features: int
    widths: tuple
    activation: str = 'relu'
    task: str = 'regr'
    prior: str = 'normal'
    prior_loc: float = 0.0
    prior_scale: float = 1.0
    n_batches: float = 1.0
    layer_order: tuple = field(default=None)

    def __post_init__(self):
        object.__setattr__(self, 'widths', tuple(int(w) for w in self.widths))
        if self.layer_order is None:
            # ravel_pytree flattens dicts in sorted key order ('layer10' < 'layer2')
            object.__setattr__(self, 'layer_order',
                               tuple(sorted(range(len(self.widths)), key=lambda i: f'layer{i}')))
        if self.activation not in capi.ACTIVATIONS:
            raise NotImplementedError(f'activation {self.activation!r} is not supported by the CUDA path')
        if len(self.widths) > capi.MILE_MAX_LAYERS:
            raise NotImplementedError(f'at most {capi.MILE_MAX_LAYERS} layers')

    @property
    def dims(self):
        return (self.n_features,) + self.widths

    @property
    def n_params(self):
        d = self.dims
        return sum(d[i] * d[i + 1] + d[i + 1] for i in range(len(self.widths)))

    def offsets(self):
        d = self.dims
        nl = len(self.widths)
        bias_off, kern_off, off = [0] * nl, [0] * nl, 0
        for l in self.layer_order:
            bias_off[l] = off
            off += d[l + 1]
            kern_off[l] = off
            off += d[l] * d[l + 1]
        return bias_off, kern_off

    def hidden_layer_mask(self) -> np.ndarray:
        """bool [d]: True for the parameters of the hidden layers (everything except layer0 and the last layer) -- the
        ones `partition_params` (partition_sampling.py:290-302) freezes."""
        b, k = self.offsets()
        dims = self.dims
        m = np.zeros(self.n_params, bool)
        for l in range(1, len(self.widths) - 1):
            m[b[l]:b[l] + dims[l + 1]] = True
            m[k[l]:k[l] + dims[l] * dims[l + 1]] = True
        return m

    def to_desc(self) -> capi.ModelDesc:
        m = capi.ModelDesc()
        m.n_features = self.n_features
        m.n_layers = len(self.widths)
        b, k = self.offsets()
        for l, w in enumerate(self.widths):
            m.widths[l], m.bias_off[l], m.kernel_off[l] = w, b[l], k[l]
        m.activation = capi.ACTIVATIONS[self.activation]
        m.task = capi.TASKS[self.task]
        m.prior = capi.PRIORS[self.prior.lower()]
        m.prior_loc, m.prior_scale, m.n_batches = self.prior_loc, self.prior_scale, self.n_batches
        return m

    # --- pytree <-> flat (ravel_pytree order) ----------------------------------------
    def ravel(self, tree: dict) -> np.ndarray:
        """{'fcn': {'layer{i}': {'bias','kernel'}}} with optional leading batch axes -> [..., d]."""
        inner = tree['fcn'] if 'fcn' in tree else tree
        d = self.dims
        parts = []
        for l in self.layer_order:
            lay = inner[f'layer{l}']
            b = np.asarray(lay['bias'], dtype=np.float32)
            k = np.asarray(lay['kernel'], dtype=np.float32)
            lead = b.shape[:-1]
            parts.append(b.reshape(lead + (d[l + 1],)))
            parts.append(k.reshape(lead + (d[l] * d[l + 1],)))
        return np.ascontiguousarray(np.concatenate(parts, axis=-1))

    def unravel(self, theta: np.ndarray) -> dict:
        b_off, k_off = self.offsets()
        d = self.dims
        lead = theta.shape[:-1]
        out = {}
        for l in range(len(self.widths)):
            out[f'layer{l}'] = {
                'bias': np.array(theta[..., b_off[l]:b_off[l] + d[l + 1]]),
                'kernel': np.array(theta[..., k_off[l]:k_off[l] + d[l] * d[l + 1]]).reshape(lead + (d[l], d[l + 1])),
            }
        return {'fcn': out}


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _dev_ptr(t):
    """torch CUDA tensor (contiguous) -> void*; None -> NULL."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'device tensors must be contiguous CUDA tensors'
    return C.c_void_p(t.data_ptr())


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Ensemble:
    """One ensemble wave of `n_chains` chains on one GPU."""

    def __init__(self, spec: FCNSpec, n_chains: int, device: int = 0, **options):
        self.spec, self.n_chains, self.device = spec, int(n_chains), int(device)
        self.lib = capi.load()
        h = C.c_void_p()
        desc = spec.to_desc()
        capi.check(self.lib.mile_create(C.byref(desc), self.n_chains, self.device, C.byref(h)))
        self.h = h
        self.d = self.lib.mile_n_params(self.h)
        assert self.d == spec.n_params
        self.n_train = self.n_test = 0
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, 'h', None):
            self._pre_close()
            self.lib.mile_destroy(self.h)
            self.h = None

    def _pre_close(self):
        pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        capi.check(self.lib.mile_set_option(self.h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        return int(self.lib.mile_get_option(self.h, key.encode()))

    @property
    def launches(self) -> int:
        return int(self.lib.mile_launch_count(self.h))

    def synchronize(self):
        capi.check(self.lib.mile_synchronize(self.h))

    # ---- data ---------------------------------------------------------------------------
    def _y(self, y):
        return np.ascontiguousarray(y, dtype=np.int32 if self.spec.task.startswith('class') else np.float32)

    def set_data(self, X, y):
        X, y = _f32(X), self._y(y)
        assert X.ndim == 2 and X.shape[1] == self.spec.n_features and y.shape == (X.shape[0],)
        capi.check(self.lib.mile_set_data_host(self.h, capi.host_ptr(X), capi.host_ptr(y), X.shape[0]))
        self.n_train = X.shape[0]

    def set_test(self, X, y):
        X, y = _f32(X), self._y(y)
        assert X.ndim == 2 and X.shape[1] == self.spec.n_features and y.shape == (X.shape[0],)
        capi.check(self.lib.mile_set_test_host(self.h, capi.host_ptr(X), capi.host_ptr(y), X.shape[0]))
        self.n_test = X.shape[0]

    # ---- a1: value_and_grad -------------------------------------------------------------
    def value_and_grad(self, theta):
        theta = _f32(theta).reshape(-1, self.d)
        n = theta.shape[0]
        lp = np.empty(n, np.float32)
        g = np.empty((n, self.d), np.float32)
        capi.check(self.lib.mile_logpost_value_and_grad_host(self.h, capi.host_ptr(theta), n, capi.host_ptr(lp),
                                                             capi.host_ptr(g)))
        return lp, g

    # ---- state --------------------------------------------------------------------------
    def init(self, theta0, z0=None, seed: int = 0):
        theta0 = _f32(theta0).reshape(self.n_chains, self.d)
        z0 = None if z0 is None else _f32(z0).reshape(self.n_chains, self.d)
        capi.check(self.lib.mile_mclmc_init_host(self.h, capi.host_ptr(theta0), capi.host_ptr(z0), seed))

    def get_state(self, out=None):
        """(theta [C,d], u [C,d], logdensity [C], grad [C,d]) as host arrays; `out` = caller-owned float32 arrays of these
        shapes (e.g. pinned) to be filled instead of fresh ones."""
        C_, d = self.n_chains, self.d
        if out is not None:
            th, u, lp, g = out
            assert all(a.dtype == np.float32 and a.flags['C_CONTIGUOUS'] for a in out)
            assert th.shape == u.shape == g.shape == (C_, d) and lp.shape == (C_,)
        else:
            th, u, g = (np.empty((C_, d), np.float32) for _ in range(3))
            lp = np.empty(C_, np.float32)
        capi.check(self.lib.mile_get_state_host(self.h, capi.host_ptr(th), capi.host_ptr(u), capi.host_ptr(lp),
                                                capi.host_ptr(g)))
        return th, u, lp, g

    def set_state(self, theta=None, u=None, lp=None, grad=None):
        capi.check(self.lib.mile_set_state_host(self.h, capi.host_ptr(_f32(theta)), capi.host_ptr(_f32(u)),
                                                capi.host_ptr(_f32(lp)), capi.host_ptr(_f32(grad))))

    # ---- sampling -----------------------------------------------------------------------
    def sample(self, n_steps, step_size, L, *, step_base=0, n_thinning=1, z=None, seed=0, keep=True, info=False,
               lppd=False, out=None):
        """Host-buffer sampling call (the e2e path): returns (samples [S,C,d] | None, info [n,C,3] | None).
        `out`: optional caller-owned float32 host array of at least S*C*d elements for the kept samples (e.g. the numpy
        view of a pinned torch tensor: the device-to-host copy then runs at PCIe speed instead of through a staging
        buffer); the returned samples are a view of it."""
        C_, d = self.n_chains, self.d
        eps = _f32(np.broadcast_to(step_size, (C_,)))
        Ls = _f32(np.broadcast_to(L, (C_,)))
        first = -(-step_base // n_thinning)
        last = (step_base + n_steps - 1) // n_thinning
        n_slots = max(0, last - first + 1) if keep and n_steps > 0 else 0
        if keep and out is not None:
            assert out.dtype == np.float32 and out.flags['C_CONTIGUOUS'] and out.size >= n_slots * C_ * d
            samples = out.reshape(-1)[:n_slots * C_ * d].reshape(n_slots, C_, d)
        else:
            samples = np.empty((n_slots, C_, d), np.float32) if keep else None
        inf = np.empty((n_steps, C_, 3), np.float32) if info else None
        z = None if z is None else _f32(z)
        capi.check(self.lib.mile_mclmc_sample_host(self.h, n_steps, step_base, n_thinning, capi.host_ptr(eps),
                                                   capi.host_ptr(Ls), capi.host_ptr(z), seed, capi.host_ptr(samples),
                                                   n_slots, capi.host_ptr(inf), int(lppd)))
        return samples, inf

    def sample_device(self, n_steps, step_size_dev, L_dev, *, step_base=0, n_thinning=1, sample_base=0, z_dev=None,
                      seed=0, samples_dev=None, n_slots=0, info_dev=None, lppd=False):
        """Device-buffer sampling call on torch's current stream (no host transfer, asynchronous)."""
        capi.check(self.lib.mile_mclmc_sample(self.h, n_steps, step_base, n_thinning, sample_base,
                                              _dev_ptr(step_size_dev), _dev_ptr(L_dev), _dev_ptr(z_dev), seed,
                                              _dev_ptr(samples_dev), n_slots, _dev_ptr(info_dev), int(lppd),
                                              _stream_ptr()))

    # ---- tuning -------------------------------------------------------------------------
    @staticmethod
    def tune_cfg(tune1, tune2, desired_energy_var_start, desired_energy_var_end, trust_in_estimate,
                 num_effective_samples) -> capi.TuneCfg:
        return capi.TuneCfg(int(tune1), int(tune2), desired_energy_var_start, desired_energy_var_end,
                            trust_in_estimate, float(num_effective_samples))

    def tune_reset(self, step_size_init: float):
        capi.check(self.lib.mile_tune_reset(self.h, step_size_init, None))
        self.synchronize()

    def tune(self, n_steps, step_base, cfg: capi.TuneCfg, z=None, seed=0, info=False):
        inf = np.empty((n_steps, self.n_chains, 4), np.float32) if info else None
        z = None if z is None else _f32(z)
        capi.check(self.lib.mile_mclmc_tune_host(self.h, n_steps, step_base, C.byref(cfg), capi.host_ptr(z), seed,
                                                 capi.host_ptr(inf)))
        return inf

    def tune_finish_phase2(self):
        capi.check(self.lib.mile_tune_finish_phase2(self.h, None))
        self.synchronize()

    def get_tuning(self, moments=False):
        C_, d = self.n_chains, self.d
        eps, L, emax = (np.empty(C_, np.float32) for _ in range(3))
        mx = np.empty((C_, d), np.float32) if moments else None
        mx2 = np.empty((C_, d), np.float32) if moments else None
        capi.check(self.lib.mile_get_tuning_host(self.h, capi.host_ptr(eps), capi.host_ptr(L), capi.host_ptr(emax),
                                                 capi.host_ptr(mx), capi.host_ptr(mx2)))
        return (eps, L, emax, mx, mx2) if moments else (eps, L, emax)

    def set_tuning(self, step_size=None, L=None):
        C_ = self.n_chains
        e = None if step_size is None else _f32(np.broadcast_to(step_size, (C_,)))
        l = None if L is None else _f32(np.broadcast_to(L, (C_,)))
        capi.check(self.lib.mile_set_tuning_host(self.h, capi.host_ptr(e), capi.host_ptr(l)))

    # ---- phase 3 of the warmup: capture + effective sample size on the device (warmup.py:408-465) --------
    @staticmethod
    def _idx(a):
        return None if a is None else np.ascontiguousarray(a, dtype=np.int32)

    def phase3_ess(self, n_steps, step_size, L, *, seed=0, param_idx=None, sample_idx=None):
        """n_steps sampling steps from the current state with every position kept on the device, then the effective sample
        size of every (chain, selected parameter) series -> [C, n_selected]."""
        C_ = self.n_chains
        eps = _f32(np.broadcast_to(step_size, (C_,)))
        Ls = _f32(np.broadcast_to(L, (C_,)))
        pi, si = self._idx(param_idx), self._idx(sample_idx)
        out = np.empty((C_, self.d if pi is None else pi.size), np.float32)
        capi.check(self.lib.mile_mclmc_phase3_ess(self.h, int(n_steps), capi.host_ptr(eps), capi.host_ptr(Ls), seed,
                                                  capi.host_ptr(pi), 0 if pi is None else pi.size,
                                                  capi.host_ptr(si), 0 if si is None else si.size, capi.host_ptr(out)))
        return out

    def ess_positions(self, positions, *, param_idx=None, sample_idx=None, pooled=False):
        """The same estimator on host positions [n, C, d] (blackjax.diagnostics.effective_sample_size): one chain per series
        -> [C, n_selected]; pooled=True: all chains of a parameter together, as the report computes it -> [n_selected]."""
        pos = _f32(positions)
        assert pos.ndim == 3 and pos.shape[1:] == (self.n_chains, self.d)
        pi, si = self._idx(param_idx), self._idx(sample_idx)
        nsel = self.d if pi is None else pi.size
        out = np.empty(nsel if pooled else (self.n_chains, nsel), np.float32)
        fn = self.lib.mile_ess_pooled_host if pooled else self.lib.mile_ess_positions_host
        capi.check(fn(self.h, capi.host_ptr(pos), pos.shape[0], capi.host_ptr(pi), 0 if pi is None else pi.size,
                      capi.host_ptr(si), 0 if si is None else si.size, capi.host_ptr(out)))
        return out

    # ---- NUTS branch (sampling.py:70-81,107-210; warmup.py:27-152) ------------------------------
    NUTS_INFO_FIELDS = ('num_integration_steps', 'acceptance_rate', 'num_trajectory_expansions', 'is_divergent', 'energy',
                        'is_turning', 'logdensity', 'step_size')

    def nuts_init(self, theta0, max_num_doublings: int = 10, divergence_threshold: float = 1000.0,
                  target_acceptance_rate: float = 0.8, initial_step_size: float = 1.0):
        """hmc.init (logdensity + gradient at theta0) and window_adaptation's init (unit metric, dual averaging at
        initial_step_size)."""
        theta0 = _f32(theta0).reshape(self.n_chains, self.d)
        cfg = capi.NutsCfg(int(max_num_doublings), divergence_threshold, target_acceptance_rate, initial_step_size)
        self._nuts_D = int(max_num_doublings)
        capi.check(self.lib.mile_nuts_init_host(self.h, capi.host_ptr(theta0), C.byref(cfg)))

    def nuts_uni_len(self) -> int:
        D = getattr(self, '_nuts_D', 10)
        return 2 * D + 2 ** D

    def _nuts_run(self, n_steps, step_base, schedule, n_thinning, z, uni, seed, keep, info, lppd):
        C_, d = self.n_chains, self.d
        first = -(-step_base // n_thinning)
        last = (step_base + n_steps - 1) // n_thinning
        n_slots = max(0, last - first + 1) if keep and n_steps > 0 else 0
        if schedule is not None and keep:
            n_slots = n_steps                      # warm-up: the position before every transition
        samples = np.empty((n_slots, C_, d), np.float32) if keep else None
        inf = np.empty((n_steps, C_, 8), np.float32) if info else None
        z = None if z is None else _f32(z).reshape(n_steps, C_, d)
        uni = None if uni is None else _f32(uni).reshape(n_steps, C_, self.nuts_uni_len())
        capi.check(self.lib.mile_nuts_run_host(self.h, n_steps, step_base, capi.host_ptr(schedule), n_thinning,
                                               capi.host_ptr(z), capi.host_ptr(uni), seed, capi.host_ptr(samples), n_slots,
                                               capi.host_ptr(inf), int(lppd)))
        return samples, inf

    def nuts_warmup(self, n_steps, schedule, *, step_base=0, z=None, uni=None, seed=0, info=False, keep=False):
        """n_steps warm-up transitions, each followed by adapt_step; schedule = [(stage, is_middle_window_end)] of THESE
        steps (window_adaptation.build_schedule).  Returns info [n,C,8] | None, or with keep=True (positions [n,C,d] BEFORE
        each transition, info)."""
        sched = np.ascontiguousarray([int(st) | (2 if end else 0) for st, end in schedule], dtype=np.uint8)
        assert sched.shape == (n_steps,)
        pos, inf = self._nuts_run(n_steps, step_base, sched, 1, z, uni, seed, keep, info, False)
        return (pos, inf) if keep else inf

    def nuts_finish_warmup(self):
        capi.check(self.lib.mile_nuts_finish_warmup(self.h, None))
        self.synchronize()

    def nuts_sample(self, n_steps, *, step_base=0, n_thinning=1, z=None, uni=None, seed=0, keep=True, info=False, lppd=False):
        """(samples [S,C,d] | None, info [n,C,8] | None): positions after every n_thinning-th transition."""
        return self._nuts_run(n_steps, step_base, None, n_thinning, z, uni, seed, keep, info, lppd)

    def nuts_params(self):
        eps = np.empty(self.n_chains, np.float32)
        imm = np.empty((self.n_chains, self.d), np.float32)
        capi.check(self.lib.mile_nuts_get_params_host(self.h, capi.host_ptr(eps), capi.host_ptr(imm)))
        return eps, imm

    def set_nuts_params(self, step_size=None, inverse_mass_matrix=None):
        e = None if step_size is None else _f32(np.broadcast_to(step_size, (self.n_chains,)))
        m = None if inverse_mass_matrix is None else _f32(np.broadcast_to(inverse_mass_matrix, (self.n_chains, self.d)))
        capi.check(self.lib.mile_nuts_set_params_host(self.h, capi.host_ptr(e), capi.host_ptr(m)))

    # ---- partition sampling (partition_sampling.py; trainer.py:613-659) ------------------------
    def set_frozen_mask(self, frozen):
        """frozen: bool [d] (True = the parameter keeps its value, sees no prior, gets no gradient / momentum / noise) or
        None to clear.  The dimension of the MCLMC dynamics becomes the number of sampled parameters."""
        a = None if frozen is None else np.ascontiguousarray(np.asarray(frozen).reshape(self.d), dtype=np.uint8)
        capi.check(self.lib.mile_set_frozen_mask_host(self.h, capi.host_ptr(a)))

    # ---- diagonal preconditioning (warmup.py:385-401; blackjax sqrt_diag_cov) -----------------
    def precondition_from_moments(self):
        """sqrt_diag_cov = sqrt(E[x^2] - E[x]^2) from the phase-2 streaming moments, L = sqrt(d); active from now on."""
        capi.check(self.lib.mile_precondition_from_moments(self.h, None))
        self.synchronize()

    def set_sqrt_diag_cov(self, sdc):
        """[C,d] (or [d], broadcast) preconditioner; None clears it."""
        a = None if sdc is None else _f32(np.broadcast_to(sdc, (self.n_chains, self.d)))
        capi.check(self.lib.mile_set_sqrt_diag_cov_host(self.h, capi.host_ptr(a)))

    def get_sqrt_diag_cov(self):
        out = np.empty((self.n_chains, self.d), np.float32)
        capi.check(self.lib.mile_get_sqrt_diag_cov_host(self.h, capi.host_ptr(out)))
        return out

    # ---- LPPD / predict -----------------------------------------------------------------
    def lppd_reset(self):
        capi.check(self.lib.mile_lppd_reset(self.h, None))
        self.synchronize()

    def lppd_accumulate(self, theta):
        """theta host [n,d] (row c -> chain c) folded into the online logsumexp state."""
        import torch
        theta = _f32(theta).reshape(-1, self.d)
        t = torch.from_numpy(theta).to(f'cuda:{self.device}')
        capi.check(self.lib.mile_lppd_accumulate(self.h, _dev_ptr(t), theta.shape[0], _stream_ptr()))
        torch.cuda.current_stream().synchronize()

    def lppd_state(self):
        m = np.empty((self.n_chains, self.n_test), np.float32)
        s = np.empty((self.n_chains, self.n_test), np.float32)
        cnt = C.c_int64()
        capi.check(self.lib.mile_lppd_state_host(self.h, capi.host_ptr(m), capi.host_ptr(s), C.byref(cnt)))
        return m, s, cnt.value

    def predict(self, theta, which='test'):
        import torch
        theta = _f32(theta).reshape(-1, self.d)
        n = theta.shape[0]
        N = self.n_test if which == 'test' else self.n_train
        K = self.spec.widths[-1]
        t = torch.from_numpy(theta).to(f'cuda:{self.device}')
        out = torch.empty((n, N, K), dtype=torch.float32, device=t.device)
        capi.check(self.lib.mile_predict(self.h, _dev_ptr(t), n, 1 if which == 'test' else 0, _dev_ptr(out),
                                         _stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return out.cpu().numpy()


    # ---- deep-ensemble warm-start training (trainer.py:330-538) ----------------------------------------
    def _to_dev(self, a, dtype=np.float32):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(f'cuda:{self.device}')

    def train_init(self, theta0):
        """Parameters of all members [C, d]; optimizer moments and step counts are zeroed."""
        import torch
        t = self._to_dev(np.reshape(theta0, (self.n_chains, self.d)))
        capi.check(self.lib.mile_train_init(self.h, _dev_ptr(t), _stream_ptr()))
        torch.cuda.current_stream().synchronize()

    @staticmethod
    def opt_cfg(name='adamw', learning_rate=1e-3, b1=0.9, b2=0.999, eps=1e-8, weight_decay=None, **unknown) -> capi.OptCfg:
        """optax.<name>(**parameters) of the reference's OptimizerConfig (src/config/warmstart.py:17-41) with optax's
        defaults (adamw: weight_decay 1e-4; adam / sgd: none)."""
        if unknown:
            raise NotImplementedError(f'optimizer parameters {sorted(unknown)} are not supported on the CUDA path')
        name = str(name).lower()
        if name not in capi.OPTIMIZERS:
            raise NotImplementedError(f'optimizer {name!r} is not supported on the CUDA path')
        wd = (1e-4 if weight_decay is None else weight_decay) if name == 'adamw' else 0.0
        return capi.OptCfg(capi.OPTIMIZERS[name], learning_rate, b1, b2, eps, wd)

    def train_epoch(self, batch_idx, opt: capi.OptCfg, stopped=None, metrics=True):
        """One epoch = one launch.  batch_idx [n_batches, B] int32 rows of the training split (shared by all members);
        stopped [C] bool.  Returns [n_batches, C, 2] = (loss, RMSE | accuracy) per step, NaN for stopped members."""
        import torch
        batch_idx = np.ascontiguousarray(batch_idx, dtype=np.int32)
        nb, B = batch_idx.shape
        if batch_idx.size and (batch_idx.min() < 0 or batch_idx.max() >= self.n_train):
            raise ValueError('batch index out of range')
        bi = self._to_dev(batch_idx, np.int32)
        st = None if stopped is None else self._to_dev(np.asarray(stopped, dtype=np.uint8), np.uint8)
        out = torch.empty((nb, self.n_chains, 2), dtype=torch.float32, device=bi.device) if metrics else None
        capi.check(self.lib.mile_train_epoch(self.h, _dev_ptr(bi), nb, B, C.byref(opt), _dev_ptr(st), _dev_ptr(out),
                                             _stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return None if out is None else out.cpu().numpy()

    def eval_metrics(self, theta=None, which='test'):
        """predict_regr / predict_class: (mean loss, RMSE | accuracy) per parameter row over a split -> [n, 2]."""
        import torch
        t = None if theta is None else self._to_dev(np.reshape(theta, (-1, self.d)))
        n = self.n_chains if t is None else t.shape[0]
        out = torch.empty((n, 2), dtype=torch.float32, device=f'cuda:{self.device}')
        capi.check(self.lib.mile_eval_metrics(self.h, _dev_ptr(t), n, 1 if which == 'test' else 0, _dev_ptr(out), _stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return out.cpu().numpy()

    def train_state(self):
        import torch
        dev = f'cuda:{self.device}'
        th, m, v = (torch.empty((self.n_chains, self.d), dtype=torch.float32, device=dev) for _ in range(3))
        t = torch.empty(self.n_chains, dtype=torch.int32, device=dev)
        capi.check(self.lib.mile_train_get_state(self.h, _dev_ptr(th), _dev_ptr(m), _dev_ptr(v), _dev_ptr(t), _stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return th.cpu().numpy(), m.cpu().numpy(), v.cpu().numpy(), t.cpu().numpy()


class ShardedEnsemble(Ensemble):
    """Data-sharded variant (covertype): every rank holds all chains and 1/world of the
    training rows; gradients are all-reduced over NCCL at every evaluation.  Call `set_data` with the LOCAL shard."""

    def __init__(self, spec: FCNSpec, n_chains: int, device: int = 0, rank: int | None = None, world: int | None = None,
                 **options):
        super().__init__(spec, n_chains, device, **options)
        import torch
        import torch.distributed as dist
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
            world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world = rank, world
        uid = np.zeros(128, np.uint8)
        if world > 1:
            if rank == 0:
                capi.check(self.lib.mile_nccl_unique_id(capi.host_ptr(uid)))
            t = torch.from_numpy(uid).to(f'cuda:{device}' if dist.get_backend() == 'nccl' else 'cpu')
            dist.broadcast(t, 0)
            uid = np.ascontiguousarray(t.cpu().numpy())
        capi.check(self.lib.mile_shard_init(self.h, capi.host_ptr(uid), rank, world))
        # peer-memory all-reduce over NVLink (CUDA IPC) for the step loop; MILE_SHARD_P2P=0 keeps ncclAllReduce
        import os
        if 1 < world <= 8 and os.environ.get('MILE_SHARD_P2P', '1') != '0' and dist.get_backend() == 'nccl':
            h = np.zeros(64, np.uint8)
            capi.check(self.lib.mile_shard_p2p_handle(self.h, capi.host_ptr(h)))
            t = torch.from_numpy(h).to(f'cuda:{device}')
            allh = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allh, t)
            hs = np.ascontiguousarray(torch.stack(allh).cpu().numpy())
            ok = self.lib.mile_shard_p2p_open(self.h, capi.host_ptr(hs)) == 0
            # all ranks or none: a rank without peer access to everybody sends the whole job back to ncclAllReduce
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f'cuda:{device}')
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if ok:
                    self.set_option('p2p', 0)
                return
            # with the mapping open the whole step loop is one persistent kernel per rank (flagged-word exchange over
            # NVLink inside it); MILE_SHARD_FUSED=0 keeps one launch per phase
            if os.environ.get('MILE_SHARD_FUSED', '1') == '0':
                self.set_option('shard_fused', 0)

    def _pre_close(self):
        # no rank may free its exchange region while a peer can still read it
        try:
            import torch
            import torch.distributed as dist
            if self.world > 1 and self.get_option('p2p') == 1 and dist.is_initialized():
                torch.cuda.synchronize()
                dist.barrier()
        except Exception:
            pass

    @staticmethod
    def shard_rows(n_rows: int, rank: int, world: int) -> slice:
        per = -(-n_rows // world)
        return slice(rank * per, min(n_rows, (rank + 1) * per))

    def _dev(self, a):
        import torch
        return None if a is None else torch.from_numpy(_f32(a)).to(f'cuda:{self.device}')

    def init(self, theta0, z0=None, seed: int = 0):
        import torch
        th, z = self._dev(np.reshape(theta0, (self.n_chains, self.d))), self._dev(z0)
        capi.check(self.lib.mile_shard_mclmc_init(self.h, _dev_ptr(th), _dev_ptr(z), seed, _stream_ptr()))
        torch.cuda.current_stream().synchronize()

    def sample(self, n_steps, step_size, L, *, step_base=0, n_thinning=1, z=None, seed=0, keep=True, info=False,
               lppd=False):
        import torch
        if lppd:
            raise NotImplementedError('fused LPPD is not part of the sharded step loop; use lppd_accumulate')
        C_, d = self.n_chains, self.d
        eps, Ls = self._dev(np.broadcast_to(step_size, (C_,))), self._dev(np.broadcast_to(L, (C_,)))
        first, last = -(-step_base // n_thinning), (step_base + n_steps - 1) // n_thinning
        n_slots = max(0, last - first + 1) if keep and n_steps > 0 else 0
        dev = eps.device
        smp = torch.empty((n_slots, C_, d), dtype=torch.float32, device=dev) if keep else None
        inf = torch.empty((n_steps, C_, 3), dtype=torch.float32, device=dev) if info else None
        zd = self._dev(z)
        capi.check(self.lib.mile_shard_mclmc_sample(self.h, n_steps, step_base, n_thinning, first, _dev_ptr(eps),
                                                    _dev_ptr(Ls), _dev_ptr(zd), seed, _dev_ptr(smp), n_slots,
                                                    _dev_ptr(inf), _stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return (None if smp is None else smp.cpu().numpy()), (None if inf is None else inf.cpu().numpy())

    def sample_device(self, n_steps, step_size_dev, L_dev, *, step_base=0, n_thinning=1, sample_base=0, z_dev=None,
                      seed=0, samples_dev=None, n_slots=0, info_dev=None, lppd=False):
        """Device-buffer form of `sample` on torch's current stream (no host transfer, asynchronous)."""
        if lppd:
            raise NotImplementedError('fused LPPD is not part of the sharded step loop; use lppd_accumulate')
        capi.check(self.lib.mile_shard_mclmc_sample(self.h, n_steps, step_base, n_thinning, sample_base,
                                                    _dev_ptr(step_size_dev), _dev_ptr(L_dev), _dev_ptr(z_dev), seed,
                                                    _dev_ptr(samples_dev), n_slots, _dev_ptr(info_dev), _stream_ptr()))

    def tune(self, n_steps, step_base, cfg: capi.TuneCfg, z=None, seed=0, info=False):
        import torch
        dev = f'cuda:{self.device}'
        inf = torch.empty((n_steps, self.n_chains, 4), dtype=torch.float32, device=dev) if info else None
        zd = self._dev(z)
        capi.check(self.lib.mile_shard_mclmc_tune(self.h, n_steps, step_base, C.byref(cfg), _dev_ptr(zd), seed,
                                                  _dev_ptr(inf), _stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return None if inf is None else inf.cpu().numpy()


def lppd_from_state(m: np.ndarray, s: np.ndarray, total_samples: int) -> float:
    """Merge per-chain online (max, sum-exp) states into LPPD = mean_n logsumexp_{c,s}(lp) - log(C*S)
    (src/inference/metrics.py:296-312)."""
    M = m.max(axis=0)
    safe = np.where(np.isfinite(M), M, 0.0)
    tot = (s.astype(np.float64) * np.exp(m.astype(np.float64) - safe)).sum(axis=0)
    return float((safe + np.log(tot / total_samples)).mean())
