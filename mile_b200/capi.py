"""ctypes binding of include/mile_b200.h.  The product path: there is no CPU fallback --
if the CUDA library is missing or no CUDA device is present, calls fail loudly."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

MILE_MAX_LAYERS = 12
ACTIVATIONS = {'identity': 0, 'relu': 1, 'sigmoid': 2, 'tanh': 3, 'gelu': 4, 'leaky_relu': 5}
TASKS = {'regr': 0, 'regression': 0, 'class': 1, 'classification': 1}
PRIORS = {'normal': 0, 'standardnormal': 0, 'laplace': 1}

# every symbol include/mile_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    'mile_last_error', 'mile_version', 'mile_create', 'mile_destroy', 'mile_n_params', 'mile_set_option',
    'mile_get_option', 'mile_set_data', 'mile_set_data_host', 'mile_set_test', 'mile_set_test_host',
    'mile_logpost_value_and_grad', 'mile_logpost_value_and_grad_host', 'mile_mclmc_init', 'mile_mclmc_init_host',
    'mile_set_state_host', 'mile_get_state_host', 'mile_get_state', 'mile_mclmc_sample', 'mile_mclmc_sample_host',
    'mile_tune_reset', 'mile_mclmc_tune', 'mile_mclmc_tune_host', 'mile_tune_finish_phase2', 'mile_get_tuning_host',
    'mile_set_tuning_host', 'mile_tuning_ptrs', 'mile_lppd_reset', 'mile_lppd_accumulate', 'mile_lppd_state_host',
    'mile_predict', 'mile_launch_count', 'mile_synchronize', 'mile_measure_fp32_peak',
    'mile_mclmc_phase3_ess', 'mile_ess_positions', 'mile_ess_positions_host', 'mile_ess_pooled_host',
    'mile_nuts_init', 'mile_nuts_init_host', 'mile_nuts_warmup', 'mile_nuts_finish_warmup', 'mile_nuts_sample', 'mile_nuts_run_host',
    'mile_nuts_get_params_host', 'mile_nuts_set_params_host',
    'mile_write_npz_batch', 'mile_set_frozen_mask_host', 'mile_precondition_from_moments', 'mile_set_sqrt_diag_cov_host', 'mile_get_sqrt_diag_cov_host',
    'mile_train_init', 'mile_train_epoch', 'mile_eval_metrics', 'mile_train_get_state',
    'mile_nccl_unique_id', 'mile_shard_init', 'mile_shard_p2p_handle', 'mile_shard_p2p_open', 'mile_shard_mclmc_init', 'mile_shard_mclmc_sample', 'mile_shard_mclmc_tune',
]


class ModelDesc(C.Structure):
    _fields_ = [
        ('n_features', C.c_int32), ('n_layers', C.c_int32),
        ('widths', C.c_int32 * MILE_MAX_LAYERS), ('bias_off', C.c_int32 * MILE_MAX_LAYERS),
        ('kernel_off', C.c_int32 * MILE_MAX_LAYERS),
        ('activation', C.c_int32), ('task', C.c_int32), ('prior', C.c_int32),
        ('prior_loc', C.c_float), ('prior_scale', C.c_float), ('n_batches', C.c_float),
    ]


class TuneCfg(C.Structure):
    _fields_ = [
        ('tune1_steps', C.c_int32), ('tune2_steps', C.c_int32),
        ('desired_energy_var_start', C.c_float), ('desired_energy_var_end', C.c_float),
        ('trust_in_estimate', C.c_float), ('num_effective_samples', C.c_float),
    ]


class NutsCfg(C.Structure):
    _fields_ = [('max_num_doublings', C.c_int32), ('divergence_threshold', C.c_float), ('target_acceptance_rate', C.c_float),
                ('initial_step_size', C.c_float)]


class OptCfg(C.Structure):
    _fields_ = [('kind', C.c_int32), ('learning_rate', C.c_float), ('b1', C.c_float), ('b2', C.c_float), ('eps', C.c_float),
                ('weight_decay', C.c_float)]


OPTIMIZERS = {'adamw': 0, 'adam': 1, 'sgd': 2}


class MileError(RuntimeError):
    pass


_LIB = None


def lib_path() -> Path:
    return Path(__file__).resolve().parent / '_lib' / 'libmile_b200.so'


def load():
    """Load libmile_b200.so (built by mile_b200.build / __graft_entry__.build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.exists():
        raise MileError(f'{p} is missing: run `python -m mile_b200.build` (there is no CPU fallback)')
    lib = C.CDLL(str(p))
    vp, i32, i64, u64, fp = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p
    lib.mile_last_error.restype = C.c_char_p
    lib.mile_version.restype = C.c_int
    lib.mile_create.argtypes = [C.POINTER(ModelDesc), i32, i32, C.POINTER(vp)]
    lib.mile_destroy.argtypes = [vp]
    lib.mile_destroy.restype = None
    lib.mile_n_params.argtypes = [vp]
    lib.mile_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.mile_get_option.argtypes = [vp, C.c_char_p]
    lib.mile_get_option.restype = i64
    for name in ('mile_set_data', 'mile_set_test'):
        getattr(lib, name).argtypes = [vp, fp, vp, i64, vp]
    for name in ('mile_set_data_host', 'mile_set_test_host'):
        getattr(lib, name).argtypes = [vp, fp, vp, i64]
    lib.mile_logpost_value_and_grad.argtypes = [vp, fp, i32, fp, fp, vp]
    lib.mile_logpost_value_and_grad_host.argtypes = [vp, fp, i32, fp, fp]
    lib.mile_mclmc_init.argtypes = [vp, fp, fp, u64, vp]
    lib.mile_mclmc_init_host.argtypes = [vp, fp, fp, u64]
    lib.mile_set_state_host.argtypes = [vp, fp, fp, fp, fp]
    lib.mile_get_state_host.argtypes = [vp, fp, fp, fp, fp]
    lib.mile_get_state.argtypes = [vp, fp, fp, fp, fp, vp]
    lib.mile_mclmc_sample.argtypes = [vp, i32, i64, i32, i64, fp, fp, fp, u64, fp, i64, fp, i32, vp]
    lib.mile_mclmc_sample_host.argtypes = [vp, i32, i64, i32, fp, fp, fp, u64, fp, i64, fp, i32]
    lib.mile_tune_reset.argtypes = [vp, C.c_float, vp]
    lib.mile_mclmc_tune.argtypes = [vp, i32, i64, C.POINTER(TuneCfg), fp, u64, fp, vp]
    lib.mile_mclmc_tune_host.argtypes = [vp, i32, i64, C.POINTER(TuneCfg), fp, u64, fp]
    lib.mile_tune_finish_phase2.argtypes = [vp, vp]
    lib.mile_get_tuning_host.argtypes = [vp, fp, fp, fp, fp, fp]
    lib.mile_set_tuning_host.argtypes = [vp, fp, fp]
    lib.mile_tuning_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.mile_lppd_reset.argtypes = [vp, vp]
    lib.mile_lppd_accumulate.argtypes = [vp, fp, i32, vp]
    lib.mile_lppd_state_host.argtypes = [vp, fp, fp, C.POINTER(i64)]
    lib.mile_predict.argtypes = [vp, fp, i32, i32, fp, vp]
    lib.mile_launch_count.argtypes = [vp]
    lib.mile_launch_count.restype = i64
    lib.mile_synchronize.argtypes = [vp]
    lib.mile_measure_fp32_peak.argtypes = [i32, i32, C.POINTER(C.c_double)]
    lib.mile_train_init.argtypes = [vp, fp, vp]
    lib.mile_train_epoch.argtypes = [vp, vp, i32, i32, C.POINTER(OptCfg), vp, fp, vp]
    lib.mile_eval_metrics.argtypes = [vp, fp, i32, i32, fp, vp]
    lib.mile_train_get_state.argtypes = [vp, fp, fp, fp, vp, vp]
    lib.mile_mclmc_phase3_ess.argtypes = [vp, i32, fp, fp, u64, vp, i32, vp, i32, fp]
    lib.mile_ess_positions.argtypes = [vp, fp, i32, vp, i32, vp, i32, fp, vp]
    lib.mile_ess_positions_host.argtypes = [vp, fp, i32, vp, i32, vp, i32, fp]
    lib.mile_ess_pooled_host.argtypes = [vp, fp, i32, vp, i32, vp, i32, fp]
    lib.mile_nuts_init.argtypes = [vp, fp, C.POINTER(NutsCfg), vp]
    lib.mile_nuts_init_host.argtypes = [vp, fp, C.POINTER(NutsCfg)]
    lib.mile_nuts_warmup.argtypes = [vp, i32, i64, vp, fp, fp, u64, fp, fp, vp]
    lib.mile_nuts_finish_warmup.argtypes = [vp, vp]
    lib.mile_nuts_sample.argtypes = [vp, i32, i64, i32, i64, fp, fp, u64, fp, i64, fp, i32, vp]
    lib.mile_nuts_run_host.argtypes = [vp, i32, i64, vp, i32, fp, fp, u64, fp, i64, fp, i32]
    lib.mile_nuts_get_params_host.argtypes = [vp, fp, fp]
    lib.mile_nuts_set_params_host.argtypes = [vp, fp, fp]
    lib.mile_write_npz_batch.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, i32]
    lib.mile_set_frozen_mask_host.argtypes = [vp, vp]
    lib.mile_precondition_from_moments.argtypes = [vp, vp]
    lib.mile_set_sqrt_diag_cov_host.argtypes = [vp, fp]
    lib.mile_get_sqrt_diag_cov_host.argtypes = [vp, fp]
    lib.mile_nccl_unique_id.argtypes = [vp]
    lib.mile_shard_init.argtypes = [vp, vp, i32, i32]
    lib.mile_shard_p2p_handle.argtypes = [vp, vp]
    lib.mile_shard_p2p_open.argtypes = [vp, vp]
    lib.mile_shard_mclmc_init.argtypes = [vp, fp, fp, u64, vp]
    lib.mile_shard_mclmc_sample.argtypes = [vp, i32, i64, i32, i64, fp, fp, fp, u64, fp, i64, fp, vp]
    lib.mile_shard_mclmc_tune.argtypes = [vp, i32, i64, C.POINTER(TuneCfg), fp, u64, fp, vp]
    _LIB = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise MileError(load().mile_last_error().decode())


def host_ptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(C.c_void_p)
