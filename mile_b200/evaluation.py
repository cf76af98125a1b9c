"""Posterior-predictive evaluation with the reference's names (mirror of src/inference/evaluation.py:14-137,334-544).

The reference runs one un-jitted Flax forward per (chain, sample) in a Python loop (evaluation.py:38-42,378-400);
here every (chain, sample) parameter vector is one row of a [C*S, d] matrix and the CUDA library's PREDICT mode
(`mile_predict`, include/mile_b200.h) evaluates them in one launch per chunk.  The metric reductions that follow
are small numpy expressions over the [C, S, N, K] predictions, as in the reference.  No CPU forward pass exists
here: without the CUDA library these functions raise.
"""
from __future__ import annotations

import numpy as np

from . import metrics
from .engine import Ensemble, FCNSpec


def _task_name(task) -> str:
    t = str(getattr(task, 'value', task)).lower()
    return 'regr' if t.startswith('regr') else 'class'


def _is_regression(task) -> bool:
    return _task_name(task) == 'regr'


def _leaves(tree):
    if isinstance(tree, dict):
        for k in sorted(tree):
            yield from _leaves(tree[k])
    else:
        yield np.asarray(tree)


def count_chains(params) -> int:
    """evaluation.py:370 (src/training/utils.py count_chains): leading axis of every leaf."""
    return next(_leaves(params)).shape[0]


def count_samples(params) -> int:
    return next(_leaves(params)).shape[1]


def _spec(module, n_features: int, task) -> FCNSpec:
    return FCNSpec(int(n_features), module.hidden_structure, module.activation, _task_name(task))


def _forward_rows(spec: FCNSpec, theta: np.ndarray, x: np.ndarray, device: int | None, chunk: int) -> np.ndarray:
    """theta [n, d], x [B, F] -> network outputs [n, B, K] through the CUDA PREDICT launch."""
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    x = np.ascontiguousarray(x, np.float32)
    n, K = theta.shape[0], spec.widths[-1]
    out = np.empty((n, x.shape[0], K), np.float32)
    chunk = max(1, min(chunk, n, max(1, (1 << 28) // max(1, x.shape[0] * K))))   # <= 1 GiB of outputs per launch
    if spec.n_params > 20000:
        # wide models run layer by layer with HBM-resident activations [chunk, rows, sum(widths)]: keep them under ~2 GiB
        chunk = max(1, min(chunk, (1 << 29) // max(1, x.shape[0] * sum(spec.widths))))
    ens = Ensemble(spec, chunk, device=device)
    try:
        dummy_y = np.zeros(x.shape[0], np.float32 if spec.task.startswith('regr') else np.int32)
        ens.set_test(x, dummy_y)
        for a in range(0, n, chunk):
            out[a:a + chunk] = ens.predict(theta[a:a + chunk], which='test')
    finally:
        ens.close()
    return out


def predict_from_samples(model, samples, x, task='regr', device: int | None = None, chunk: int = 4096, **kwargs) -> np.ndarray:
    """evaluation.py:14-43: samples = ParamTree with ONE leading sample axis -> predictions [n_samples, B, K]."""
    x = np.asarray(x, np.float32)
    spec = _spec(model, x.shape[-1], task)
    return _forward_rows(spec, spec.ravel(samples).reshape(-1, spec.n_params), x, device, chunk)


def sample_from_predictions(predictions: np.ndarray, task, rng_key=42) -> np.ndarray:
    """evaluation.py:46-72.  Draws come from numpy's Generator seeded by `rng_key` (the reference uses
    jax.random with the same distributions; the draws themselves are not bit-comparable)."""
    from .types import key_to_seed
    rng = np.random.default_rng(key_to_seed(rng_key))
    if _is_regression(task):
        loc = predictions[..., 0]
        scale = np.clip(np.exp(predictions[..., 1]), 1e-6, 1e6)
        return rng.standard_normal(loc.shape).astype(np.float32) * scale + loc
    g = rng.gumbel(size=predictions.shape)
    return np.argmax(predictions + g, axis=-1)


def calibration_error(nominal_coverage, observed_coverage, weights=None):
    """evaluation.py:75-86."""
    sq = np.square(np.asarray(nominal_coverage) - np.asarray(observed_coverage))
    return float(np.sqrt(np.mean(sq if weights is None else np.asarray(weights) * sq)))


def coverage_weighting(nominal_coverage, kappa: float = 1.0):
    """evaluation.py:89-95."""
    c = np.asarray(nominal_coverage, np.float64) ** kappa
    return c / c.sum()


def get_quantiles(coverage: float):
    """evaluation.py:98-100."""
    return np.array([0.5 - coverage / 2, 0.5 + coverage / 2])


def calculate_coverage(nominal_coverages, Y_test, preds):
    """evaluation.py:103-137: central credible intervals of the pooled (chain, sample) predictive draws."""
    pooled = np.asarray(preds).reshape(-1, np.shape(preds)[-1])
    Y_test = np.asarray(Y_test)
    cov = []
    for nc in nominal_coverages:
        lo, hi = np.quantile(pooled, get_quantiles(nc), axis=0)
        cov.append(np.mean((lo <= Y_test) & (hi >= Y_test)))
    return np.array(cov)


def predict_bde(params, module, features, labels, batch_size=None, verbose: bool = True, task='regr',
                device: int | None = None, **kwargs) -> np.ndarray:
    """evaluation.py:334-407: params leaves [n_chains, n_samples, ...] -> predictions [n_chains, n_samples, B, K].
    `batch_size` is accepted for signature compatibility; the CUDA launch streams the whole split."""
    features = np.asarray(features, np.float32)
    assert features.shape[0] == np.shape(labels)[0], 'Labels and Features must match on the first dimension'
    n_chains, n_samples = count_chains(params), count_samples(params)
    if verbose:
        print(f'| Predicting with {n_chains} chains each with {n_samples} samples')
    spec = _spec(module, features.shape[-1], task)
    theta = spec.ravel(params).reshape(n_chains * n_samples, spec.n_params)
    out = _forward_rows(spec, theta, features, device, kwargs.get('chunk', 4096))
    return out.reshape(n_chains, n_samples, features.shape[0], -1)


def _mode(a: np.ndarray, n_classes: int, axes) -> np.ndarray:
    """Most frequent class over `axes` (scipy.stats.mode semantics: smallest value wins ties)."""
    onehot = (a[..., None] == np.arange(n_classes)).sum(axis=axes)
    return onehot.argmax(axis=-1)


def evaluate_bde(params, module, features, labels, task, batch_size=None, verbose: bool = True, metrics_dict=None,
                 rng_key=42, nominal_coverages=None, device: int | None = None, **kwargs):
    """evaluation.py:410-544: returns (predictions [C, S, B, K], metrics dict with lppd / nll / rmse|acc and the
    optional coverage entries)."""
    metrics_dict = {} if metrics_dict is None else metrics_dict
    labels = np.asarray(labels)
    regression = _is_regression(task)
    if not regression:
        labels = labels.astype(np.int32)
    lv = predict_bde(params, module, features, labels, batch_size, verbose, task=task, device=device, **kwargs)
    preds = sample_from_predictions(lv, task=task, rng_key=rng_key)
    if not regression:
        if nominal_coverages is not None:
            print('Calculating coverage is not supported for classification yet.')
        pw = metrics.pointwise_lppd(lv, labels, task='class')
        K = lv.shape[-1]
        metrics_dict['lppd'] = metrics.lppd(pw)
        metrics_dict['nll'] = float(-pw.mean())
        metrics_dict['acc'] = float(np.mean(labels == _mode(preds, K, (0, 1))))
        if verbose:
            print(f"Bayesian Deep Ensemble Performance | LPPD: {metrics_dict['lppd']:.3f}, ACC: {metrics_dict['acc']:.4f}")
            print('_' * 50)
            for i in range(lv.shape[0]):
                acc = np.mean(labels == _mode(preds[i], K, (0,)))
                print(f'Chain {i} | LPPD: {metrics.lppd(pw[i:i + 1]):.3f}, ACC: {acc:.4f}')
        return lv, metrics_dict
    nan_chains = np.isnan(preds).any(axis=(1, 2))
    if nan_chains.any():
        print(f'Warning: Chains {np.where(nan_chains)[0]} have NaN predictions')
    ok = ~nan_chains
    pw = metrics.pointwise_lppd(lv[ok], labels, task='regr')
    metrics_dict['lppd'] = metrics.lppd(pw)
    metrics_dict['nll'] = float(-pw.mean())
    metrics_dict['rmse'] = float(np.sqrt(np.mean((labels - preds[ok].mean(axis=(0, 1))) ** 2)))
    if verbose:
        print(f"Bayesian Deep Ensemble Performance | LPPD: {metrics_dict['lppd']:.3f}, RMSE: {metrics_dict['rmse']:.4f}")
        print('_' * 50)
        pw_all = metrics.pointwise_lppd(lv, labels, task='regr')
        for i in range(lv.shape[0]):
            rmse = np.sqrt(np.mean((labels - preds[i].mean(axis=0)) ** 2))
            print(f'Chain {i} | LPPD: {metrics.lppd(pw_all[i:i + 1]):.3f}, RMSE: {rmse:.4f}')
    if nominal_coverages is not None:
        coverage = calculate_coverage(nominal_coverages, labels, preds[ok])
        metrics_dict['cal_error'] = calibration_error(np.array(nominal_coverages), coverage)
        for i, cov in enumerate(nominal_coverages):
            metrics_dict[f'coverage_{cov}'] = float(coverage[i])
        if verbose:
            print('_' * 50)
            print(f"Calibration Error: {metrics_dict['cal_error']:.4f}")
    return lv, metrics_dict
