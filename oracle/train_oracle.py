"""CPU restatement of the deep-ensemble warm-start training step (TEST INFRASTRUCTURE ONLY; nothing under mile_b200/
imports this file).  Follows src/training/trainer.py:662-760 (single_step_regr / single_step_class: mean Gaussian-NLL /
softmax cross-entropy over the minibatch, `state.apply_gradients`), src/inference/metrics.py:315-333 (GaussianNLLLoss)
and optax 0.2.3's adamw / adam / sgd update rules (pinned in the reference's pyproject.toml; restated from the published
algorithm: scale_by_adam with bias correction, add_decayed_weights, scale by -learning_rate).  Pinned in
tests/test_oracle_training.py against an independent torch implementation (autograd + torch.optim.AdamW / Adam / SGD).
"""
from __future__ import annotations

import numpy as np

from . import mile_oracle as o


def batch_loss_and_grad(spec, theta, Xb, yb):
    """Mean minibatch loss and its gradient; (loss, grad [d], aux) with aux = RMSE (regression) or accuracy."""
    dt = theta.dtype
    lp, g = o.logpost_value_and_grad(spec, theta, Xb.astype(dt), yb)
    # log-posterior = log-prior + sum log-lik (n_batches = 1): strip the prior to get the likelihood part
    pr, gp = o.log_prior(spec, theta)
    B = Xb.shape[0]
    loss = -(lp - pr) / B
    grad = -(g - gp) / B
    out = o.forward(spec, theta, Xb.astype(dt))
    if spec.task == 'regr':
        aux = float(np.sqrt(np.mean((yb.astype(dt) - out[:, 0]) ** 2)))
    else:
        aux = float(np.mean(np.argmax(out, axis=1) == yb))
    return float(loss), grad, aux


class OptState:
    def __init__(self, d, dtype=np.float64):
        self.m, self.v, self.t = np.zeros(d, dtype), np.zeros(d, dtype), 0


def opt_update(kind, theta, grad, st: OptState, lr, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4):
    """optax.adamw / adam / sgd: returns the new parameters (st updated in place)."""
    if kind == 'sgd':
        return theta - lr * grad
    st.t += 1
    st.m = b1 * st.m + (1 - b1) * grad
    st.v = b2 * st.v + (1 - b2) * grad * grad
    mhat, vhat = st.m / (1 - b1 ** st.t), st.v / (1 - b2 ** st.t)
    upd = mhat / (np.sqrt(vhat) + eps)
    if kind == 'adamw':
        upd = upd + wd * theta
    return theta - lr * upd


def train_epoch(spec, theta, st: OptState, X, y, batch_idx, kind='adamw', **opt):
    """One epoch over batch_idx [n_batches, B]; returns (theta, metrics [n_batches, 2])."""
    mets = []
    for idx in batch_idx:
        loss, g, aux = batch_loss_and_grad(spec, theta, X[idx], y[idx])
        mets.append((loss, aux))
        theta = opt_update(kind, theta, g, st, **opt)
    return theta, np.asarray(mets)


def eval_metrics(spec, theta, X, y):
    """predict_regr / predict_class (trainer.py:763-868): mean loss and RMSE | accuracy over a whole split."""
    dt = theta.dtype
    out = o.forward(spec, theta, X.astype(dt))
    if spec.task == 'regr':
        sigma = np.clip(np.exp(out[:, 1]), 1e-5, 1e6)
        res = y.astype(dt) - out[:, 0]
        nll = 0.5 * np.log(2 * np.pi * sigma ** 2) + res ** 2 / (2 * sigma ** 2)
        return float(nll.mean()), float(np.sqrt(np.mean(res ** 2)))
    m = out.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(out - m).sum(axis=1, keepdims=True)))[:, 0]
    ce = lse - out[np.arange(len(y)), y]
    return float(ce.mean()), float(np.mean(np.argmax(out, axis=1) == y))


def earlystop(losses: np.ndarray, patience: int) -> np.ndarray:
    """trainer.py:920-938 (losses [n_members, n_epochs])."""
    if losses.shape[-1] < patience:
        return np.zeros(len(losses), bool)
    # (jax clamps the out-of-range index -(patience + 1) to column 0 when exactly `patience` losses are recorded)
    ref = losses[:, -(patience + 1)][:, None] if losses.shape[-1] > patience else losses[:, :1]
    return np.all(losses[:, -patience:] >= ref, axis=1)
