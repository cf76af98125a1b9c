"""Deep-ensemble warm-start training (mirror of src/training/trainer.py:330-538: train_warmstart / train_de_member, and
of the step / predict functions l.662-868).  Every epoch of all members is ONE launch of the persistent training kernel
(csrc/mile_train.cuh: minibatch gather + forward / loss / backward + fused AdamW), validation and the final test pass are
one metrics launch each; early stopping, the batch permutations and the file output (`warmstart/params_{i}.npz`, `tree`)
stay on the host with the reference's semantics."""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np

from .engine import Ensemble, FCNSpec
from .utils import save_params

logger = logging.getLogger(__name__)


def earlystop(losses: np.ndarray, patience: int) -> np.ndarray:
    """trainer.py:920-938: stop a member when its last `patience` validation losses are all >= the one before them.
    (Like the reference, the window needs patience + 1 recorded losses; with fewer it indexes from the end.)"""
    losses = np.asarray(losses)
    if losses.shape[-1] < patience:
        return np.zeros(len(losses), bool)
    ref = losses[:, -(patience + 1)][:, None] if losses.shape[-1] > patience else losses[:, :1]
    return np.all(losses[:, -patience:] >= ref, axis=1)


def init_params(spec: FCNSpec, rng: np.random.Generator, n_members: int) -> np.ndarray:
    """flax.linen.Dense defaults (src/flax_building_blocks/basic.py:49-60): kernel ~ lecun_normal (truncated normal,
    variance 1 / fan_in), bias = 0.  -> [n_members, d] in ravel_pytree order.  (JAX's threefry streams cannot be
    reproduced without jax; the DISTRIBUTION is the reference's.)"""
    b_off, k_off = spec.offsets()
    dims = spec.dims
    th = np.zeros((n_members, spec.n_params), np.float32)
    for l in range(len(spec.widths)):
        fan_in, out = dims[l], dims[l + 1]
        std = np.sqrt(1.0 / fan_in) / 0.87962566103423978          # jax.nn.initializers.variance_scaling, truncated_normal
        z = rng.standard_normal((n_members, fan_in * out))
        bad = np.abs(z) > 2.0
        while bad.any():                                            # truncate at +-2 sigma by resampling
            z[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(z) > 2.0
        th[:, k_off[l]:k_off[l] + fan_in * out] = (z * std).astype(np.float32)
    return th


def make_batches(n_rows: int, batch_size: int, rng: np.random.Generator) -> np.ndarray:
    """TabularLoader._iter with a batch size (src/dataset/tabular.py:180-199): drop the remainder, permute, split into
    n_batches equal parts.  The reference hands every device the same permutation key, i.e. the same batches."""
    nb = n_rows // batch_size
    perm = rng.permutation(nb * batch_size).astype(np.int32)
    return perm.reshape(nb, batch_size)


def train_de_members(ens: Ensemble, theta0: np.ndarray, valid, test, *, optimizer: dict, max_epochs: int, batch_size: int,
                     patience: int | None, rng: np.random.Generator):
    """train_de_member for all members of `ens` at once (ens.set_data holds the training split).
    valid / test = (X, y) or None.  Returns (params [C, d], metrics dict of arrays)."""
    C = ens.n_chains
    opt = ens.opt_cfg(**optimizer)
    ens.train_init(theta0)
    stopped = np.zeros(C, bool)
    valid_losses = np.zeros((C, 0), np.float32)
    m_train, m_valid = [], []
    if valid is not None:
        ens.set_test(*valid)
    epoch = -1
    for epoch in range(int(max_epochs)):
        if stopped.all():
            break                                                   # trainer.py:436-437
        batches = make_batches(ens.n_train, batch_size or ens.n_train, rng)
        m_train.append(ens.train_epoch(batches, opt, stopped=stopped))
        if valid is not None:
            mv = ens.eval_metrics(None, 'test')                     # full-batch validation pass after every epoch
            mv[stopped] = np.nan
            m_valid.append(mv)
            valid_losses = np.concatenate([valid_losses, mv[:, :1]], axis=1)
            if patience:
                stopped = stopped | earlystop(valid_losses, patience)
            logger.info(f'Epoch {epoch} | Validation Loss: {mv[:, 0]} | early stopping status: {stopped}')
    params = ens.train_state()[0]
    m_test = None
    if test is not None:
        ens.set_test(*test)
        m_test = ens.eval_metrics(None, 'test')
    metrics = {'train': np.concatenate(m_train) if m_train else np.zeros((0, C, 2), np.float32),
               'valid': np.stack(m_valid) if m_valid else np.zeros((0, C, 2), np.float32),
               'test': m_test, 'epochs': epoch + 1, 'stopped': stopped}
    return params, metrics


def train_warmstart(spec: FCNSpec, train, valid, test, exp_dir, step_ids, *, optimizer: dict, max_epochs: int,
                    batch_size: int, patience: int | None, seed: int = 0, device: int = 0):
    """BDETrainer.train_warmstart without checkpoints (trainer.py:330-364): trains len(step_ids) members and writes
    <exp_dir>/warmstart/params_{i}.npz (+ `tree` / tree.json) in the layout `start_sampling` reads."""
    rng = np.random.default_rng(seed)
    step_ids = [int(s) for s in step_ids]
    ens = Ensemble(spec, len(step_ids), device=device)
    try:
        ens.set_data(*train)
        theta0 = init_params(spec, rng, len(step_ids))
        params, metrics = train_de_members(ens, theta0, valid, test, optimizer=optimizer, max_epochs=max_epochs,
                                           batch_size=batch_size, patience=patience, rng=rng)
    finally:
        ens.close()
    warm = Path(exp_dir) / 'warmstart'
    for k, cid in enumerate(step_ids):
        save_params(warm, spec.unravel(params[k]), cid)
        logger.info(f'\t| Deep Ensemble {cid} saved at {warm}')
    return params, metrics
