"""Partition sampling (mirror of src/training/partition_sampling.py:32-330 with the log-density of
src/training/trainer.py:613-659): only the first and the last layer of the FCN are sampled, the hidden layers stay at
their warm-start values.  On the CUDA path this is the ordinary sampler with a frozen-parameter mask
(`mile_set_frozen_mask_host`): frozen parameters keep their value, contribute no prior term and get zero gradient,
momentum and noise, and the dimension of the MCLMC dynamics is the number of sampled parameters.  The saved positions are
the merged networks, as in the reference (partition_sampling.py:117-126)."""
from __future__ import annotations

from pathlib import Path

from .config import SamplerConfig
from .probabilistic import merge_partition, unwrap_posterior
from .sampling import inference_loop


def partition_params(params: dict):
    """partition_sampling.py:290-302: ({'fcn': first + last layer}, {'fcn': hidden layers})."""
    n = len(params['fcn'])
    io = {k: v for k, v in params['fcn'].items() if k in ('layer0', f'layer{n - 1}')}
    hidden = {k: v for k, v in params['fcn'].items() if k not in io}
    return {'fcn': io}, {'fcn': hidden}


def partition_inference_loop(unnorm_log_posterior, config: SamplerConfig, rng_key, init_params: dict, step_ids,
                             saving_path: Path, saving_path_warmup: Path | None = None):
    """partition_sampling.py:32-223.  `unnorm_log_posterior` = partial(prob_model.log_unnormalized_posterior_partition,
    x=train_x, y=train_y) (trainer.py:593-598); `init_params` the FULL stacked parameter tree (it is partitioned here, as
    the reference's warmup does, partition_sampling.py:268)."""
    model, _, _ = unwrap_posterior(unnorm_log_posterior)
    frozen = model.spec.hidden_layer_mask()
    if not frozen.any():
        raise ValueError('partition sampling needs at least one hidden layer to freeze')
    return inference_loop(unnorm_log_posterior, config, rng_key, init_params, step_ids, saving_path, saving_path_warmup,
                          _frozen=frozen)


__all__ = ['partition_params', 'partition_inference_loop', 'merge_partition']
