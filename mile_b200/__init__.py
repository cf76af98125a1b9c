"""mile_b200 -- B200-native MCLMC ensemble sampling path for MILE (zhiyuan-yang/MILE).

Only the hot path named by BASELINE.json:north_star lives here: the C-ABI CUDA library
(csrc/, built into _lib/libmile_b200.so) and the host-side mirror of the reference's
Python seams (sampling.inference_loop, kernels.KERNELS['mclmc' | 'nuts'], warmup.custom_mclmc_warmup,
nuts.custom_window_adaptation, probabilistic.ProbabilisticModel, ...).  There is no CPU fallback.
"""
from .engine import Ensemble, FCNSpec, ShardedEnsemble, lppd_from_state  # noqa: F401
from .config import PriorConfig, Sampler, SamplerConfig  # noqa: F401
from .kernels import KERNELS, mclmc  # noqa: F401   (KERNELS['nuts']: the name `nuts` is the submodule mile_b200.nuts)
from .models import FCN  # noqa: F401
from .priors import Prior, PriorDist  # noqa: F401
from .probabilistic import ProbabilisticModel  # noqa: F401
from .sampling import inference_loop, warmup_mclmc, warmup_nuts  # noqa: F401
from .nuts import custom_window_adaptation  # noqa: F401
from .warmup import custom_mclmc_warmup  # noqa: F401
from .partition_sampling import partition_inference_loop, partition_params  # noqa: F401
from .evaluation import evaluate_bde, predict_bde  # noqa: F401

__all__ = ['Ensemble', 'FCNSpec', 'ShardedEnsemble', 'lppd_from_state', 'PriorConfig', 'Sampler', 'SamplerConfig', 'KERNELS', 'mclmc',
           'FCN', 'Prior', 'PriorDist', 'ProbabilisticModel', 'inference_loop', 'warmup_mclmc', 'warmup_nuts', 'custom_mclmc_warmup', 'custom_window_adaptation',
           'evaluate_bde', 'predict_bde', 'partition_inference_loop', 'partition_params']
