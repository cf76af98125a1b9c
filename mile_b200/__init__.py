"""mile_b200 -- B200-native MCLMC ensemble sampling path for MILE (zhiyuan-yang/MILE).

Only the hot path named by BASELINE.json:north_star lives here: the C-ABI CUDA library
(csrc/, built into _lib/libmile_b200.so) and the host-side mirror of the reference's
Python seams (sampling.inference_loop, kernels.KERNELS['mclmc'], warmup.custom_mclmc_warmup,
probabilistic.ProbabilisticModel, ...).  There is no CPU fallback.
"""
from .engine import Ensemble, FCNSpec, lppd_from_state  # noqa: F401

__all__ = ['Ensemble', 'FCNSpec', 'lppd_from_state']
