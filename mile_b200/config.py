"""Sampler configuration (mirror of src/config/sampler.py:10-201 and the YAML `training.sampler` block)."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field, fields

from .priors import Prior, PriorDist


class Sampler(str, enum.Enum):
    """src/config/sampler.py:10-57."""
    NUTS = 'nuts'
    HMC = 'hmc'
    MCLMC = 'mclmc'

    def get_kernel(self):
        from .kernels import KERNELS
        return KERNELS[self.value]


@dataclass(frozen=True)
class PriorConfig:
    """src/config/sampler.py:60-92."""
    name: PriorDist = PriorDist.StandardNormal
    parameters: dict = field(default_factory=dict)

    def get_prior(self) -> Prior:
        return PriorDist(self.name).get_prior(**self.parameters)


@dataclass(frozen=True)
class SamplerConfig:
    """src/config/sampler.py:95-201 (same field names and defaults)."""
    name: Sampler = Sampler.NUTS
    epoch_wise_sampling: bool = False
    params_frozen: list = field(default_factory=list)
    warmup_steps: int = 50
    n_chains: int = 2
    n_samples: int = 1000
    use_warmup_as_init: bool = True
    n_thinning: int = 1
    diagonal_preconditioning: bool = False
    desired_energy_var_start: float = 5e-4
    desired_energy_var_end: float = 1e-4
    trust_in_estimate: float = 1.5
    num_effective_samples: int = 100
    step_size_init: float = 0.005
    keep_warmup: bool = False
    prior_config: PriorConfig = field(default_factory=PriorConfig)
    partition_sampling: bool = False

    def __post_init__(self):
        object.__setattr__(self, 'name', Sampler(self.name))
        if isinstance(self.prior_config, dict):
            pc = dict(self.prior_config)
            object.__setattr__(self, 'prior_config', PriorConfig(PriorDist(pc.get('name', 'StandardNormal')),
                                                                 dict(pc.get('parameters') or {})))

    @classmethod
    def from_dict(cls, d: dict) -> 'SamplerConfig':
        known = {f.name for f in fields(cls)}
        extra = set(d) - known
        if extra:  # the reference rejects unknown keys (src/config/base.py:394-395)
            raise ValueError(f'unknown sampler config keys: {sorted(extra)}')
        return cls(**d)

    @classmethod
    def from_yaml(cls, path) -> 'SamplerConfig':
        """Reads `training.sampler` of a reference experiment YAML (experiments/*/mclmc.yaml)."""
        import yaml
        with open(path) as f:
            cfg = yaml.safe_load(f)
        return cls.from_dict(cfg['training']['sampler'])

    @property
    def prior(self) -> Prior:
        return self.prior_config.get_prior()

    @property
    def kernel(self):
        return self.name.get_kernel()

    @property
    def _dir_name(self) -> str:
        return 'samples'
