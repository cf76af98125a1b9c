"""Model descriptors (mirror of src/models/tabular/fcn.py:11-28 + src/config/models/fcn.py:7-30).  The
network itself only ever runs inside the CUDA library; this object carries its shape."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class FCN:
    """Fully connected network: `hidden_structure` INCLUDES the output width (e.g. [16,16,16,2])."""
    hidden_structure: tuple
    activation: str = 'relu'
    use_bias: bool = True

    def __post_init__(self):
        object.__setattr__(self, 'hidden_structure', tuple(int(h) for h in self.hidden_structure))
        if not self.use_bias:
            raise NotImplementedError('use_bias=False is not supported by the CUDA path')

    def init(self, rng: np.random.Generator, n_features: int, scale: float = 1.0) -> dict:
        """Random ParamTree {'fcn': {'layer{i}': {'kernel','bias'}}} (lecun-normal kernels, zero biases)."""
        tree, fan_in = {}, n_features
        for i, h in enumerate(self.hidden_structure):
            tree[f'layer{i}'] = {
                'kernel': (rng.standard_normal((fan_in, h)) * scale / np.sqrt(fan_in)).astype(np.float32),
                'bias': np.zeros(h, np.float32)}
            fan_in = h
        return {'fcn': tree}
