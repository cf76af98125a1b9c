"""Priors (mirror of src/training/priors.py: PriorDist, Prior.from_name, log_prior closures)."""
from __future__ import annotations

import enum
import math
from typing import Callable, NamedTuple

import numpy as np


class PriorDist(str, enum.Enum):
    """src/training/priors.py:12-56."""
    NORMAL = 'Normal'
    StandardNormal = 'StandardNormal'
    LAPLACE = 'Laplace'

    def get_prior(self, **parameters):
        return Prior.from_name(self, **parameters)


class Prior(NamedTuple):
    """src/training/priors.py:58-98.  `kind/loc/scale` describe the prior to the CUDA library;
    `log_prior` is the host closure of the reference signature (flat numpy evaluation)."""
    f_init: Callable
    log_prior: Callable
    name: str
    kind: str = 'normal'
    loc: float = 0.0
    scale: float = 1.0

    @classmethod
    def from_name(cls, name, **parameters):
        name = PriorDist(name)
        if name == PriorDist.StandardNormal:
            return cls(f_init_normal(), log_prior_normal(), PriorDist.StandardNormal.value, 'normal', 0.0, 1.0)
        loc, scale = float(parameters.get('loc', 0.0)), float(parameters.get('scale', 1.0))
        if name == PriorDist.NORMAL:
            return cls(f_init_normal(loc, scale), log_prior_normal(loc, scale), name.value, 'normal', loc, scale)
        if name == PriorDist.LAPLACE:
            return cls(f_init_laplace(loc, scale), log_prior_laplace(loc, scale), name.value, 'laplace', loc, scale)
        raise NotImplementedError(f'Prior Distribution for {name} is not yet implemented.')


def _leaves(params):
    if isinstance(params, dict):
        out = []
        for k in sorted(params):
            out.extend(_leaves(params[k]))
        return out
    return [np.asarray(params, dtype=np.float32).reshape(-1)]


def f_init_normal(loc: float = 0.0, scale: float = 1.0) -> Callable:
    def f_init(rng: np.random.Generator, shape, dtype=np.float32):
        return (rng.standard_normal(shape) * scale).astype(dtype)
    return f_init


def log_prior_normal(loc: float = 0.0, scale: float = 1.0) -> Callable:
    """priors.py:101-108."""
    def log_prior(params) -> np.float32:
        x = np.concatenate(_leaves(params))
        return np.sum(-(np.log(np.float32(2 * math.pi) * np.float32(scale) ** 2)
                        + np.square(x - np.float32(loc)) / np.float32(scale) ** 2) / np.float32(2), dtype=np.float32)
    return log_prior


def f_init_laplace(loc: float = 0.0, scale: float = 1.0) -> Callable:
    def f_init(rng: np.random.Generator, shape, dtype=np.float32):
        return (rng.laplace(size=shape) * scale + loc).astype(dtype)
    return f_init


def log_prior_laplace(loc: float = 0.0, scale: float = 1.0) -> Callable:
    """priors.py:121-128."""
    def log_prior(params) -> np.float32:
        x = np.concatenate(_leaves(params))
        return np.sum(-np.log(np.float32(2 * scale)) - np.abs(x - np.float32(loc)) / np.float32(scale), dtype=np.float32)
    return log_prior
