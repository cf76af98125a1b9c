"""ProbabilisticModel (mirror of src/training/probabilistic.py:17-138).  `log_unnormalized_posterior`
keeps the reference signature; its value -- and, for the samplers, its gradient -- is computed by the fused
CUDA value_and_grad kernel (mile_logpost_value_and_grad).  No host / CPU evaluation path exists."""
from __future__ import annotations

import functools
import logging
import weakref

import numpy as np

from .engine import Ensemble, FCNSpec
from .models import FCN
from .priors import Prior

logger = logging.getLogger(__name__)

_TASKS = {'regr': 'regr', 'regression': 'regr', 'class': 'class', 'classification': 'class'}


class ProbabilisticModel:
    """Convert a frequentist FCN description into the Bayesian log-posterior the samplers use."""

    def __init__(self, module: FCN, params: dict, prior: Prior, task, n_batches: int = 1):
        self.task = _TASKS[str(getattr(task, 'value', task)).lower()]
        self.module = module
        self.prior = prior
        self.n_batches = n_batches
        inner = params['fcn'] if 'fcn' in params else params
        k0 = np.asarray(inner['layer0']['kernel'])
        self.n_features = int(k0.shape[-2])
        self.n_params = int(sum(np.asarray(v).size for lay in inner.values() for v in lay.values())
                            // max(1, int(np.prod(k0.shape[:-2]))))
        self._engines = weakref.WeakValueDictionary()
        self._test = None
        logger.info(f'Initialized ProbModelBuilder for {self.task} task')

    def __str__(self):
        return (f'{self.__class__.__name__}:\n | Task: {self.task}\n | Params: {self.n_params}'
                f' | Batches: {self.n_batches}\n | Prior: {self.prior.name}')

    @property
    def minibatch(self):
        return self.n_batches > 1

    @property
    def spec(self) -> FCNSpec:
        return FCNSpec(self.n_features, self.module.hidden_structure, self.module.activation, self.task,
                       self.prior.kind, self.prior.loc, self.prior.scale, float(self.n_batches))

    def attach_test_split(self, x, y):
        """Optional: test split for the fused posterior-predictive LPPD (src/inference/evaluation.py:378-400)."""
        self._test = (np.asarray(x, np.float32), np.asarray(y))

    def log_prior(self, params):
        return self.prior.log_prior(params)

    # ---- engine management ---------------------------------------------------------------------
    def make_ensemble(self, n_chains: int, x, y, device: int | None = None, **options) -> Ensemble:
        if device is None:
            # torch's current device when torch is in use by the caller; a plain run never imports torch (seconds of start-up)
            import sys
            torch = sys.modules.get('torch')
            device = torch.cuda.current_device() if torch is not None and torch.cuda.is_available() else 0
        ens = Ensemble(self.spec, n_chains, device=device, **options)
        ens.set_data(np.asarray(x, np.float32), np.asarray(y))
        if self._test is not None:
            ens.set_test(*self._test)
        return ens

    def log_unnormalized_posterior_partition(self, input_output_layers, hidden_layers, x, y, **kwargs):
        """trainer.py:651-659: log prior of the SAMPLED layers (first and last) + log-likelihood of the merged network."""
        spec = self.spec
        theta = spec.ravel(merge_partition(input_output_layers, hidden_layers))
        batched = theta.ndim == 2
        theta = theta.reshape(-1, spec.n_params)
        ens = self.make_ensemble(theta.shape[0], x, y)
        try:
            ens.set_frozen_mask(spec.hidden_layer_mask())
            lp, _ = ens.value_and_grad(theta)
        finally:
            ens.close()
        return lp if batched else lp[0]

    def log_unnormalized_posterior(self, position, x, y, **kwargs):
        """probabilistic.py:115-138.  position: ParamTree (optionally with a leading chain axis)."""
        spec = self.spec
        theta = spec.ravel(position)
        batched = theta.ndim == 2
        theta = theta.reshape(-1, spec.n_params)
        ens = self.make_ensemble(theta.shape[0], x, y)
        try:
            lp, _ = ens.value_and_grad(theta)
        finally:
            ens.close()
        return lp if batched else lp[0]


def merge_partition(input_output_layers: dict, hidden_layers: dict) -> dict:
    """{'fcn': {**input_output['fcn'], **hidden['fcn']}} (partition_sampling.py:117,144)."""
    return {'fcn': {**input_output_layers['fcn'], **hidden_layers['fcn']}}


_ADOPTED = weakref.WeakKeyDictionary()     # reference ProbabilisticModel object -> (n_features, adapter)


def _closure_vars(fn) -> dict:
    code, cells = getattr(fn, '__code__', None), getattr(fn, '__closure__', None)
    if code is None or not cells:
        return {}
    out = {}
    for name, cell in zip(code.co_freevars, cells):
        try:
            out[name] = cell.cell_contents
        except ValueError:
            pass
    return out


def adopt_reference_model(owner, x) -> 'ProbabilisticModel':
    """Describe the REFERENCE's own `src.training.probabilistic.ProbabilisticModel` (probabilistic.py:17-47) to the CUDA
    library without importing it: `module.config.hidden_structure / activation / use_bias` (src/config/models/fcn.py:8-30),
    `task`, `n_batches`, and the prior from `prior.name` plus the `loc` / `scale` its `log_prior` closure was built with
    (src/training/priors.py:100-128).  Lets `inference_loop` take the closure the reference's trainer builds
    (trainer.py:576-580) as it is."""
    n_features = int(np.shape(x)[-1])
    hit = _ADOPTED.get(owner)
    if hit is not None and hit[0] == n_features:
        return hit[1]
    cfg = getattr(getattr(owner, 'module', None), 'config', None)
    if cfg is None or not hasattr(cfg, 'hidden_structure') or not hasattr(cfg, 'activation'):
        raise TypeError('the log-posterior closure belongs to a model whose `module.config` is not an FCN configuration '
                        '(hidden_structure, activation): only the FCN of src/models/tabular/fcn.py runs on the CUDA path')
    act = cfg.activation
    module = FCN(tuple(cfg.hidden_structure), str(getattr(act, 'value', act)), bool(getattr(cfg, 'use_bias', True)))
    pr = owner.prior
    name = str(getattr(pr.name, 'value', pr.name))
    cv = _closure_vars(pr.log_prior)
    params = {} if name == 'StandardNormal' else {k: float(cv[k]) for k in ('loc', 'scale') if k in cv}
    from .priors import PriorDist
    prior = PriorDist(name).get_prior(**params)
    hs = module.hidden_structure
    dims = (n_features,) + hs
    shape_tree = {'fcn': {f'layer{i}': {'kernel': np.zeros((dims[i], dims[i + 1]), np.float32),
                                          'bias': np.zeros(dims[i + 1], np.float32)} for i in range(len(hs))}}
    task = owner.task
    pm = ProbabilisticModel(module, shape_tree, prior, task, n_batches=getattr(owner, 'n_batches', 1))
    try:
        _ADOPTED[owner] = (n_features, pm)
    except TypeError:
        pass
    return pm


def unwrap_posterior(fn):
    """Recognise `partial(prob_model.log_unnormalized_posterior, x=train_x, y=train_y)`
    (src/training/trainer.py:576-580) and return (prob_model, x, y).  `prob_model` is this package's ProbabilisticModel, or
    the reference's own one (recognised by its attributes and described through `adopt_reference_model`).  Anything else
    cannot be routed to the fused CUDA kernel and is rejected loudly (there is no tracing / CPU fallback)."""
    if isinstance(fn, functools.partial):
        target = fn.func
        owner = getattr(target, '__self__', None)
        kw = fn.keywords or {}
        if getattr(target, '__name__', '') in ('log_unnormalized_posterior', 'log_unnormalized_posterior_partition') \
                and 'x' in kw and 'y' in kw:
            if isinstance(owner, ProbabilisticModel):
                return owner, kw['x'], kw['y']
            if owner is not None and all(hasattr(owner, a) for a in ('module', 'prior', 'task', 'n_batches')):
                return adopt_reference_model(owner, kw['x']), kw['x'], kw['y']
    raise TypeError('mile_b200 samplers need `functools.partial(prob_model.log_unnormalized_posterior, x=..., y=...)` '
                    'of a ProbabilisticModel (this package\'s or the reference\'s): arbitrary Python log-densities cannot '
                    'run on the CUDA path and there is no CPU fallback')
