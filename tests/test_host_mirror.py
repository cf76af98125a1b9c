"""CPU tests of the host-side mirror of the reference's Python seams (no compute calls)."""
import functools
import json
from pathlib import Path

import numpy as np
import pytest
import torch
from scipy import stats

from oracle import mile_oracle as o

GOLDEN = Path(__file__).resolve().parent / 'golden'


def test_sampler_config_parses_every_reference_mclmc_yaml():
    """tests/golden/sampler_configs.json = the `training.sampler` blocks of the reference's MCLMC and NUTS YAMLs."""
    from mile_b200 import FCN, Sampler, SamplerConfig
    cfgs = json.loads((GOLDEN / 'sampler_configs.json').read_text())
    assert len(cfgs) >= 10
    for name, c in cfgs.items():
        sc = SamplerConfig.from_dict(c['sampler'])
        assert sc.name in (Sampler.MCLMC, Sampler.NUTS) and sc.warmup_steps > 0 and sc.n_thinning >= 1
        assert sc.prior.kind in ('normal', 'laplace')
        assert callable(sc.kernel)
        if c['model'] and c['model'].get('model') == 'FCN':
            FCN(tuple(c['model']['hidden_structure']), c['model']['activation'], c['model'].get('use_bias', True))
    assert sum(SamplerConfig.from_dict(c['sampler']).name == Sampler.NUTS for c in cfgs.values()) == 9
    nut = SamplerConfig.from_dict(cfgs['illustrative_example_readme/nuts.yaml']['sampler'])
    assert (nut.warmup_steps, nut.n_chains, nut.n_samples, nut.n_thinning) == (100, 12, 1000, 1)
    ill = SamplerConfig.from_dict(cfgs['illustrative_example_readme/mclmc.yaml']['sampler'])
    assert (ill.warmup_steps, ill.n_chains, ill.n_samples, ill.n_thinning) == (50000, 12, 10000, 10)
    assert (ill.desired_energy_var_start, ill.desired_energy_var_end, ill.step_size_init) == (0.5, 0.1, 0.01)
    with pytest.raises(ValueError):
        SamplerConfig.from_dict({'name': 'mclmc', 'not_a_field': 1})


def test_priors_match_scipy():
    from mile_b200 import PriorDist
    rng = np.random.default_rng(0)
    tree = {'fcn': {'layer0': {'kernel': rng.standard_normal((3, 4)).astype(np.float32),
                               'bias': rng.standard_normal(4).astype(np.float32)}}}
    flat = np.concatenate([tree['fcn']['layer0']['bias'], tree['fcn']['layer0']['kernel'].ravel()])
    p = PriorDist.StandardNormal.get_prior()
    assert abs(p.log_prior(tree) - stats.norm.logpdf(flat).sum()) < 1e-4
    p = PriorDist('Normal').get_prior(loc=0.5, scale=2.0)
    assert abs(p.log_prior(tree) - stats.norm.logpdf(flat, 0.5, 2.0).sum()) < 1e-4 and (p.kind, p.loc, p.scale) == ('normal', 0.5, 2.0)
    p = PriorDist('Laplace').get_prior(loc=0.1, scale=0.7)
    assert abs(p.log_prior(tree) - stats.laplace.logpdf(flat, 0.1, 0.7).sum()) < 1e-4 and p.kind == 'laplace'


def test_keys_and_split_are_deterministic():
    from mile_b200.types import key_to_seed, split
    assert key_to_seed(42) == 42
    assert key_to_seed(np.array([0, 42], np.uint32)) == 42
    a, b, c = split(42, 3)
    assert len({a, b, c}) == 3 and split(42, 3) == [a, b, c] and split(43, 3)[0] != a


def test_unwrap_posterior_accepts_only_the_reference_closure():
    from mile_b200 import FCN, PriorDist, ProbabilisticModel
    from mile_b200.probabilistic import unwrap_posterior
    rng = np.random.default_rng(0)
    module = FCN((16, 16, 2), 'relu')
    params = module.init(rng, 5)
    pm = ProbabilisticModel(module, params, PriorDist.StandardNormal.get_prior(), 'regr')
    assert pm.n_params == 5 * 16 + 16 + 16 * 16 + 16 + 16 * 2 + 2 and pm.n_features == 5
    X, y = rng.standard_normal((10, 5)), rng.standard_normal(10)
    m, x2, y2 = unwrap_posterior(functools.partial(pm.log_unnormalized_posterior, x=X, y=y))
    assert m is pm and x2 is X and y2 is y
    with pytest.raises(TypeError):
        unwrap_posterior(lambda position: 0.0)
    with pytest.raises(TypeError):
        unwrap_posterior(functools.partial(pm.log_unnormalized_posterior, x=X))
    from mile_b200 import KERNELS
    with pytest.raises(NotImplementedError):
        KERNELS['hmc'](None)
    assert callable(KERNELS['nuts']) and callable(KERNELS['mclmc'])


def test_param_io_layout_matches_reference(tmp_path):
    """warmstart/params_{i}.npz + samples/{chain}/sample_{n}.npz, members in leaf order
    'fcn.layer0.bias','fcn.layer0.kernel',... (src/training/callbacks.py:36-43, utils.py:69-175)."""
    from mile_b200 import FCN
    from mile_b200.callbacks import save_position
    from mile_b200.utils import load_params_batch, load_samples_from_dir, save_params
    rng = np.random.default_rng(1)
    module = FCN((16, 2))
    trees = [module.init(rng, 5) for _ in range(3)]
    for i, t in enumerate(trees):
        save_params(tmp_path / 'warmstart', t, i)
    with np.load(tmp_path / 'warmstart' / 'params_0.npz') as z:
        assert z.files == ['fcn.layer0.bias', 'fcn.layer0.kernel', 'fcn.layer1.bias', 'fcn.layer1.kernel']
    batch = load_params_batch([tmp_path / 'warmstart' / f'params_{i}.npz' for i in (2, 0, 1)])
    assert batch['fcn']['layer0']['kernel'].shape == (3, 5, 16)
    np.testing.assert_array_equal(batch['fcn']['layer1']['bias'][1], trees[1]['fcn']['layer1']['bias'])
    single = load_params_batch([tmp_path / 'warmstart' / 'params_1.npz'])
    assert single['fcn']['layer0']['kernel'].shape == (5, 16)
    for c in (0, 1):
        for n in (0, 10, 20):
            save_position(trees[c], tmp_path / 'samples', np.asarray(c), n)
    s = load_samples_from_dir(tmp_path / 'samples')
    assert s['fcn']['layer0']['kernel'].shape == (2, 3, 5, 16)     # (n_chains, n_samples, ...)
    with np.load(tmp_path / 'samples' / '1' / 'sample_10.npz') as z:
        assert z.files[0] == 'fcn.layer0.bias'


@pytest.mark.parametrize('n,dim', [(500, 3), (501, 2), (64, 5)])
def test_torch_ess_matches_oracle(n, dim):
    from mile_b200.diagnostics import effective_sample_size
    rng = np.random.default_rng(n)
    x = np.zeros((1, n, dim))
    e = rng.standard_normal((n, dim))
    for t in range(1, n):
        x[0, t] = 0.7 * x[0, t - 1] + e[t]
    want = np.atleast_1d(o.effective_sample_size(x))
    got = effective_sample_size(torch.from_numpy(x)).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-9)
    x2 = rng.standard_normal((3, n, dim))
    np.testing.assert_allclose(effective_sample_size(torch.from_numpy(x2)).numpy(),
                               np.atleast_1d(o.effective_sample_size(x2)), rtol=1e-9)


def test_metrics_match_oracle():
    from mile_b200 import metrics
    rng = np.random.default_rng(3)
    spec = o.ModelSpec(3, (4, 2), 'relu', 'regr')
    lv = rng.standard_normal((2, 6, 9, 2))
    y = rng.standard_normal(9)
    pw = metrics.pointwise_lppd(lv, y, 'regr')
    np.testing.assert_allclose(pw, o.pointwise_lppd(spec, lv, y), rtol=1e-12)
    assert abs(metrics.lppd(pw) - o.lppd(pw)) < 1e-12
    np.testing.assert_allclose(metrics.running_lppd(pw), o.running_lppd(pw), rtol=1e-12)


def test_shard_rows_partition_is_exact():
    """Row shards of the data-sharded variant (ShardedEnsemble.shard_rows) tile [0, N) without gaps or overlap."""
    from mile_b200.engine import ShardedEnsemble
    for n_rows in (1, 7, 232404, 12165):
        for world in (1, 2, 3, 8):
            covered = np.zeros(n_rows, np.int32)
            for r in range(world):
                sl = ShardedEnsemble.shard_rows(n_rows, r, world)
                covered[sl] += 1
            assert np.all(covered == 1)


def test_synthetic_workloads_are_seeded_and_shaped():
    from mile_b200 import synthetic as syn
    for name, (N, Nt, F, widths, act, task) in syn.CONFIGS.items():
        if N > 40000:
            continue
        X, y, Xt, yt = syn.synthetic_data(name)
        assert X.shape == (N, F) and Xt.shape == (Nt, F) and y.shape == (N,) and X.dtype == np.float32
        assert (y.dtype == np.float32) == (task == 'regr')
        X2, y2, _, _ = syn.synthetic_data(name)
        assert np.array_equal(X, X2) and np.array_equal(y, y2)
        spec = syn.workload_spec(name)
        assert spec.n_features == F and tuple(spec.widths) == tuple(widths)
        th = syn.synthetic_theta0(spec.n_params, 3)
        assert th.shape == (3, spec.n_params) and not np.array_equal(th[0], th[1])


def test_unwrap_posterior_adopts_the_reference_models_own_closure():
    """The reference's trainer builds `partial(self.prob_model.log_unnormalized_posterior, x=..., y=...)` from ITS
    ProbabilisticModel (trainer.py:576-580; probabilistic.py:17-47: attributes task, module, n_params, n_batches, prior;
    module.config = FCNConfig(hidden_structure, activation enum, use_bias), src/config/models/fcn.py:8-30; prior =
    NamedTuple(f_init, log_prior, name) whose log_prior closes over loc / scale, priors.py:60-128).  jax/flax are not
    installable here, so stand-ins with exactly that attribute structure are adopted."""
    import enum
    from typing import Callable, NamedTuple
    from mile_b200.probabilistic import ProbabilisticModel, unwrap_posterior

    class Activation(str, enum.Enum):
        RELU = 'relu'
        TANH = 'tanh'

    class PriorDist(str, enum.Enum):
        NORMAL = 'Normal'
        StandardNormal = 'StandardNormal'
        LAPLACE = 'Laplace'

    class Task(str, enum.Enum):
        REGRESSION = 'regr'
        CLASSIFICATION = 'class'

    class RefPrior(NamedTuple):
        f_init: Callable
        log_prior: Callable
        name: str

    def log_prior_laplace(loc=0.0, scale=1.0):
        def log_prior(params):
            return loc + scale        # (the real one calls jax; only its closure matters here)
        return log_prior

    class FCNConfig:
        def __init__(self, hs, act):
            self.hidden_structure, self.activation, self.use_bias = hs, act, True

    class RefFCN:
        def __init__(self, config):
            self.config = config

    class RefProbabilisticModel:
        def __init__(self, module, prior, task, n_batches=1):
            self.task, self.module, self.n_params, self.n_batches, self.prior = task, module, 0, n_batches, prior

        def log_unnormalized_posterior(self, position, x, y, **kwargs):
            raise AssertionError('the JAX log-density must never be called by the CUDA path')

    ref = RefProbabilisticModel(RefFCN(FCNConfig([16, 16, 2], Activation.TANH)),
                                RefPrior(None, log_prior_laplace(loc=0.25, scale=1.5), PriorDist.LAPLACE), Task.REGRESSION)
    X, y = np.zeros((7, 5), np.float32), np.zeros(7, np.float32)
    pm, x2, y2 = unwrap_posterior(functools.partial(ref.log_unnormalized_posterior, x=X, y=y))
    assert isinstance(pm, ProbabilisticModel) and x2 is X and y2 is y
    spec = pm.spec
    assert (spec.n_features, spec.widths, spec.activation, spec.task) == (5, (16, 16, 2), 'tanh', 'regr')
    assert (spec.prior, spec.prior_loc, spec.prior_scale, spec.n_batches) == ('laplace', 0.25, 1.5, 1.0)
    assert pm.n_params == 5 * 16 + 16 + 16 * 16 + 16 + 16 * 2 + 2
    # adopted once per model object
    assert unwrap_posterior(functools.partial(ref.log_unnormalized_posterior, x=X, y=y))[0] is pm
    std = RefProbabilisticModel(RefFCN(FCNConfig([8, 3], Activation.RELU)),
                                RefPrior(None, log_prior_laplace(), PriorDist.StandardNormal), Task.CLASSIFICATION)
    s2 = unwrap_posterior(functools.partial(std.log_unnormalized_posterior, x=X, y=y))[0].spec
    assert (s2.prior, s2.prior_loc, s2.prior_scale, s2.task, s2.widths) == ('normal', 0.0, 1.0, 'class', (8, 3))


def test_nuts_host_mirror_signatures_and_loud_failures():
    """warmup.py:27-36 / sampling.py:220-229 argument names are kept; what the CUDA path does not implement fails loudly
    before any device work (no CPU fallback)."""
    import inspect
    from mile_b200 import FCN, PriorDist, ProbabilisticModel, custom_window_adaptation, warmup_nuts
    from mile_b200.probabilistic import unwrap_posterior
    sig = list(inspect.signature(custom_window_adaptation).parameters)
    assert sig[:7] == ['algorithm', 'logdensity_fn', 'is_mass_matrix_diagonal', 'initial_step_size', 'target_acceptance_rate',
                       'progress_bar', 'saving_path']
    assert list(inspect.signature(warmup_nuts).parameters)[:8] == ['kernel', 'config', 'rng_key', 'init_params', 'step_ids',
                                                                    'unnorm_log_posterior', 'n_devices', 'saving_path']
    module = FCN((16, 2), 'relu')
    pm = ProbabilisticModel(module, module.init(np.random.default_rng(0), 5), PriorDist.StandardNormal.get_prior(), 'regr')
    lp = functools.partial(pm.log_unnormalized_posterior, x=np.zeros((4, 5), np.float32), y=np.zeros(4, np.float32))
    with pytest.raises(NotImplementedError):
        custom_window_adaptation(None, lp, is_mass_matrix_diagonal=False)
    assert callable(custom_window_adaptation(None, lp).run)

    class NotAnFCN:
        task, prior, n_batches = 'regr', None, 1
        module = object()

        def log_unnormalized_posterior(self, position, x, y):
            return 0.0
    with pytest.raises(TypeError):
        unwrap_posterior(functools.partial(NotAnFCN().log_unnormalized_posterior, x=np.zeros((2, 3)), y=np.zeros(2)))
