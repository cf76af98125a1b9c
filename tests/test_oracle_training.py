"""Pin of the warm-start training oracle (oracle/train_oracle.py) against an INDEPENDENT implementation: the same MLP,
Gaussian-NLL / cross-entropy loss and optimizer written with torch autograd and torch.optim (AdamW / Adam / SGD have the
update rule of optax.adamw / adam / sgd: decoupled decay lr * wd * theta, eps outside the square root, bias correction)."""
import numpy as np
import pytest
import torch

from oracle import mile_oracle as o
from oracle import train_oracle as t


def torch_model(spec, theta):
    b_off, k_off = spec.offsets()
    dims = spec.dims
    ps = []
    for l in range(len(spec.widths)):
        i, out = dims[l], dims[l + 1]
        W = torch.tensor(theta[k_off[l]:k_off[l] + i * out].reshape(i, out), dtype=torch.float64, requires_grad=True)
        b = torch.tensor(theta[b_off[l]:b_off[l] + out], dtype=torch.float64, requires_grad=True)
        ps.append((W, b))
    return ps


def torch_loss(spec, ps, X, y):
    h = torch.tensor(X, dtype=torch.float64)
    for l, (W, b) in enumerate(ps):
        h = h @ W + b
        if l < len(ps) - 1:
            h = torch.relu(h) if spec.activation == 'relu' else torch.sigmoid(h)
    if spec.task == 'regr':
        sigma = torch.exp(h[:, 1]).clamp(1e-6, 1e6).clamp(min=1e-5)
        yy = torch.tensor(y, dtype=torch.float64)
        return (0.5 * torch.log(2 * torch.pi * sigma ** 2) + (yy - h[:, 0]) ** 2 / (2 * sigma ** 2)).mean()
    return torch.nn.functional.cross_entropy(h, torch.tensor(y, dtype=torch.int64))


def flat(spec, ps):
    b_off, k_off = spec.offsets()
    out = np.zeros(spec.n_params)
    for l, (W, b) in enumerate(ps):
        out[k_off[l]:k_off[l] + W.numel()] = W.detach().numpy().ravel()
        out[b_off[l]:b_off[l] + b.numel()] = b.detach().numpy()
    return out


@pytest.mark.parametrize('task,kind', [('regr', 'adamw'), ('regr', 'adam'), ('regr', 'sgd'), ('class', 'adamw')])
def test_training_oracle_matches_torch(task, kind):
    K = 2 if task == 'regr' else 4
    spec = o.ModelSpec(6, (8, 5, K), 'relu' if task == 'regr' else 'sigmoid', task)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((96, 6))
    y = rng.standard_normal(96) if task == 'regr' else rng.integers(0, K, 96).astype(np.int32)
    th0 = rng.standard_normal(spec.n_params) * 0.4
    batches = rng.permutation(96).reshape(6, 16)
    opt = dict(lr=3e-3, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4)
    st = t.OptState(spec.n_params)
    th, mets = t.train_epoch(spec, th0.copy(), st, X, y, batches, kind=kind, **opt)
    ps = torch_model(spec, th0)
    params = [p for wb in ps for p in wb]
    topt = {'adamw': lambda: torch.optim.AdamW(params, lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4),
            'adam': lambda: torch.optim.Adam(params, lr=3e-3, betas=(0.9, 0.999), eps=1e-8),
            'sgd': lambda: torch.optim.SGD(params, lr=3e-3)}[kind]()
    for k, idx in enumerate(batches):
        topt.zero_grad()
        loss = torch_loss(spec, ps, X[idx], y[idx])
        assert abs(loss.item() - mets[k, 0]) <= 1e-9 * max(1.0, abs(loss.item()))
        loss.backward()
        topt.step()
    np.testing.assert_allclose(th, flat(spec, ps), rtol=1e-8, atol=1e-10)
    lossv, aux = t.eval_metrics(spec, th, X, y)
    assert abs(lossv - torch_loss(spec, ps, X, y).item()) <= 1e-9 * max(1.0, abs(lossv))


def test_earlystop_matches_reference_semantics():
    # trainer.py:920-938: stop when the last `patience` losses are all >= the loss before them
    L = np.array([[3.0, 2.0, 2.5, 2.6, 2.7], [3.0, 2.0, 1.9, 1.8, 1.7]])
    assert t.earlystop(L[:, :2], 3).tolist() == [False, False]
    assert t.earlystop(L, 3).tolist() == [True, False]
    from mile_b200.training import earlystop
    assert earlystop(L, 3).tolist() == [True, False]
