"""Chain-batched CPU restatement of the MCLMC sampling hot path on torch-CPU (fp32): mode (ii) of BASELINE.md section 3,
"batched over chains, single process, all host cores through BLAS".

TEST INFRASTRUCTURE / CPU BASELINE ONLY (same role and caveat as oracle/mile_oracle.py: PARITY UNPINNED).  Nothing under
mile_b200/ imports this file.  It exists because the row-at-a-time C port (oracle/mile_oracle.c, one chain per OpenMP
thread) leaves most of a many-core host idle and runs each chain well below what a GEMM library reaches; this variant
expresses every layer of all chains as one batched GEMM ([C, N, in] x [C, in, out]) the way XLA-CPU sees the reference's
`jax.vmap`-free per-device program, and lets torch's intra-op thread pool use every core.  bench.py times both modes
and reports the faster one.

Follows, function by function:
  value_and_grad     src/training/probabilistic.py:92-138, src/training/priors.py:101-128,
                     src/flax_building_blocks/basic.py:41-61 (hand-differentiated, like the other oracles)
  esh_update / step  blackjax 1.2.2 mcmc/integrators.py + mcmc/mclmc.py
  run_sampling       src/training/sampling.py:134-177
Checked against oracle/mile_oracle.py in tests/test_oracle_torch_batched.py.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

B1 = 0.1931833275037836


def _views(spec, theta: torch.Tensor):
    """theta [C, d] -> per layer (W [C, in, out], b [C, 1, out]) views in ravel_pytree order."""
    C = theta.shape[0]
    b_off, k_off = spec.offsets()
    dims = spec.dims
    out = []
    for l in range(len(dims) - 1):
        i, o = dims[l], dims[l + 1]
        out.append((theta[:, k_off[l]:k_off[l] + i * o].view(C, i, o), theta[:, b_off[l]:b_off[l] + o].view(C, 1, o)))
    return out


def _act(name: str, z: torch.Tensor):
    if name == 'relu':
        a = torch.relu(z)
        return a, (z > 0).to(z.dtype)
    if name == 'sigmoid':
        a = torch.sigmoid(z)
        return a, a * (1 - a)
    if name == 'tanh':
        a = torch.tanh(z)
        return a, 1 - a * a
    if name == 'leaky_relu':
        return torch.where(z >= 0, z, 0.01 * z), torch.where(z >= 0, torch.ones_like(z), torch.full_like(z, 0.01))
    if name == 'identity':
        return z, torch.ones_like(z)
    raise NotImplementedError(name)


def value_and_grad(spec, theta: torch.Tensor, X: torch.Tensor, y: torch.Tensor):
    """theta [C, d] -> (log-posterior [C], gradient [C, d]) for all chains at once."""
    C, d = theta.shape
    N = X.shape[-2]
    layers = _views(spec, theta)
    # X may come pre-expanded [C, N, F] (contiguous): batched GEMMs on a stride-0 batch axis take a very slow path
    h = X if X.dim() == 3 else X.unsqueeze(0).expand(C, N, X.shape[1]).contiguous()
    acts, dacts = [h], []
    for l, (W, b) in enumerate(layers):
        z = torch.baddbmm(b, h, W)
        if l < len(layers) - 1:
            h, da = _act(spec.activation, z)
            acts.append(h); dacts.append(da)
    out = z
    if spec.task.startswith('regr'):
        mu, s = out[..., 0], out[..., 1]
        e = torch.exp(s)
        sigma = e.clamp(1e-6, 1e6)
        inside = ((e > 1e-6) & (e < 1e6)).to(out.dtype)
        s2 = sigma * sigma
        res = y.unsqueeze(0) - mu
        q = res * res / s2
        ll = -(torch.log(2 * math.pi * s2) + q) / 2
        bad = torch.isnan(ll)
        delta = torch.zeros_like(out)
        delta[..., 0] = torch.where(bad, torch.zeros_like(q), res / s2)
        delta[..., 1] = torch.where(bad, torch.zeros_like(q), (q - 1) * inside)
        ll = torch.where(bad, torch.zeros_like(ll), ll).sum(1)
    else:
        lsm = torch.log_softmax(out, dim=-1)
        yi = y.to(torch.int64).unsqueeze(0).expand(C, N).unsqueeze(-1)
        llr = lsm.gather(-1, yi)[..., 0]
        bad = torch.isnan(llr)
        delta = -torch.exp(lsm)
        delta.scatter_add_(-1, yi, torch.ones_like(llr).unsqueeze(-1))
        delta = torch.where(bad.unsqueeze(-1), torch.zeros_like(delta), delta)
        ll = torch.where(bad, torch.zeros_like(llr), llr).sum(1)
    delta = delta * spec.n_batches
    grad = torch.empty_like(theta)
    b_off, k_off = spec.offsets()
    dims = spec.dims
    for l in range(len(layers) - 1, -1, -1):
        i, o = dims[l], dims[l + 1]
        grad[:, k_off[l]:k_off[l] + i * o] = torch.bmm(acts[l].transpose(1, 2), delta).reshape(C, i * o)
        grad[:, b_off[l]:b_off[l] + o] = delta.sum(1)
        if l > 0:
            delta = torch.bmm(delta, layers[l][0].transpose(1, 2)) * dacts[l - 1]
    loc, sc = spec.prior_loc, spec.prior_scale
    dlt = theta - loc
    if spec.prior.lower() in ('normal', 'standardnormal'):
        pv = (-(math.log(2 * math.pi * sc * sc) + dlt * dlt / (sc * sc)) / 2).sum(1)
        grad = grad - dlt / (sc * sc)
    else:
        pv = (-math.log(2 * sc) - dlt.abs() / sc).sum(1)
        grad = grad - torch.sign(dlt) / sc
    return pv + ll * spec.n_batches, grad


def _esh(u, g, eps, coef):
    """B-step for all chains (literal blackjax formula); returns (u', kinetic-energy change [C])."""
    d = u.shape[1]
    gn = g.norm(dim=1, keepdim=True)
    e = g / gn.clamp_min(1e-13)
    p = (u * e).sum(1, keepdim=True)
    delta = eps * coef * gn / (d - 1)
    zeta = torch.exp(-delta)
    raw = e * (1 - zeta) * (1 + zeta + p * (1 - zeta)) + 2 * zeta * u
    un = raw / raw.norm(dim=1, keepdim=True)
    dk = (delta - math.log(2.0) + torch.log(1 + p + (1 - p) * zeta * zeta)) * (d - 1)
    return un, dk[:, 0]


def mclmc_step(f, theta, u, lp, g, eps, L, z):
    """One MCLMC kernel step for all chains; eps, L [C, 1]; z [C, d] normal draws."""
    d = theta.shape[1]
    u, dk = _esh(u, g, eps, B1)
    theta = theta + 0.5 * eps * u
    lp1, g = f(theta)
    u, dk2 = _esh(u, g, eps, 1 - 2 * B1)
    theta = theta + 0.5 * eps * u
    lp2, g = f(theta)
    u, dk3 = _esh(u, g, eps, B1)
    nu = torch.sqrt((torch.exp(2 * eps / L) - 1) / d)
    w = u + nu * z
    u = w / w.norm(dim=1, keepdim=True)
    dK = dk + dk2 + dk3
    return theta, u, lp2, g, (lp2, dK, dK - lp2 + lp)


def run_sampling_timed(spec, X, y, theta0, n_steps, eps, L, threads: int | None = None, seed: int = 0):
    """n_steps MCLMC steps for all chains (noise from torch's CPU generator).  Returns (seconds, final theta)."""
    if threads:
        torch.set_num_threads(int(threads))
    with torch.no_grad():
        Xt = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
        yt = torch.from_numpy(np.ascontiguousarray(y))
        yt = yt.to(torch.float32) if spec.task.startswith('regr') else yt.to(torch.int64)
        theta = torch.from_numpy(np.ascontiguousarray(theta0, dtype=np.float32)).clone()
        C, d = theta.shape
        Xt = Xt.unsqueeze(0).expand(C, *Xt.shape).contiguous()     # closure constant, replicated once per chain
        f = lambda t: value_and_grad(spec, t, Xt, yt)
        gen = torch.Generator().manual_seed(seed)
        lp, g = f(theta)
        z0 = torch.randn(C, d, generator=gen)
        u = z0 / z0.norm(dim=1, keepdim=True)
        e = torch.full((C, 1), float(eps)) if np.isscalar(eps) else torch.as_tensor(eps, dtype=torch.float32).view(C, 1)
        Lt = torch.full((C, 1), float(L)) if np.isscalar(L) else torch.as_tensor(L, dtype=torch.float32).view(C, 1)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            theta, u, lp, g, _ = mclmc_step(f, theta, u, lp, g, e, Lt, torch.randn(C, d, generator=gen))
        dt = time.perf_counter() - t0
    return dt, theta.numpy()
