#!/usr/bin/env python
"""bench.py -- chain-steps/s of the MCLMC ensemble sampling hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement
                                                             # of the reference (oracle/), all host threads

A bench "step" is ONE launch of the persistent sampler kernel: `inner` MCLMC steps for every chain of the ensemble
wave (two full-batch gradient evaluations per chain-step), thinned samples captured in HBM and the test-set
posterior-predictive logsumexp folded in at every kept sample.  Default workload: `airfoil_3x16`, 12 chains -- the
configuration BASELINE.json's >= 50x target is quoted on (experiments/illustrative_example_readme/mclmc.yaml:31-47 of the
reference: 12 chains, FCN [16,16,16,2], 1052 x 5 training matrix), synthetic data of that shape.
With --gpus N every rank runs its own block of chains on the same split (weak scaling: the ensemble shards by chain
with no data-path collective; the per-chain test-set logsumexp states are merged once at the end over NCCL, timed
inside `e2e`).  At N = 1 the line also carries `extra.workloads`: the other named shapes of BASELINE.json, each timed the
same way on a shorter run; at N > 1 it carries the row-sharded covertype step loop (strong scaling, NCCL all-reduce per
gradient evaluation) with its parity against a single-GPU evaluation.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (config key, chains per GPU, inner steps per launch)
    'airfoil_3x16': ('airfoil_3x16', 12, 2000),
    'bikesharing_2x16': ('bikesharing_2x16', 10, 500),
    'protein_2x16': ('protein_2x16', 10, 200),
    'covertype_ref': ('covertype_ref', 12, 500),
    'airfoil_3x16_1024': ('airfoil_3x16', 1024, 200),
    'wide_4x256': ('wide_4x256', 8, 10),          # HBM-resident chain-batched GEMM path
    # BASELINE.json configs[2]: every rank holds all 12 chains and 1/N of the 232 404 rows; one NCCL all-reduce of the
    # packed [C, d+1] (gradient, log-lik) per evaluation.  Strong scaling: total work is fixed as N grows.
    'covertype_full': ('covertype_full', 12, 20),
}
DEFAULT_WORKLOAD = 'airfoil_3x16'
EXTRA_WORKLOADS = ('bikesharing_2x16', 'protein_2x16', 'airfoil_3x16_1024', 'wide_4x256')
N_THINNING = 10  # every reference MCLMC YAML (experiments/*/mclmc.yaml: n_thinning 10)


def load_tensor_peak():
    """Measured dense bf16 tensor throughput (MEASURED_PEAKS.json), else the profiling recipe's fallback."""
    try:
        pk = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())
        for k in ('bf16_tflops', 'bf16_tflops_sustained', 'tensor_bf16_tflops'):
            if k in pk:
                return float(pk[k]), f'MEASURED_PEAKS.json {k}'
    except Exception:
        pass
    return 1590.0, 'fallback (B200_PROFILING.md)'


def load_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        j = json.loads(p.read_text())
        return j.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def load_traffic(workload: str, inner: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the sampler kernel, from the committed ncu capture of
    this command (profiles/traffic.json: {workload: {inner, bytes, source}}); None when no capture matches."""
    try:
        t = json.loads((ROOT / 'profiles' / 'traffic.json').read_text()).get(workload)
        if t and int(t.get('inner', -1)) == int(inner):
            return int(t['bytes']), t.get('source')
    except Exception:
        pass
    return None, None


def flops_per_chain_step(n_rows, dims):
    """12*N*W FLOPs per chain-step (two fwd+bwd passes, W = sum in*out)."""
    W = sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    return 12.0 * n_rows * W


def flops_executed_per_chain_step(n_rows, dims):
    """What a gradient evaluation actually needs: no input gradient for layer 0 (12 N W minus 4 N F h1 per step)."""
    return flops_per_chain_step(n_rows, dims) - 4.0 * n_rows * dims[0] * dims[1]


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement.  Three ways of using the host are tried and the fastest is reported:
#   chains       C port (oracle/mile_oracle.c), one chain per OpenMP thread -- the reference's one-virtual-XLA-device-
#                per-chain pmap (train.py:16, src/training/sampling.py:181-184)
#   chains+rows  the same with the spare cores splitting each chain's rows (hosts with more cores than chains)
#   batched      oracle/torch_batched.py: every layer of all chains as one batched GEMM on torch-CPU, all cores through
#                the BLAS torch links (mode ii)
# ----------------------------------------------------------------------------------------------
def _cpu_modes(C, cores):
    par = min(C, cores)
    modes = [('chains', dict(threads=par, row_threads=1))]
    if cores // par >= 2:
        modes.append(('chains+rows', dict(threads=par, row_threads=cores // par)))
    modes.append(('batched', None))
    return modes


def cpu_chain_steps_per_s(workload: str, budget_s: float = 12.0, max_steps: int = 200000, waves: int = 1,
                          fixed_steps: int | None = None, only_mode: str | None = None):
    """Times the CPU restatement of the reference on a bounded sample of the same workload: all chains of the job, as
    many MCLMC steps as fit the budget (or `fixed_steps`), on all the host threads the fastest mode can use."""
    from oracle import mile_oracle as o
    key, C, _ = WORKLOADS[workload]
    C = C * max(1, waves)          # reference arm at N GPUs: the same N x C chains our arm runs (weak scaling)
    spec = o.make_spec(key)
    X, y, _, _ = o.synthetic_data(key)
    th0 = o.synthetic_theta0(spec, C, scale=0.3 if key != 'wide_4x256' else 0.05)
    cores = os.cpu_count() or 1
    eps, L = 0.02, float(np.sqrt(spec.n_params))
    try:
        from oracle import c_oracle
        have_c = c_oracle.available()
    except Exception:
        have_c = False

    def run(mode, kw, n):
        t0 = time.perf_counter()
        if mode == 'batched':
            from oracle import torch_batched as tb
            dt, _ = tb.run_sampling_timed(spec, X, y, th0, n, eps, L, threads=cores)
            return dt
        c_oracle.run_sampling_timed(spec, X, y, th0, n, eps, L, **kw)
        return time.perf_counter() - t0

    rates = {}
    for mode, kw in _cpu_modes(C, cores):
        if only_mode and mode != only_mode:
            continue
        if mode != 'batched' and not have_c:
            continue
        if mode == 'batched' and spec.n_params > 50000:
            continue                          # activations of all chains at once would not fit a sensible budget
        try:
            n_cal = 2
            dt = run(mode, kw, n_cal)
            if dt < 0.2:                      # very fast workloads: calibrate on a longer run
                n_cal = int(min(200, max(4, 0.5 / max(dt / 2, 1e-6))))
                dt = run(mode, kw, n_cal)
            rates[mode] = C * n_cal / dt
        except Exception as e:   # noqa: BLE001
            rates[mode] = 0.0
            print(f'[bench] cpu mode {mode} failed: {e!r}', file=sys.stderr)
    if not rates or max(rates.values()) <= 0:
        # numpy port, single process (no C compiler on the host)
        f = lambda t: o.logpost_value_and_grad(spec, t, X, y)
        rng = np.random.default_rng(0)
        d = spec.n_params
        st = o.mclmc_init(f, th0[0], rng.standard_normal(d).astype(np.float32))
        n, t0 = 0, time.perf_counter()
        while n < max_steps and time.perf_counter() - t0 < budget_s:
            st, _ = o.mclmc_step(f, st, eps, L, rng.standard_normal(d).astype(np.float32))
            n += 1
        dt = time.perf_counter() - t0
        return {'value': n / dt, 'unit': 'chain-steps/s', 'cores': 1, 'kind': 'port', 'mode': 'numpy', 'n_steps': n,
                'seconds': dt, 'sample': f'1 chain x {n} MCLMC steps of {workload} (numpy restatement)'}
    best = max(rates, key=rates.get)
    kw = dict(_cpu_modes(C, cores))[best]
    n = fixed_steps or int(max(2, min(max_steps, budget_s * rates[best] / C)))
    dt = run(best, kw, n)
    used = cores if best == 'batched' else kw['threads'] * kw['row_threads']
    desc = {'chains': 'C restatement, OpenMP, one chain per thread',
            'chains+rows': f'C restatement, OpenMP, one chain per thread x {kw["row_threads"] if kw else 0} threads sharing its rows',
            'batched': 'torch-CPU restatement, chain-batched GEMMs on all cores'}[best]
    return {'value': C * n / dt, 'unit': 'chain-steps/s', 'cores': used, 'host_cores': cores, 'kind': 'port',
            'mode': best, 'modes_tried': {k: round(v, 1) for k, v in rates.items()}, 'n_steps': n, 'seconds': dt,
            'sample': f'{C} chains x {n} MCLMC steps of {workload} ({desc}; fastest of the modes tried)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    W = args.workload
    key, C, inner = WORKLOADS[W]
    from oracle import mile_oracle as o
    spec = o.make_spec(key)
    Ntr = o.CONFIGS[key][0]
    waves = 1 if key == 'covertype_full' else max(1, args.gpus)   # row-sharded workload: the chains do not multiply
    total = args.warmup + args.steps
    # one reference "step" = a bounded sample of the workload: ref_inner MCLMC steps of all chains, sized from a short
    # calibration so that the whole --steps/--warmup run ends within about two minutes
    cal = cpu_chain_steps_per_s(W, budget_s=3.0, waves=waves)
    ref_inner = int(max(2, min(inner, cal['value'] * (90.0 / max(1, total)) / (C * waves))))
    secs, last = [], cal
    for i in range(total):
        r = cpu_chain_steps_per_s(W, waves=waves, fixed_steps=ref_inner, only_mode=cal.get('mode'))
        if i >= args.warmup:
            secs.append(r['seconds'])
        last = r
    t = float(np.sum(secs))
    v = C * waves * ref_inner * len(secs) / t
    line = {
        'impl': 'reference', 'metric': 'chain-steps/sec', 'value': v, 'unit': 'chain-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / max(1, len(secs)), 'higher_is_better': True,
        'scaling': 'strong' if key == 'covertype_full' else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': W, 'chains_per_gpu': C, 'chains_total': C * waves, 'inner_steps_per_launch': ref_inner,
                   'n_train': Ntr, 'n_params': spec.n_params,
                   'note': 'CPU restatement of the reference (JAX/BlackJAX not installable here); all chains of the N-GPU job; '
                           'a reference step is a bounded sample of ref_inner MCLMC steps (not the GPU arm\'s inner count)'},
        'grad_evals_per_s': 2 * v,
        'cpu_baseline': dict(last, value=v, modes_tried=cal.get('modes_tried')),
        'e2e': {'value': v, 'unit': 'chain-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
MMA_PER_TILE = {(4, 8): 84, (4, 16): 96, (3, 8): 48, (3, 16): 60}   # 3xTF32 mma.sync per 16-row tile and evaluation


def measure(W, args, steps, warmup, rank, world, local, with_cpu, peaks):
    """Times one workload; returns the dict of its bench line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from mile_b200 import Ensemble
    from mile_b200 import synthetic as syn   # seeded synthetic inputs (the oracle is only used by the cpu_baseline leg)
    from mile_b200.distributed import merge_lppd_states

    key, C, inner = WORKLOADS[W]
    if args.inner and W == args.workload:
        inner = args.inner
    spec = syn.workload_spec(key)
    X, y, Xt, yt = syn.synthetic_data(key, seed=1234)   # same split on every rank; the CHAINS are what shards
    d = spec.n_params
    sharded = key == 'covertype_full'
    dev = torch.device(f'cuda:{local}')
    if sharded:
        from mile_b200 import ShardedEnsemble
        ens = ShardedEnsemble(spec, C, device=local, rank=rank, world=world)
        rows = ShardedEnsemble.shard_rows(X.shape[0], rank, world)
        Xfull, yfull = X, y
        X, y = np.ascontiguousarray(X[rows]), np.ascontiguousarray(y[rows])
    else:
        ens = Ensemble(spec, C, device=local)
        ens.set_option('chain_base', rank * C)      # global chain ids: every rank draws its own noise streams
    if W == args.workload:
        if args.cluster:
            ens.set_option('cluster_size', args.cluster)
        if args.sync_mode >= 0:
            ens.set_option('sync_mode', args.sync_mode)
        if args.tile_rows:
            ens.set_option('tile_rows', args.tile_rows)
        if args.tensor >= 0:
            ens.set_option('tensor', args.tensor)
        if args.fast >= 0:
            ens.set_option('fast', args.fast)
        if args.steploop >= 0:
            ens.set_option('steploop', args.steploop)
        if args.kslices > 0:
            ens.set_option('kslices', args.kslices)
    ens.set_data(X, y)
    wide = bool(ens.get_option('wide'))
    fused_lppd = not sharded          # (the wide path folds through its forward GEMMs, the others inside the sampler kernel)
    if fused_lppd:
        ens.set_test(Xt, yt)      # fused posterior-predictive LPPD fold at every kept sample
    crank = 0 if sharded else rank        # sharded: every rank carries the SAME chains (same seeds, same noise)
    th0 = syn.synthetic_theta0(d, C, seed0=1000 + 100 * crank, scale=0.3 if key != 'wide_4x256' else 0.05)
    ens.init(th0, seed=17)
    # short tuning run -> frozen (eps, L); fallback eps=0.02, L=sqrt(d)
    eps = np.full(C, 0.02, np.float32)
    L = np.full(C, np.sqrt(d), np.float32)
    if not args.no_tune:
        ens.tune_reset(0.01)
        nt1, nt2 = (40, 10) if (wide or sharded) else (800, 100)
        tc = ens.tune_cfg(nt1, nt2, 0.5, 0.1, 1.5, 100)
        ens.tune(nt1 + nt2, 0, tc, seed=99)
        ens.tune_finish_phase2()
        e, l, _ = ens.get_tuning()
        if np.all(np.isfinite(e)) and np.all(e > 0) and np.all(np.isfinite(l)) and np.all(l > 0):
            eps, L = e, l
    eps_d, L_d = torch.from_numpy(eps).to(dev), torch.from_numpy(L).to(dev)
    n_slots = inner // N_THINNING
    samples_d = torch.empty((n_slots, C, d), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def one_step(i):
        ens.sample_device(inner, eps_d, L_d, step_base=i * inner, n_thinning=N_THINNING,
                          sample_base=i * n_slots, seed=1234, samples_dev=samples_d, n_slots=n_slots, lppd=fused_lppd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        flush.zero_()
        one_step(i)
    barrier()
    if fused_lppd:
        ens.lppd_reset()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = ens.launches
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step(warmup + i)
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    launches = ens.launches - l0
    clk = clocks.stop()
    t_dev = sum(kernel_ms) * 1e-3
    if world > 1:
        t = torch.tensor([t_dev], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev = float(t.item())
    nrep = 1 if sharded else world        # sharded: the ranks share one set of chains
    value = nrep * C * inner * steps / t_dev
    lppd_val, lppd_total = None, 0
    if fused_lppd:
        m_, s_, cnt_ = ens.lppd_state()
        lppd_val, lppd_total = merge_lppd_states(m_, s_, cnt_, device=dev)

    sharded_parity = None
    if sharded and world > 1:
        # rank 0 compares the N-rank gradient state with a single-GPU full-data evaluation at the same positions
        th_s, _, lp_s, g_s = ens.get_state()
        if rank == 0:
            ref = Ensemble(spec, C, device=local)
            ref.set_data(Xfull, yfull)
            lp_f, g_f = ref.value_and_grad(th_s)
            ref.close()
            sharded_parity = {'grad_rel': float(np.linalg.norm(g_s - g_f) / np.linalg.norm(g_f)),
                              'logdensity_rel': float(np.max(np.abs(lp_s - lp_f) / np.abs(lp_f)))}
        barrier()      # (the ranks of the fused step loop wait on each other inside the kernel: start the next phase together)

    # ---- e2e: the same metric through the host-buffer C-ABI call (H2D + D2H inside), plus the one exchange step of the
    #      path at N > 1: the NCCL merge of the per-chain test-set logsumexp states ------------------------------------
    st = ens.get_state()
    Xp = torch.from_numpy(X).pin_memory().numpy()
    h2d = X.nbytes + y.nbytes + 3 * C * d * 4 + C * 4 + 2 * C * 4
    d2h = n_slots * C * d * 4 + 3 * C * d * 4 + C * 4
    e2e_steps = max(1, min(steps, 5))
    # one untimed pass first: the host-buffer entry points allocate their pinned staging buffers on first use
    ens.set_data(Xp, y); ens.set_state(*st)
    ens.sample(inner, eps, L, step_base=0, n_thinning=N_THINNING, seed=4320, lppd=False)
    ens.set_state(*st)
    # the step's result is read back into pinned host memory (the sample tensor is the bulk of the D2H bytes)
    out_pinned = None if sharded else torch.empty(max(1, -(-inner // N_THINNING)) * C * d, dtype=torch.float32, pin_memory=True).numpy()
    skw = {} if sharded else {'out': out_pinned}
    pin = lambda *shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    st_pinned = (pin(C, d), pin(C, d), pin(C), pin(C, d))        # chain state travels through pinned buffers as well
    for dst, src in zip(st_pinned, st):
        dst[...] = src
    st = st_pinned
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ens.set_data(Xp, y)
        ens.set_state(*st)
        smp, _ = ens.sample(inner, eps, L, step_base=0, n_thinning=N_THINNING, seed=4321 + i, lppd=fused_lppd, **skw)
        st = ens.get_state(out=st_pinned)
    if fused_lppd and world > 1:
        m_, s_, cnt_ = ens.lppd_state()
        merge_lppd_states(m_, s_, cnt_, device=dev)
        d2h += 2 * m_.nbytes // max(1, e2e_steps)
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = nrep * C * inner * e2e_steps / t_e2e
    finite = bool(np.all(np.isfinite(smp)))

    line = None
    if rank == 0:
        Ntr = Xfull.shape[0] if sharded else X.shape[0]
        fl = flops_per_chain_step(Ntr, spec.dims)
        per_gpu_steps_per_s = value / world     # (sharded: each GPU does 1/world of every chain-step's rows)
        achieved_tf = per_gpu_steps_per_s * fl / 1e12
        fast = int(ens.get_option('fast'))
        G = ens.get_option('cluster_size')
        resident = bool(ens.get_option('resident'))
        mma_path = fast == 2 and not wide and not sharded
        if sharded:
            path = ('data-sharded: ONE persistent kernel per rank for the whole step loop; per gradient evaluation a local reduce-scatter through L2, then every CTA pushes its summed slice as flagged 8-byte words into every rank\'s exchange region (CUDA IPC peer memory over NVLink) and polls its own copy'
                    if ens.get_option('shard_fused') == 1 else
                    'data-sharded: gradient kernel -> integrator kernel that sums the ranks\' [C, d+1] partials out of peer memory (CUDA IPC over NVLink), twice per step'
                    if ens.get_option('p2p') == 1 else
                    'data-sharded: gradient kernel -> ncclAllReduce([C, d+1]) -> integrator kernel, twice per step')
        elif wide:
            path = 'wide: HBM-resident chain-batched GEMMs, ' + ('tcgen05 3xTF32' if ens.get_option('tensor') else 'FP32 SIMT')
        else:
            path = {2: 'MmaGE: register-chained 3xTF32 mma.sync evaluator + integrator-warp step loop',
                    1: 'FastGE: warp-specialised FFMA layer pipeline', 0: 'GenericGE: 4x4 register tiles'}[fast]
        hbm_bytes_per_launch = C * (4 * d * 4 + n_slots * d * 4)        # state r/w + kept samples
        hbm_peak, peak_src = peaks['hbm']
        line = {
            'metric': 'chain-steps/sec', 'value': value, 'unit': 'chain-steps/s', 'n_gpus': world,
            'steps': steps, 'warmup': warmup, 'ms_per_step': 1e3 * t_dev / steps,
            'higher_is_better': True, 'scaling': 'strong' if sharded else 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': W, 'chains_per_gpu': C, 'inner_steps_per_launch': inner, 'n_train': int(Ntr),
                       'n_features': spec.n_features, 'hidden_structure': list(spec.widths), 'n_params': d,
                       'n_thinning': N_THINNING, 'noise': 'in-kernel Philox4x32-10, streams keyed by global chain id',
                       'cluster_size': G, 'sync_mode': ens.get_option('sync_mode'), 'tile_rows': ens.get_option('tile_rows'),
                       'x_resident_in_smem': resident, 'kernel_path': path, 'peer_memory_allreduce': bool(ens.get_option('p2p') == 1), 'fused_multi_rank_step_loop': bool(ens.get_option('shard_fused') == 1),
                       'l2': 'flushed between timed iterations (256 MiB write); ' +
                             ('working set is SMEM-resident' if resident else 'X streams from the padded HBM copy through L2 every evaluation'),
                       'step_size_mean': float(eps.mean()), 'L_mean': float(L.mean()),
                       'parallelism': (f'rows x{world} + all-reduce of [C, d+1] per gradient evaluation (peer memory over NVLink inside the kernel; NCCL when the peer mapping is unavailable)' if sharded else f'chains x{world}')},
            'grad_evals_per_s': 2 * value,
            'samples_finite': finite,
            'lppd': {'value': lppd_val, 'samples': int(lppd_total),
                     'note': 'test-set logsumexp folded in-kernel at every kept sample; merged across ranks over NCCL (timed inside e2e)'},
            'wall_s_timed_region': t_wall,
            'gpu_launches': int(launches),
            'clocks': clk,
            'e2e': {'value': e2e_value, 'unit': 'chain-steps/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps,
                    'path': 'Ensemble.set_data + set_state + sample (mile_*_host C-ABI calls, host numpy buffers; X, the chain state and the kept-sample buffer are pinned) + get_state'
                            + (' + NCCL merge of the LPPD states' if fused_lppd and world > 1 else '')},
        }
        if sharded_parity is not None:
            line['sharded_parity_rel'] = sharded_parity
        traffic, tsrc = load_traffic(W, inner)
        roof = {'bound': 'fp32', 'achieved': achieved_tf, 'peak': peaks['fp32'], 'unit': 'TFLOP/s',
                'frac': achieved_tf / peaks['fp32'] if peaks['fp32'] else None, 'traffic': traffic,
                'kernel': 'mile_mma_step_kernel' if mma_path else 'mile_mclmc_kernel',
                'flops_per_chain_step': fl, 'flops_executed_per_chain_step': flops_executed_per_chain_step(Ntr, spec.dims),
                'peak_source': peaks['fp32_src'],
                'note': 'SURVEY.md 8(d): these shapes are bound by on-chip arithmetic / latency (working set in SMEM/L2), '
                        'neither by HBM nor by the tcgen05 pipe; achieved = algorithmic fp32 FLOPs (12 N W per chain-step) '
                        'against the measured FP32 FMA peak; the HBM view is given for completeness',
                'avg_launch_ms': float(np.mean(kernel_ms)),
                'hbm': {'achieved_gbs': hbm_bytes_per_launch / (np.mean(kernel_ms) * 1e-3) / 1e9,
                        'peak_gbs': hbm_peak, 'peak_source': peak_src,
                        'note': 'tiny by design: theta/u/g and the X slice stay in shared memory for the whole launch'}}
        if traffic is not None:
            roof['traffic_note'] = tsrc
        if mma_path:
            NL, FP = len(spec.widths), (8 if spec.n_features <= 8 else 16)
            g_ = max(1, int(G))
            per = -(-int(Ntr) // g_)
            tiles = sum(-(-max(0, min(per, Ntr - r * per)) // 16) for r in range(g_))
            mma_flops = 2.0 * tiles * MMA_PER_TILE[(NL, FP)] * 2048.0          # per chain-step (two evaluations)
            roof['tensor_pipe'] = {'instruction': 'mma.sync.m16n8k8 tf32 (3 per fp32 product)',
                                   'issued_tflops': per_gpu_steps_per_s * mma_flops / 1e12, 'peak_tflops': peaks['mma'],
                                   'frac': per_gpu_steps_per_s * mma_flops / 1e12 / peaks['mma'] if peaks['mma'] else None,
                                   'peak_source': 'live mma.sync micro-benchmark (mile_measure_fp32_peak variant 2)'}
        if wide and ens.get_option('tensor'):
            tpk, tsrc2 = load_tensor_peak()
            roof.update({'bound': 'tensor', 'peak': tpk, 'frac': achieved_tf / tpk, 'peak_source': tsrc2,
                         'kernel': 'wide_gemm_tc2_kernel (tcgen05 kind::tf32, 3 MMAs per product for fp32-level accuracy)',
                         'note': 'achieved = ALGORITHMIC fp32 FLOPs (12 N W per chain-step); the 3xTF32 split issues 3x that '
                                 'on the tensor pipe, and TF32 dense peak is half the bf16 figure used as denominator'})
        line['roofline'] = roof
        if with_cpu:
            line['cpu_baseline'] = cpu_chain_steps_per_s(W, budget_s=with_cpu)
    ens.close()
    del flush, samples_d
    torch.cuda.empty_cache()
    return line


def measure_nuts(local: int, workload: str = DEFAULT_WORKLOAD, chains: int = 12, n_warm: int = 300, n_samp: int = 100):
    """The NUTS branch (csrc/mile_nuts.cuh) on the default workload: window adaptation, then sampling transitions with
    in-kernel Philox noise, timed through the host-buffer ABI call (wall clock around a synchronous call).  Reported as
    gradient evaluations per second because a transition's length is data dependent."""
    import numpy as np
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200 import synthetic as syn
    from mile_b200.nuts import build_schedule
    spec = syn.workload_spec(workload)
    X, y, _, _ = syn.synthetic_data(workload, seed=1234)
    fs = FCNSpec(spec.n_features, spec.widths, spec.activation, spec.task)
    ens = Ensemble(fs, chains, device=local)
    try:
        ens.set_data(X, y)
        th0 = (0.3 * np.random.default_rng(0).standard_normal((chains, fs.n_params))).astype(np.float32)
        ens.nuts_init(th0, max_num_doublings=10, initial_step_size=1.0)
        sched = build_schedule(n_warm)
        l0 = ens.launches
        t0 = time.perf_counter()
        done, ev_w = 0, 0
        while done < n_warm:
            n = min(100, n_warm - done)
            ev_w += int(ens.nuts_warmup(n, sched[done:done + n], step_base=done, seed=1, info=True)[..., 0].sum())
            done += n
        ens.nuts_finish_warmup()
        t_w = time.perf_counter() - t0
        t0 = time.perf_counter()
        samples, info = ens.nuts_sample(n_samp, seed=2, info=True)
        t_s = time.perf_counter() - t0
        ev_s = int(info[..., 0].sum())
        return {'workload': workload, 'chains': chains, 'kernel': 'mile_nuts_kernel (whole transitions per launch)',
                'warmup': {'transitions': n_warm, 'grad_evals': ev_w, 'grad_evals_per_s': ev_w / t_w, 'seconds': t_w},
                'sampling': {'transitions': n_samp, 'grad_evals': ev_s, 'grad_evals_per_s': ev_s / t_s, 'seconds': t_s,
                             'mean_tree_size': float(info[..., 0].mean()), 'mean_acceptance': float(info[..., 1].mean()),
                             'divergent_transitions': int(info[..., 3].sum())},
                'gpu_launches': ens.launches - l0, 'samples_finite': bool(np.isfinite(samples).all()),
                'note': 'host-buffer ABI calls, wall clock; MCLMC on this workload runs 2 gradient evaluations per chain-step'}
    finally:
        ens.close()


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a CUDA device: there is no CPU fallback'
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    from mile_b200 import capi
    from mile_b200 import build as _b
    _b.build()
    lib = capi.load()
    peaks = {'hbm': load_peaks(), 'fp32': None, 'fp32_src': None, 'mma': None}
    if rank == 0:
        # FP32 peak measured live (scalar FFMA and packed FFMA2, best of both) + the evaluator's tensor instruction
        import ctypes
        pk = [ctypes.c_double(), ctypes.c_double(), ctypes.c_double()]
        for v in (0, 1, 2):
            capi.check(lib.mile_measure_fp32_peak(local, v, ctypes.byref(pk[v])))
        peaks['fp32'] = max(pk[0].value, pk[1].value)
        peaks['fp32_src'] = f'live FMA micro-benchmark (FFMA {pk[0].value:.1f}, FFMA2 {pk[1].value:.1f} TFLOP/s)'
        peaks['mma'] = pk[2].value
    line = measure(args.workload, args, args.steps, args.warmup, rank, world, local,
                   (0 if (args.no_cpu or world > 1) else 12.0), peaks)
    extra = {}
    if not args.no_extra and args.workload == DEFAULT_WORKLOAD:
        names = EXTRA_WORKLOADS if world == 1 else ('covertype_full',)
        for W in names:
            try:
                sub = measure(W, args, min(args.steps, 3), min(args.warmup, 3), rank, world, local,
                              (0 if (args.no_cpu or world > 1) else 4.0), peaks)
            except Exception as e:   # noqa: BLE001
                sub = {'error': repr(e)} if rank == 0 else None
            if rank == 0 and sub is not None:
                keep = ('value', 'unit', 'ms_per_step', 'scaling', 'e2e', 'gpu_launches', 'samples_finite',
                        'cpu_baseline', 'sharded_parity_rel', 'error')
                s = {k: sub[k] for k in keep if k in sub}
                if 'roofline' in sub:
                    s['roofline'] = {k: sub['roofline'][k] for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'kernel', 'tensor_pipe')
                                     if k in sub['roofline']}
                if 'config' in sub:
                    s['config'] = {k: sub['config'][k] for k in ('chains_per_gpu', 'inner_steps_per_launch', 'n_train', 'n_params',
                                                                 'cluster_size', 'kernel_path', 'parallelism')}
                extra[W] = s
    nuts = None
    if rank == 0 and world == 1 and not args.no_extra and args.workload == DEFAULT_WORKLOAD:
        try:
            nuts = measure_nuts(local)
        except Exception as e:   # noqa: BLE001
            nuts = {'error': repr(e)}
    if rank == 0:
        if extra:
            line['extra'] = {'workloads': extra,
                             'note': 'the other named shapes of BASELINE.json, same timing method on a shorter run (<= 3 steps)'}
            if nuts is not None:
                line['extra']['nuts'] = nuts
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--inner', type=int, default=0)
    ap.add_argument('--cluster', type=int, default=0)
    ap.add_argument('--sync-mode', type=int, default=-1, help='CTAs of a chain exchange through: 0 = DSMEM (thread-block cluster), 1 = flagged words in L2')
    ap.add_argument('--tile-rows', type=int, default=0)
    ap.add_argument('--tensor', type=int, default=-1, help='wide path: 1 = tcgen05 3xTF32 GEMM core, 0 = FP32 SIMT core')
    ap.add_argument('--fast', type=int, default=-1, help='narrow-MLP evaluator: 2 = 3xTF32 register MMA (default), 1 = FFMA layer pipeline, 0 = generic tiles')
    ap.add_argument('--steploop', type=int, default=-1, help='1 = integrator-warp step loop of the tensor evaluator (default), 0 = generic loop')
    ap.add_argument('--kslices', type=int, default=0, help='wide path: split-K slices of the dW GEMMs (0 = auto)')
    ap.add_argument('--no-tune', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip extra.workloads (the other named shapes)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
