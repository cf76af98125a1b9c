#!/usr/bin/env python
"""bench.py -- chain-steps/s of the MCLMC ensemble sampling hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement
                                                             # of the reference (oracle/), all host threads

A bench "step" is ONE launch of the persistent sampler kernel: `inner` MCLMC steps for every
chain of the ensemble wave (two full-batch gradient evaluations per chain-step), thinned samples
captured in HBM.  Workload (N=1): BASELINE.json configs[1] -- UCI bikesharing shape, FCN 2x16,
10 chains per split on one B200, synthetic data.  With --gpus N every rank
runs its own block of 10 chains on the same split (weak scaling: the ensemble shards by chain with no
data-path collective; the per-chain test-set logsumexp states are merged once at the end over NCCL).
Prints ONE JSON line on rank 0.
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
    # name: (oracle config key, chains per GPU, inner steps per launch)
    'bikesharing_2x16': ('bikesharing_2x16', 10, 500),
    'airfoil_3x16': ('airfoil_3x16', 12, 2000),
    'protein_2x16': ('protein_2x16', 10, 200),
    'covertype_ref': ('covertype_ref', 12, 500),
    'airfoil_3x16_1024': ('airfoil_3x16', 1024, 200),
    'wide_4x256': ('wide_4x256', 8, 10),          # HBM-resident chain-batched GEMM path (no fused LPPD fold)
    # BASELINE.json configs[2]: every rank holds all 12 chains and 1/N of the 232 404 rows; one NCCL all-reduce of the
    # packed [C, d+1] (gradient, log-lik) per evaluation.  Strong scaling: total work is fixed as N grows.
    'covertype_full': ('covertype_full', 12, 20),
}
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
    return 1626.1, 'fallback: measured cuBLAS bf16 figure quoted in SURVEY.md 8(d)'


def flops_per_chain_step(n_rows, dims):
    """12*N*W FLOPs per chain-step (two fwd+bwd passes, W = sum in*out)."""
    W = sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    return 12.0 * n_rows * W


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


def load_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        j = json.loads(p.read_text())
        return j.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement (C + OpenMP when built, numpy otherwise)
# ----------------------------------------------------------------------------------------------
def cpu_chain_steps_per_s(workload: str, budget_s: float = 12.0, max_steps: int = 200, waves: int = 1):
    """Times the CPU restatement of the reference on a bounded sample of the same workload:
    all chains of one wave, one chain per host thread (mirrors one-virtual-device-per-chain pmap,
    train.py:16), as many MCLMC steps as fit the budget."""
    from oracle import mile_oracle as o
    key, C, _ = WORKLOADS[workload]
    C = C * max(1, waves)          # reference arm at N GPUs: the same N x C chains our arm runs (weak scaling)
    spec = o.make_spec(key)
    X, y, _, _ = o.synthetic_data(key)
    th0 = o.synthetic_theta0(spec, C)
    try:
        from oracle import c_oracle
        have_c = c_oracle.available()
    except Exception:
        have_c = False
    if have_c:
        cores = min(C, os.cpu_count() or 1)
        n = 2
        t0 = time.perf_counter()
        c_oracle.run_sampling_timed(spec, X, y, th0, n, 0.02, float(np.sqrt(spec.n_params)), threads=cores)
        dt = time.perf_counter() - t0
        n = int(max(2, min(max_steps, budget_s / max(dt / 2, 1e-6))))
        t0 = time.perf_counter()
        c_oracle.run_sampling_timed(spec, X, y, th0, n, 0.02, float(np.sqrt(spec.n_params)), threads=cores)
        dt = time.perf_counter() - t0
        return {'value': C * n / dt, 'unit': 'chain-steps/s', 'cores': cores, 'kind': 'port',
                'sample': f'{C} chains x {n} MCLMC steps of {workload} (C restatement, OpenMP one chain per thread)'}
    # numpy port, single process
    f = lambda t: o.logpost_value_and_grad(spec, t, X, y)
    rng = np.random.default_rng(0)
    d = spec.n_params
    st = o.mclmc_init(f, th0[0], rng.standard_normal(d).astype(np.float32))
    n, t0 = 0, time.perf_counter()
    while n < max_steps and time.perf_counter() - t0 < budget_s:
        st, _ = o.mclmc_step(f, st, 0.02, float(np.sqrt(d)), rng.standard_normal(d).astype(np.float32))
        n += 1
    dt = time.perf_counter() - t0
    return {'value': n / dt, 'unit': 'chain-steps/s', 'cores': 1, 'kind': 'port',
            'sample': f'1 chain x {n} MCLMC steps of {workload} (numpy restatement, BLAS threads as configured)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    W = args.workload
    key, C, inner = WORKLOADS[W]
    from oracle import mile_oracle as o
    spec = o.make_spec(key)
    Ntr = o.CONFIGS[key][0]
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        waves = 1 if key == 'covertype_full' else max(1, args.gpus)   # row-sharded workload: the chains do not multiply
        r = cpu_chain_steps_per_s(W, budget_s=max(2.0, 20.0 / max(1, args.steps + args.warmup)), waves=waves)
        if i >= args.warmup:
            vals.append(r['value'])
        base = r
    v = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': 'chain-steps/sec', 'value': v, 'unit': 'chain-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * C * waves * inner / v, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': W, 'chains_per_gpu': C, 'chains_total': C * waves, 'inner_steps_per_launch': inner, 'n_train': Ntr,
                   'n_params': spec.n_params, 'note': 'CPU restatement of the reference (JAX/BlackJAX not installable here); '
                                                      'one chain per host thread, all chains of the N-GPU job'},
        'grad_evals_per_s': 2 * v,
        'cpu_baseline': dict(base, value=v),
        'e2e': {'value': v, 'unit': 'chain-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
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
    from mile_b200 import Ensemble, FCNSpec, capi
    from mile_b200 import build as _b
    _b.build()
    from mile_b200 import synthetic as syn   # seeded synthetic inputs (the oracle is only used by the cpu_baseline leg)

    W = args.workload
    key, C, inner = WORKLOADS[W]
    if args.inner:
        inner = args.inner
    spec = ospec = syn.workload_spec(key)
    X, y, Xt, yt = syn.synthetic_data(key, seed=1234)   # same split on every rank; the CHAINS are what shards
    d = spec.n_params
    sharded = key == 'covertype_full'
    if sharded:
        from mile_b200 import ShardedEnsemble
        ens = ShardedEnsemble(spec, C, device=local, rank=rank, world=world)
        rows = ShardedEnsemble.shard_rows(X.shape[0], rank, world)
        Xfull, yfull = X, y
        X, y = np.ascontiguousarray(X[rows]), np.ascontiguousarray(y[rows])
    else:
        ens = Ensemble(spec, C, device=local)
    if args.cluster:
        ens.set_option('cluster_size', args.cluster)
    if args.tile_rows:
        ens.set_option('tile_rows', args.tile_rows)
    if args.tensor >= 0:
        ens.set_option('tensor', args.tensor)
    if args.fast >= 0:
        ens.set_option('fast', args.fast)
    ens.set_data(X, y)
    fused_lppd = key != 'wide_4x256' and not sharded
    if fused_lppd:
        ens.set_test(Xt, yt)      # fused posterior-predictive LPPD fold at every kept sample
    crank = 0 if sharded else rank        # sharded: every rank carries the SAME chains (same seeds, same noise)
    th0 = syn.synthetic_theta0(d, C, seed0=1000 + 100 * crank, scale=0.3 if key != 'wide_4x256' else 0.05)
    ens.init(th0, seed=17 + crank)
    # short tuning run -> frozen (eps, L); fallback eps=0.02, L=sqrt(d)
    eps = np.full(C, 0.02, np.float32)
    L = np.full(C, np.sqrt(d), np.float32)
    if not args.no_tune:
        ens.tune_reset(0.01)
        nt1, nt2 = (40, 10) if key in ('wide_4x256', 'covertype_full') else (800, 100)
        tc = ens.tune_cfg(nt1, nt2, 0.5, 0.1, 1.5, 100)
        ens.tune(nt1 + nt2, 0, tc, seed=99 + crank)
        ens.tune_finish_phase2()
        e, l, _ = ens.get_tuning()
        if np.all(np.isfinite(e)) and np.all(e > 0) and np.all(np.isfinite(l)) and np.all(l > 0):
            eps, L = e, l
    dev = torch.device(f'cuda:{local}')
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

    for i in range(args.warmup):
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
    for i in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step(args.warmup + i)
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
    chain_steps = nrep * C * inner * args.steps
    value = chain_steps / t_dev
    # the one exchange step of the path: merge the per-chain online logsumexp states (NCCL all-gather)
    from mile_b200.distributed import merge_lppd_states
    t0 = time.perf_counter()
    lppd_val, lppd_total = None, 0
    if fused_lppd:
        m_, s_, cnt_ = ens.lppd_state()
        lppd_val, lppd_total = merge_lppd_states(m_, s_, cnt_, device=dev)
    lppd_ms = 1e3 * (time.perf_counter() - t0)

    # ---- e2e: the same metric through the host-buffer C-ABI call (H2D + D2H inside) --------
    st = ens.get_state()
    Xp = torch.from_numpy(X).pin_memory().numpy()
    h2d = X.nbytes + y.nbytes + 3 * C * d * 4 + C * 4 + 2 * C * 4
    d2h = n_slots * C * d * 4 + 3 * C * d * 4 + C * 4
    e2e_steps = max(1, min(args.steps, 5))
    # one untimed pass first: the host-buffer entry points allocate their pinned staging buffers on first use
    ens.set_data(Xp, y); ens.set_state(*st)
    ens.sample(inner, eps, L, step_base=0, n_thinning=N_THINNING, seed=4320, lppd=False)
    ens.set_state(*st)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ens.set_data(Xp, y)
        ens.set_state(*st)
        smp, _ = ens.sample(inner, eps, L, step_base=0, n_thinning=N_THINNING, seed=4321 + i, lppd=fused_lppd)
        st = ens.get_state()
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = nrep * C * inner * e2e_steps / t_e2e
    finite = bool(np.all(np.isfinite(smp)))

    if rank == 0:
        # FP32 peak measured live (scalar FFMA and packed FFMA2), best of both
        import ctypes
        pk = [ctypes.c_double(), ctypes.c_double()]
        for v in (0, 1):
            capi.check(ens.lib.mile_measure_fp32_peak(local, v, ctypes.byref(pk[v])))
        fp32_peak = max(pk[0].value, pk[1].value)
        hbm_peak, peak_src = load_peaks()
        Ntr = Xfull.shape[0] if sharded else X.shape[0]
        fl = flops_per_chain_step(Ntr, ospec.dims)
        per_gpu_steps_per_s = value / world     # (sharded: each GPU does 1/world of every chain-step's rows)
        achieved_tf = per_gpu_steps_per_s * fl / 1e12
        hbm_bytes_per_launch = C * (4 * d * 4 + n_slots * d * 4) + X.nbytes * 0  # state r/w + samples
        line = {
            'metric': 'chain-steps/sec', 'value': value, 'unit': 'chain-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if sharded else 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': W, 'chains_per_gpu': C, 'inner_steps_per_launch': inner, 'n_train': int(Ntr),
                       'n_features': ospec.n_features, 'hidden_structure': list(ospec.widths), 'n_params': d,
                       'n_thinning': N_THINNING, 'noise': 'in-kernel Philox4x32-10',
                       'cluster_size': ens.get_option('cluster_size'), 'tile_rows': ens.get_option('tile_rows'),
                       'x_resident_in_smem': bool(ens.get_option('resident')),
                       'kernel_path': 'data-sharded: gradient kernel -> ncclAllReduce([C, d+1]) -> integrator kernel, twice per step' if sharded else (('wide: HBM-resident chain-batched GEMMs, ' + ('tcgen05 3xTF32' if ens.get_option('tensor') else 'FP32 SIMT')) if ens.get_option('wide') else ('FastGE pipeline' if ens.get_option('fast') else 'GenericGE')),
                       'l2': 'flushed between timed iterations (256 MiB write); working set is SMEM-resident',
                       'step_size_mean': float(eps.mean()), 'L_mean': float(L.mean()), 'parallelism': (f'rows x{world} + NCCL all-reduce of [C, d+1] per gradient evaluation' if sharded else f'chains x{world}')},
            'grad_evals_per_s': 2 * value,
            'samples_finite': finite,
            'lppd': {'value': lppd_val, 'samples': int(lppd_total), 'merge_ms': lppd_ms,
                     'note': 'test-set logsumexp folded in-kernel at every kept sample; merged across ranks after the timed region'},
            'wall_s_timed_region': t_wall,
            'gpu_launches': int(launches),
            'clocks': clk,
            'e2e': {'value': e2e_value, 'unit': 'chain-steps/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps,
                    'path': 'Ensemble.set_data + set_state + sample (mile_*_host C-ABI calls, host numpy buffers) + get_state'},
        }
        roof = {'bound': 'fp32', 'achieved': achieved_tf, 'peak': fp32_peak, 'unit': 'TFLOP/s',
                'frac': achieved_tf / fp32_peak if fp32_peak else None, 'traffic': None,
                'kernel': 'mile_mclmc_kernel', 'flops_per_chain_step': fl,
                'peak_source': f'live FMA micro-benchmark (FFMA {pk[0].value:.1f}, FFMA2 {pk[1].value:.1f} TFLOP/s)',
                'note': 'SURVEY.md 8(d): these shapes are bound by the FP32 FMA pipe (working set in SMEM/L2), '
                        'neither by HBM nor by the tensor pipe; the HBM view is given below for completeness',
                'avg_launch_ms': float(np.mean(kernel_ms)),
                'hbm': {'achieved_gbs': hbm_bytes_per_launch / (np.mean(kernel_ms) * 1e-3) / 1e9,
                        'peak_gbs': hbm_peak, 'peak_source': peak_src,
                        'note': 'tiny by design: theta/u/g and the X slice stay in shared memory for the whole launch'}}
        if ens.get_option('wide') and ens.get_option('tensor'):
            tpk, tsrc = load_tensor_peak()
            roof.update({'bound': 'tensor', 'peak': tpk, 'frac': achieved_tf / tpk, 'peak_source': tsrc,
                         'kernel': 'wide_gemm_tc_kernel (tcgen05 kind::tf32, 3 MMAs per product for fp32-level accuracy)',
                         'note': 'achieved = ALGORITHMIC fp32 FLOPs (12 N W per chain-step); the 3xTF32 split issues 3x that '
                                 'on the tensor pipe, and TF32 dense peak is half the bf16 figure used as denominator'})
        if W == 'bikesharing_2x16' and inner == 500 and not sharded:
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full (profiles/r1q_ncu_fastge_final.txt)
            roof['traffic'] = 1339904 + 109824
            roof['traffic_note'] = 'bytes per launch of mile_mclmc_kernel from the committed ncu capture of this command'
        line['roofline'] = roof
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_chain_steps_per_s(W)
        print(json.dumps(line))
    ens.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='bikesharing_2x16', choices=sorted(WORKLOADS))
    ap.add_argument('--inner', type=int, default=0)
    ap.add_argument('--cluster', type=int, default=0)
    ap.add_argument('--tile-rows', type=int, default=0)
    ap.add_argument('--tensor', type=int, default=-1, help='wide path: 1 = tcgen05 3xTF32 GEMM core, 0 = FP32 SIMT core')
    ap.add_argument('--fast', type=int, default=-1, help='narrow-MLP evaluator: 2 = 3xTF32 register MMA (default), 1 = FFMA layer pipeline, 0 = generic tiles')
    ap.add_argument('--no-tune', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
