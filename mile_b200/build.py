"""Build the CUDA library in-tree: nvcc -> mile_b200/_lib/libmile_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / 'csrc'
LIBDIR = PKG / '_lib'
LIB = LIBDIR / 'libmile_b200.so'
SOURCES = ['mile_api.cu', 'mile_microbench.cu', 'mile_npz.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / f for f in SOURCES] + sorted(CSRC.glob('*.cuh')) + sorted((PKG.parent / 'include').glob('*.h'))
    return any(f.resolve().stat().st_mtime > t for f in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source of the package for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, '-o', str(LIB), *[str(CSRC / s) for s in SOURCES], '-ldl', '-lz']
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
