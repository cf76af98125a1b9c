"""Developer tool: per-launch time of the wide-path GEMM cores on one [M x K] x [K x N] problem, plus (with the
-DMILE_PROFILE build in tools/_prof) the in-kernel wait/work cycle split of CTA 0 of the TMA-fed core."""
import sys, os, ctypes as C; sys.path.insert(0, '.')
import numpy as np
from pathlib import Path
os.environ['MILE_DEBUG_TIMING'] = '1'
from mile_b200 import capi
prof = Path('tools/_prof/libmile_b200.so')
if len(sys.argv) > 1 and sys.argv[1] == 'prof':
    capi.lib_path = lambda: prof.resolve()
elif len(sys.argv) > 1:
    capi.lib_path = lambda: Path(sys.argv[1]).resolve()
lib = capi.load()
lib.mile_debug_wide_gemm.argtypes = [C.c_int32]*7 + [C.c_void_p]*3
rng = np.random.default_rng(0)
NAMES = {16: 'producer wait empty', 17: 'producer issue', 18: 'mma wait converted', 19: 'mma issue+commit', 20: 'cvt wait full',
         21: 'cvt work', 22: 'epi wait tmem_full', 23: 'epi work'}
for (M, N, K) in [(12165, 256, 256), (12165 * 8, 256, 256)]:
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
    for core in (1, 2):
        Cc = np.empty((M, N), np.float32)
        if hasattr(lib, 'mile_debug_read_profile'):
            p = (C.c_ulonglong * 32)(); lib.mile_debug_read_profile(p, 1)
        rc = lib.mile_debug_wide_gemm(0, core, M, N, K, 0, 0, A.ctypes.data, B.ctypes.data, Cc.ctypes.data)
        if hasattr(lib, 'mile_debug_read_profile') and core == 2:
            lib.mile_debug_read_profile(p, 1)
            print({NAMES[i]: int(p[i] / 11) for i in NAMES})
