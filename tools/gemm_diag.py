import sys, ctypes as C; sys.path.insert(0, '.')
import numpy as np
from pathlib import Path
from mile_b200 import capi
# mile_debug_wide_gemm only exists in the -DMILE_PROFILE build (tools/phase_profile.py builds tools/_prof/)
capi.lib_path = lambda: Path('tools/_prof/libmile_b200.so').resolve()
lib = capi.load()
lib.mile_debug_wide_gemm.argtypes = [C.c_int32]*7 + [C.c_void_p]*3
rng = np.random.default_rng(0)
for (M, N, K) in [(128, 256, 64), (700, 256, 256), (256, 256, 704)]:
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
            Ah = np.ascontiguousarray(A.T if a_mn else A); Bh = np.ascontiguousarray(B if b_mn else B.T)
            ref = A.astype(np.float64) @ B.astype(np.float64)
            out = []
            for core in (0, 1, 2):
                Cc = np.empty((M, N), np.float32)
                rc = lib.mile_debug_wide_gemm(0, core, M, N, K, a_mn, b_mn, Ah.ctypes.data, Bh.ctypes.data, Cc.ctypes.data)
                out.append('%d:%.2e' % (rc, np.linalg.norm(Cc - ref) / np.linalg.norm(ref)))
            print((M, N, K), 'a_mn', a_mn, 'b_mn', b_mn, ' '.join(out), flush=True)
