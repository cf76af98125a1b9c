import sys, ctypes as C; sys.path.insert(0, '.')
import numpy as np
from pathlib import Path
from mile_b200 import capi
# mile_debug_wide_gemm only exists in the -DMILE_PROFILE build (tools/phase_profile.py builds tools/_prof/)
capi.lib_path = lambda: Path('tools/_prof/libmile_b200.so').resolve()
lib = capi.load()
lib.mile_debug_wide_gemm.argtypes = [C.c_int32]*7 + [C.c_void_p]*3
lib.mile_last_error.restype = C.c_char_p
np.set_printoptions(linewidth=200, precision=3, suppress=True)
M, N, K = 128, 256, 32
def run(A, B):   # A (M,K), B (K,N) logical
    Ah = np.ascontiguousarray(A.T); Bh = np.ascontiguousarray(B)   # a_mn: [K][M]; b_mn: [K][N]
    Cc = np.empty((M, N), np.float32)
    rc = lib.mile_debug_wide_gemm(0, 2, M, N, K, 1, 1, Ah.ctypes.data, Bh.ctypes.data, Cc.ctypes.data)
    if rc: print('rc', rc, lib.mile_last_error().decode())
    return Cc
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
Cc = run(A, B); ref = A @ B
print('random: |C|', np.abs(Cc).mean(), '|ref|', np.abs(ref).mean(), 'rel', np.linalg.norm(Cc - ref) / np.linalg.norm(ref))
# A = delta on k=0 with value m+1 ; B(0,n) = n+1 -> C[m][n] = (m+1)(n+1)
A = np.zeros((M, K), np.float32); A[:, 0] = np.arange(M) + 1
B = np.zeros((K, N), np.float32); B[0, :] = np.arange(N) + 1
Cc = run(A, B)
print(Cc[:4, :8]); print(Cc[32:36, 32:40]); print('nonzero count', np.count_nonzero(Cc))
# k = 5
A = np.zeros((M, K), np.float32); A[:, 5] = np.arange(M) + 1
B = np.zeros((K, N), np.float32); B[5, :] = np.arange(N) + 1
Cc = run(A, B)
print(Cc[:4, :8]); print('nonzero count', np.count_nonzero(Cc))
A = np.zeros((M, K), np.float32); A[:, 5] = np.arange(M) + 1
B = np.zeros((K, N), np.float32); B[6, :] = np.arange(N) + 1
Cc = run(A, B)
print('mismatched k: nonzero count', np.count_nonzero(Cc), Cc[:2, :6])
