"""GPU probe: does programmatic dependent launch shorten a chain of dependent tcgen05 GEMMs inside a CUDA graph?"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib
L = _lib.lib()
for (M, N, K) in [(128, 64, 64), (2048, 512, 512), (2048, 1024, 512)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for pdl in (0, 1, 0, 1):
        L.st_set_pdl(pdl)
        ms = C.c_double()
        _lib.check(L.st_bench_gemm(M, N, K, 1, 200, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), C.byref(ms)))
        print(f"M={M} N={N} K={K} pdl={pdl}: {ms.value * 1e3:.2f} us per launch in a 200-kernel graph chain", flush=True)
