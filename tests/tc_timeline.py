"""GPU probe: clock64 timeline of CTA (0,0) of one tcgen05 GEMM launch (debug aid, not a test)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib
L = _lib.lib()
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
for (M, N, K) in [(128, 512, 512), (512, 512, 512), (1024, 512, 512), (2048, 512, 512), (2048, 1536, 512), (2048, 1024, 512), (16384, 512, 512)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    call = lambda e: _lib.check(L.st_selftest_gemm(M, N, K, e, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
    call(1); call(2); call(2)
    torch.cuda.synchronize()
    _lib.check(L.st_debug_timeline(dbg.data_ptr()))
    call(2)
    torch.cuda.synchronize()
    _lib.check(L.st_debug_timeline(None))
    d = dbg.cpu().tolist(); t0 = d[0]
    nkb = K // 64
    print(f"M={M} N={N} K={K}: setup {d[1]-t0}, tma issue {[x-t0 for x in d[8:8+nkb]]}, full ready {[x-t0 for x in d[24:24+nkb]]}, "
          f"mma issued {d[2]-t0}, acc ready {d[3]-t0}, phase1 done {d[6]-t0}, pair sync {d[7]-t0}, epilogue done {d[4]-t0}, end {d[5]-t0}")
    call(1)
