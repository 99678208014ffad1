"""GPU probe (not a test): which resource bounds the tcgen05 GEMM main loop?  Times the product kernel with the MMAs
skipped (TMA only), the TMA skipped (MMA only) and MMA issue-order variants; prints per-launch device time in a
200-launch graph chain and the K-block cadence of CTA (0,0)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib
L = _lib.lib()
L.st_debug_probe.argtypes = [C.c_int]
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
EXTRA = int(os.environ.get('ST_PROBE_EXTRA', '0'))
names = {0: "product", 1: "TMA only", 2: "MMA only", 6: "MMA only grouped", 10: "MMA only 1 acc", 4: "grouped", 8: "1 acc"}
for (M, N, K) in [(128, 512, 512), (2048, 512, 512), (2048, 1024, 512), (2048, 1536, 512), (2048, 512, 1024)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for probe in ((0, 1, 2) if os.environ.get('ST_PROBE_SHORT') else (0, 1, 2, 6, 10, 4, 8)):
        _lib.check(L.st_debug_probe(probe | EXTRA))
        ms = C.c_double(0)
        _lib.check(L.st_bench_gemm(M, N, K, 1, 200, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), C.byref(ms)))
        call = lambda e: _lib.check(L.st_selftest_gemm(M, N, K, e, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
        call(1); call(2); call(2)
        torch.cuda.synchronize()
        _lib.check(L.st_debug_timeline(dbg.data_ptr()))
        call(2)
        torch.cuda.synchronize()
        _lib.check(L.st_debug_timeline(None))
        d = dbg.cpu().tolist(); t0 = d[0]; nkb = K // 64
        fr = [x - t0 for x in d[24:24 + min(nkb, 16)]]
        cad = [b_ - a_ for a_, b_ in zip(fr[:-1], fr[1:])]
        print(f"M={M} N={N} K={K} probe={probe:2d} ({names[probe]:16s}): {ms.value*1e3:7.2f} us/launch in chain; first full {fr[0]}, cadence {cad}, acc ready {d[3]-t0}, end {d[5]-t0}")
        call(1)
    _lib.check(L.st_debug_probe(0))
