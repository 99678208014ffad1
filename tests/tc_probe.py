"""GPU probe (not a pytest file): accuracy and speed of the two GEMM engines on the path's shapes.
    python tests/tc_probe.py [out.json]
Accuracy is against an fp64 product of the same fp32 inputs; speed is CUDA-event time over 20 launches."""
import ctypes as C
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib

SHAPES = [(128, 64, 64), (128, 128, 64), (256, 512, 512), (2048, 512, 512), (2048, 1536, 512), (2048, 1024, 512), (2048, 512, 1024),
          (2048, 1536, 1536 // 3 * 3), (1024, 512, 1536), (4096, 512, 512), (8192, 1536, 512), (32768, 512, 1536), (200, 78, 512), (131072, 512, 512)]


def run(engine, M, N, K, A, W, b, out, reps=20):
    L = _lib.lib()
    call = lambda e: _lib.check(L.st_selftest_gemm(M, N, K, e, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
    call(engine)                                   # engine 1 builds fresh weight planes
    f = (lambda: call(2)) if engine == 1 else (lambda: call(0))   # 2 = tcgen05 with the planes kept (steady state)
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    if engine == 1:
        call(1)                                    # drop the cached planes again: the address will be reused
    ms = C.c_double()
    _lib.check(L.st_bench_gemm(M, N, K, engine, 50, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), C.byref(ms)))
    return ms.value                                # device time per launch inside a CUDA graph


def main():
    res = []
    for (M, N, K) in SHAPES:
        g = torch.Generator().manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g).cuda()
        W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b = torch.randn(N, generator=g).cuda()
        ref = (A.double() @ W.double().t() + b.double())
        row = {"M": M, "N": N, "K": K}
        for name, eng in (("simt", 0), ("tc", 1)):
            out = torch.zeros(M, N, device="cuda")
            try:
                ms = run(eng, M, N, K, A, W, b, out)
                err = float((out.double() - ref).abs().max())
                row[name] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "max_abs_err": err}
            except Exception as ex:      # noqa: BLE001
                row[name] = {"error": str(ex)}
        print(json.dumps(row), flush=True)
        res.append(row)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
