"""GPU probe (not a test): the stride-1 conv modes of the tcgen05 engine against fp64, with integer-valued operands (the fp16 lo planes
are exactly zero, so only the addressing is tested) and with random ones (hi + lo), staged taps on / off."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib
torch.set_grad_enabled(False)
L_ = _lib.lib()


def conv_ref(A, W, taps, dil):
    B, L, C = A.shape
    N = W.shape[0]
    w = W[:, :taps * C].reshape(N, taps, C).permute(0, 2, 1).double()          # [N, C, taps]
    y = torch.nn.functional.conv1d(A.double().permute(0, 2, 1), w, padding=dil * (taps - 1) // 2, dilation=dil)
    return y.permute(0, 2, 1).reshape(B * L, N)


def run(B, L, C, N, taps, dil, integer, flags):
    g = torch.Generator().manual_seed(B + L + C + taps + dil)
    if integer:
        A = torch.randint(-8, 9, (B, L, C), generator=g).float()
        W = torch.randint(-4, 5, (N, taps * C), generator=g).float()
    else:
        A = torch.randn(B, L, C, generator=g)
        W = torch.randn(N, taps * C, generator=g) / (taps * C) ** 0.5
    ref = conv_ref(A, W, taps, dil)
    out = torch.empty(B * L, N, device="cuda")
    _lib.check(L_.st_debug_probe(flags))
    Ad, Wd = A.cuda().contiguous(), W.cuda().contiguous()
    _lib.check(L_.st_selftest_conv(B, L, C, N, taps, dil, 1, Ad.data_ptr(), Wd.data_ptr(), None, out.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    _lib.check(L_.st_debug_probe(0))
    d = (out.cpu().double() - ref).abs()
    bad = (d > 1e-3 * max(1.0, float(ref.abs().max()))).reshape(B, L, N).any(-1)
    rows = bad.nonzero()
    return float(d.max()), int(bad.sum()), rows[:6].tolist()


for (B, L, C, N, taps, dil) in [(2, 300, 64, 64, 15, 1), (2, 300, 128, 128, 15, 1), (1, 256, 512, 512, 3, 9), (1, 256, 512, 512, 3, 3), (3, 400, 64, 64, 3, 1)]:
    for integer in (True, False):
        for flags, name in ((8192, "per-tap fetch"), (0, "staged taps")):
            mx, nbad, rows = run(B, L, C, N, taps, dil, integer, flags)
            print(f"B{B} L{L} C{C} N{N} k{taps} d{dil} {'int ' if integer else 'rand'} {name:24s} max-abs {mx:.3e} bad rows {nbad} {rows}")
