#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=${1:-r2h}
timeout 300 python tests/conv_probe.py 2>&1 | grep -v "base offset" | tail -12
cat > /tmp/taps_check.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from syntalker_b200 import _lib, synth
from syntalker_b200.denoiser import MDM
from oracle import mdm as omdm
torch.set_grad_enabled(False)
L = _lib.lib()
W = synth.mdm_state_dict("beatx", seed=0)
m = MDM(None).load_state_dict(W)
inp = synth.make_inputs(4, seed=1, variant="beatx")
ref = omdm.wav_encoder(W, inp["audio"])
y = {k: inp[k].cuda() for k in ("audio", "word", "seed")}
for flags, name in ((8192, "per-tap fetch"), (0, "staged taps")):
    _lib.check(L.st_debug_probe(flags))
    m.encode_cond(y, force=True)
    at = torch.empty(4, 128, 512, device="cuda")
    _lib.check(L.st_debug_cond_taps(m.handle, at.data_ptr(), None, 4, _lib.stream_ptr()))
    torch.cuda.synchronize()
    print(f"{name:20s} WavEncoder max-abs vs oracle {float((at[:, :, :256].cpu() - ref).abs().max()):.3e}")
    i32 = synth.make_inputs(32, seed=2, variant="beatx"); y32 = {k: i32[k].cuda() for k in ("audio", "word", "seed")}
    for _ in range(3): m.encode_cond(y32, force=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): m.encode_cond(y32, force=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:20s} cond encode, 32 clips: {e0.elapsed_time(e1) / 10:.3f} ms")
_lib.check(L.st_debug_probe(0))
PY
timeout 200 python /tmp/taps_check.py > $out/${tag}_taps.log 2>&1; cat $out/${tag}_taps.log | tail -6
timeout 200 python tests/trace_pipeline.py > $out/${tag}_trace_pipeline.log 2>&1; grep -E "^cond sequence|^--" $out/${tag}_trace_pipeline.log | cut -c1-500
timeout 600 python -m pytest tests -q -m gpu -x > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
