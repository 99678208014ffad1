"""GPU probe (not a test): clock64 timeline of CTA (0,0) of the fused qkv + attention kernel inside one denoiser evaluation."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.denoiser import MDM
B = 64
torch.set_grad_enabled(False)
L = _lib.lib()
if os.environ.get("ST_PROBE"):
    _lib.check(L.st_debug_probe(int(os.environ["ST_PROBE"])))
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx", seed=0))
inp = synth.make_inputs(B, seed=1, variant="beatx")
y = {k: inp[k].cuda() for k in ("audio", "word", "seed")}
x = inp["noise"].cuda(); t = torch.full((B,), 500, dtype=torch.int64, device="cuda")
for _ in range(3):
    model(x, t, y)
torch.cuda.synchronize()
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
_lib.check(L.st_debug_timeline(dbg.data_ptr()))
model(x, t, y)
torch.cuda.synchronize()
_lib.check(L.st_debug_timeline(None))
d = dbg.cpu().tolist()
t0 = d[40]
names = ["start", "acc ready", "q/k/v in smem", "scores done", "cluster barrier passed", "softmax done", "PV + planes staged", "TMA store done"]
print("fused qkv+attention CTA(0,0) timeline (cycles from CTA start):")
for n, v in zip(names, d[40:48]):
    print(f"  {n:24s} {v - t0}")
