"""GPU probe (not a pytest file): BASELINE config 5, RVQ decode-only, 3 body parts, B x 128 frames.
    python tests/decode_bench.py [B] [engine]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.vq import RVQVAE

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
engine = sys.argv[2] if len(sys.argv) > 2 else "tc"
torch.set_grad_enabled(False)
_lib.set_engine(engine)
vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
g = torch.Generator().manual_seed(5)
lats = [(5.0 * torch.randn(B, 32, 512, generator=g)).cuda() for _ in range(3)]
work = [l.clone() for l in lats]


def step():
    for w, l in zip(work, lats):
        w.copy_(l)
    return [v.latent2origin(w)[0] for v, w in zip(vqs, work)]


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flops = B * 3.899e9
print(json.dumps({"config": "RVQ decode-only, 3 body parts", "B": B, "engine": engine, "ms": ms, "frames_per_s": B * 128 / (ms / 1e3),
                  "algorithmic_tflops": flops / (ms / 1e3) / 1e12}))
