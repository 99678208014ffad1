"""GPU probe (not a test): wall time per window batch of (a) the device-resident call, (b) the blocking host-buffer call,
(c) the two-slot host-buffer pipeline; n batches back to back, no L2 flush, CUDA events around the whole run."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
from syntalker_b200.pipeline import Window330
from syntalker_b200.vq import RVQVAE
B, n = 32, 10
torch.set_grad_enabled(False)
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
diff = create_gaussian_diffusion(use_ddim=True)
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
win = Window330(w, diff, *vqs, B=B, use_ddim=True)
d_in = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
y_dev = {"scale": torch.ones(1) * 2.0, "style_feature": d_in["style_feature"]}
y_host = {"scale": torch.ones(1) * 2.0, "style_feature": inp["style_feature"]}
pin = {k: inp[k].contiguous().pin_memory() for k in ("audio", "word", "seed", "noise")}


def timed(fn):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


def dev():
    for _ in range(n):
        win.run_device(d_in["audio"], d_in["word"], d_in["seed"], d_in["noise"], y=y_dev)


def blocking():
    for _ in range(n):
        win.run(pin["audio"], pin["word"], pin["seed"], pin["noise"], y=y_host)


def piped():
    for i in range(n):
        win.begin(i % 2, pin["audio"], pin["word"], pin["seed"], pin["noise"], y=y_host)
        if i >= 1:
            win.wait((i - 1) % 2)
    win.wait((n - 1) % 2)


for name, fn in (("device-resident", dev), ("host, blocking", blocking), ("host, two in flight", piped)):
    ev, wall = timed(fn)
    print(f"{name:22s}: {ev:7.3f} ms per batch by events, {wall:7.3f} ms by wall clock")
