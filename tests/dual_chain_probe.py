"""GPU probe (not a test): the two CFG evaluations of a step as two chains on two streams (st_debug_probe bit 4096) against the stacked
single chain -- same result? -- and the window-batch time both ways (B = 32, DDIM-50, CFG)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion

torch.set_grad_enabled(False)
L = _lib.lib()
B = 32
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
y = {k: inp[k].cuda() for k in ("audio", "word", "seed", "style_feature")}; y["scale"] = torch.ones(1) * 2.0
x0 = inp["noise"].cuda()
res = {}
for flags in (0, 4096):
    _lib.check(L.st_debug_probe(flags))
    d = create_gaussian_diffusion(use_ddim=True)
    run = lambda: d.ddim_sample_loop(w, (B, 1536, 1, 32), noise=x0, clip_denoised=False, model_kwargs={"y": y}, consume_rng=False)
    for _ in range(3):
        out = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = run()
    e1.record(); torch.cuda.synchronize()
    res[flags] = (out.clone(), e0.elapsed_time(e1) / 5)
    print(f"probe {flags}: {res[flags][1]:.3f} ms per 50-step loop (cond encode included)")
_lib.check(L.st_debug_probe(0))
print("max-abs dual vs single:", float((res[0][0] - res[4096][0]).abs().max()), "bitwise equal:", bool(torch.equal(res[0][0], res[4096][0])))
