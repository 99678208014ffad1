"""GPU probe (not a test): the sampling loop (DDIM-50, CFG) with the trunk layers as chains (one cluster launch per step) against one
launch per layer (the default; st_debug_probe bit 131072 switches the chains on), at a batch whose row tiles fit the GPU as clusters of 8 in one wave (B = 30: 15 row
tiles) and at the bench's B = 32 (16 row tiles: two waves when only 15 clusters of 8 can be resident)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
torch.set_grad_enabled(False)
L = _lib.lib()
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
diff = create_gaussian_diffusion(timestep_respacing="ddim50")
for B in [int(b) for b in os.environ.get("ST_BS", "30,32").split(",")]:
    inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
    d = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
    y = {"audio": d["audio"], "word": d["word"], "seed": d["seed"], "style_feature": d["style_feature"], "scale": torch.ones(1) * 2.0}
    outs = {}
    for flags in (int(os.environ.get('ST_PROBE', '0')),) * 2:      # one setting per process: captured graphs are cached per schedule
        _lib.check(L.st_debug_probe(flags))
        run = lambda: diff.ddim_sample_loop(w, (B, 1536, 1, 32), noise=d["noise"], clip_denoised=False, model_kwargs={"y": y})
        for _ in range(3): out = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = run()
        e1.record(); torch.cuda.synchronize()
        outs[flags] = out
        print(f"B={B} {'layer chains        ' if flags & 131072 else 'one launch per layer'}: {e0.elapsed_time(e1) / 5:.3f} ms per 50-step loop", flush=True)
    print(f"B={B} checksum {float(out.double().abs().sum()):.6f}")
