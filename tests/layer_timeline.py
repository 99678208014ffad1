"""GPU probe (not a test): clock64 timeline of CTA (0,0) of one trunk-kernel layer INSIDE the captured sampling loop
(steady state, last launch of the selected shape): python tests/layer_timeline.py"""
import os, subprocess, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) < 3:
    for n, k, name in ((512, 512, "proj"), (1024, 512, "fc1"), (512, 1024, "fc2"), (1536, 512, "qkv+attention / output"), (512, 1536, "W_x")):
        out = subprocess.run([sys.executable, __file__, str(n), str(k)], capture_output=True, text=True).stdout.strip().splitlines()
        print(f"{name:24s}", out[-1] if out else "?")
    sys.exit(0)
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
N, K = int(sys.argv[1]), int(sys.argv[2])
B = int(os.environ.get("ST_B", "32"))
torch.set_grad_enabled(False)
L = _lib.lib()
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
if os.environ.get("ST_PROBE"):          # 32768: the timeline of the LAST CTA of the grid instead of CTA (0,0)
    _lib.check(L.st_debug_probe(int(os.environ["ST_PROBE"])))
_lib.check(L.st_debug_timeline_select(N, K))
_lib.check(L.st_debug_timeline(dbg.data_ptr()))          # before the step graph is captured: the captured launches carry the pointer
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
diff = create_gaussian_diffusion(timestep_respacing="ddim10")
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
d = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
y = {"audio": d["audio"], "word": d["word"], "seed": d["seed"], "style_feature": d["style_feature"], "scale": torch.ones(1) * 2.0}
for _ in range(4):
    diff.ddim_sample_loop(w, (B, 1536, 1, 32), noise=d["noise"], clip_denoised=False, model_kwargs={"y": y})
torch.cuda.synchronize()
t = dbg.cpu().tolist(); t0 = t[0]
nkb = K // 64
print(f"N={N} K={K}: set-up done {t[32]-t0}, dependency resolved {t[1]-t0}, first TMA issued {t[8]-t0}, first operands {t[24]-t0}, last K block ready {t[24+min(nkb,16)-1]-t0}, "
      f"accumulator ready {t[3]-t0}, staged {t[6]-t0}, stores read {t[4]-t0}, CTA end {t[5]-t0}"
      f" | producer: at its branch {t[33]-t0}, proxy fence done {t[34]-t0}, weight tiles requested {t[35]-t0}; stores waited for {t[36]-t0}, CTA barrier passed {t[37]-t0}")
