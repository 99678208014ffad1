"""Target for compute-sanitizer (not a pytest file): one tiny pass over every kernel of the path.
    compute-sanitizer --tool memcheck python tests/sanitize_target.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
from syntalker_b200.pipeline import Window330, LongClip330
from syntalker_b200.vq import RVQVAE
torch.set_grad_enabled(False)
_lib.check(_lib.lib().st_set_graphs(int(os.environ.get("ST_GRAPHS", "0"))))
if os.environ.get("ST_PROBE"):
    _lib.check(_lib.lib().st_debug_probe(int(os.environ["ST_PROBE"])))
B = int(os.environ.get("ST_B", "3"))
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
diff = create_gaussian_diffusion(timestep_respacing="ddim10")
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
d = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
y = {"scale": torch.ones(1) * 2.0, "style_feature": d["style_feature"]}
win = Window330(w, diff, *vqs, B=B, use_ddim=True)
pose, trans, sample = win.run_device(d["audio"], d["word"], d["seed"], d["noise"], y=y)
torch.cuda.synchronize()
lat = vqs[0].map2latent(torch.randn(2, 64, 78).cuda())
torch.cuda.synchronize()
print("ok", float(pose.abs().max()), float(lat.abs().max()))
