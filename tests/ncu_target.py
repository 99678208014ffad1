"""Target for ncu (not a pytest file): one config-2 window batch with the profiler range around the
steady-state pass.  ncu --profile-from-start off ... python tests/ncu_target.py [engine] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
from syntalker_b200.pipeline import Window330
from syntalker_b200.vq import RVQVAE

engine = sys.argv[1] if len(sys.argv) > 1 else "tc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = sys.argv[3] if len(sys.argv) > 3 else "ddim50"
torch.set_grad_enabled(False)
_lib.set_engine(engine)
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
diff = create_gaussian_diffusion(timestep_respacing=steps)
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
d = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
y = {"scale": torch.ones(1) * 2.0, "style_feature": d["style_feature"]}
win = Window330(w, diff, *vqs, B=B, use_ddim=True)
for _ in range(3):
    win.run_device(d["audio"], d["word"], d["seed"], d["noise"], y=y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
win.run_device(d["audio"], d["word"], d["seed"], d["noise"], y=y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
