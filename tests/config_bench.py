"""GPU probe (not a pytest file): device time of the other BASELINE configs (3: 1000-step DDPM at 32 clips per GPU,
4: h3d body-part CFG at B=64, 5: RVQ decode-only at B=1024), CUDA events, inputs resident in HBM, 3 warm-up runs.
    python tests/config_bench.py [out.json]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import TwoClassifierFreeSampleModel_Bodypart
from syntalker_b200.denoiser import MDM
from syntalker_b200.denoiser_h3d import MDM as MDM_H3D
from syntalker_b200.diffusion import create_gaussian_diffusion
from syntalker_b200.pipeline import Window330, Window623
from syntalker_b200.vq import RVQVAE

torch.set_grad_enabled(False)
res = []


def timed(fn, n=3, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- config 3: diffusion_rvqvae_128, 1000-step p_sample_loop, 32 clips (one GPU's shard of B=256), decode + 330-d ----
B = 32
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx", seed=0))
vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
inp = synth.make_inputs(B, seed=1, variant="beatx")
d_in = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise")}
diff = create_gaussian_diffusion()                      # 1000 steps
win = Window330(model, diff, *vqs, B=B, use_ddim=False)
tape = torch.randn(1000, B, 1536, 1, 32, device="cuda")  # the reference draws randn_like(x) every step (gaussian_diffusion.py:541)
ms3 = timed(lambda: win.run_device(d_in["audio"], d_in["word"], d_in["seed"], d_in["noise"], noise_tape=tape), n=2, warm=2)
res.append({"config": 3, "what": "1000-step DDPM p_sample_loop + decode + 330-d, 32 clips on one GPU (shard of B=256)", "ms": ms3,
            "frames_per_s": B * 128 / (ms3 / 1e3), "us_per_diffusion_step": ms3 * 1e3 / 1000})
del tape, win
# ---- config 4: diffusion_h3d, upper + lower prompts + audio, B = 64, DDIM-50, body-part CFG (9 evaluations -> 4) ----
B = 64
m_h = MDM_H3D(None).load_state_dict(synth.mdm_state_dict("h3d", seed=0))
vqs_h = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_H3D]
inp = synth.make_inputs(B, seed=41, variant="h3d")
sf = {"upper_mask": inp["style_upper"].cuda(), "hands_mask": None, "lower_mask": inp["style_lower"].cuda()}
d50 = create_gaussian_diffusion(use_ddim=True)
w623 = Window623(TwoClassifierFreeSampleModel_Bodypart(m_h), d50, *vqs_h)
d_in = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise")}
ms4 = timed(lambda: w623.run(d_in["audio"], d_in["word"], d_in["seed"], d_in["noise"], sf), n=3)
res.append({"config": 4, "what": "denoiser_h3d, body-part CFG (4 de-duplicated evaluations per step), DDIM-50, B=64, decode + 623-d", "ms": ms4,
            "frames_per_s": B * 128 / (ms4 / 1e3)})
# ---- config 5: RVQ decode-only, 3 body parts, B = 1024 x 128 frames ----
B = 1024
g = torch.Generator().manual_seed(5)
lats = [(5.0 * torch.randn(B, 32, 512, generator=g)).cuda() for _ in range(3)]
work = [l.clone() for l in lats]


def dec():
    for w, l in zip(work, lats):
        w.copy_(l)
    return [v.latent2origin(w)[0] for v, w in zip(vqs, work)]


ms5 = timed(dec, n=5)
res.append({"config": 5, "what": "RVQ decode-only, 3 body parts, B=1024 x 128 frames", "ms": ms5, "frames_per_s": B * 128 / (ms5 / 1e3),
            "algorithmic_tflops": B * 3.899e9 / (ms5 / 1e3) / 1e12})
for r in res:
    print(json.dumps(r))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
