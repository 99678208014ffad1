"""Round-2 golden fixtures from the REAL reference modules (build container only; same shims as make_golden.py).

    python tests/golden/make_golden_r2.py [ddpm|cfg|h3d|wav ...]     # default: all; writes tests/golden/*.npz

 * ddpm1000      B=1, `create_gaussian_diffusion()` (1000 steps) through the reference's own `p_sample_loop`
                 (gaussian_diffusion.py:607-739, p_sample :505-557). The per-step `th.randn_like` draws are taken from a
                 dedicated seeded CPU generator (torch's CPU SDPA kernel advances the GLOBAL generator inside every model
                 call, so draws from the global stream are not replayable); a test rebuilds the identical tape with
                 `ddpm_tape(seed, S, shape)` below. Intermediate samples via the loop's own `dump_steps`.
 * cfg_two       `TwoClassifierFreeSampleModel` over denoiser_h3d.MDM (cfg_sampler.py:31-54)
 * cfg_h3d_text  `ClassifierFreeSampleModel` over denoiser_h3d.MDM, and the same with `eval=True` (cfg_sampler.py:10-28;
                 the eval variant is what h3d_diffusion_new_trainer.py:922 builds)
 * cfg_bodypart1 `ClassifierFreeSampleModel_Bodypart` (cfg_sampler.py:125-167), B=1 (the wrapper hard-codes [1,256] zeros)
 * h3d_decode    latent2origin of the three HumanML3D decoders (D = 156 / 360 / 107) and the 623-d scatter with the joint
                 masks of h3d_diffusion_new_trainer.py:194-221, assembled as :604-607
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up the shims, sys.path and the reference imports)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from syntalker_b200 import synth  # noqa: E402
from diffusion.cfg_sampler import (ClassifierFreeSampleModel, ClassifierFreeSampleModel_Bodypart,  # noqa: E402
                                   TwoClassifierFreeSampleModel)
from diffusion.model_util import create_gaussian_diffusion  # noqa: E402
from diffusion import gaussian_diffusion as gd  # noqa: E402

DDPM_SEED = 777
DDPM_DUMPS = [0, 99, 499, 899, 999]


def ddpm_tape(seed, S, shape):
    """eps_k in draw order (t = S-1 .. 0): what the patched randn_like hands to the reference loop."""
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randn(shape, generator=g) for _ in range(S)])


def h3d_masks():
    """h3d_diffusion_new_trainer.py:194-221."""
    up, ha = [], []
    for i in [3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21]:
        up.extend([4 + (i - 1) * 3 + c for c in range(3)])
        up.extend([4 + 51 * 3 + (i - 1) * 6 + c for c in range(6)])
        up.extend([4 + 51 * 9 + i * 3 + c for c in range(3)])
    for i in range(22, 52):
        ha.extend([4 + (i - 1) * 3 + c for c in range(3)])
        ha.extend([4 + 51 * 3 + (i - 1) * 6 + c for c in range(6)])
        ha.extend([4 + 51 * 9 + i * 3 + c for c in range(3)])
    lo = list(range(0, 4)) + list(range(619, 623))
    for i in [0, 1, 2, 4, 5, 7, 8, 10, 11]:
        if i > 0:
            lo.extend([4 + (i - 1) * 3 + c for c in range(3)])
            lo.extend([4 + 51 * 3 + (i - 1) * 6 + c for c in range(6)])
        lo.extend([4 + 51 * 9 + i * 3 + c for c in range(3)])
    return up, ha, lo


def do_ddpm():
    m = mg.build_mdm("beatx")
    inp = synth.make_inputs(1, seed=1, variant="beatx")
    y = mg.y_of(inp, "beatx")
    d = create_gaussian_diffusion()
    assert d.num_timesteps == 1000
    g = torch.Generator().manual_seed(DDPM_SEED)
    real = torch.randn_like
    torch.randn_like = lambda x, **k: torch.randn(x.shape, generator=g)
    assert gd.th.randn_like is torch.randn_like
    try:
        dump = d.p_sample_loop(m, (1, 1536, 1, 32), noise=inp["noise"], clip_denoised=False, model_kwargs={"y": y},
                               dump_steps=DDPM_DUMPS)
    finally:
        torch.randn_like = real
    mg.save("ddpm1000", seed=np.int64(DDPM_SEED), dump_steps=np.array(DDPM_DUMPS), **{f"x_after_{i}": s for i, s in zip(DDPM_DUMPS, dump)})


def do_cfg():
    h3d = mg.build_mdm("h3d")
    inp = synth.make_inputs(2, seed=3, variant="h3d")
    t = torch.tensor([700, 40], dtype=torch.int64)
    y = mg.y_of(inp, "h3d")
    y["style_feature"] = inp["style_upper"]
    y2 = dict(y); y2["scale_audio"] = torch.ones(1) * 1.5; y2["scale_prompt"] = torch.tensor([3.0, 0.5])
    out = TwoClassifierFreeSampleModel(h3d)(inp["noise"], t, y2)
    mg.save("cfg_two", out=out, t=t, scale_audio=y2["scale_audio"], scale_prompt=y2["scale_prompt"])
    y3 = dict(y); y3["scale"] = torch.ones(1) * 2.5
    out_t = ClassifierFreeSampleModel(h3d)(inp["noise"], t, dict(y3))
    out_e = ClassifierFreeSampleModel(h3d, eval=True)(inp["noise"], t, dict(y3))
    mg.save("cfg_h3d_text", out=out_t, out_eval=out_e, t=t, scale=y3["scale"])
    inp1 = synth.make_inputs(1, seed=4, variant="h3d")
    y4 = mg.y_of(inp1, "h3d")
    y4["style_feature"] = {"upper_mask": inp1["style_upper"], "hands_mask": None, "lower_mask": inp1["style_lower"]}
    y4["scale"] = torch.ones(1) * 2.5
    t1 = torch.tensor([300], dtype=torch.int64)
    out_b = ClassifierFreeSampleModel_Bodypart(h3d)(inp1["noise"], t1, y4)
    y5 = mg.y_of(inp1, "h3d")
    y5["style_feature"] = {"upper_mask": inp1["style_upper"], "hands_mask": None, "lower_mask": inp1["style_lower"]}
    out_be = ClassifierFreeSampleModel_Bodypart(h3d, eval=True)(inp1["noise"], t1, y5)
    mg.save("cfg_bodypart1", out=out_b, out_eval=out_be, t=t1, scale=y4["scale"])


def do_h3d():
    g = torch.Generator().manual_seed(6)
    rv, recs = {}, []
    for dpart in synth.PART_DIMS_H3D:
        vq = mg.RVQVAE(mg.args_for("beatx"), dpart, 512, 512, 512, 2, 2, 512, 3, 3, "relu", None)
        vq.load_state_dict(synth.rvq_state_dict(dpart, seed=0), strict=True)
        vq.eval()
        lat = 5.0 * torch.randn(2, 32, 512, generator=g)
        rv[f"lat{dpart}"] = lat.clone()
        xq, idx, _, _ = vq.quantizer(lat.permute(0, 2, 1).clone(), sample_codebook_temp=0.5)
        rec = vq.latent2origin(lat.clone())[0]
        rv[f"idx{dpart}"] = idx
        rv[f"rec{dpart}"] = rec
        recs.append(rec)
    up, ha, lo = h3d_masks()
    assert (len(up), len(ha), len(lo)) == synth.PART_DIMS_H3D and len(set(up + ha + lo)) == 623
    rec_pose = torch.zeros(2, 128, 623)
    rec_pose[..., up] = recs[0]
    rec_pose[..., ha] = recs[1]
    rec_pose[..., lo] = recs[2]
    mg.save("h3d_decode", rec_pose=rec_pose, mask_upper=np.array(up), mask_hands=np.array(ha), mask_lower=np.array(lo), **rv)


if __name__ == "__main__":
    what = sys.argv[1:] or ["cfg", "h3d", "ddpm"]
    for w in what:
        {"ddpm": do_ddpm, "cfg": do_cfg, "h3d": do_h3d}[w]()
