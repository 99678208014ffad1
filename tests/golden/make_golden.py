"""Generate the golden fixtures by running the REAL reference modules (build container only).

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

The reference cannot travel to the GPU box, so its outputs on seeded synthetic weights/inputs
(syntalker_b200/synth.py) are committed here. The four import shims are the ones SURVEY.md §8(c) lists;
none touches arithmetic. The L4 trainer glue cannot be imported (needs smplx/librosa/...), so the 330-d
assembly golden is produced by the trainer's statements (diffusion_rvqvae_trainer.py:484-531) re-typed
around the reference's own utils/rotation_conversions.py and mean_std/*.npy.
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SYNTALKER_REF", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
os.chdir(REF)
for m in ("lmdb", "fasttext"):
    sys.modules[m] = types.ModuleType(m)
torch.Tensor.cuda = lambda self, *a, **k: self
from dataloaders.build_vocab import Vocab  # noqa: E402
import __main__  # noqa: E402
__main__.Vocab = Vocab

from syntalker_b200 import synth  # noqa: E402
from models.denoiser import MDM  # noqa: E402
from models.denoiser_h3d import MDM as MDM_H3D  # noqa: E402
from models.vq.model import RVQVAE  # noqa: E402
from diffusion.model_util import create_gaussian_diffusion  # noqa: E402
from diffusion.respace import SpacedDiffusion, space_timesteps  # noqa: E402
from diffusion import gaussian_diffusion as gd  # noqa: E402
from diffusion.cfg_sampler import ClassifierFreeSampleModel, TwoClassifierFreeSampleModel_Bodypart  # noqa: E402
import utils.rotation_conversions as rc  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(8)


def args_for(variant):
    return SimpleNamespace(vqvae_type="rvqvae", use_motionclip=(variant == "beatx_motionclip"),
                           audio_rep="onset+amplitude", audio_f=256, word_f=256,
                           data_path=REF + "/datasets/BEAT_SMPL/beat_v2.0.0/beat_english_v2.0.0/",
                           t_fix_pre=False, vqvae_squeeze_scale=4, num_quantizers=6, shared_codebook=False,
                           quantize_dropout_prob=0.2, mu=0.99)


def build_mdm(variant):
    cls = MDM_H3D if variant == "h3d" else MDM
    m = cls(args_for(variant))
    m.load_state_dict(synth.mdm_state_dict(variant, seed=0), strict=True)
    return m.eval()


def build_vq(dim):
    vq = RVQVAE(args_for("beatx"), dim, 512, 512, 512, 2, 2, 512, 3, 3, "relu", None)
    vq.load_state_dict(synth.rvq_state_dict(dim, seed=0), strict=True)
    return vq.eval()


def spaced(spec):
    """create_gaussian_diffusion() (model_util.py:8-50) with an arbitrary respacing."""
    betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
    return SpacedDiffusion(use_timesteps=space_timesteps(1000, spec), betas=betas,
                           model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                           loss_type=gd.LossType.MSE, rescale_timesteps=False, lambda_vel=0.0, lambda_rcxyz=0.0,
                           lambda_fc=0.0)


def y_of(inp, variant):
    y = {"audio": inp["audio"], "word": inp["word"], "seed": inp["seed"]}
    if "style_feature" in inp:
        y["style_feature"] = inp["style_feature"]
    return y


def save(name, **arrs):
    out = {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


def main():
    # ---- 1. schedule tables --------------------------------------------------------------------------
    tabs = {}
    for tag, d in (("ddim50", create_gaussian_diffusion(use_ddim=True)), ("ddpm1000", create_gaussian_diffusion()),
                   ("ddim10", spaced("ddim10")), ("sec20", spaced([20]))):
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_log_variance_clipped", "posterior_mean_coef1",
                  "posterior_mean_coef2"):
            tabs[f"{tag}.{k}"] = getattr(d, k)
        tabs[f"{tag}.timestep_map"] = np.array(d.timestep_map, dtype=np.int64)
    save("schedule", **tabs)

    # ---- 2. single denoiser evaluations --------------------------------------------------------------
    models = {}
    for variant in synth.VARIANTS:
        m = models[variant] = build_mdm(variant)
        inp = synth.make_inputs(2, seed=1, variant=variant)
        y = y_of(inp, variant)
        if variant == "h3d":
            y["style_feature"] = inp["style_upper"]
        t = torch.tensor([980, 20], dtype=torch.int64)
        taps = {}
        hooks = [m.mytimmblocks[i].register_forward_hook(lambda _m, _i, o, i=i: taps.__setitem__(f"block{i}", o.clone()))
                 for i in (0, 7)]
        out = m(inp["noise"], t, y)
        for h in hooks:
            h.remove()
        extra = {}
        if variant == "beatx":
            extra = {"block0": taps["block0"], "block7": taps["block7"],
                     "wav": m.WavEncoder(inp["audio"])[:, ::8].contiguous()}
        if variant != "beatx":
            yu = dict(y); yu["uncond"] = True
            extra["out_uncond"] = m(inp["noise"], t, yu)
        if variant == "h3d":
            ya = dict(y); ya["uncond_audio"] = True
            extra["out_uncond_audio"] = m(inp["noise"], t, ya)
        save(f"mdm_{variant}", out=out, t=t, **extra)

    # ---- 3. CFG wrappers -----------------------------------------------------------------------------
    inp = synth.make_inputs(2, seed=1, variant="beatx_motionclip")
    y = y_of(inp, "beatx_motionclip"); y["scale"] = torch.ones(1) * 2.0
    out = ClassifierFreeSampleModel(models["beatx_motionclip"])(inp["noise"], torch.tensor([500, 500]), y)
    save("cfg_text", out=out)
    inp = synth.make_inputs(1, seed=2, variant="h3d")
    y = y_of(inp, "h3d")
    y["style_feature"] = {"upper_mask": inp["style_upper"], "hands_mask": None, "lower_mask": inp["style_lower"]}
    out = TwoClassifierFreeSampleModel_Bodypart(models["h3d"])(inp["noise"], torch.tensor([300]), y)
    save("cfg_bodypart", out=out)

    # ---- 4. sampling loops ---------------------------------------------------------------------------
    m = models["beatx"]
    inp = synth.make_inputs(1, seed=1, variant="beatx")
    y = y_of(inp, "beatx")
    d10 = spaced("ddim10")
    s10 = d10.ddim_sample_loop(m, (1, 1536, 1, 32), noise=inp["noise"], clip_denoised=False, model_kwargs={"y": y})
    d20 = spaced([20])
    torch.manual_seed(123)
    p20 = d20.p_sample_loop(m, (1, 1536, 1, 32), noise=inp["noise"], clip_denoised=False, model_kwargs={"y": y})
    save("loops", ddim10=s10, ddpm_sec20=p20)

    # ---- 5. RVQ decode -------------------------------------------------------------------------------
    vqs = {d: build_vq(d) for d in synth.PART_DIMS_BEATX}
    g = torch.Generator().manual_seed(5)
    rv = {}
    recs = []
    for d in synth.PART_DIMS_BEATX:
        lat = 5.0 * torch.randn(2, 32, 512, generator=g)
        rv[f"lat{d}"] = lat.clone()
        x = lat.permute(0, 2, 1)
        xq, idx, _, _ = vqs[d].quantizer(x.clone(), sample_codebook_temp=0.5)
        rec = vqs[d].latent2origin(lat.clone())[0]
        rv[f"idx{d}"] = idx; rv[f"rec{d}"] = rec; rv[f"xq{d}"] = xq[:, ::8].contiguous()
        recs.append(rec)
    save("rvq", **rv)

    # ---- 6. 330-d assembly (trainer statements around the reference's rotation_conversions) ------------
    mean = np.load(REF + "/mean_std/beatx_2_330_mean.npy"); std = np.load(REF + "/mean_std/beatx_2_330_std.npy")
    tmean = np.load(REF + "/mean_std/beatx_2_trans_mean.npy"); tstd = np.load(REF + "/mean_std/beatx_2_trans_std.npy")
    np.savez(os.path.join(ROOT, "syntalker_b200", "data", "beatx_mean_std.npz"),
             mean=mean.astype(np.float32), std=std.astype(np.float32),
             trans_mean=tmean.astype(np.float32), trans_std=tstd.astype(np.float32))

    def masks(joints):
        return [i * 6 + c for i in joints for c in range(6)]
    UJ = [3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21]; HJ = list(range(25, 55)); LJ = [0, 1, 2, 4, 5, 7, 8, 10, 11]
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))

    def assemble(rec_upper, rec_hands, rec_lower, jaw):
        rec_trans_v = rec_lower[..., -3:] * T(tstd) + T(tmean)
        rec_trans = torch.cumsum(rec_trans_v, dim=-2)
        rec_trans[..., 1] = rec_trans_v[..., 1]
        rec_lower = rec_lower[..., :-3]
        rec_upper = rec_upper * T(std[masks(UJ)]) + T(mean[masks(UJ)])
        rec_hands = rec_hands * T(std[masks(HJ)]) + T(mean[masks(HJ)])
        rec_lower = rec_lower * T(std[masks(LJ)]) + T(mean[masks(LJ)])
        bs, n = rec_lower.shape[:2]
        pose = torch.zeros(bs * n, 165)
        for part, J in ((rec_upper, UJ), (rec_hands, HJ), (rec_lower[:, :, :54], LJ)):
            aa = rc.matrix_to_axis_angle(rc.rotation_6d_to_matrix(part.reshape(bs, n, len(J), 6))).reshape(bs * n, len(J) * 3)
            sel = torch.tensor([3 * j + c for j in J for c in range(3)])
            for i in range(bs * n):
                pose[i, sel] = aa[i]
        pose[:, 66:69] = jaw.reshape(bs * n, 3)
        pose = rc.matrix_to_rotation_6d(rc.axis_angle_to_matrix(pose.reshape(bs * n, 55, 3))).reshape(bs, n, 330)
        return pose, rec_trans

    jaw = 0.1 * torch.randn(2, 128, 3, generator=g)
    pose, trans = assemble(recs[0], recs[1], recs[2], jaw)
    save("pose", rec_pose=pose, rec_trans=trans, jaw=jaw)

    # ---- 7. end to end, config 1: B=1, DDIM-10 -> x5 -> latent2origin x3 -> 330-d ---------------------
    sample = s10.squeeze().permute(1, 0).unsqueeze(0)
    lats = [sample[..., :512] * 5.0, sample[..., 512:1024] * 5.0, sample[..., 1024:] * 5.0]
    recs1 = [vqs[d].latent2origin(l.clone())[0] for d, l in zip(synth.PART_DIMS_BEATX, lats)]
    pose1, trans1 = assemble(recs1[0], recs1[1], recs1[2], torch.zeros(1, 128, 3))
    save("e2e_config1", rec_pose=pose1, rec_trans=trans1)

    # ---- 8. h3d 623-d scatter mask sizes -------------------------------------------------------------
    print("done")


if __name__ == "__main__":
    main()
