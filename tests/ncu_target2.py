"""Target for `ncu --set full` captures of the kernels outside the headline loop (not a pytest file).
    ncu --profile-from-start off ... python tests/ncu_target2.py decode|encode|step|ddpm
 decode: st_rvq_decode x3 + 330-d assembly at B = 32 (vq_select, trunk kernel in conv mode, generic tcgen05 kernel, pose330, trans)
 encode: RVQVAE.map2latent at B = 32 (gemm_simt for the D-channel / strided convs) + the conditioning encoder (wav_first, strided tcgen05 convs)
 step:   one DDIM-5 loop at B = 32 with CFG (tokens_step, the K = 1536 two-tensor-map GEMM, the fused qkv + attention launch)
 ddpm:   one 50-step chunk of the DDPM z recursion at B = 32 (W_x eps GEMM, transposes, tokens_step with the noise term)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
from syntalker_b200.pipeline import load_mean_std, pose_assemble_330
from syntalker_b200.vq import RVQVAE

what = sys.argv[1] if len(sys.argv) > 1 else "decode"
B = 32
torch.set_grad_enabled(False)
vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
g = torch.Generator().manual_seed(5)
ms = {k: v.cuda() for k, v in load_mean_std().items()}
if what == "decode":
    lats = [(5.0 * torch.randn(B, 32, 512, generator=g)).cuda() for _ in range(3)]

    def fn():
        recs = [v.latent2origin(l.clone())[0] for v, l in zip(vqs, lats)]
        return pose_assemble_330(recs[0], recs[1], recs[2], ms, None)
elif what == "encode":
    poses = [torch.randn(B, 128, d, generator=g).cuda() for d in synth.PART_DIMS_BEATX]
    model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx", seed=0))
    inp = synth.make_inputs(B, seed=1, variant="beatx")
    y = {k: inp[k].cuda() for k in ("audio", "word", "seed")}

    def fn():
        model.encode_cond(y, force=True)
        return [v.map2latent(p) for v, p in zip(vqs, poses)]
else:
    variant = "beatx_motionclip" if what == "step" else "beatx"
    model = MDM(None).load_state_dict(synth.mdm_state_dict(variant, seed=0))
    inp = synth.make_inputs(B, seed=1, variant=variant)
    y = {k: inp[k].cuda() for k in ("audio", "word", "seed")}
    x0 = inp["noise"].cuda()
    if what == "step":
        y["style_feature"] = inp["style_feature"].cuda(); y["scale"] = torch.ones(1) * 2.0
        w = ClassifierFreeSampleModel(model)
        d = create_gaussian_diffusion(timestep_respacing="ddim5")
        fn = lambda: d.ddim_sample_loop(w, (B, 1536, 1, 32), noise=x0, clip_denoised=False, model_kwargs={"y": y}, consume_rng=False)
    else:
        d = create_gaussian_diffusion(timestep_respacing=[50])
        fn = lambda: d.p_sample_loop(model, (B, 1536, 1, 32), noise=x0, clip_denoised=False, model_kwargs={"y": y})
for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
