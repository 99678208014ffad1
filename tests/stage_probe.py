"""GPU probe: CUDA-event time of each stage of a config-2 window batch (cond encode / sampling loop / decode / pose)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
from syntalker_b200.pipeline import load_mean_std, pose_assemble_330
from syntalker_b200.vq import RVQVAE

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.set_grad_enabled(False)
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
diff = create_gaussian_diffusion(use_ddim=True)
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
d = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
y = {"audio": d["audio"], "word": d["word"], "seed": d["seed"], "style_feature": d["style_feature"], "scale": torch.ones(1) * 2.0}
ms = load_mean_std()


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


for it in range(4):
    t0 = ev()
    model.encode_cond(y, force=True)
    t1 = ev()
    sample = diff.ddim_sample_loop(w, (B, 1536, 1, 32), noise=d["noise"], clip_denoised=False, model_kwargs={"y": y})
    t2 = ev()
    lat = sample[:, :, 0, :].permute(0, 2, 1) * 5.0
    recs = [vqs[k].latent2origin(lat[..., 512 * k:512 * (k + 1)].contiguous())[0] for k in range(3)]
    t3 = ev()
    pose, trans = pose_assemble_330(recs[0], recs[1], recs[2], ms)
    t4 = ev()
    torch.cuda.synchronize()
    print(f"iter {it}: cond {t0.elapsed_time(t1):.2f} ms | sample {t1.elapsed_time(t2):.2f} ms | decode x3 {t2.elapsed_time(t3):.2f} ms | pose {t3.elapsed_time(t4):.2f} ms", flush=True)
