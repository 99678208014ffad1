"""Oracle: the trainer's window loop over a long clip, batched over clips.

Follows diffusion_rvqvae_trainer.py:413-431 (window count, slices of words / audio, seed hand-off from
`last_sample[:, -pre_frames:]`), :445-474 (sample, `squeeze().permute(1,0)`, keep all 32 tokens of window 0 and the last
28 of every later one), :476-482 (concatenate, x vqvae_latent_scale, ONE latent2origin per body part over the whole
clip) and :484-531 (330-d assembly, oracle/pose.py). The reference's glue is written for bs = 1; rows are independent,
so the oracle runs B clips side by side (same arithmetic per row).
"""
import torch

from . import diffusion as odiff, pose as opose, rvq as orvq

POSE_LENGTH, PRE_FRAMES, SQUEEZE = 128, 4, 4
ROUND_L = POSE_LENGTH - PRE_FRAMES * SQUEEZE          # 112 (trainer:415)
AUDIO_PER_FRAME = 16000 // 30                         # 533 (trainer:422)


def n_windows(n_frames):
    return (n_frames - PRE_FRAMES * SQUEEZE) // ROUND_L                                   # roundt, trainer:413


def window_inputs(audio_long, word_long, i):
    w = word_long[:, i * ROUND_L:(i + 1) * ROUND_L + PRE_FRAMES * SQUEEZE]                # trainer:420
    a = audio_long[:, i * AUDIO_PER_FRAME * ROUND_L:(i + 1) * AUDIO_PER_FRAME * ROUND_L + AUDIO_PER_FRAME * PRE_FRAMES * SQUEEZE]
    return a, w


def long_clip_latents(sched, model_fn, audio_long, word_long, seed0, x_init, extra_y=None):
    """x_init [R,B,1536,1,32] -> stitched latents [B, 32 + 28 (R-1), 1536] (DDIM, eta = 0)."""
    R = x_init.shape[0]
    keep, last = [], None
    for i in range(R):
        a, w = window_inputs(audio_long, word_long, i)
        seed = seed0 if i == 0 else last[:, -PRE_FRAMES:, :]                              # trainer:428-431
        y = dict(extra_y or {})
        y.update({"audio": a, "word": w, "seed": seed})
        sample = odiff.ddim_sample_loop(sched, model_fn, x_init[i], y)                    # [B,1536,1,32]
        tok = sample[:, :, 0, :].permute(0, 2, 1)                                         # trainer:457, per row
        last = tok.clone()
        keep.append(tok if i == 0 else tok[:, PRE_FRAMES:])                               # trainer:462-469
    return torch.cat(keep, dim=1)


def long_clip_330(sched, model_fn, vq_weights, audio_long, word_long, seed0, x_init, ms, jaw_aa=None, latent_scale=5.0, extra_y=None):
    lat = long_clip_latents(sched, model_fn, audio_long, word_long, seed0, x_init, extra_y)
    parts = [lat[..., :512] * latent_scale, lat[..., 512:1024] * latent_scale, lat[..., 1024:] * latent_scale]   # trainer:476-478
    recs, idxs = zip(*[orvq.latent2origin(w, p) for w, p in zip(vq_weights, parts)])
    pose, trans = opose.assemble_330(recs[0], recs[1], recs[2], ms, jaw_aa)
    return pose, trans, lat, idxs
