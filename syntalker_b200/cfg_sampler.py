"""Drop-ins for diffusion/cfg_sampler.py: the four classifier-free-guidance wrappers.

Each keeps the reference's constructor and `forward(x, timesteps, y)`; instead of calling the model
2 / 3 / 9 times from Python, it hands the native library a guidance descriptor, which runs the
de-duplicated evaluations as one batched pass and mixes them with the reference's fp32 formula
(cfg_sampler.py:28, :54, :86-114).
"""
from __future__ import annotations

import torch

from . import _lib
from .denoiser import Guidance

mask_dict = {"upper_mask": list(range(0, 512)), "hands_mask": list(range(512, 1024)), "lower_mask": list(range(1024, 1536))}


def unwrap(model):
    """(base MDM, outermost CFG wrapper or None) behind any nesting of nn.DataParallel / DDP (`.module`, train.py:87-94) and the
    CFG wrappers (`.model`): the trainers wrap the DataParallel model in a CFG wrapper (h3d_diffusion_new_trainer.py:922)."""
    wrapper, m = None, model
    for _ in range(16):
        if isinstance(m, _Wrapper):
            wrapper = wrapper or m
            m = m.model
        elif hasattr(m, "module") and isinstance(m, torch.nn.Module) and not isinstance(m, _mdm_type()):
            m = m.module
        else:
            break
    return m, wrapper


def _mdm_type():
    from .denoiser import MDM
    return MDM


class _Wrapper(torch.nn.Module):
    def __init__(self, model, eval=False):
        super().__init__()
        self.model = model
        self.eval_metric = eval

    @property
    def base(self):
        return unwrap(self)[0]

    def guidance(self, y):
        raise NotImplementedError

    def styles(self, y):
        return None

    def forward(self, x, timesteps, y=None):
        base = self.base
        st = self.styles(y)
        if st is not None:
            base.encode_cond(y, styles=st)
        return base.forward(x, timesteps, y, _guidance=self.guidance(y))


class ClassifierFreeSampleModel(_Wrapper):
    """cfg_sampler.py:10-28: out_uncond + scale * (out - out_uncond); eval=True returns out_uncond."""

    def guidance(self, y):
        if self.eval_metric:
            flags = _lib.ST_FLAG_UNCOND | (_lib.ST_FLAG_UNCOND_AUDIO if self.base.variant == "h3d" else 0)
            return Guidance(_lib.ST_CFG_NONE, flags=flags)
        return Guidance(_lib.ST_CFG_TEXT, scale=y["scale"])


class TwoClassifierFreeSampleModel(_Wrapper):
    """cfg_sampler.py:31-54: uu + s_audio*(u_text - uu) + s_prompt*(u_audio - uu)."""

    def guidance(self, y):
        return Guidance(_lib.ST_CFG_TWO, scale=y["scale_audio"], scale2=y["scale_prompt"])


class TwoClassifierFreeSampleModel_Bodypart(_Wrapper):
    """cfg_sampler.py:57-117: per body part pick the prompt and (s_audio, s_prompt), keep that channel range.
    y['style_feature'] = {'upper_mask': [B,256]|None, 'hands_mask': ..., 'lower_mask': ...}."""

    def __init__(self, model, eval=False):
        super().__init__(model, eval)
        self.latent_dim = 1536
        self.audio_scale = 1
        self.prompt_scale = 4

    def styles(self, y):
        sf = y["style_feature"]
        if self.eval_metric:
            return [sf["lower_mask"], None, None]
        return [sf.get("upper_mask"), sf.get("hands_mask"), sf.get("lower_mask")]

    def guidance(self, y):
        if self.eval_metric:     # cfg_sampler.py:77-80: TwoCFG with (audio_scale, 0) and a null prompt
            return Guidance(_lib.ST_CFG_TWO, scale=[float(self.audio_scale)], scale2=[0.0])
        return Guidance(_lib.ST_CFG_BODYPART, audio_scale=self.audio_scale, prompt_scale=self.prompt_scale)


class ClassifierFreeSampleModel_Bodypart(_Wrapper):
    """cfg_sampler.py:125-167: per prompted body part one evaluation (that prompt, audio masked), one evaluation with the null
    prompt (audio kept) = out_uncond; parts without a prompt keep out_uncond; out_uncond + scale * (out - out_uncond).
    eval=True returns model(x, t, uncond=True) (:143-146)."""

    def __init__(self, model, eval=False):
        super().__init__(model, eval)
        self.latent_dim = 1536

    def styles(self, y):
        sf = y["style_feature"]
        if self.eval_metric:
            return [sf["lower_mask"], None, None]
        return [sf.get("upper_mask"), sf.get("hands_mask"), sf.get("lower_mask")]

    def guidance(self, y):
        if self.eval_metric:
            return Guidance(_lib.ST_CFG_NONE, flags=_lib.ST_FLAG_UNCOND)
        return Guidance(_lib.ST_CFG_BODYPART1, scale=y["scale"])
