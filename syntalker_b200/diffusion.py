"""Drop-in for diffusion/model_util.py `create_gaussian_diffusion` and diffusion/respace.py
`SpacedDiffusion` (sampling half of diffusion/gaussian_diffusion.py `GaussianDiffusion`).

`diffusion.p_sample_loop(model, shape, noise=..., clip_denoised=False, model_kwargs={'y': ...}, ...)` and
`diffusion.ddim_sample_loop(...)` keep the reference's signatures (gaussian_diffusion.py:607-670, 888-935)
and return the final sample [B,1536,1,32]; the whole loop runs inside libsyntalker_b200.so (st_sample).
Options no trainer uses on this path raise NotImplementedError instead of silently differing.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np
import torch

from . import _lib, schedule as _sch
from .cfg_sampler import _Wrapper, unwrap
from .denoiser import MDM, Guidance, _dev_f32


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


get_named_beta_schedule = _sch.get_named_beta_schedule
space_timesteps = _sch.space_timesteps


class SpacedDiffusion(_sch.Tables):
    """respace.py:64-121 over gaussian_diffusion.py:103-197; only START_X / FIXED_SMALL (what the factory builds)."""

    def __init__(self, use_timesteps, betas, model_mean_type=ModelMeanType.START_X, model_var_type=ModelVarType.FIXED_SMALL,
                 loss_type=None, rescale_timesteps=False, **_unused):
        if model_mean_type != ModelMeanType.START_X or model_var_type != ModelVarType.FIXED_SMALL or rescale_timesteps:
            raise NotImplementedError("only START_X / FIXED_SMALL / rescale_timesteps=False (diffusion/model_util.py:15-24)")
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(betas)
        new_betas, self.timestep_map = _sch.respace(betas, self.use_timesteps)
        super().__init__(new_betas)
        self.model_mean_type, self.model_var_type, self.rescale_timesteps = model_mean_type, model_var_type, rescale_timesteps
        self._sched = {}

    # ---- native schedule handles, one per (mode, eta) ----
    def _native(self, mode, eta=0.0):
        key = (mode, float(eta))
        if key not in self._sched:
            coef = _sch.ddim_coefs(self, eta) if mode == _lib.ST_MODE_DDIM else _sch.ddpm_coefs(self)
            tmap = np.asarray(self.timestep_map, dtype=np.int32)
            h = C.c_void_p()
            _lib.check(_lib.lib().st_schedule_create(self.num_timesteps, mode, tmap.ctypes.data, np.ascontiguousarray(coef).ctypes.data,
                                                    C.byref(h)))
            self._sched[key] = (h, coef)
        return self._sched[key]

    def __del__(self):
        try:
            for h, _ in self._sched.values():
                _lib.lib().st_schedule_destroy(h)
        except Exception:
            pass

    @staticmethod
    def _split(model):
        base, wrapper = unwrap(model)
        if not isinstance(base, MDM):
            raise TypeError("model must be a syntalker_b200 MDM (optionally inside nn.DataParallel / DDP and / or a syntalker_b200 CFG "
                            "wrapper); any other torch nn.Module cannot be run by the native sampler and there is no fallback")
        return base, wrapper

    def _run(self, mode, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, eta, skip_timesteps,
             init_image, randomize_class, cond_fn_with_grad, const_noise, consume_rng, noise_tape=None):
        if clip_denoised:
            raise NotImplementedError("clip_denoised=True: every caller on this path passes False (trainer:447, h3d:563, demo:512)")
        if denoised_fn is not None or cond_fn is not None or cond_fn_with_grad or randomize_class:
            raise NotImplementedError("denoised_fn / cond_fn / randomize_class are not used on the sampling path")
        if skip_timesteps or init_image is not None:
            raise NotImplementedError("skip_timesteps / init_image are not used on the sampling path")
        base, wrapper = self._split(model)
        y = dict((model_kwargs or {}).get("y", {}))
        dev = base.device
        assert isinstance(shape, (tuple, list))
        B = int(shape[0])
        if tuple(shape) != (B, 1536, 1, 32):
            raise ValueError(f"shape must be (B,1536,1,32), got {tuple(shape)}")
        if 'inpainting_mask' in y:
            raise NotImplementedError("inpainting is not used by the trainers (gaussian_diffusion.py:316-320)")
        with torch.cuda.device(dev):
            img = noise if noise is not None else torch.randn(*shape, device=dev)      # gaussian_diffusion.py:700-703
            img = _dev_f32(img, "noise", dev)
            styles = wrapper.styles(y) if wrapper is not None else None
            base.encode_cond(y, styles=styles)
            g = wrapper.guidance(y) if wrapper is not None else Guidance(
                _lib.ST_CFG_NONE, flags=(_lib.ST_FLAG_UNCOND if y.get("uncond", False) else 0) |
                (_lib.ST_FLAG_UNCOND_AUDIO if y.get("uncond_audio", False) else 0))
            sched, coef = self._native(mode, eta)
            S = self.num_timesteps
            sig_col = 2 if mode == _lib.ST_MODE_DDPM else 4
            need_noise = bool(np.any(coef[:, sig_col] != 0))
            out = torch.empty_like(img)
            L, h, gs = _lib.lib(), base.handle, g.struct(B)
            if noise_tape is not None:           # extension: caller-supplied eps_k, [S,B,1536,1,32] in draw order
                tape = _dev_f32(noise_tape, "noise_tape", dev)
                if tuple(tape.shape) != (S, B, 1536, 1, 32):
                    raise ValueError(f"noise_tape must be [{S},{B},1536,1,32], got {tuple(tape.shape)}")
                _lib.check(L.st_sample(h, sched, C.byref(gs), img.data_ptr(), tape.data_ptr(), B, out.data_ptr(), _lib.stream_ptr()))
            elif need_noise:
                # the reference draws th.randn_like(x) once per step, t = S-1..0, also at t == 0 (:541, :781): same draws, same
                # order, handed to the native loop one chunk of steps at a time (a whole 1000-step tape would be B x 197 MB)
                G = int(L.st_sample_chunk(sched))
                buf = torch.empty((G, B, 1536, 1, 32), device=dev, dtype=torch.float32)
                _lib.check(L.st_sample_begin(h, sched, C.byref(gs), img.data_ptr(), B, _lib.stream_ptr()))
                for _ in range(S // G):
                    for i in range(G):
                        buf[i].normal_()
                        if const_noise:
                            buf[i] = buf[i][[0]].repeat(B, 1, 1, 1)      # gaussian_diffusion.py:543-544
                    _lib.check(L.st_sample_run(h, G, buf.data_ptr(), _lib.stream_ptr()))
                _lib.check(L.st_sample_end(h, out.data_ptr(), _lib.stream_ptr()))
            else:
                if consume_rng:                  # ddim_sample draws an (unused, sigma = 0) randn_like every step (:781)
                    for _ in range(S):
                        torch.empty_like(img).normal_()
                _lib.check(L.st_sample(h, sched, C.byref(gs), img.data_ptr(), None, B, out.data_ptr(), _lib.stream_ptr()))
        return out

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                      device=None, progress=False, skip_timesteps=0, init_image=None, randomize_class=False,
                      cond_fn_with_grad=False, dump_steps=None, const_noise=False, noise_tape=None):
        if dump_steps is not None:
            raise NotImplementedError("dump_steps is not used on the sampling path (trainer:453 passes None)")
        return self._run(_lib.ST_MODE_DDPM, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, 0.0,
                         skip_timesteps, init_image, randomize_class, cond_fn_with_grad, const_noise, True, noise_tape)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                         device=None, progress=False, eta=0.0, skip_timesteps=0, init_image=None, randomize_class=False,
                         cond_fn_with_grad=False, dump_steps=None, const_noise=False, consume_rng=True, noise_tape=None):
        """`consume_rng` (extension, default True like the reference): draw the per-step randn_like the reference's ddim_sample
        draws even at eta = 0 (gaussian_diffusion.py:781), so that the torch generator stands where it would after the reference's
        loop (e.g. the next window's start noise); False skips those S small launches."""
        if dump_steps is not None:
            raise NotImplementedError()          # gaussian_diffusion.py:912-913
        if const_noise is True:
            raise NotImplementedError()          # gaussian_diffusion.py:914-915
        return self._run(_lib.ST_MODE_DDIM, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, eta,
                         skip_timesteps, init_image, randomize_class, cond_fn_with_grad, False, consume_rng, noise_tape)

    def training_losses(self, *a, **k):
        raise NotImplementedError("training is out of scope for syntalker_b200 (SURVEY.md §8)")


def create_gaussian_diffusion(DiffusionClass=SpacedDiffusion, use_ddim=False, timestep_respacing=None):
    """diffusion/model_util.py:8-50: cosine schedule, 1000 steps, predict x0, fixed-small sigma; 'ddim50' when
    use_ddim. `timestep_respacing` (extension) overrides the spacing, e.g. 'ddim10' for BASELINE config 1."""
    steps = 1000
    if timestep_respacing is None:
        timestep_respacing = "ddim50" if use_ddim else [steps]
    betas = get_named_beta_schedule("cosine", steps, 1.0)
    return DiffusionClass(use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas,
                          model_mean_type=ModelMeanType.START_X, model_var_type=ModelVarType.FIXED_SMALL,
                          loss_type=None, rescale_timesteps=False)
