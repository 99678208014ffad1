"""The window pipeline a trainer's `_g_test` runs per 128-frame window batch, as one native call with HOST
buffers: H2D -> conditioning -> sampling loop (+CFG) -> x latent_scale -> latent2origin x3 -> 330-d -> D2H
(diffusion_rvqvae_trainer.py:433-531). This is the call bench.py times for the end-to-end number.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .cfg_sampler import _Wrapper, unwrap
from .denoiser import Guidance


def load_mean_std():
    """mean_std/beatx_2_330_{mean,std}.npy and beatx_2_trans_{mean,std}.npy (fixtures of the reference)."""
    d = np.load(os.path.join(os.path.dirname(__file__), "data", "beatx_mean_std.npz"))
    return {k: torch.from_numpy(d[k].astype(np.float32)) for k in d.files}


def pose_assemble_330(rec_upper, rec_hands, rec_lower, ms, jaw_aa=None):
    """Device tensors in, device tensors out (st_pose_assemble_330)."""
    dev = rec_upper.device
    B, n, _ = rec_upper.shape
    f = lambda t: t.to(dev).float().contiguous()
    up, ha, lo = f(rec_upper), f(rec_hands), f(rec_lower)
    mean, std, tm, ts = f(ms["mean"]), f(ms["std"]), f(ms["trans_mean"]), f(ms["trans_std"])
    jaw = f(jaw_aa) if jaw_aa is not None else None
    pose = torch.empty((B, n, 330), device=dev)
    trans = torch.empty((B, n, 3), device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().st_pose_assemble_330(up.data_ptr(), ha.data_ptr(), lo.data_ptr(), mean.data_ptr(), std.data_ptr(),
                                                  tm.data_ptr(), ts.data_ptr(), jaw.data_ptr() if jaw is not None else None,
                                                  B, n, pose.data_ptr(), trans.data_ptr(), _lib.stream_ptr()))
    return pose, trans


def pose_assemble_623(rec_upper, rec_hands, rec_lower):
    dev = rec_upper.device
    B, n, _ = rec_upper.shape
    f = lambda t: t.to(dev).float().contiguous()
    up, ha, lo = f(rec_upper), f(rec_hands), f(rec_lower)
    pose = torch.empty((B, n, 623), device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().st_pose_assemble_623(up.data_ptr(), ha.data_ptr(), lo.data_ptr(), B, n, pose.data_ptr(), _lib.stream_ptr()))
    return pose


class Window330:
    """Pinned host staging + one st_generate_330_host call per window batch."""

    def __init__(self, model, diffusion, vq_upper, vq_hands, vq_lower, B, use_ddim=True, eta=0.0, latent_scale=5.0, ms=None):
        self.base, self.wrapper = unwrap(model)          # through nn.DataParallel / DDP and the CFG wrappers
        self.diffusion, self.B = diffusion, B
        self.vqs = (vq_upper, vq_hands, vq_lower)
        self.mode = _lib.ST_MODE_DDIM if use_ddim else _lib.ST_MODE_DDPM
        self.eta, self.latent_scale = eta, float(latent_scale)
        self.ms = ms or load_mean_std()
        pin = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype).pin_memory()
        sdim = 512 if self.base.variant == "beatx_motionclip" else 256
        self.h = {
            "audio": pin(B, 68224, 2), "word": pin(B, 128, dtype=torch.int32), "seed": pin(B, 4, 1536),
            "style": [pin(B, sdim) for _ in range(3)], "x_init": pin(B, 1536, 1, 32), "jaw": pin(B, 128, 3),
            "pose": pin(B, 128, 330), "trans": pin(B, 128, 3), "sample": pin(B, 1536, 1, 32),
        }
        self.h_ms = {k: v.clone().contiguous() for k, v in self.ms.items()}
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run_device(self, audio, word, seed, x_init, y=None, styles=None, jaw_aa=None, noise_tape=None, out=None):
        """Same pipeline with every input already resident in HBM (CUDA tensors); no host sync.
        Returns (rec_pose, rec_trans, sample) CUDA tensors (st_generate_330)."""
        B, dev = self.B, self.base.device
        y = dict(y or {})
        if styles is None and self.base.variant != "beatx":
            sf = y.get("style_feature")
            styles = [sf.get("upper_mask"), sf.get("hands_mask"), sf.get("lower_mask")] if isinstance(sf, dict) else [sf, None, None]
        styles = styles or [None, None, None]
        c = _lib.StCond()
        c.audio, c.word, c.seed = audio.data_ptr(), word.data_ptr(), seed.data_ptr()
        assert audio.is_cuda and audio.dtype == torch.float32 and word.dtype == torch.int32 and seed.is_contiguous()
        for k in range(3):
            c.style[k] = styles[k].data_ptr() if (styles[k] is not None and self.base.variant != "beatx") else None
        if not hasattr(self, "_ms_dev"):
            m = self.ms
            self._ms_dev = torch.cat([m["mean"], m["std"], m["trans_mean"], m["trans_std"]]).float().to(dev).contiguous()
        if out is None:
            out = (torch.empty((B, 128, 330), device=dev), torch.empty((B, 128, 3), device=dev), torch.empty((B, 1536, 1, 32), device=dev))
        g = self.wrapper.guidance(y) if self.wrapper is not None else Guidance(_lib.ST_CFG_NONE)
        sched, _ = self.diffusion._native(self.mode, self.eta)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().st_generate_330(
                self.base.handle, sched, C.byref(g.struct(B)), self.vqs[0].handle, self.vqs[1].handle, self.vqs[2].handle,
                C.byref(c), x_init.data_ptr(), noise_tape.data_ptr() if noise_tape is not None else None,
                jaw_aa.data_ptr() if jaw_aa is not None else None, self._ms_dev.data_ptr(), B, self.latent_scale,
                out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), _lib.stream_ptr()))
        self.base._cond_key.key = None
        return out

    def run(self, audio, word, seed, x_init, y=None, styles=None, jaw_aa=None, noise_tape=None, want_sample=False):
        """All inputs are CPU tensors. Returns (rec_pose [B,128,330], rec_trans [B,128,3]) pinned CPU tensors."""
        self.begin(0, audio, word, seed, x_init, y=y, styles=styles, jaw_aa=jaw_aa, noise_tape=noise_tape, want_sample=want_sample)
        return self.wait(0)

    def begin(self, slot, audio, word, seed, x_init, y=None, styles=None, jaw_aa=None, noise_tape=None, want_sample=False):
        """Queue one window batch on staging set `slot` (0 / 1) and return: H2D on a copy stream, compute on the current
        stream, D2H on a second copy stream (st_generate_330_host_begin). Two batches may be in flight, so the inputs of
        batch i + 1 cross PCIe while batch i computes; `wait(slot)` returns the slot's pinned results."""
        B, h = self.B, self.h
        pin = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype).pin_memory()
        if "slots" not in h:
            h["slots"] = [{"pose": h["pose"], "trans": h["trans"], "sample": h["sample"], "style": h["style"], "jaw": h["jaw"]},
                          {"pose": pin(B, 128, 330), "trans": pin(B, 128, 3), "sample": pin(B, 1536, 1, 32),
                           "style": [pin(*t.shape) for t in h["style"]], "jaw": pin(B, 128, 3)}]
        hs = h["slots"][slot]

        def stage(name, t, dtype=torch.float32):
            """A caller's pinned, contiguous tensor of the right dtype is handed to the native call as it is; anything else is
            copied into this object's pinned staging buffer first (slot 0 only: a second batch in flight needs its own memory)."""
            if t.dtype == dtype and t.is_pinned() and t.is_contiguous() and t.numel() == h[name].numel():
                return t
            if slot != 0:
                raise ValueError(f"begin(slot=1) needs pinned, contiguous {dtype} inputs ('{name}' is not)")
            h[name].copy_(t.to(dtype).reshape(h[name].shape))
            return h[name]

        t_audio, t_word = stage("audio", audio), stage("word", word, torch.int32)
        t_seed, t_x = stage("seed", seed), stage("x_init", x_init)
        y = dict(y or {})
        if styles is None and self.base.variant != "beatx":
            sf = y.get("style_feature")
            styles = [sf.get("upper_mask"), sf.get("hands_mask"), sf.get("lower_mask")] if isinstance(sf, dict) else [sf, None, None]
        styles = styles or [None, None, None]
        inp = _lib.StHostInputs()
        inp.audio, inp.word, inp.seed, inp.x_init = t_audio.data_ptr(), t_word.data_ptr(), t_seed.data_ptr(), t_x.data_ptr()
        nb = h["audio"].numel() * 4 + h["word"].numel() * 4 + h["seed"].numel() * 4 + h["x_init"].numel() * 4 + 666 * 4
        for k in range(3):
            if styles[k] is not None and self.base.variant != "beatx":
                s = styles[k]
                hs["style"][k].copy_(s.expand(B, -1) if s.shape[0] == 1 else s)
                inp.style[k] = hs["style"][k].data_ptr()
                nb += hs["style"][k].numel() * 4
            else:
                inp.style[k] = None
        if jaw_aa is not None:
            hs["jaw"].copy_(jaw_aa); inp.jaw_aa = hs["jaw"].data_ptr(); nb += hs["jaw"].numel() * 4
        if noise_tape is not None:
            noise_tape = noise_tape.contiguous()
            inp.noise_tape = noise_tape.data_ptr(); nb += noise_tape.numel() * 4
        m = self.h_ms
        inp.mean, inp.std, inp.trans_mean, inp.trans_std = m["mean"].data_ptr(), m["std"].data_ptr(), m["trans_mean"].data_ptr(), m["trans_std"].data_ptr()
        g = self.wrapper.guidance(y) if self.wrapper is not None else Guidance(_lib.ST_CFG_NONE)
        sched, _ = self.diffusion._native(self.mode, self.eta)
        hs["keep"] = (inp, g, noise_tape, t_audio, t_word, t_seed, t_x)       # alive until wait()
        with torch.cuda.device(self.base.device):
            _lib.check(_lib.lib().st_generate_330_host_begin(
                self.base.handle, sched, C.byref(g.struct(B)), self.vqs[0].handle, self.vqs[1].handle, self.vqs[2].handle,
                C.byref(inp), B, self.latent_scale, hs["pose"].data_ptr(), hs["trans"].data_ptr(),
                hs["sample"].data_ptr() if want_sample else None, slot, _lib.stream_ptr()))
        self.base._cond_key.key = None          # the native call re-encoded the cache from its own staging buffers
        self.h2d_bytes = nb
        self.d2h_bytes = (hs["pose"].numel() + hs["trans"].numel() + (hs["sample"].numel() if want_sample else 0)) * 4

    def wait(self, slot):
        """Block until the batch queued on `slot` is complete; returns its (rec_pose, rec_trans) pinned CPU tensors."""
        hs = self.h["slots"][slot] if "slots" in self.h else None
        with torch.cuda.device(self.base.device):
            _lib.check(_lib.lib().st_generate_330_host_wait(self.base.handle, slot))
        if hs is None:
            raise _lib.StError("wait() before begin()")
        hs.pop("keep", None)
        return hs["pose"], hs["trans"]


class LongClip330:
    """The trainer's window loop over a long clip (diffusion_rvqvae_trainer.py:413-531) as ONE native call
    (st_generate_long_330): R overlapping 128-frame windows, seed hand-off from window to window on the device, the
    latents stitched (32 + 28 per further window), decoded once per body part and assembled into 330-d features.
    All tensors are CUDA tensors; B long clips run side by side."""

    ROUND_L, PRE = 112, 16                 # pose_length - pre_frames * 4, pre_frames * 4
    HOP_AUDIO = (16000 // 30) * 112

    def __init__(self, model, diffusion, vq_upper, vq_hands, vq_lower, use_ddim=True, eta=0.0, latent_scale=5.0, ms=None):
        self.base, self.wrapper = unwrap(model)          # through nn.DataParallel / DDP and the CFG wrappers
        self.diffusion = diffusion
        self.vqs = (vq_upper, vq_hands, vq_lower)
        self.mode = _lib.ST_MODE_DDIM if use_ddim else _lib.ST_MODE_DDPM
        self.eta, self.latent_scale = eta, float(latent_scale)
        self.ms = ms or load_mean_std()

    @classmethod
    def windows(cls, n_frames):
        """roundt of the trainer (trainer:413-415)."""
        return (n_frames - cls.PRE) // cls.ROUND_L

    def run(self, audio_long, word_long, seed0, x_init, y=None, styles=None, jaw_aa=None, noise_tape=None, want_latents=False):
        """audio_long [B,La,2] fp32, word_long [B,Nw] int32, seed0 [B,4,1536], x_init [R,B,1536,1,32] (one start noise per
        window). Returns (rec_pose [B,n,330], rec_trans [B,n,3][, latents [B,n/4,1536]]), n = 4 (32 + 28 (R-1))."""
        dev = self.base.device
        R, B = x_init.shape[0], x_init.shape[1]
        for t in (audio_long, word_long, seed0, x_init):
            if not (t.is_cuda and t.is_contiguous()):
                raise ValueError("LongClip330.run takes contiguous CUDA tensors")
        if audio_long.dtype != torch.float32 or word_long.dtype != torch.int32 or audio_long.shape[0] != B or word_long.shape[0] != B:
            raise ValueError("audio_long must be fp32 [B,La,2] and word_long int32 [B,Nw]")
        y = dict(y or {})
        if styles is None and self.base.variant != "beatx":
            sf = y.get("style_feature")
            styles = [sf.get("upper_mask"), sf.get("hands_mask"), sf.get("lower_mask")] if isinstance(sf, dict) else [sf, None, None]
        styles = styles or [None, None, None]
        sarr = (C.c_void_p * 3)(*[(s.data_ptr() if (s is not None and self.base.variant != "beatx") else None) for s in styles])
        if not hasattr(self, "_ms_dev"):
            m = self.ms
            self._ms_dev = torch.cat([m["mean"], m["std"], m["trans_mean"], m["trans_std"]]).float().to(dev).contiguous()
        Ttot = 32 + 28 * (R - 1)
        pose = torch.empty((B, 4 * Ttot, 330), device=dev)
        trans = torch.empty((B, 4 * Ttot, 3), device=dev)
        lat = torch.empty((B, Ttot, 1536), device=dev) if want_latents else None
        g = self.wrapper.guidance(y) if self.wrapper is not None else Guidance(_lib.ST_CFG_NONE)
        sched, _ = self.diffusion._native(self.mode, self.eta)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().st_generate_long_330(
                self.base.handle, sched, C.byref(g.struct(B)), self.vqs[0].handle, self.vqs[1].handle, self.vqs[2].handle,
                audio_long.data_ptr(), audio_long.shape[1], word_long.data_ptr(), word_long.shape[1], seed0.data_ptr(), sarr,
                x_init.data_ptr(), noise_tape.data_ptr() if noise_tape is not None else None,
                jaw_aa.data_ptr() if jaw_aa is not None else None, self._ms_dev.data_ptr(), B, R, self.latent_scale,
                pose.data_ptr(), trans.data_ptr(), lat.data_ptr() if lat is not None else None, _lib.stream_ptr()))
        self.base._cond_key.key = None
        return (pose, trans, lat) if want_latents else (pose, trans)


class Window623:
    """HumanML3D-623 window (h3d_diffusion_new_trainer.py:548-607): conditioning -> DDIM-50 with the body-part
    CFG wrapper -> x latent_scale -> latent2origin x3 (156 / 360 / 107) -> 623-d scatter. Device tensors in and out;
    every stage is a native call (st_cond_encode, st_sample, st_rvq_decode, st_pose_assemble_623)."""

    def __init__(self, model, diffusion, vq_upper, vq_hands, vq_lower, latent_scale=5.0):
        self.model, self.diffusion = model, diffusion
        self.vqs = (vq_upper, vq_hands, vq_lower)
        self.latent_scale = float(latent_scale)

    def run(self, audio, word, seed, x_init, style_feature, use_ddim=True):
        B = x_init.shape[0]
        y = {"audio": audio, "word": word, "seed": seed, "style_feature": style_feature}
        loop = self.diffusion.ddim_sample_loop if use_ddim else self.diffusion.p_sample_loop
        sample = loop(self.model, (B, 1536, 1, 32), noise=x_init, clip_denoised=False, model_kwargs={"y": y})
        dev = sample.device
        tok = torch.empty((B, 32, 1536), device=dev)
        recs = []
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().st_sample_to_tokens(sample.data_ptr(), B, 32, 1.0, tok.data_ptr(), _lib.stream_ptr()))
            for k, vq in enumerate(self.vqs):
                rec = torch.empty((B, 128, vq.input_width), device=dev)
                _lib.check(_lib.lib().st_rvq_decode(vq.handle, tok.data_ptr() + 4 * 512 * k, 1536, self.latent_scale, B, 32,
                                                   rec.data_ptr(), None, None, _lib.stream_ptr()))
                recs.append(rec)
        return pose_assemble_623(recs[0], recs[1], recs[2]), sample
