"""Drop-in for models/denoiser.py `MDM` (and, via denoiser_h3d.py, models/denoiser_h3d.py `MDM`).

Same call boundary as the reference: `MDM(args)`, `load_state_dict(state_dict)`, `model(x, timesteps, y=dict)`
with x [B,1536,1,32] fp32 on the GPU, timesteps [B] int64 (original 0..999 ids), y holding 'audio'
[B,68224,2], 'word' [B,128] int, 'seed' [B,4,1536], optional 'style_feature', 'uncond', 'uncond_audio'
(models/denoiser.py:132-196). The arithmetic runs in libsyntalker_b200.so; without it every call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, packer
from .synth import AUDIO_LEN, LATENT_C, N_FRAMES, N_TOKENS


def _dev_f32(t: torch.Tensor, name: str, device) -> torch.Tensor:
    if t.device != device:
        t = t.to(device, non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class CondKey:
    """Identity of the y tensors last encoded, so a Python-side step loop (the reference's own
    gaussian_diffusion loop calling model(x,t,y) once per step) pays for the conditioning once.

    The key is (data_ptr, _version, shape, dtype, device) of the CALLER's tensors, and `held` keeps strong references to
    exactly those tensors for as long as the key is current: an address can only be handed to a new tensor after the old one
    is freed, so while they are held a matching key really means "the same, unmodified tensors" (without the references the
    allocator recycles the addresses of freed batches at version 0 and a different batch would hit the cache)."""

    def __init__(self):
        self.key = None
        self.held = None

    @staticmethod
    def of(*tensors):
        return tuple((t.data_ptr(), t._version, tuple(t.shape), t.dtype, str(t.device)) if t is not None else None for t in tensors)


_BUFFER_KEYS = ("running_mean", "running_var", "num_batches_tracked", ".pe", "inv_freq")   # buffers of the reference module; the rest are parameters


def _mangle(k: str) -> str:
    return k.replace(".", "/")          # nn.Module refuses '.' in parameter names; state_dict() / named_parameters() hand back the original keys


class MDM(torch.nn.Module):
    """An `nn.Module` like the reference's: it holds the reference state dict (same keys, same shapes) as parameters / buffers on its
    device, so `next(model.parameters()).device` (gaussian_diffusion.py:697), `state_dict()` / `load_state_dict()` (with or without the
    `module.` prefix, utils/other_tools.py:771-790), `named_parameters()`, `.cuda()` / `.to()`, `.eval()` and wrapping in
    `nn.DataParallel` / DDP (train.py:87-94) behave as they do for `models.denoiser.MDM`.  The arithmetic never touches those
    tensors: `load_state_dict` packs them (packer.py) into the native handle, and `forward` runs in libsyntalker_b200.so."""
    variant_default = "beatx"

    def __init__(self, args=None, state_dict=None, device: Optional[torch.device] = None):
        super().__init__()
        self.args = args
        use_mc = bool(getattr(args, "use_motionclip", False)) if args is not None else False
        self.variant = "beatx_motionclip" if (use_mc and self.variant_default == "beatx") else self.variant_default
        self._device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else torch.device("cpu")
        self.use_motionclip = self.variant != "beatx"
        self._h = None
        self._h_device = None
        self._keys = []                # reference state-dict keys in order
        self._cond_key = CondKey()
        self._keep = None
        self._vocab_rows = 0
        self.eval()
        if state_dict is not None:
            self.load_state_dict(state_dict)

    @property
    def device(self):
        return self._device

    # ---- nn.Module surface the trainers touch (train.py:85-94, other_tools.py:771-790) ----
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self._load(packer.strip_module_prefix(state_dict), strict)
        return self

    def _load(self, sd, strict=True):
        if self._device.type != "cuda":
            raise _lib.StError("syntalker_b200.MDM needs a CUDA device: there is no CPU path")
        found = packer.detect_variant(sd)
        if found != self.variant:
            if strict and self.args is not None:
                raise RuntimeError(f"state dict is for variant '{found}' but the model was built as '{self.variant}'")
            self.variant = found
            self.use_motionclip = found != "beatx"
        for k in self._keys:                                   # drop a previous load
            (self._parameters if _mangle(k) in self._parameters else self._buffers).pop(_mangle(k), None)
        self._keys = list(sd.keys())
        for k, v in sd.items():
            t = v.detach().to(self._device)
            if any(k.endswith(b) for b in _BUFFER_KEYS) or not t.is_floating_point():
                self.register_buffer(_mangle(k), t.clone())
            else:
                self.register_parameter(_mangle(k), torch.nn.Parameter(t.clone(), requires_grad=False))
        self._vocab_rows = int(sd["text_pre_encoder_body.weight"].shape[0])
        self._build()

    def _build(self):
        """(Re)create the native handle from the tensors this module holds, on the device they live on."""
        sd = self._ref_state()
        packed = packer.pack_mdm(sd, self.variant)
        arr, keep = _lib.tensor_array(packed)
        h = C.c_void_p()
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().st_model_create(arr, len(packed), _lib.ST_VARIANT[self.variant], C.byref(h)))
        self._free()
        self._h, self._h_device = h, self._device
        self._cond_key = CondKey()

    def _ref_state(self):
        out = {}
        for k in self._keys:
            m = _mangle(k)
            out[k] = self._parameters[m] if m in self._parameters else self._buffers[m]
        return out

    # state_dict(): the reference's keys (this hook also serves wrappers, which add their own prefix: 'module.', 'model.')
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for k, v in self._ref_state().items():
            destination[prefix + k] = v if keep_vars else v.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        sub = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        if not sub:
            if strict:
                missing_keys.append(prefix + "*")
            return
        try:
            self._load(sub, strict)
        except Exception as e:      # nn.Module.load_state_dict reports, it does not raise from the hooks
            error_msgs.append(str(e))

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        for k in self._keys:
            if _mangle(k) in self._parameters:
                yield (prefix + ("." if prefix else "") + k, self._parameters[_mangle(k)])

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        for k in self._keys:
            if _mangle(k) in self._buffers:
                yield (prefix + ("." if prefix else "") + k, self._buffers[_mangle(k)])

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse) if recurse is not True else super()._apply(fn)
        # .cuda() / .to(device): the native handle lives on one device; follow the tensors (dtype changes are not followed: the
        # path computes in its own arithmetic whatever the storage type of the reference tensors)
        if self._keys:
            dev = self._ref_state()[self._keys[0]].device
            if dev.type == "cuda" and dev != self._h_device:
                self._device = dev
                self._build()
        return self

    def _free(self):
        if getattr(self, "_h", None):
            _lib.lib().st_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("syntalker_b200.MDM is the sampling path only (SURVEY.md §8: training is out of scope)")
        return super().train(False)

    @property
    def handle(self):
        if self._h is None:
            raise _lib.StError("MDM has no weights: call load_state_dict() first")
        return self._h

    def set_engine(self, name):
        """GEMM engine of THIS model ('simt' = exact fp32, 'tc' = tcgen05 split-fp16, None = follow _lib.set_engine's process default)."""
        e = {"simt": _lib.ST_ENGINE_SIMT, "tc": _lib.ST_ENGINE_TC, None: -1}[name]
        _lib.check(_lib.lib().st_model_set_engine(self.handle, e))
        return self

    # ---- conditioning ----
    def encode_cond(self, y, styles=None, force: bool = False):
        """st_cond_encode on y['audio'], y['word'], y['seed'] (+ style vectors). Cached on tensor identity."""
        dev = self.device
        audio = _dev_f32(y["audio"], "audio", dev)
        B = audio.shape[0]
        if tuple(audio.shape) != (B, AUDIO_LEN, 2):
            raise ValueError(f"y['audio'] must be [B,{AUDIO_LEN},2], got {tuple(audio.shape)}")
        word = y["word"]
        if tuple(word.shape) != (B, N_FRAMES):
            raise ValueError(f"y['word'] must be [B,{N_FRAMES}], got {tuple(word.shape)}")
        if word.device.type == "cpu" and word.numel() and (int(word.min()) < 0 or int(word.max()) >= self._vocab_rows):
            raise IndexError(f"y['word'] holds ids outside [0,{self._vocab_rows}) (nn.Embedding would raise, denoiser.py:152)")
        word = word.to(dev, non_blocking=True).to(torch.int32).contiguous()
        seed = _dev_f32(y["seed"], "seed", dev)
        if seed.numel() != B * 4 * LATENT_C:
            raise ValueError(f"y['seed'] must be [B,4,{LATENT_C}], got {tuple(seed.shape)}")
        if styles is None:
            sf = y.get("style_feature") if self.variant != "beatx" else None
            styles = [sf, None, None] if not isinstance(sf, dict) else \
                [sf.get("upper_mask"), sf.get("hands_mask"), sf.get("lower_mask")]
        sdim = 512 if self.variant == "beatx_motionclip" else 256
        originals = (y["audio"], y["word"], y["seed"]) + tuple(s if self.variant != "beatx" else None for s in styles)
        key = CondKey.of(*originals) + (B,)
        if not force and key == self._cond_key.key:
            return B
        st = []
        for s in styles:
            if s is None or self.variant == "beatx":
                st.append(None)
                continue
            s = _dev_f32(s, "style", dev)
            if s.shape[0] == 1 and B > 1:
                s = s.expand(B, -1).contiguous()
            if tuple(s.shape) != (B, sdim):
                raise ValueError(f"style vector must be [B,{sdim}], got {tuple(s.shape)}")
            st.append(s)
        c = _lib.StCond()
        c.audio, c.word, c.seed = audio.data_ptr(), word.data_ptr(), seed.data_ptr()
        for k in range(3):
            c.style[k] = st[k].data_ptr() if st[k] is not None else None
        self._keep = (audio, word, seed, st)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().st_cond_encode(self.handle, C.byref(c), B, _lib.stream_ptr()))
        self._cond_key.key = key
        self._cond_key.held = originals
        return B

    # ---- the model call ----
    def forward(self, x, timesteps, y=None, _guidance=None):
        y = dict(y or {})
        B = self.encode_cond(y)
        if tuple(x.shape) != (B, LATENT_C, 1, N_TOKENS):
            raise ValueError(f"x must be [{B},{LATENT_C},1,{N_TOKENS}], got {tuple(x.shape)}")
        xs = _dev_f32(x, "x", self.device)
        if timesteps.numel() != B:
            raise ValueError(f"timesteps must be [B={B}], got {tuple(timesteps.shape)}")
        if timesteps.device.type == "cpu" and B and (int(timesteps.min()) < 0 or int(timesteps.max()) >= 1000):
            raise ValueError("timesteps must lie in [0,1000)")    # device tensors are clamped by the kernel instead: no sync per call
        t = timesteps.to(self.device).to(torch.int64).contiguous()
        g = _guidance
        if g is None:
            g = Guidance(_lib.ST_CFG_NONE, flags=(_lib.ST_FLAG_UNCOND if y.get("uncond", False) else 0) |
                         (_lib.ST_FLAG_UNCOND_AUDIO if y.get("uncond_audio", False) else 0))
        out = torch.empty_like(xs)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().st_denoise(self.handle, xs.data_ptr(), t.data_ptr(), C.byref(g.struct(B)), out.data_ptr(),
                                            B, _lib.stream_ptr()))
        return out


class Guidance:
    """Host description of the CFG wrapper around the model (diffusion/cfg_sampler.py) -> st_guidance."""

    def __init__(self, mode, flags=0, scale=None, scale2=None, audio_scale=1.0, prompt_scale=4.0):
        self.mode, self.flags, self.scale, self.scale2 = mode, flags, scale, scale2
        self.audio_scale, self.prompt_scale = float(audio_scale), float(prompt_scale)
        self._keep = None

    @staticmethod
    def _host(v, B):
        if v is None:
            return None
        v = torch.as_tensor(v, dtype=torch.float32).detach().cpu().reshape(-1)
        if v.numel() == 1:
            v = v.expand(B)
        if v.numel() != B:
            raise ValueError(f"guidance scale must have 1 or B={B} entries, got {v.numel()}")
        return v.contiguous()

    def struct(self, B):
        s, s2 = self._host(self.scale, B), self._host(self.scale2, B)
        self._keep = (s, s2)
        g = _lib.StGuidance()
        g.mode, g.flags = self.mode, self.flags
        g.scale = s.data_ptr() if s is not None else None
        g.scale2 = s2.data_ptr() if s2 is not None else None
        g.audio_scale, g.prompt_scale = self.audio_scale, self.prompt_scale
        return g
