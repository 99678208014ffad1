"""Drop-in for models/vq/model.py `RVQVAE`: `vq.latent2origin(x)` (decode side) and `vq.map2latent(x)` (encoder side).

`RVQVAE(args, input_width, nb_code, code_dim, output_emb_width, down_t, stride_t, width, depth,
dilation_growth_rate, activation, norm)` keeps the reference's constructor (diffusion_rvqvae_trainer.py:106-140);
only the configuration the trainers build (512 codes x 512 dims x 6 layers, width 512, down_t 2, depth 3,
growth 3, relu, no norm) is implemented. `latent2origin` returns `(rec, None, None)`: callers consume `[0]`
only (trainer:480-482); commit loss and perplexity are training statistics.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, packer


class RVQVAE:
    def __init__(self, args=None, input_width=263, nb_code=512, code_dim=512, output_emb_width=512, down_t=2, stride_t=2,
                 width=512, depth=3, dilation_growth_rate=3, activation="relu", norm=None, device=None):
        cfg = (nb_code, code_dim, output_emb_width, down_t, stride_t, width, depth, dilation_growth_rate, activation, norm)
        if cfg != (512, 512, 512, 2, 2, 512, 3, 3, "relu", None):
            raise NotImplementedError(f"RVQVAE configuration {cfg} is not the one the trainers build")
        nq = getattr(args, "num_quantizers", 6) if args is not None else 6
        if nq != 6 or (args is not None and getattr(args, "shared_codebook", False)):
            raise NotImplementedError("only 6 independent codebooks (diffusion_rvqvae_trainer.py:89-92)")
        self.input_width = int(input_width)
        self.code_dim = code_dim
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else torch.device("cpu")
        self._h = None

    def load_state_dict(self, state_dict, strict=True):
        if self.device.type != "cuda":
            raise _lib.StError("syntalker_b200.RVQVAE needs a CUDA device: there is no CPU path")
        packed = packer.pack_rvq(state_dict)
        if packed["dec.6.b"].numel() != self.input_width:
            raise RuntimeError(f"checkpoint decodes to {packed['dec.6.b'].numel()} channels, model expects {self.input_width}")
        arr, keep = _lib.tensor_array(packed)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().st_vq_create(arr, len(packed), self.input_width, C.byref(h)))
        self._free()
        self._h = h
        return self

    def _free(self):
        if getattr(self, "_h", None):
            _lib.lib().st_vq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def cuda(self, *a, **k):
        return self

    @property
    def handle(self):
        if self._h is None:
            raise _lib.StError("RVQVAE has no weights: call load_state_dict() first")
        return self._h

    def set_engine(self, name):
        """GEMM engine of THIS decoder ('simt', 'tc', or None = the process default)."""
        _lib.check(_lib.lib().st_vq_set_engine(self.handle, {"simt": _lib.ST_ENGINE_SIMT, "tc": _lib.ST_ENGINE_TC, None: -1}[name]))
        return self

    def latent2origin(self, x, return_indices=False):
        """x [B,T/4,512] fp32 on the GPU (already x vqvae_latent_scale) -> (rec [B,T,D], None, None).
        Like the reference (residual_vq.py:146) the call leaves the final residual in `x` when x is a
        contiguous fp32 CUDA tensor."""
        if x.dim() != 3 or x.shape[-1] != 512:
            raise ValueError(f"latent must be [B,T/4,512], got {tuple(x.shape)}")
        B, T4, _ = x.shape
        inplace = x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        xs = x if inplace else x.to(self.device).float().contiguous()
        rec = torch.empty((B, 4 * T4, self.input_width), device=self.device, dtype=torch.float32)
        idx = torch.empty((B, T4, 6), device=self.device, dtype=torch.int64) if return_indices else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().st_rvq_decode(self.handle, xs.data_ptr(), 512, 1.0, B, T4, rec.data_ptr(),
                                               idx.data_ptr() if idx is not None else None, xs.data_ptr(), _lib.stream_ptr()))
        if return_indices:
            return rec, None, None, idx
        return rec, None, None

    def map2latent(self, x):
        """x [B,T,D] normalised pose features (T a multiple of 4) -> [B,T/4,512] encoder output before quantisation
        (models/vq/model.py:95-100; diffusion_rvqvae_trainer.py:290-294 divides it by vqvae_latent_scale)."""
        if x.dim() != 3 or x.shape[-1] != self.input_width or x.shape[1] % 4:
            raise ValueError(f"pose must be [B,T,{self.input_width}] with T a multiple of 4, got {tuple(x.shape)}")
        B, T, _ = x.shape
        xs = x.to(self.device).float().contiguous()
        lat = torch.empty((B, T // 4, 512), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().st_rvq_encode(self.handle, xs.data_ptr(), B, T, lat.data_ptr(), _lib.stream_ptr()))
        return lat
