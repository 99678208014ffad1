"""Oracle: RVQVAE.latent2origin = 6-stage residual nearest-code search + conv decoder; RVQVAE.map2latent = conv encoder.

Follows models/vq/model.py:102-109, residual_vq.py:99-169 (eval: no dropout), quantizer.py:67-84,132-158
(distance, argmax(-d) first-index ties, straight-through x + (x_d - x)), encdec.py:37-68, resnet.py:12-84.
"""
import torch
import torch.nn.functional as F


def residual_quantize(W, x):
    """x [B,512,T] -> (quantized [B,512,T], idx [B,T,6] int64). The reference subtracts in place on the
    caller's tensor (residual_vq.py:146); the oracle works on a copy and returns the final residual too."""
    B, C, T = x.shape
    residual = x.clone()
    out = torch.zeros_like(x)
    idxs = []
    for q in range(6):
        cb = W[f"quantizer.layers.{q}.codebook"]
        r = residual.permute(0, 2, 1).reshape(B * T, C)                 # 'n c t -> (n t) c'
        k_w = cb.t()
        dist = torch.sum(r ** 2, dim=-1, keepdim=True) - 2 * torch.matmul(r, k_w) + torch.sum(k_w ** 2, dim=0, keepdim=True)
        idx = (-dist).argmax(dim=-1)
        x_d = F.embedding(idx, cb)
        x_d = r + (x_d - r)                                             # quantizer.py:150
        x_d = x_d.view(B, T, C).permute(0, 2, 1).contiguous()
        residual = residual - x_d
        out = out + x_d
        idxs.append(idx.view(B, T))
    return out, torch.stack(idxs, dim=-1), residual


def decoder(W, h):
    """encdec.py:51-68: [B,512,T] -> [B,4T,D]."""
    c = lambda name, h, dil=1, pad=1: F.conv1d(h, W[name + ".weight"], W[name + ".bias"], padding=pad, dilation=dil)
    h = F.relu(c("decoder.model.0", h))
    for i in (2, 3):
        for j, dil in enumerate((9, 3, 1)):                             # reverse_dilation=True (resnet.py:78-80)
            p = f"decoder.model.{i}.0.model.{j}"
            r = c(p + ".conv1", F.relu(h), dil, dil)
            r = c(p + ".conv2", F.relu(r), 1, 0)
            h = r + h
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        h = c(f"decoder.model.{i}.2", h)
    h = F.relu(c("decoder.model.4", h))
    h = c("decoder.model.6", h)
    return h.permute(0, 2, 1)


def encoder(W, x):
    """encdec.py:5-34: [B,D,T] -> [B,512,T/4]. Conv k3 + ReLU; 2 x (Conv k4 stride 2 pad 1, 3 pre-activation res blocks with
    dilations 9,3,1 -- Resnet1D's reverse_dilation defaults to True, resnet.py:72-80); Conv k3."""
    c = lambda name, h, dil=1, pad=1, stride=1: F.conv1d(h, W[name + ".weight"], W[name + ".bias"], stride=stride, padding=pad, dilation=dil)
    h = F.relu(c("encoder.model.0", x))
    for i in (2, 3):
        h = c(f"encoder.model.{i}.0", h, 1, 1, 2)
        for j, dil in enumerate((9, 3, 1)):
            p = f"encoder.model.{i}.1.model.{j}"
            r = c(p + ".conv1", F.relu(h), dil, dil)
            r = c(p + ".conv2", F.relu(r), 1, 0)
            h = r + h
    return c("encoder.model.4", h)


def map2latent(W, pose):
    """models/vq/model.py:95-100. pose [B,T,D] (normalised features) -> latent [B,T/4,512] (before quantisation)."""
    return encoder(W, pose.permute(0, 2, 1)).permute(0, 2, 1)


def latent2origin(W, lat):
    """models/vq/model.py:102-109. lat [B,T,512] (already x vqvae_latent_scale) -> rec [B,4T,D], idx [B,T,6]."""
    xq, idx, _ = residual_quantize(W, lat.permute(0, 2, 1))
    return decoder(W, xq), idx
