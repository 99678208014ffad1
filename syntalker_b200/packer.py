"""Weight packer: reference state dicts -> the tensors libsyntalker_b200.so takes.

All folding is done in float64 and rounded once to fp32:
  * BatchNorm1d (eval, running stats) folded into the WavEncoder convs  (layer.py:173-184);
    conv weights re-laid out [C_out, tap, C_in] for channels-last implicit GEMM, rows padded to 4 floats.
  * Embedding(11195,300) -> Linear(300,256) folded into one [11195,256] table (denoiser.py:152-153).
  * input_process (poseEmbedding), input_process2 and input_process3 folded into
        tokens = W_x . x_t + vt[t] + W_cm . pool4([audio|word]) + W_seed . seed + bias_all (+ W_style . style)
    (denoiser.py:147-174): W_in2 = [Wa|Wb|Wc], P = W_in3[:, :512] (or I):
        W_x = P Wb W_pe,  vt[t] = P Wa e_t(t),  W_seed = P Wa W_embed_text,  W_cm = P Wc W_mix,
        bias_all = P (Wa b_seed + Wc b_mix + b_in2 + Wb b_pe) + b_in3,  W_style = W_in3[:, 512:].
  * the timestep MLP (denoiser.py:231-245) is tabulated for t = 0..999.
  * rotary cos/sin tables computed with the reference's fp32 torch ops (denoiser.py:324-334).
RVQ-VAE: codebooks + |c|^2 (quantizer.py:72-74) and the decoder convs in [C_out, tap, C_in] layout.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .synth import WAV_BLOCKS


def _f64(t):
    return t.detach().cpu().double().numpy()


def _t32(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32)))


def _conv_layout(w, scale=None):
    """torch Conv1d weight [C_out, C_in, k] (float64 ndarray) -> [C_out, k*C_in] tap-major, rows padded to x4."""
    if scale is not None:
        w = w * scale[:, None, None]
    co, ci, k = w.shape
    flat = np.transpose(w, (0, 2, 1)).reshape(co, k * ci)
    ld = (k * ci + 3) // 4 * 4
    out = np.zeros((co, ld), dtype=np.float64)
    out[:, :k * ci] = flat
    return out


def strip_module_prefix(sd):
    """DataParallel checkpoints carry a 'module.' prefix (utils/other_tools.py:771-790)."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def detect_variant(sd) -> str:
    if "uncon_text_embeddings" in sd:
        return "h3d"
    return "beatx_motionclip" if "input_process3.weight" in sd else "beatx"


def pack_mdm(sd: Dict[str, torch.Tensor], variant: str | None = None) -> Dict[str, torch.Tensor]:
    sd = strip_module_prefix(sd)
    variant = variant or detect_variant(sd)
    out: Dict[str, torch.Tensor] = {}
    # ---- WavEncoder: fold BN, channels-last layout ----
    for i, (_cin, _cout, _s, _p, ds) in enumerate(WAV_BLOCKS):
        p = f"WavEncoder.feat_extractor.{i}."
        for conv, bn, name in (("conv1", "bn1", "conv1"), ("conv2", "bn2", "conv2"), ("downsample.0", "downsample.1", "ds")):
            if name == "ds" and not ds:
                continue
            w, b = _f64(sd[p + conv + ".weight"]), _f64(sd[p + conv + ".bias"])
            g, beta = _f64(sd[p + bn + ".weight"]), _f64(sd[p + bn + ".bias"])
            mu, var = _f64(sd[p + bn + ".running_mean"]), _f64(sd[p + bn + ".running_var"])
            s = g / np.sqrt(var + 1e-5)
            out[f"wav.{i}.{name}.w"] = _t32(_conv_layout(w, s))
            out[f"wav.{i}.{name}.b"] = _t32((b - mu) * s + beta)
    # ---- word path ----
    E = _f64(sd["text_pre_encoder_body.weight"])
    out["word_table"] = _t32(E @ _f64(sd["text_encoder_body.weight"]).T + _f64(sd["text_encoder_body.bias"]))
    # ---- input folding ----
    W2, b2 = _f64(sd["input_process2.weight"]), _f64(sd["input_process2.bias"])
    Wa, Wb, Wc = W2[:, :512], W2[:, 512:1024], W2[:, 1024:1280]
    Wpe, bpe = _f64(sd["input_process.poseEmbedding.weight"]), _f64(sd["input_process.poseEmbedding.bias"])
    Wseed, bseed = _f64(sd["embed_text.weight"]), _f64(sd["embed_text.bias"])
    Wmix, bmix = _f64(sd["mix_audio_text.weight"]), _f64(sd["mix_audio_text.bias"])
    if variant == "beatx":
        P, b3 = np.eye(512), np.zeros(512)
    else:
        W3, b3 = _f64(sd["input_process3.weight"]), _f64(sd["input_process3.bias"])
        P = W3[:, :512]
        out["w_style"] = _t32(W3[:, 512:])
        if variant == "h3d":
            out["null_sv"] = _t32(W3[:, 512:] @ _f64(sd["uncon_text_embeddings"])[0])
    out["w_x"] = _t32(P @ Wb @ Wpe)
    out["w_seed"] = _t32(P @ Wa @ Wseed)
    out["w_cm"] = _t32(P @ Wc @ Wmix)
    out["bias_all"] = _t32(P @ (Wa @ bseed + Wc @ bmix + b2 + Wb @ bpe) + b3)
    # ---- timestep MLP table ----
    pe = _f64(sd["sequence_pos_encoder.pe"])[:1000, 0, :]
    h = pe @ _f64(sd["embed_timestep.time_embed.0.weight"]).T + _f64(sd["embed_timestep.time_embed.0.bias"])
    h = h / (1.0 + np.exp(-h))
    et = h @ _f64(sd["embed_timestep.time_embed.2.weight"]).T + _f64(sd["embed_timestep.time_embed.2.bias"])
    out["vt_table"] = _t32(et @ (P @ Wa).T)
    # ---- rotary tables, fp32 like the reference ----
    inv_freq = sd["rel_pos.inv_freq"].detach().cpu().float()
    freqs = torch.einsum("i,j->ij", torch.arange(32).type_as(inv_freq), inv_freq)
    out["rope_cos"] = freqs.cos().contiguous()
    out["rope_sin"] = freqs.sin().contiguous()
    # ---- transformer blocks ----
    for i in range(8):
        p = f"mytimmblocks.{i}."
        for src, dst in (("norm1.weight", "ln1.g"), ("norm1.bias", "ln1.b"), ("attn.qkv.weight", "qkv.w"),
                         ("attn.proj.weight", "proj.w"), ("attn.proj.bias", "proj.b"), ("norm2.weight", "ln2.g"),
                         ("norm2.bias", "ln2.b"), ("mlp.fc1.weight", "fc1.w"), ("mlp.fc1.bias", "fc1.b"),
                         ("mlp.fc2.weight", "fc2.w"), ("mlp.fc2.bias", "fc2.b")):
            out[f"blk.{i}.{dst}"] = sd[p + src].detach().cpu().float().contiguous()
    # LayerNorm folded into the Linear that consumes it (tcgen05 engine):
    #   W (gamma*(x-mu)/sigma + beta) + b = (1/sigma) (W' x - mu s) + c,   W' = W*gamma,  s = W' 1,  c = W beta + b
    # so the GEMM runs on the raw residual stream and the epilogue applies the per-row (mu, 1/sigma).
    for i in range(8):
        p = f"mytimmblocks.{i}."
        for norm, lin, dst, has_b in (("norm1", "attn.qkv", "qkv", False), ("norm2", "mlp.fc1", "fc1", True)):
            g, be = _f64(sd[p + norm + ".weight"]), _f64(sd[p + norm + ".bias"])
            Wl = _f64(sd[p + lin + ".weight"])
            Wg = Wl * g[None, :]
            out[f"blk.{i}.{dst}.wg"] = _t32(Wg)
            out[f"blk.{i}.{dst}.s"] = _t32(Wg.sum(axis=1))
            out[f"blk.{i}.{dst}.c"] = _t32(Wl @ be + (_f64(sd[p + lin + ".bias"]) if has_b else 0.0))
        # rows of the qkv layer regrouped as [head h][half j][q 64 | k 64 | v 64] (row = which*512 + h*128 + j*64 + d,
        # transformer.py:85-86): GEMM tile n = 2h + j of the fused qkv + attention kernel is then 192 consecutive rows
        perm = np.concatenate([w * 512 + h * 128 + j * 64 + np.arange(64) for h in range(4) for j in range(2) for w in range(3)])
        for suf in ("wg", "s", "c"):
            out[f"blk.{i}.qkv.{suf}p"] = out[f"blk.{i}.qkv.{suf}"][torch.from_numpy(perm)].contiguous()
    out["out.w"] = sd["output_process.poseFinal.weight"].detach().cpu().float().contiguous()
    out["out.b"] = sd["output_process.poseFinal.bias"].detach().cpu().float().contiguous()
    # ---- deterministic DDIM keeps the loop in token space ("z recursion") ----
    # x_{k-1} = a_k x0_hat + b_k x_k is linear (gaussian_diffusion.py:772-790 with sigma = 0), x0_hat = W_out h + b_out per
    # evaluation (denoiser.py:287-301) and the next step only needs W_x x_{k-1}:
    #   W_x x_{k-1} = a_k (W_x W_out) h_mix + a_k W_x b_out + b_k (W_x x_k)
    # so between steps one 512 x 512 GEMM replaces the 512 -> 1536 output GEMM, the state update and the 1536 -> 512 input GEMM.
    Wx64 = P @ Wb @ Wpe
    out["w_xo"] = _t32(Wx64 @ _f64(sd["output_process.poseFinal.weight"]))
    out["c_xo"] = _t32(Wx64 @ _f64(sd["output_process.poseFinal.bias"]))
    # ... and the last block's fc2 folds into the same GEMM (x + fc2(g) is linear in (x, g), transformer.py:197-198):
    #   W_xo (x_mid + W_fc2 g + b_fc2) = [W_xo | W_xo W_fc2] [x_mid ; g] + W_xo b_fc2           (K = 512 + 1024)
    Wxo64 = Wx64 @ _f64(sd["output_process.poseFinal.weight"])
    Wf2, bf2 = _f64(sd["mytimmblocks.7.mlp.fc2.weight"]), _f64(sd["mytimmblocks.7.mlp.fc2.bias"])
    out["w_xo2"] = _t32(np.concatenate([Wxo64, Wxo64 @ Wf2], axis=1))
    out["c_xo2"] = _t32(Wxo64 @ bf2)
    return out


def pack_rvq(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """'net' state dict of RVQVAE -> quantiser, decoder and (when the checkpoint has them) encoder tensors."""
    sd = strip_module_prefix(sd)
    out: Dict[str, torch.Tensor] = {}
    if "encoder.model.0.weight" in sd:                       # encdec.py:5-34, channels-last tap-major like the decoder
        enc = [("0", "0"), ("4", "4")]
        for i in (2, 3):
            enc.append((f"{i}.0", f"{i}.0"))
            for j in range(3):
                for c in ("conv1", "conv2"):
                    enc.append((f"{i}.1.{j}.{c}", f"{i}.1.model.{j}.{c}"))
        for dst, src in enc:
            key = "encoder.model." + src
            out[f"enc.{dst}.w"] = _t32(_conv_layout(_f64(sd[key + ".weight"])))
            out[f"enc.{dst}.b"] = sd[key + ".bias"].detach().cpu().float().contiguous()
    for q in range(6):
        cb = sd[f"quantizer.layers.{q}.codebook"].detach().cpu().float().contiguous()
        out[f"cb.{q}"] = cb
        out[f"cnorm.{q}"] = torch.sum(cb.t() ** 2, dim=0).contiguous()        # quantizer.py:74, same op
    pairs = [("0", "0"), ("4", "4"), ("6", "6")]
    for i in (2, 3):
        pairs.append((f"{i}.2", f"{i}.2"))
        for j in range(3):
            for c in ("conv1", "conv2"):
                pairs.append((f"{i}.0.{j}.{c}", f"{i}.0.model.{j}.{c}"))
    for dst, src in pairs:
        key = "decoder.model." + src
        out[f"dec.{dst}.w"] = _t32(_conv_layout(_f64(sd[key + ".weight"])))
        out[f"dec.{dst}.b"] = sd[key + ".bias"].detach().cpu().float().contiguous()
    # nearest x2 upsample followed by the k=3 conv (encdec.py:59-61) == two 2-tap convs on the low-res rows:
    #   out[2u] = W0 in[u-1] + (W1+W2) in[u];   out[2u+1] = (W0+W1) in[u] + W2 in[u+1]       (tap-major [C_out, 2*C_in])
    for i in (2, 3):
        w = _f64(sd[f"decoder.model.{i}.2.weight"])                       # [C_out, C_in, 3]
        b = sd[f"decoder.model.{i}.2.bias"].detach().cpu().float().contiguous()
        even = np.concatenate([w[:, :, 0], w[:, :, 1] + w[:, :, 2]], axis=1)
        odd = np.concatenate([w[:, :, 0] + w[:, :, 1], w[:, :, 2]], axis=1)
        out[f"dec.{i}.2.even.w"], out[f"dec.{i}.2.even.b"] = _t32(even), b
        out[f"dec.{i}.2.odd.w"], out[f"dec.{i}.2.odd.b"] = _t32(odd), b.clone()
    return out
