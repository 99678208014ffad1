"""Seeded synthetic weights and inputs for the SynTalker sampling hot path.

There is no network for checkpoints or datasets, so benches, tests and golden fixtures all use
random-init weights of the reference architecture and synthetic inputs of the reference shapes.
State-dict key names and shapes are exactly the reference's (probed from `models/denoiser.py:12-106`,
`models/denoiser_h3d.py:12-111`, `models/vq/model.py:7-38`), so a dict made here loads with
`strict=True` into the reference modules (tests/golden/make_golden.py does that) and a real
checkpoint loads into the packer unchanged.

Init follows torch defaults (Linear/Conv1d: U(-1/sqrt(fan_in), +1/sqrt(fan_in)) for weight and bias)
but perturbs LayerNorm/BatchNorm affine terms and running stats and fills the RVQ codebooks with
N(0,1) so that no branch of the arithmetic is trivially the identity.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

VOCAB_SIZE = 11195          # rows of weights/vocab.pkl (models/denoiser.py:68-71)
AUDIO_LEN = 68224           # 16000 // 30 * 128 samples per 128-frame window (diffusion_rvqvae_trainer.py:422)
N_FRAMES = 128
N_TOKENS = 32               # 128 frames / vqvae_squeeze_scale 4
LATENT_C = 1536             # 3 body parts x 512
VARIANTS = ("beatx", "beatx_motionclip", "h3d")
# decoder widths per body part: upper / hands / lower(+trans)  (diffusion_rvqvae_trainer.py:106-140)
PART_DIMS_BEATX = (78, 180, 57)
PART_DIMS_H3D = (156, 360, 107)

# (C_in, C_out, stride, pad of first conv, has conv shortcut)  models/denoiser.py:308-315
WAV_BLOCKS = ((2, 64, 5, 1700, True), (64, 64, 6, 0, True), (64, 64, 1, 7, False),
              (64, 128, 6, 0, True), (128, 128, 1, 7, False), (128, 256, 3, 0, True))


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _linear(sd, g, name, out_f, in_f, bias=True):
    b = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(g, (out_f, in_f), b)
    if bias:
        sd[name + ".bias"] = _uniform(g, (out_f,), b)


def _conv(sd, g, name, out_c, in_c, k):
    b = 1.0 / math.sqrt(in_c * k)
    sd[name + ".weight"] = _uniform(g, (out_c, in_c, k), b)
    sd[name + ".bias"] = _uniform(g, (out_c,), b)


def _bn(sd, g, name, c):
    sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
    sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
    sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
    sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
    sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def _ln(sd, g, name, c):
    sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
    sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)


def sinusoid_table(max_len: int = 5000, d: int = 512) -> torch.Tensor:
    """pe[p,2i]=sin(p*w_i), pe[p,2i+1]=cos(p*w_i), fp32 ops in the reference's order
    (models/denoiser.py:215-222). Shape [max_len, 1, d] like the registered buffer."""
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(1).contiguous()


def mdm_state_dict(variant: str = "beatx", seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init state dict of `MDM` with the reference's keys (SURVEY.md §5 weight contract)."""
    assert variant in VARIANTS, variant
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for i, (cin, cout, _s, _p, ds) in enumerate(WAV_BLOCKS):
        p = f"WavEncoder.feat_extractor.{i}."
        _conv(sd, g, p + "conv1", cout, cin, 15)
        _bn(sd, g, p + "bn1", cout)
        _conv(sd, g, p + "conv2", cout, cout, 15)
        _bn(sd, g, p + "bn2", cout)
        if ds:
            _conv(sd, g, p + "downsample.0", cout, cin, 15)
            _bn(sd, g, p + "downsample.1", cout)
    _linear(sd, g, "text_encoder_body", 256, 300)
    sd["text_pre_encoder_body.weight"] = 0.5 * torch.randn(VOCAB_SIZE, 300, generator=g)
    pe = sinusoid_table()
    sd["sequence_pos_encoder.pe"] = pe
    for i in range(8):
        p = f"mytimmblocks.{i}."
        _ln(sd, g, p + "norm1", 512)
        _linear(sd, g, p + "attn.qkv", 1536, 512, bias=False)
        _linear(sd, g, p + "attn.proj", 512, 512)
        _ln(sd, g, p + "norm2", 512)
        _linear(sd, g, p + "mlp.fc1", 1024, 512)
        _linear(sd, g, p + "mlp.fc2", 512, 1024)
    sd["embed_timestep.sequence_pos_encoder.pe"] = pe
    _linear(sd, g, "embed_timestep.time_embed.0", 512, 512)
    _linear(sd, g, "embed_timestep.time_embed.2", 512, 512)
    _linear(sd, g, "embed_style", 64, 6)                 # defined, never used (denoiser.py:91)
    _linear(sd, g, "embed_text", 512, 4 * LATENT_C)
    _linear(sd, g, "output_process.poseFinal", LATENT_C, 512)
    sd["rel_pos.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    _linear(sd, g, "input_process.poseEmbedding", 512, LATENT_C)
    _linear(sd, g, "input_process2", 512, 1280)
    if variant == "beatx_motionclip":
        _linear(sd, g, "input_process3", 512, 1024)
    elif variant == "h3d":
        _linear(sd, g, "input_process3", 512, 768)
        sd["uncon_text_embeddings"] = 0.5 * torch.randn(1, 256, generator=g)
        sd["uncon_audio_embeddings"] = torch.zeros(1, 256)   # defined, never used
    _linear(sd, g, "mix_audio_text", 256, 512)
    return sd


def rvq_state_dict(out_dim: int, seed: int = 0, with_encoder: bool = True) -> Dict[str, torch.Tensor]:
    """Random-init `RVQVAE` state dict ('net' of the checkpoint, rvq_beatx_train.py:404)."""
    g = torch.Generator().manual_seed(1000 + seed * 7 + out_dim)
    sd: Dict[str, torch.Tensor] = {}

    def resnet(prefix):
        for j in range(3):
            _conv(sd, g, f"{prefix}.model.{j}.conv1", 512, 512, 3)
            _conv(sd, g, f"{prefix}.model.{j}.conv2", 512, 512, 1)

    if with_encoder:
        _conv(sd, g, "encoder.model.0", 512, out_dim, 3)
        for i in (2, 3):
            _conv(sd, g, f"encoder.model.{i}.0", 512, 512, 4)
            resnet(f"encoder.model.{i}.1")
        _conv(sd, g, "encoder.model.4", 512, 512, 3)
    _conv(sd, g, "decoder.model.0", 512, 512, 3)
    for i in (2, 3):
        resnet(f"decoder.model.{i}.0")
        _conv(sd, g, f"decoder.model.{i}.2", 512, 512, 3)
    _conv(sd, g, "decoder.model.4", 512, 512, 3)
    _conv(sd, g, "decoder.model.6", out_dim, 512, 3)
    for q in range(6):
        sd[f"quantizer.layers.{q}.codebook"] = torch.randn(512, 512, generator=g)
    return sd


def make_inputs(B: int, seed: int = 1, variant: str = "beatx") -> Dict[str, torch.Tensor]:
    """Synthetic per-window inputs with the loader's shapes/dtypes."""
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(B, LATENT_C, 1, N_TOKENS, generator=g)
    amp = torch.rand(B, AUDIO_LEN, generator=g)
    onset = (torch.rand(B, AUDIO_LEN, generator=g) < 2e-4).float()
    audio = torch.stack([amp, onset], dim=-1).contiguous()
    word = torch.randint(0, VOCAB_SIZE, (B, N_FRAMES), generator=g, dtype=torch.int64).to(torch.int32)
    seed_lat = torch.randn(B, 4, LATENT_C, generator=g)
    out = {"noise": noise, "audio": audio, "word": word, "seed": seed_lat}
    if variant == "beatx":
        out["style_feature"] = torch.zeros(B, 512)
    elif variant == "beatx_motionclip":
        out["style_feature"] = torch.randn(B, 512, generator=g)
    else:
        out["style_upper"] = torch.randn(B, 256, generator=g)
        out["style_lower"] = torch.randn(B, 256, generator=g)
    return out


def mean_std_beatx(seed: int = 7):
    """Stand-in for mean_std/beatx_2_330_{mean,std}.npy and beatx_2_trans_{mean,std}.npy
    (diffusion_rvqvae_trainer.py:188-196) when the real files are not at hand."""
    g = torch.Generator().manual_seed(seed)
    return {
        "mean": 0.3 * torch.randn(330, generator=g),
        "std": 0.2 + 0.5 * torch.rand(330, generator=g),
        "trans_mean": 0.01 * torch.randn(3, generator=g),
        "trans_std": 0.02 + 0.05 * torch.rand(3, generator=g),
    }
