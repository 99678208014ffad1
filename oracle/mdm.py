"""Oracle: the MDM denoiser forward and the CFG wrappers, as the reference computes them.

Follows models/denoiser.py:132-196 (BEAT-X), models/denoiser_h3d.py:148-225 (h3d),
models/timm_transformer/transformer.py:83-104,145-151,195-198, models/utils/layer.py:173-184,
diffusion/cfg_sampler.py. Written as pure functions over a state dict `W` (reference key names).
Deliberately keeps the reference's cost structure: the audio/word encoders run on EVERY call.
"""
import torch
import torch.nn.functional as F

WAV_CFG = ((5, 1700, True), (6, 0, True), (1, 7, False), (6, 0, True), (1, 7, False), (3, 0, True))


def _bn(W, p, h):
    return F.batch_norm(h, W[p + ".running_mean"], W[p + ".running_var"], W[p + ".weight"], W[p + ".bias"],
                        False, 0.0, 1e-5)


def wav_encoder(W, audio):
    """models/denoiser.py:304-322 + layer.py:173-184. audio [B,L,2] -> [B,128,256]."""
    h = audio.transpose(1, 2)
    for i, (stride, pad, ds) in enumerate(WAV_CFG):
        p = f"WavEncoder.feat_extractor.{i}."
        sc = h
        h = F.conv1d(h, W[p + "conv1.weight"], W[p + "conv1.bias"], stride=stride, padding=pad)
        h = F.leaky_relu(_bn(W, p + "bn1", h), 0.01)
        h = F.conv1d(h, W[p + "conv2.weight"], W[p + "conv2.bias"], stride=1, padding=7)
        h = _bn(W, p + "bn2", h)
        if ds:
            sc = F.conv1d(sc, W[p + "downsample.0.weight"], W[p + "downsample.0.bias"], stride=stride, padding=pad)
            sc = _bn(W, p + "downsample.1", sc)
        h = F.leaky_relu(h + sc, 0.01)
    return h.transpose(1, 2)


def timestep_embed(W, t):
    """models/denoiser.py:244-245: pe[t] -> Linear -> SiLU -> Linear; returns [1,B,512]."""
    pe = W["sequence_pos_encoder.pe"][t]                     # [B,1,512]
    h = F.linear(pe, W["embed_timestep.time_embed.0.weight"], W["embed_timestep.time_embed.0.bias"])
    h = F.silu(h)
    h = F.linear(h, W["embed_timestep.time_embed.2.weight"], W["embed_timestep.time_embed.2.bias"])
    return h.permute(1, 0, 2)


def rope(xseq):
    """models/denoiser.py:178-186,324-343. xseq [T,B,512] -> [B,T,512] rotated in 8 groups of 64."""
    T, B, _ = xseq.shape
    x = xseq.permute(1, 0, 2).reshape(B, T, 8, 64).permute(0, 2, 1, 3).reshape(B * 8, T, 64)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    pos = torch.arange(T).type_as(inv_freq)
    freqs = torch.einsum("i,j->ij", pos, inv_freq)
    freqs = torch.cat((freqs, freqs), dim=-1)                # [T,64]
    x1, x2 = x[..., :32], x[..., 32:]
    rot = torch.cat((-x2, x1), dim=-1)
    x = x * freqs.cos() + rot * freqs.sin()
    return x.reshape(B, 8, T, 64).permute(0, 2, 1, 3).reshape(B, T, 512)


def block(W, p, x):
    """transformer.py:195-198 with Attention.forward:83-104 and Mlp.forward:145-151 (eval mode)."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), W[p + "norm1.weight"], W[p + "norm1.bias"], 1e-5)
    qkv = F.linear(h, W[p + "attn.qkv.weight"]).reshape(B, N, 3, 4, C // 4).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    a = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0)
    a = a.transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(a, W[p + "attn.proj.weight"], W[p + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), W[p + "norm2.weight"], W[p + "norm2.bias"], 1e-5)
    h = F.gelu(F.linear(h, W[p + "mlp.fc1.weight"], W[p + "mlp.fc1.bias"]))
    return x + F.linear(h, W[p + "mlp.fc2.weight"], W[p + "mlp.fc2.bias"])


def mdm_forward(W, x, t, y, variant="beatx", taps=None):
    """One denoiser evaluation. x [B,1536,1,T], t [B] int64 (ORIGINAL timestep ids), y dict.
    BEAT-X: models/denoiser.py:132-196; h3d: models/denoiser_h3d.py:148-225.
    `taps` (optional dict) receives intermediate tensors for stage-by-stage parity tests."""
    B, C, _, T = x.shape
    emb_t = timestep_embed(W, t)                                              # [1,B,512]
    force_mask = bool(y.get("uncond", False))
    emb_seed = F.linear(y["seed"].reshape(B, -1), W["embed_text.weight"], W["embed_text.bias"])
    audio, word = y["audio"], y["word"]
    if variant == "h3d" and y.get("uncond_audio", False):                     # denoiser_h3d.py:161,173-179
        audio = torch.zeros_like(audio)
        word = torch.zeros_like(word)
    a = wav_encoder(W, audio).permute(1, 0, 2)                                # [128,B,256]
    w = F.embedding(word.long(), W["text_pre_encoder_body.weight"])
    w = F.linear(w, W["text_encoder_body.weight"], W["text_encoder_body.bias"]).permute(1, 0, 2)
    at = F.linear(torch.cat([a, w], dim=2), W["mix_audio_text.weight"], W["mix_audio_text.bias"])
    at = F.avg_pool1d(at.permute(1, 2, 0), 4).permute(2, 0, 1)                # [T,B,256]
    x_ = F.linear(x.permute(3, 0, 1, 2).reshape(T, B, C),
                  W["input_process.poseEmbedding.weight"], W["input_process.poseEmbedding.bias"])
    xseq = torch.cat((x_, at), dim=2)
    st = (emb_seed + emb_t).repeat(T, 1, 1)
    xseq = F.linear(torch.cat((st, xseq), dim=2), W["input_process2.weight"], W["input_process2.bias"])
    if variant in ("beatx_motionclip", "h3d"):
        style = y["style_feature"]
        if force_mask:
            if variant == "h3d":
                style = W["uncon_text_embeddings"].repeat(B, 1)               # denoiser_h3d.py:116-125
            else:
                style = torch.zeros_like(style)                               # denoiser.py:109-112
        elif style.shape[0] != B:
            style = style.expand(B, -1)
        xseq = torch.cat((xseq, style.unsqueeze(0).repeat(T, 1, 1)), dim=2)
        xseq = F.linear(xseq, W["input_process3.weight"], W["input_process3.bias"])
    xseq = rope(xseq)
    if taps is not None:
        taps["cond_at"] = at.permute(1, 0, 2).contiguous()
        taps["tokens_in"] = xseq.clone()
    for i in range(8):
        xseq = block(W, f"mytimmblocks.{i}.", xseq)
        if taps is not None:
            taps[f"block{i}"] = xseq.clone()
    out = F.linear(xseq.permute(1, 0, 2), W["output_process.poseFinal.weight"], W["output_process.poseFinal.bias"])
    return out.reshape(T, B, C, 1).permute(1, 2, 3, 0)


# ---- classifier-free-guidance wrappers (diffusion/cfg_sampler.py) ------------------------------------

def cfg_text(model_fn, x, t, y, eval_metric=False):
    """ClassifierFreeSampleModel.forward, cfg_sampler.py:17-28 (eval=True returns the unconditional output, :25-26)."""
    yc = dict(y); yc["uncond_audio"] = True
    out = model_fn(x, t, yc)
    yu = dict(yc); yu["uncond"] = True
    out_u = model_fn(x, t, yu)
    if eval_metric:
        return out_u
    return out_u + y["scale"].view(-1, 1, 1, 1) * (out - out_u)


def cfg_two(model_fn, x, t, y):
    """TwoClassifierFreeSampleModel.forward, cfg_sampler.py:38-54."""
    yuu = dict(y); yuu["uncond_audio"] = True; yuu["uncond"] = True
    o_uu = model_fn(x, t, yuu)
    yua = dict(y); yua["uncond_audio"] = True
    o_ua = model_fn(x, t, yua)
    yut = dict(y); yut["uncond"] = True
    o_ut = model_fn(x, t, yut)
    return o_uu + y["scale_audio"].view(-1, 1, 1, 1) * (o_ut - o_uu) + y["scale_prompt"].view(-1, 1, 1, 1) * (o_ua - o_uu)


PART_SLICES = {"upper_mask": slice(0, 512), "hands_mask": slice(512, 1024), "lower_mask": slice(1024, 1536)}


def cfg_bodypart(model_fn, x, t, y, audio_scale=1.0, prompt_scale=4.0):
    """TwoClassifierFreeSampleModel_Bodypart.forward, cfg_sampler.py:67-117: 3 parts x 3 evaluations."""
    out = torch.zeros_like(x)
    for key, value in y["style_feature"].items():
        yp = dict(y)
        if value is None:
            yp["style_feature"] = torch.zeros(1, 256)
            sa, sp = audio_scale, 0.0
        else:
            yp["style_feature"] = value
            sa, sp = (1.0, prompt_scale) if key in "upper_mask" else (0.0, prompt_scale)
        yp["scale_audio"] = torch.ones(1) * sa
        yp["scale_prompt"] = torch.ones(1) * sp
        o = cfg_two(model_fn, x, t, yp)
        sl = PART_SLICES[key]
        out[:, sl] = out[:, sl] + o[:, sl]
    return out


def cfg_bodypart1(model_fn, x, t, y, eval_metric=False):
    """ClassifierFreeSampleModel_Bodypart.forward, cfg_sampler.py:133-167: per prompted part one evaluation with that prompt and the
    audio masked, one evaluation with the null prompt (audio kept); out_uncond + scale * (out - out_uncond)."""
    yu = dict(y); yu["uncond"] = True
    if eval_metric:                                                       # :143-146
        yu["style_feature"] = y["style_feature"]["lower_mask"]
        return model_fn(x, t, yu)
    out = torch.zeros_like(x)
    covered = torch.zeros(x.shape[1], dtype=torch.bool)
    for key, value in y["style_feature"].items():
        if value is None:
            continue
        yp = dict(y); yp["style_feature"] = value; yp["uncond_audio"] = True
        o = model_fn(x, t, yp)
        sl = PART_SLICES[key]
        out[:, sl] = out[:, sl] + o[:, sl]
        covered[sl] = True
    yu["style_feature"] = torch.zeros(1, 256)
    out_u = model_fn(x, t, yu)
    out[:, ~covered] = out[:, ~covered] + out_u[:, ~covered]
    return out_u + y["scale"].view(-1, 1, 1, 1) * (out - out_u)
