"""TEST INFRASTRUCTURE: a torch-CPU walk through the *packed* dataflow the CUDA library executes
(syntalker_b200/packer.py tensors), used only to validate the packer's algebra on a box
without a GPU. It is never imported by the product."""
import torch
import torch.nn.functional as F

WAV = ((2, 64, 5, 1700, True), (64, 64, 6, 0, True), (64, 64, 1, 7, False), (64, 128, 6, 0, True),
       (128, 128, 1, 7, False), (128, 256, 3, 0, True))


def _conv(P, name, h, cin, k, stride=1, pad=0, dil=1):
    w = P[name + ".w"][:, :k * cin].reshape(-1, k, cin).permute(0, 2, 1).contiguous()
    return F.conv1d(h, w, P[name + ".b"], stride=stride, padding=pad, dilation=dil)


def wav(P, audio):
    h = audio.transpose(1, 2)
    for i, (cin, cout, s, p, ds) in enumerate(WAV):
        h1 = F.leaky_relu(_conv(P, f"wav.{i}.conv1", h, cin, 15, s, p), 0.01)
        sc = _conv(P, f"wav.{i}.ds", h, cin, 15, s, p) if ds else h
        h = F.leaky_relu(_conv(P, f"wav.{i}.conv2", h1, cout, 15, 1, 7) + sc, 0.01)
    return h.transpose(1, 2)                                   # [B,128,256]


def cond(P, audio, word, seed, null_audio=False):
    B = audio.shape[0]
    if null_audio:
        audio, word = torch.zeros_like(audio), torch.zeros_like(word)
    at = torch.cat([wav(P, audio), P["word_table"][word.long()]], dim=-1)          # [B,128,512]
    pooled = at.reshape(B, 32, 4, 512).mean(dim=2)
    cst = pooled @ P["w_cm"].t() + P["bias_all"]
    g2 = seed.reshape(B, -1) @ P["w_seed"].t()
    return cst, g2


def tokens(P, z, t, cst, g2, sv=None):
    """z = W_x x_t [B,32,512] -> rotary-embedded tokens (the conditioning terms are step-invariant)."""
    B = z.shape[0]
    z = z + P["vt_table"][t][:, None, :] + cst + g2[:, None, :]
    if sv is not None:
        z = z + sv[:, None, :]
    zz = z.reshape(B, 32, 8, 64)
    cos, sin = P["rope_cos"][None, :, None, :], P["rope_sin"][None, :, None, :]
    x1, x2 = zz[..., :32], zz[..., 32:]
    return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1).reshape(B, 32, 512)


def blocks(P, h, split_last=False):
    """split_last: stop in front of the last block's fc2 and return (residual stream, GELU output) instead."""
    B = h.shape[0]
    for i in range(8):
        p = f"blk.{i}."
        a = F.layer_norm(h, (512,), P[p + "ln1.g"], P[p + "ln1.b"], 1e-5)
        qkv = (a @ P[p + "qkv.w"].t()).reshape(B, 32, 3, 4, 128).permute(2, 0, 3, 1, 4)
        att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * 128 ** -0.5, dim=-1) @ qkv[2]
        h = h + att.transpose(1, 2).reshape(B, 32, 512) @ P[p + "proj.w"].t() + P[p + "proj.b"]
        a = F.layer_norm(h, (512,), P[p + "ln2.g"], P[p + "ln2.b"], 1e-5)
        g = F.gelu(a @ P[p + "fc1.w"].t() + P[p + "fc1.b"])
        if split_last and i == 7:
            return h, g
        h = h + g @ P[p + "fc2.w"].t() + P[p + "fc2.b"]
    return h


def trunk(P, x, t, cst, g2, sv=None):
    """x [B,1536,1,32] -> model output [B,1536,1,32] for one evaluation."""
    xs = x[:, :, 0, :].permute(0, 2, 1)                                            # [B,32,1536]
    h = blocks(P, tokens(P, xs @ P["w_x"].t(), t, cst, g2, sv))
    o = h @ P["out.w"].t() + P["out.b"]                                            # [B,32,1536]
    return o.permute(0, 2, 1).unsqueeze(2)


def ddim_z_loop(P, coef, t_model, x_init, evals, mix):
    """Deterministic DDIM as the library runs it on the tcgen05 engine: the loop carries z = W_x x_k, not x_k.
    coef [S,5] = schedule.ddim_coefs rows {a, b, c1, c2, sigma = 0}; evals = [(cst, g2, sv), ...] one per planned evaluation in
    the order of step_update; mix(list of per-evaluation tensors) = the guidance combination (linear, cfg_sampler.py:28,54)."""
    S, B = len(t_model), x_init.shape[0]
    z = x_init[:, :, 0, :].permute(0, 2, 1) @ P["w_x"].t()
    for k in range(S - 1, -1, -1):
        t = torch.full((B,), int(t_model[k]), dtype=torch.int64)
        if k == 0:                                                                  # alpha_bar_prev = 1: x <- x0_hat
            h_mix = mix([blocks(P, tokens(P, z, t, *e)) for e in evals])
            return (h_mix @ P["out.w"].t() + P["out.b"]).permute(0, 2, 1).unsqueeze(2)
        # every other step ends in ONE GEMM over [residual stream | GELU output] of the last block (K = 512 + 1024)
        p_mix = mix([torch.cat(blocks(P, tokens(P, z, t, *e), split_last=True), dim=-1) @ P["w_xo2"].t() + P["c_xo2"] for e in evals])
        a, b, c1, c2 = (float(v) for v in coef[k][:4])
        alpha, beta = c1 - c2 / b, c2 * a / b                                       # x_{k-1} = alpha x0_hat + beta x_k
        z = beta * z + alpha * (p_mix + P["c_xo"])


def rvq_decoder(P, xq, out_dim):
    """xq [B,512,T] -> [B,4T,D] through the packed decoder convs."""
    c3 = lambda n, h, dil=1: _conv(P, "dec." + n, h, 512, 3, 1, dil, dil)
    h = F.relu(c3("0", xq))
    for i in (2, 3):
        for j, dil in enumerate((9, 3, 1)):
            r = c3(f"{i}.0.{j}.conv1", F.relu(h), dil)
            h = _conv(P, f"dec.{i}.0.{j}.conv2", F.relu(r), 512, 1) + h
        h = c3(f"{i}.2", F.interpolate(h, scale_factor=2, mode="nearest"))
    h = F.relu(c3("4", h))
    return c3("6", h).permute(0, 2, 1)
