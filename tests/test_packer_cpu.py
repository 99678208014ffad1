"""The packer's folding algebra, checked on CPU: packed dataflow (tests/packed_emulator.py) vs the oracle."""
import numpy as np
import pytest
import torch

import packed_emulator as emu
from oracle import mdm as omdm
from oracle import rvq as orvq
from syntalker_b200 import packer, synth

torch.set_grad_enabled(False)


@pytest.mark.parametrize("variant", synth.VARIANTS)
def test_packed_forward_matches_oracle(variant):
    W = synth.mdm_state_dict(variant, seed=0)
    P = packer.pack_mdm(W)
    assert packer.detect_variant(W) == variant
    inp = synth.make_inputs(2, seed=4, variant=variant)
    t = torch.tensor([980, 0])
    y = {k: inp[k] for k in ("audio", "word", "seed")}
    style = inp.get("style_feature", inp.get("style_upper"))
    y["style_feature"] = style
    cst, g2 = emu.cond(P, inp["audio"], inp["word"], inp["seed"])
    sv = None if variant == "beatx" else style @ P["w_style"].t()
    ref = omdm.mdm_forward(W, inp["noise"], t, y, variant)
    got = emu.trunk(P, inp["noise"], t, cst, g2, sv)
    assert float((ref - got).abs().max()) < 1e-4
    if variant != "beatx":
        yu = dict(y); yu["uncond"] = True
        ref_u = omdm.mdm_forward(W, inp["noise"], t, yu, variant)
        svu = P["null_sv"][None].expand(2, -1) if variant == "h3d" else None
        assert float((ref_u - emu.trunk(P, inp["noise"], t, cst, g2, svu)).abs().max()) < 1e-4
    if variant == "h3d":
        ya = dict(y); ya["uncond_audio"] = True
        cstn, _ = emu.cond(P, inp["audio"], inp["word"], inp["seed"], null_audio=True)
        assert float((cstn[0] - cstn[1]).abs().max()) < 1e-7          # a per-model constant
        ref_a = omdm.mdm_forward(W, inp["noise"], t, ya, variant)
        assert float((ref_a - emu.trunk(P, inp["noise"], t, cstn, g2, sv)).abs().max()) < 1e-4


@pytest.mark.parametrize("plan", ["none", "text", "two"])
def test_ddim_z_recursion_equals_the_reference_loop(plan):
    """Deterministic DDIM keeps the loop in token space (packer `w_xo`, `c_xo`, `w_xo2`, `c_xo2`):
    z_{k-1} = beta_k z_k + alpha_k (W_xo h_mix + c_xo) must reproduce the reference's x-space loop (gaussian_diffusion.py:772-790)
    without guidance (denoiser.MDM), under ClassifierFreeSampleModel (cfg_sampler.py:17-28) and under
    TwoClassifierFreeSampleModel (cfg_sampler.py:38-54, denoiser_h3d)."""
    from oracle import diffusion as odiff
    from syntalker_b200 import schedule
    variant = {"none": "beatx", "text": "beatx_motionclip", "two": "h3d"}[plan]
    W = synth.mdm_state_dict(variant, seed=0)
    P = packer.pack_mdm(W)
    inp = synth.make_inputs(2, seed=6, variant=variant)
    y = {k: inp[k] for k in ("audio", "word", "seed")}
    model = lambda a, b, c: omdm.mdm_forward(W, a, b, c, variant)
    cst, g2 = emu.cond(P, inp["audio"], inp["word"], inp["seed"])
    if plan == "none":
        fn = model
        evals, mix = [(cst, g2, None)], lambda o: o[0]
    elif plan == "text":
        y["style_feature"] = inp["style_feature"]
        y["scale"] = torch.ones(1) * 2.0
        fn = lambda x, t, yy: omdm.cfg_text(model, x, t, yy)
        evals = [(cst, g2, inp["style_feature"] @ P["w_style"].t()), (cst, g2, None)]      # conditional, unconditional
        mix = lambda o: o[1] + 2.0 * (o[0] - o[1])
    else:
        y["style_feature"] = inp["style_upper"]
        y["scale_audio"], y["scale_prompt"] = torch.ones(1) * 1.5, torch.ones(1) * 3.0
        fn = lambda x, t, yy: omdm.cfg_two(model, x, t, yy)
        cst_n, _ = emu.cond(P, inp["audio"], inp["word"], inp["seed"], null_audio=True)
        sv_null = P["null_sv"][None].expand(2, -1)
        sv_real = inp["style_upper"] @ P["w_style"].t()
        evals = [(cst_n, g2, sv_null), (cst, g2, sv_null), (cst_n, g2, sv_real)]           # uu, ut, ua
        mix = lambda o: o[0] + 1.5 * (o[1] - o[0]) + 3.0 * (o[2] - o[0])
    sch = odiff.make_schedule(respacing="ddim4")
    ref = odiff.ddim_sample_loop(sch, fn, inp["noise"], y)
    betas, tmap = schedule.respace(schedule.get_named_beta_schedule("cosine", 1000), schedule.space_timesteps(1000, "ddim4"))
    assert list(tmap) == list(sch.timestep_map)
    coef = schedule.ddim_coefs(schedule.Tables(betas))
    got = emu.ddim_z_loop(P, coef, tmap, inp["noise"], evals, mix)
    assert float((ref - got).abs().max()) < 1e-4


def test_module_prefix_is_stripped():
    W = synth.mdm_state_dict("beatx", seed=0)
    Wm = {"module." + k: v for k, v in W.items()}
    a, b = packer.pack_mdm(W), packer.pack_mdm(Wm)
    assert all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.parametrize("dim", synth.PART_DIMS_BEATX)
def test_packed_rvq_decoder(dim):
    W = synth.rvq_state_dict(dim, seed=0)
    P = packer.pack_rvq(W)
    xq = torch.randn(2, 512, 32, generator=torch.Generator().manual_seed(3))
    assert float((orvq.decoder(W, xq) - emu.rvq_decoder(P, xq, dim)).abs().max()) < 1e-5
    assert P[f"dec.6.w"].shape == (dim, 1536)


def test_upsample_conv_even_odd_split():
    """nearest x2 upsample + k3 conv (encdec.py:59-61) == the two 2-tap convs the packer builds for the TC engine."""
    import torch.nn.functional as F
    W = synth.rvq_state_dict(78, seed=0)
    P = packer.pack_rvq(W)
    h = torch.randn(2, 512, 32, generator=torch.Generator().manual_seed(9))
    ref = F.conv1d(F.interpolate(h, scale_factor=2, mode="nearest"), W["decoder.model.2.2.weight"], W["decoder.model.2.2.bias"], padding=1)
    hp = F.pad(h, (1, 1))
    we, wo = P["dec.2.2.even.w"].reshape(512, 2, 512), P["dec.2.2.odd.w"].reshape(512, 2, 512)
    mm = lambda w, x: torch.einsum("oc,bct->bot", w, x)
    even = mm(we[:, 0], hp[:, :, 0:32]) + mm(we[:, 1], hp[:, :, 1:33]) + P["dec.2.2.even.b"][None, :, None]
    odd = mm(wo[:, 0], hp[:, :, 1:33]) + mm(wo[:, 1], hp[:, :, 2:34]) + P["dec.2.2.odd.b"][None, :, None]
    out = torch.stack([even, odd], dim=-1).reshape(2, 512, 64)
    assert float((out - ref).abs().max()) < 1e-5


def test_layernorm_folded_into_linear():
    """(1/sigma)(W' x - mu s) + c == Linear(LayerNorm(x)) for the qkv and fc1 layers (what the TC epilogue computes)."""
    import torch.nn.functional as F
    W = synth.mdm_state_dict("beatx", seed=0)
    P = packer.pack_mdm(W)
    x = 3.0 * torch.randn(64, 512, generator=torch.Generator().manual_seed(2)) + 0.7
    mu = x.mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
    for i in (0, 7):
        p = f"mytimmblocks.{i}."
        ref = F.linear(F.layer_norm(x, (512,), W[p + "norm1.weight"], W[p + "norm1.bias"], 1e-5), W[p + "attn.qkv.weight"])
        got = rstd * (x @ P[f"blk.{i}.qkv.wg"].t() - mu * P[f"blk.{i}.qkv.s"]) + P[f"blk.{i}.qkv.c"]
        assert float((ref - got).abs().max()) < 2e-5
        ref = F.linear(F.layer_norm(x, (512,), W[p + "norm2.weight"], W[p + "norm2.bias"], 1e-5), W[p + "mlp.fc1.weight"], W[p + "mlp.fc1.bias"])
        got = rstd * (x @ P[f"blk.{i}.fc1.wg"].t() - mu * P[f"blk.{i}.fc1.s"]) + P[f"blk.{i}.fc1.c"]
        assert float((ref - got).abs().max()) < 2e-5


def test_gelu_single_branch_erf_table_and_error():
    """The GELU of the trunk kernel's epilogue uses a fitted single-branch erf (tests/fit_erf.py): the table compiled into the
    kernel is the fitted one, and its fp32 evaluation is as close to float64 as the reference's own fp32 nn.GELU."""
    import os, re, sys
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import fit_erf
    src = open(os.path.join(here, "..", "syntalker_b200", "csrc", "st_gemm_tc.cu")).read()
    m = re.search(r"kErfQ\[11\] = \{([^}]*)\}", src)
    table = [float(v.strip().rstrip("f")) for v in m.group(1).split(",")]
    assert table == fit_erf.COEF
    assert np.allclose(fit_erf.fit(), fit_erf.COEF, rtol=0, atol=1e-8)
    max_abs, _ = fit_erf.errors()
    assert max_abs < 8e-7
    x = torch.linspace(-10, 10, 200001)
    ref64 = x.double() * 0.5 * (1 + torch.erf(x.double() / 2 ** 0.5))
    assert float((torch.nn.functional.gelu(x).double() - ref64).abs().max()) > max_abs      # the reference's fp32 GELU is no closer
