"""Pin the oracle (oracle/*.py) against outputs of the REAL reference (tests/golden/*.npz, made by
tests/golden/make_golden.py in the build container). Tolerances are fp32 round-off only: the oracle
and the reference call the same torch CPU kernels, but the GPU box may have a different CPU/ISA."""
import numpy as np
import pytest
import torch

from oracle import diffusion as odiff
from oracle import mdm as omdm
from oracle import pose as opose
from oracle import rvq as orvq
from syntalker_b200 import synth

torch.set_grad_enabled(False)


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


@pytest.fixture(scope="module")
def weights():
    return {v: synth.mdm_state_dict(v, seed=0) for v in synth.VARIANTS}


def y_of(inp):
    return {k: inp[k] for k in ("audio", "word", "seed", "style_feature") if k in inp}


@pytest.mark.parametrize("tag,spec", [("ddim50", "ddim50"), ("ddpm1000", [1000]), ("ddim10", "ddim10"), ("sec20", [20])])
def test_schedule_tables_bit_exact(golden, tag, spec):
    g = golden("schedule")
    s = odiff.Schedule(odiff.space_timesteps(1000, spec))
    assert list(g[f"{tag}.timestep_map"]) == s.timestep_map
    for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_log_variance_clipped", "posterior_mean_coef1",
              "posterior_mean_coef2"):
        assert np.array_equal(g[f"{tag}.{k}"], getattr(s, k)), k     # fp64, same numpy ops => identical


@pytest.mark.parametrize("variant", synth.VARIANTS)
def test_mdm_forward(golden, weights, variant):
    g = golden(f"mdm_{variant}")
    inp = synth.make_inputs(2, seed=1, variant=variant)
    y = y_of(inp)
    if variant == "h3d":
        y["style_feature"] = inp["style_upper"]
    t = torch.from_numpy(g["t"])
    taps = {}
    out = omdm.mdm_forward(weights[variant], inp["noise"], t, y, variant, taps)
    assert maxabs(out, g["out"]) < 2e-5
    if variant == "beatx":
        assert maxabs(taps["block0"], g["block0"]) < 2e-5
        assert maxabs(taps["block7"], g["block7"]) < 2e-5
        assert maxabs(omdm.wav_encoder(weights[variant], inp["audio"])[:, ::8], g["wav"]) < 1e-5
    else:
        yu = dict(y); yu["uncond"] = True
        assert maxabs(omdm.mdm_forward(weights[variant], inp["noise"], t, yu, variant), g["out_uncond"]) < 2e-5
    if variant == "h3d":
        ya = dict(y); ya["uncond_audio"] = True
        assert maxabs(omdm.mdm_forward(weights[variant], inp["noise"], t, ya, variant), g["out_uncond_audio"]) < 2e-5


def test_cfg_text(golden, weights):
    W = weights["beatx_motionclip"]
    inp = synth.make_inputs(2, seed=1, variant="beatx_motionclip")
    y = y_of(inp); y["scale"] = torch.ones(1) * 2.0
    fn = lambda x, t, yy: omdm.mdm_forward(W, x, t, yy, "beatx_motionclip")
    out = omdm.cfg_text(fn, inp["noise"], torch.tensor([500, 500]), y)
    assert maxabs(out, golden("cfg_text")["out"]) < 5e-5


def test_cfg_identity_without_motionclip(weights):
    """ClassifierFreeSampleModel over denoiser.MDM without motionclip is an exact no-op."""
    W = weights["beatx"]
    inp = synth.make_inputs(1, seed=3, variant="beatx")
    y = y_of(inp); y["scale"] = torch.ones(1) * 2.0
    fn = lambda x, t, yy: omdm.mdm_forward(W, x, t, yy, "beatx")
    t = torch.tensor([700])
    assert torch.equal(omdm.cfg_text(fn, inp["noise"], t, y), fn(inp["noise"], t, y))


def test_cfg_bodypart(golden, weights):
    W = weights["h3d"]
    inp = synth.make_inputs(1, seed=2, variant="h3d")
    y = y_of(inp)
    y["style_feature"] = {"upper_mask": inp["style_upper"], "hands_mask": None, "lower_mask": inp["style_lower"]}
    fn = lambda x, t, yy: omdm.mdm_forward(W, x, t, yy, "h3d")
    out = omdm.cfg_bodypart(fn, inp["noise"], torch.tensor([300]), y)
    assert maxabs(out, golden("cfg_bodypart")["out"]) < 1e-4


def test_loops(golden, weights):
    W = weights["beatx"]
    g = golden("loops")
    inp = synth.make_inputs(1, seed=1, variant="beatx")
    fn = lambda x, t, yy: omdm.mdm_forward(W, x, t, yy, "beatx")
    s10 = odiff.ddim_sample_loop(odiff.make_schedule(respacing="ddim10"), fn, inp["noise"], y_of(inp))
    assert maxabs(s10, g["ddim10"]) < 1e-4
    torch.manual_seed(123)
    p20 = odiff.p_sample_loop(odiff.make_schedule(respacing=[20]), fn, inp["noise"], y_of(inp),
                              lambda k, x: torch.randn_like(x))
    assert maxabs(p20, g["ddpm_sec20"]) < 1e-4


def test_rvq_and_pose(golden):
    g = golden("rvq")
    recs = []
    for d in synth.PART_DIMS_BEATX:
        W = synth.rvq_state_dict(d, seed=0)
        lat = torch.from_numpy(g[f"lat{d}"])
        rec, idx = orvq.latent2origin(W, lat)
        assert np.array_equal(idx.numpy(), g[f"idx{d}"])
        assert maxabs(rec, g[f"rec{d}"]) < 2e-5
        xq, _, _ = orvq.residual_quantize(W, lat.permute(0, 2, 1))
        assert maxabs(xq[:, ::8], g[f"xq{d}"]) < 1e-5
        recs.append(torch.from_numpy(g[f"rec{d}"]))
    gp = golden("pose")
    ms = {k: torch.from_numpy(v) for k, v in np.load(synth_data("beatx_mean_std.npz")).items()}
    pose, trans = opose.assemble_330(recs[0], recs[1], recs[2], ms, torch.from_numpy(gp["jaw"]))
    assert maxabs(trans, gp["rec_trans"]) < 1e-5
    d = np.abs(pose.numpy() - gp["rec_pose"])
    assert d.max() < 1e-3 and np.mean(d > 1e-5) < 1e-3      # rotation round trip is ill-conditioned near pi


def synth_data(name):
    import os
    return os.path.join(os.path.dirname(synth.__file__), "data", name)


def test_e2e_config1(golden, weights):
    """BASELINE config 1: B=1, DDIM-10 -> x5 -> latent2origin x3 -> 330-d."""
    W = weights["beatx"]
    inp = synth.make_inputs(1, seed=1, variant="beatx")
    fn = lambda x, t, yy: omdm.mdm_forward(W, x, t, yy, "beatx")
    s = odiff.ddim_sample_loop(odiff.make_schedule(respacing="ddim10"), fn, inp["noise"], y_of(inp))
    lats = opose.sample_to_parts(s, 5.0)
    recs = [orvq.latent2origin(synth.rvq_state_dict(d, seed=0), l)[0] for d, l in zip(synth.PART_DIMS_BEATX, lats)]
    ms = {k: torch.from_numpy(v) for k, v in np.load(synth_data("beatx_mean_std.npz")).items()}
    pose, trans = opose.assemble_330(recs[0], recs[1], recs[2], ms, None)
    g = golden("e2e_config1")
    assert maxabs(trans, g["rec_trans"]) < 1e-4
    assert maxabs(pose, g["rec_pose"]) < 1e-3


def test_rvq_encoder_oracle_vs_reference(golden):
    """oracle.rvq.map2latent against RVQVAE.map2latent of the real reference (tests/golden/make_golden_enc.py)."""
    g = golden("rvq_enc")
    for d in synth.PART_DIMS_BEATX:
        W = synth.rvq_state_dict(d, seed=0)
        lat = orvq.map2latent(W, torch.from_numpy(g[f"pose{d}"]))
        assert lat.shape == (2, 32, 512)
        assert float((lat - torch.from_numpy(g[f"lat{d}"])).abs().max()) <= 2e-6


def test_long_clip_window_slicing_and_seed_handoff():
    """oracle.longclip restates trainer:413-474; checked here with a stub denoiser (x0 = 0.5 x + mean(seed))."""
    from oracle import longclip as olong
    from oracle import diffusion as odiff
    R, B = 3, 2
    n = olong.ROUND_L * R + 16
    assert olong.n_windows(n) == R and olong.n_windows(n + 111) == R and olong.n_windows(n + 112) == R + 1
    g = torch.Generator().manual_seed(3)
    audio = torch.rand(B, olong.AUDIO_PER_FRAME * n, 2, generator=g)
    word = torch.arange(B * n).reshape(B, n)
    a1, w1 = olong.window_inputs(audio, word, 1)
    assert a1.shape == (B, 68224, 2) and w1.shape == (B, 128) and int(w1[0, 0]) == 112 and int(w1[0, -1]) == 239
    seeds = []

    def fn(x, t, y):
        seeds.append(y["seed"].clone())
        return 0.5 * x + y["seed"].mean(dim=(1, 2)).view(-1, 1, 1, 1)

    sched = odiff.make_schedule(respacing="ddim10")
    seed0 = torch.randn(B, 4, 1536, generator=g)
    x_init = torch.randn(R, B, 1536, 1, 32, generator=g)
    lat = olong.long_clip_latents(sched, fn, audio, word, seed0, x_init)
    assert lat.shape == (B, 32 + 28 * (R - 1), 1536)
    assert torch.equal(seeds[0], seed0)
    # window 1's seed = last 4 tokens of window 0's sample = tokens 28..31 of the stitched latents (window 0 is kept whole)
    assert torch.equal(seeds[10], lat[:, 28:32])
    # window 2's seed = last 4 tokens of window 1's sample = stitched tokens 56..59
    assert torch.equal(seeds[20], lat[:, 56:60])


# ---- round 2: goldens of the real reference for the corners round 1 only held against the oracle (tests/golden/make_golden_r2.py) ----

def _h3d_fn(weights):
    W = weights["h3d"]
    return lambda x, t, yy: omdm.mdm_forward(W, x, t, yy, "h3d")


def test_cfg_two_and_h3d_text_cfg_vs_reference(golden, weights):
    """TwoClassifierFreeSampleModel (cfg_sampler.py:31-54) and ClassifierFreeSampleModel(+eval=True, :10-28) over denoiser_h3d."""
    fn = _h3d_fn(weights)
    inp = synth.make_inputs(2, seed=3, variant="h3d")
    y = y_of(inp); y["style_feature"] = inp["style_upper"]
    g = golden("cfg_two")
    y2 = dict(y); y2["scale_audio"] = torch.from_numpy(g["scale_audio"]); y2["scale_prompt"] = torch.from_numpy(g["scale_prompt"])
    assert maxabs(omdm.cfg_two(fn, inp["noise"], torch.from_numpy(g["t"]), y2), g["out"]) < 5e-5
    g = golden("cfg_h3d_text")
    y3 = dict(y); y3["scale"] = torch.from_numpy(g["scale"])
    assert maxabs(omdm.cfg_text(fn, inp["noise"], torch.from_numpy(g["t"]), y3), g["out"]) < 5e-5
    assert maxabs(omdm.cfg_text(fn, inp["noise"], torch.from_numpy(g["t"]), y3, eval_metric=True), g["out_eval"]) < 2e-5


def test_cfg_bodypart_single_scale_vs_reference(golden, weights):
    """ClassifierFreeSampleModel_Bodypart (cfg_sampler.py:125-167), B = 1 like the reference's hard-coded [1,256] null prompt."""
    fn = _h3d_fn(weights)
    inp = synth.make_inputs(1, seed=4, variant="h3d")
    g = golden("cfg_bodypart1")
    y = y_of(inp)
    y["style_feature"] = {"upper_mask": inp["style_upper"], "hands_mask": None, "lower_mask": inp["style_lower"]}
    y["scale"] = torch.from_numpy(g["scale"])
    t = torch.from_numpy(g["t"])
    assert maxabs(omdm.cfg_bodypart1(fn, inp["noise"], t, y), g["out"]) < 5e-5
    assert maxabs(omdm.cfg_bodypart1(fn, inp["noise"], t, y, eval_metric=True), g["out_eval"]) < 2e-5


def test_h3d_decoders_and_623_scatter_vs_reference(golden):
    """latent2origin of the HumanML3D decoders (D = 156 / 360 / 107) and the scatter of h3d_diffusion_new_trainer.py:194-221,604-607."""
    g = golden("h3d_decode")
    recs = []
    for d in synth.PART_DIMS_H3D:
        W = synth.rvq_state_dict(d, seed=0)
        rec, idx = orvq.latent2origin(W, torch.from_numpy(g[f"lat{d}"]))
        assert np.array_equal(idx.numpy(), g[f"idx{d}"])
        assert maxabs(rec, g[f"rec{d}"]) < 2e-5
        recs.append(torch.from_numpy(g[f"rec{d}"]))
    up, ha, lo = opose.h3d_masks()
    assert list(up) == list(g["mask_upper"]) and list(ha) == list(g["mask_hands"]) and list(lo) == list(g["mask_lower"])
    assert torch.equal(opose.assemble_623(*recs), torch.from_numpy(g["rec_pose"]))


def ddpm_tape(seed, S, shape):
    """eps_k in draw order, as tests/golden/make_golden_r2.py fed them to the reference's p_sample_loop."""
    gen = torch.Generator().manual_seed(int(seed))
    return torch.stack([torch.randn(shape, generator=gen) for _ in range(S)])


def test_ddpm1000_first_100_steps_vs_reference(golden, weights):
    """The reference's own 1000-step p_sample_loop (gaussian_diffusion.py:607-739) with its noise draws replayed: the oracle loop is
    held to the sample after 1 and after 100 steps here (the whole loop is ~4 CPU-minutes; the GPU test runs all 1000)."""
    g = golden("ddpm1000")
    W = weights["beatx"]
    inp = synth.make_inputs(1, seed=1, variant="beatx")
    fn = lambda x, t, yy: omdm.mdm_forward(W, x, t, yy, "beatx")
    tape = ddpm_tape(g["seed"], 100, (1, 1536, 1, 32))
    sched = odiff.make_schedule()
    taps = {}

    class Stop(Exception):
        pass

    def tap(k, x0, x):
        taps[999 - k] = x.clone()
        if k == 900:
            raise Stop

    try:
        odiff.p_sample_loop(sched, fn, inp["noise"], y_of(inp), lambda k, x: tape[999 - k], tap=tap)
    except Stop:
        pass
    assert maxabs(taps[0], g["x_after_0"]) < 1e-5
    assert maxabs(taps[99], g["x_after_99"]) < 1e-4
