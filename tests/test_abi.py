"""CPU checks of the C-ABI: the library loads without a GPU, exports every symbol the header declares,
and the Python mirrors of the reference interface fail loudly instead of falling back."""
import os
import re

import numpy as np
import pytest
import torch

from syntalker_b200 import _lib, schedule
from syntalker_b200.diffusion import SpacedDiffusion, create_gaussian_diffusion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "syntalker_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == syms
    assert L.st_abi_version() == 1


def test_engine_switch_and_error_text():
    L = _lib.lib()
    assert _lib.get_engine() == "tc"                     # the product engine is the default
    assert L.st_set_engine(0) == 0 and _lib.get_engine() == "simt"
    assert L.st_set_engine(7) == -1
    assert b"unknown engine" in L.st_last_error()
    assert L.st_set_engine(1) == 0


def test_schedule_host_tables_match_golden(golden):
    g = golden("schedule")
    for tag, d in (("ddim50", create_gaussian_diffusion(use_ddim=True)), ("ddpm1000", create_gaussian_diffusion()),
                   ("ddim10", create_gaussian_diffusion(timestep_respacing="ddim10")),
                   ("sec20", create_gaussian_diffusion(timestep_respacing=[20]))):
        assert list(g[f"{tag}.timestep_map"]) == d.timestep_map
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                  "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"):
            assert np.array_equal(g[f"{tag}.{k}"], getattr(d, k)), (tag, k)


def test_step_coefficients():
    d = create_gaussian_diffusion(use_ddim=True)
    c = schedule.ddim_coefs(d)
    assert c.shape == (50, 5) and c.dtype == np.float32
    assert c[0, 2] == 1.0 and c[0, 3] == 0.0          # alpha_bar_prev = 1 at k = 0  =>  x <- x0 exactly
    assert np.all(c[:, 4] == 0)                        # eta = 0
    c_eta = schedule.ddim_coefs(d, eta=1.0)
    assert c_eta[0, 4] == 0.0 and np.all(c_eta[1:, 4] > 0)
    p = schedule.ddpm_coefs(create_gaussian_diffusion())
    assert p[0, 2] == 0.0 and np.all(p[1:, 2] > 0)     # no noise at t == 0
    assert abs(float(p[999, 0]) - float(np.float32(create_gaussian_diffusion().posterior_mean_coef1[999]))) == 0


def test_no_fallback_paths():
    d = create_gaussian_diffusion(use_ddim=True)
    with pytest.raises(TypeError):
        d.ddim_sample_loop(torch.nn.Linear(2, 2), (1, 1536, 1, 32), clip_denoised=False, model_kwargs={"y": {}})
    with pytest.raises(NotImplementedError):
        d.ddim_sample_loop(None, (1, 1536, 1, 32), clip_denoised=False, dump_steps=[1])
    with pytest.raises(NotImplementedError):
        d.ddim_sample_loop(None, (1, 1536, 1, 32), clip_denoised=False, const_noise=True)
    if not torch.cuda.is_available():
        from syntalker_b200 import synth
        from syntalker_b200.denoiser import MDM
        with pytest.raises(_lib.StError):
            MDM(None).load_state_dict(synth.mdm_state_dict("beatx"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "syntalker_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
