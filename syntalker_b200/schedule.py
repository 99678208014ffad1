"""Host-side schedule: the reference's fp64 tables and the fp32 per-step coefficients the kernels consume.

Mirrors diffusion/gaussian_diffusion.py:20-64,160-197 (cosine betas, cumprods, posterior coefficients),
diffusion/respace.py:8-61,73-87 (space_timesteps, SpacedDiffusion re-derived betas, timestep_map) and the
per-step `_extract_into_tensor(...).float()` gathers (gaussian_diffusion.py:1606-1619): each coefficient is
rounded to fp32 exactly where the reference rounds it, then combined with fp32 torch ops in the
reference's order (ddim_sample :772-790, p_sample :541-556), so the device update is bit-compatible.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.0):
    if schedule_name == "linear":
        scale = scale_betas * 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        n = num_diffusion_timesteps
        return np.array([min(1 - bar((i + 1) / n) / bar(i / n), 0.999) for i in range(n)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per, extra = divmod(num_timesteps, len(section_counts))
    start, all_steps = 0, []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            all_steps.append(start + round(cur))
            cur += frac
        start += size
    return set(all_steps)


class Tables:
    """The numpy tables of GaussianDiffusion.__init__ for a (possibly respaced) beta sequence."""

    def __init__(self, betas):
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)


def respace(base_betas, use_timesteps):
    """SpacedDiffusion.__init__ (respace.py:73-87): returns (new_betas, timestep_map)."""
    base = Tables(base_betas)
    keep = set(use_timesteps)
    last, new_betas, tmap = 1.0, [], []
    for i, ac in enumerate(base.alphas_cumprod):
        if i in keep:
            new_betas.append(1 - ac / last)
            last = ac
            tmap.append(i)
    return np.array(new_betas), tmap


def _f(x):
    return torch.tensor(float(x), dtype=torch.float64).float()      # fp64 table entry -> .float()


def ddim_coefs(tab: Tables, eta: float = 0.0) -> np.ndarray:
    """[S,5] fp32 rows {a, b, c1, c2, sigma}: e=(a*x-x0)/b; x<-x0*c1+c2*e+sigma*noise (ddim_sample :772-790)."""
    rows = []
    for k in range(tab.num_timesteps):
        a, b = _f(tab.sqrt_recip_alphas_cumprod[k]), _f(tab.sqrt_recipm1_alphas_cumprod[k])
        ab, abp = _f(tab.alphas_cumprod[k]), _f(tab.alphas_cumprod_prev[k])
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        c1 = torch.sqrt(abp)
        c2 = torch.sqrt(1 - abp - sigma ** 2)
        nz = 0.0 if k == 0 else 1.0
        rows.append([float(a), float(b), float(c1), float(c2), float(nz * sigma)])
    return np.asarray(rows, dtype=np.float32)


def ddpm_coefs(tab: Tables) -> np.ndarray:
    """[S,5] fp32 rows {coef1, coef2, sigma, 0, 0}: x<-coef1*x0+coef2*x+sigma*noise (:383, :541-556)."""
    rows = []
    for k in range(tab.num_timesteps):
        c1, c2 = _f(tab.posterior_mean_coef1[k]), _f(tab.posterior_mean_coef2[k])
        nz = torch.tensor(0.0 if k == 0 else 1.0)
        sigma = nz * torch.exp(0.5 * _f(tab.posterior_log_variance_clipped[k]))
        rows.append([float(c1), float(c2), float(sigma), 0.0, 0.0])
    return np.asarray(rows, dtype=np.float32)
