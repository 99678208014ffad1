"""Oracle: schedule tables and the two sampling loops.

Follows diffusion/gaussian_diffusion.py:20-64 (cosine betas), :160-197 (tables), :279-397
(p_mean_variance, START_X / FIXED_SMALL), :505-557 + :672-739 (DDPM), :741-791 + :937-1002 (DDIM),
:1606-1619 (_extract_into_tensor: fp64 table -> gather -> .float()), diffusion/respace.py:8-61,73-87,124-129.
"""
import math

import numpy as np
import torch


def cosine_betas(n=1000, max_beta=0.999):
    f = lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - f((i + 1) / n) / f(i / n), max_beta) for i in range(n)], dtype=np.float64)


def space_timesteps(num_timesteps, spec):
    """respace.py:8-61: 'ddimN' (integer stride) or a list / comma string of per-section counts."""
    if isinstance(spec, str):
        if spec.startswith("ddim"):
            want = int(spec[4:])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == want:
                    return sorted(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        spec = [int(x) for x in spec.split(",")]
    size_per, extra = divmod(num_timesteps, len(spec))
    start, steps = 0, []
    for i, count in enumerate(spec):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return sorted(set(steps))


class Schedule:
    """fp64 host tables of a respaced process (gaussian_diffusion.py:160-197, respace.py:73-87)."""

    def __init__(self, use_timesteps, base_betas=None):
        base_betas = cosine_betas() if base_betas is None else np.asarray(base_betas, dtype=np.float64)
        base_ac = np.cumprod(1.0 - base_betas, axis=0)
        keep = set(use_timesteps)
        last, betas, tmap = 1.0, [], []
        for i, ac in enumerate(base_ac):
            if i in keep:
                betas.append(1 - ac / last); last = ac; tmap.append(i)
        betas = np.array(betas, dtype=np.float64)
        self.timestep_map = tmap
        self.betas = betas
        self.num_timesteps = len(betas)
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)


def make_schedule(use_ddim=False, respacing=None):
    """diffusion/model_util.py:8-50: cosine/1000, 'ddim50' when use_ddim; `respacing` overrides."""
    spec = respacing if respacing is not None else ("ddim50" if use_ddim else [1000])
    return Schedule(space_timesteps(1000, spec))


def _ext(arr, k, B):
    return torch.full((B, 1, 1, 1), float(np.float32(arr[k])), dtype=torch.float32)


def ddim_sample_loop(sched, model_fn, noise, y, eta=0.0, tap=None, step_noise=None):
    """gaussian_diffusion.py:937-1002 with ddim_sample:741-791; clip_denoised=False, no cond_fn.
    model_fn(x, t_original, y) -> x0 prediction. step_noise(k, x) supplies the randn_like(x) of step k (:781); it only
    matters when eta > 0 (sigma = 0 otherwise), so with eta = 0 it may be omitted and is then not drawn."""
    x = noise
    B = x.shape[0]
    for k in range(sched.num_timesteps - 1, -1, -1):
        t = torch.full((B,), sched.timestep_map[k], dtype=torch.int64)
        x0 = model_fn(x, t, y)
        eps = (_ext(sched.sqrt_recip_alphas_cumprod, k, B) * x - x0) / _ext(sched.sqrt_recipm1_alphas_cumprod, k, B)
        ab, abp = _ext(sched.alphas_cumprod, k, B), _ext(sched.alphas_cumprod_prev, k, B)
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        x = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
        if step_noise is not None:
            x = x + (0.0 if k == 0 else 1.0) * sigma * step_noise(k, x)                  # :786-790, no noise when t == 0
        if tap is not None:
            tap(k, x0, x)
    return x


def p_sample_loop(sched, model_fn, noise, y, step_noise, tap=None):
    """gaussian_diffusion.py:672-739 with p_sample:505-557 and p_mean_variance:279-397.
    step_noise(k, x) supplies the eps the reference draws with randn_like at step k."""
    x = noise
    B = x.shape[0]
    for k in range(sched.num_timesteps - 1, -1, -1):
        t = torch.full((B,), sched.timestep_map[k], dtype=torch.int64)
        x0 = model_fn(x, t, y)
        mean = _ext(sched.posterior_mean_coef1, k, B) * x0 + _ext(sched.posterior_mean_coef2, k, B) * x
        eps = step_noise(k, x)
        nz = 0.0 if k == 0 else 1.0
        x = mean + nz * torch.exp(0.5 * _ext(sched.posterior_log_variance_clipped, k, B)) * eps
        if tap is not None:
            tap(k, x0, x)
    return x
