"""Evaluation tail of `_g_test` / `test()` (diffusion_rvqvae_trainer.py:612-732): what happens to the 330-d pose
features after the hot path -- the part that can be built without SMPL-X and without the VAESKConv evaluator weights.

  * `poses_aa165`: the `poses[n,165]` array of the result files (trainer:621-622, 702-710) from the 330-d features (native kernel).
  * `FIDAccumulator`: the reference concatenates the evaluator latents of every test sequence and calls
    `FIDCalculator.frechet_distance` (dataloaders/data_tools.py:1615-1680: np.mean / np.cov / scipy sqrtm).  Mean and covariance
    follow from (n, sum x, sum x x^T), which ADD over sequences and over ranks: each rank accumulates its shard's statistics on the
    device in float64 (st_moments_accumulate) and the ranks all-reduce 1 + D + D^2 numbers over NCCL -- the "metric reduction" of the
    north_star -- instead of gathering latents.  The evaluator network itself (`VAESKConv.map2latent`, trainer:618-619) is an injected
    callable: its weights (AESKConv_240_100.bin) and `smplx` are not in this image.
  * `L1divAccumulator`: utils/metric.py:12-27 (`L1div.run` per sequence, `avg` at the end); per-rank (sum, counter) all-reduced.
  * `save_result_npz`: the wire / disk format of trainer:702-710.
The sampling hot path never imports this module.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def poses_aa165(rec_pose: torch.Tensor) -> torch.Tensor:
    """rec_pose [..., 330] fp32 CUDA (55 joints x 6d) -> [..., 165] axis-angle (rotation_6d_to_matrix -> matrix_to_axis_angle)."""
    if rec_pose.shape[-1] != 330 or not rec_pose.is_cuda:
        raise ValueError(f"rec_pose must be a CUDA tensor [...,330], got {tuple(rec_pose.shape)} on {rec_pose.device}")
    x = rec_pose.float().contiguous()
    out = torch.empty(x.shape[:-1] + (165,), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().st_pose_330_to_aa165(x.data_ptr(), x.numel() // 330, out.data_ptr(), _lib.stream_ptr()))
    return out


def _all_reduce(t: torch.Tensor, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


class Moments:
    """(n, sum x, sum x x^T) of rows of dimension D, float64; `state` is the flat [1 + D + D*D] buffer that is all-reduced."""

    def __init__(self, dim: int, device=None):
        self.dim = int(dim)
        self.state = torch.zeros(1 + self.dim + self.dim * self.dim, dtype=torch.float64, device=device)

    def update(self, x: torch.Tensor):
        """x [*, D] on the accumulator's CUDA device (st_moments_accumulate; there is no CPU path)."""
        if not (x.is_cuda and self.state.is_cuda):
            raise _lib.StError("Moments.update needs CUDA tensors (use add_sums for statistics computed elsewhere)")
        x = x.reshape(-1, self.dim).float().contiguous()
        if x.shape[0]:
            with torch.cuda.device(x.device):
                _lib.check(_lib.lib().st_moments_accumulate(x.data_ptr(), x.shape[0], self.dim, self.state.data_ptr(), _lib.stream_ptr()))
        return self

    def add_sums(self, n, s1, s2):
        d = self.dim
        self.state[0] += float(n)
        self.state[1:1 + d] += torch.as_tensor(s1, dtype=torch.float64, device=self.state.device).reshape(d)
        self.state[1 + d:] += torch.as_tensor(s2, dtype=torch.float64, device=self.state.device).reshape(d * d)
        return self

    def all_reduce(self, group=None):
        _all_reduce(self.state, group)
        return self

    def mean_cov(self):
        """np.mean(axis=0), np.cov(rowvar=False) (ddof = 1) of everything accumulated (data_tools.py:1617-1620)."""
        d = self.dim
        st = self.state.detach().cpu().numpy()
        n, s1, s2 = st[0], st[1:1 + d], st[1 + d:].reshape(d, d)
        if n < 2:
            raise ValueError("covariance needs at least two rows")
        mu = s1 / n
        cov = (s2 - n * np.outer(mu, mu)) / (n - 1)
        return mu, cov


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """FIDCalculator.calculate_frechet_distance (data_tools.py:1628-1680, after pytorch-fid): ||mu1 - mu2||^2 + Tr(C1 + C2 - 2 sqrt(C1 C2))."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    diff = mu1 - mu2
    try:
        covmean, _ = linalg.sqrtm(sigma1.dot(sigma2), disp=False)
    except TypeError:                                  # scipy >= 1.16 dropped `disp` and returns the matrix alone
        covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError(f"Imaginary component {np.max(np.abs(covmean.imag))}")
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


class FIDAccumulator:
    """latent_out / latent_ori of trainer:618-619 as two `Moments`; `fid()` after `all_reduce()` equals the reference's
    frechet_distance over the concatenation of every rank's latents (1e+10 on a ValueError, data_tools.py:1621-1624)."""

    def __init__(self, dim: int = 240, device=None):
        self.out, self.ori = Moments(dim, device), Moments(dim, device)

    def update(self, latent_out: torch.Tensor, latent_ori: torch.Tensor):
        self.out.update(latent_out)
        self.ori.update(latent_ori)
        return self

    def all_reduce(self, group=None):
        self.out.all_reduce(group)
        self.ori.all_reduce(group)
        return self

    def fid(self) -> float:
        (mu_a, cov_a), (mu_b, cov_b) = self.out.mean_cov(), self.ori.mean_cov()
        try:
            return frechet_distance(mu_a, cov_a, mu_b, cov_b)
        except ValueError:
            return 1e+10


class L1divAccumulator:
    """utils/metric.py:12-27: run(results [n,J]) once per sequence, avg() = sum / counter; state = [sum, counter] float64."""

    def __init__(self, device=None):
        self.state = torch.zeros(2, dtype=torch.float64, device=device)

    def run(self, results: torch.Tensor):
        if not (results.is_cuda and self.state.is_cuda):
            raise _lib.StError("L1divAccumulator.run needs CUDA tensors (use add_sums for statistics computed elsewhere)")
        x = results.float().contiguous()
        if x.dim() != 2:
            raise ValueError(f"results must be [n,J], got {tuple(x.shape)}")
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().st_l1div_accumulate(x.data_ptr(), x.shape[0], x.shape[1], self.state.data_ptr(), _lib.stream_ptr()))
        return self

    def add_sums(self, total, counter):
        self.state[0] += float(total)
        self.state[1] += float(counter)
        return self

    def all_reduce(self, group=None):
        _all_reduce(self.state, group)
        return self

    def avg(self) -> float:
        s, c = self.state.detach().cpu().tolist()
        return s / c


def eval_tail_step(rec_pose, tar_pose, fid: FIDAccumulator, l1: L1divAccumulator, fid_encoder, joints_fn=None, vae_test_len=32):
    """One test sequence of trainer:612-685 on the device: rec_pose / tar_pose [bs,n,330] -> evaluator latents into `fid`
    (`fid_encoder` = VAESKConv.map2latent, any callable [bs,n',330] -> [bs,n'/k,D]), joints into `l1` (`joints_fn` = the SMPL-X joint
    regressor, [n,165] axis-angle -> [n,J]; None = the axis-angle vector itself).  Returns poses [bs*n,165] for the result file."""
    bs, n, _ = rec_pose.shape
    keep = n - n % vae_test_len
    fid.update(fid_encoder(rec_pose[:, :keep]).reshape(-1, fid.out.dim), fid_encoder(tar_pose[:, :keep]).reshape(-1, fid.ori.dim))
    aa = poses_aa165(rec_pose).reshape(bs * n, 165)
    feats = joints_fn(aa) if joints_fn is not None else aa
    for b in range(bs):
        l1.run(feats.reshape(bs, n, -1)[b])
    return aa


def save_result_npz(path, betas, poses, expressions, trans):
    """trainer:693-710: the SMPL-X parameter file consumed by the renderer / the BEAT tooling."""
    as_np = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    np.savez(path, betas=as_np(betas), poses=as_np(poses), expressions=as_np(expressions), trans=as_np(trans), model="smplx2020",
             gender="neutral", mocap_frame_rate=30)
