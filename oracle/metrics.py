"""Oracle: the rank-reducible part of the evaluation tail, restated with the reference's numpy calls.

dataloaders/data_tools.py:1615-1680 (FIDCalculator.frechet_distance / calculate_frechet_distance), utils/metric.py:12-27 (L1div),
diffusion_rvqvae_trainer.py:621-622 (330-d -> axis-angle).  Test infrastructure only."""
import numpy as np
from scipy import linalg

from . import pose as opose


def frechet_distance(samples_A, samples_B):
    """data_tools.py:1615-1625 with calculate_frechet_distance :1628-1680."""
    mu1, sigma1 = np.mean(samples_A, axis=0), np.cov(samples_A, rowvar=False)
    mu2, sigma2 = np.mean(samples_B, axis=0), np.cov(samples_B, rowvar=False)
    diff = mu1 - mu2
    try:
        covmean, _ = linalg.sqrtm(sigma1.dot(sigma2), disp=False)
    except TypeError:                                  # scipy >= 1.16 dropped `disp` and returns the matrix alone
        covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * 1e-6
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            return 1e+10
        covmean = covmean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)


class L1div:
    """utils/metric.py:12-27, statement for statement."""

    def __init__(self):
        self.counter = 0
        self.sum = 0

    def run(self, results):
        results = np.array(results, copy=True)
        self.counter += results.shape[0]
        mean = np.mean(results, 0)
        for i in range(results.shape[0]):
            results[i, :] = abs(results[i, :] - mean)
        self.sum += np.sum(results)

    def avg(self):
        return self.sum / self.counter


def poses_aa165(rec_pose):
    """trainer:621-622: rc.matrix_to_axis_angle(rc.rotation_6d_to_matrix(rec_pose.reshape(bs*n, j, 6))).reshape(bs*n, j*3)."""
    lead = rec_pose.shape[:-1]
    m = opose.rotation_6d_to_matrix(rec_pose.reshape(-1, 55, 6))
    return opose.matrix_to_axis_angle(m).reshape(*lead, 165)
