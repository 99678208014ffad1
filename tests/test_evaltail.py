"""Evaluation tail: the rank-reducible metric statistics.  CPU part: the reduction logic over a world_size-2
gloo group against the oracle's numpy restatement of the reference; GPU part: the native accumulation kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics as ometrics
from syntalker_b200 import evaltail


def _data(seed, n, d, shift=0.0):
    g = np.random.default_rng(seed)
    a = g.standard_normal((d, d)) / np.sqrt(d)
    return (g.standard_normal((n, d)) @ a + shift).astype(np.float32)


def _sums(x):
    x = x.astype(np.float64)
    return x.shape[0], x.sum(0), x.T @ x


def test_moments_reproduce_numpy_mean_and_cov():
    x = _data(0, 500, 24, 0.3)
    m = evaltail.Moments(24).add_sums(*_sums(x[:200])).add_sums(*_sums(x[200:]))
    mu, cov = m.mean_cov()
    assert np.allclose(mu, np.mean(x, axis=0), atol=1e-6) and np.allclose(cov, np.cov(x, rowvar=False), atol=1e-6)
    a, b = _data(1, 400, 24), _data(2, 300, 24, 0.5)
    ma, mb = evaltail.Moments(24).add_sums(*_sums(a)), evaltail.Moments(24).add_sums(*_sums(b))
    f = evaltail.frechet_distance(*ma.mean_cov(), *mb.mean_cov())
    assert abs(f - ometrics.frechet_distance(a, b)) < 1e-5 * max(1.0, abs(f))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out, ori = _data(10, 640, 16, 0.2), _data(11, 640, 16)
    seqs = [_data(20 + i, 50 + 7 * i, 30) for i in range(6)]
    sh = slice(rank * 320, (rank + 1) * 320)
    fid = evaltail.FIDAccumulator(16)
    fid.out.add_sums(*_sums(out[sh])); fid.ori.add_sums(*_sums(ori[sh]))
    l1 = evaltail.L1divAccumulator()
    for s in seqs[rank::world]:                                   # sequences are sharded over ranks like the clips
        o = ometrics.L1div(); o.run(s)
        l1.add_sums(o.sum, o.counter)
    fid.all_reduce(); l1.all_reduce()
    ref_l1 = ometrics.L1div()
    for s in seqs:
        ref_l1.run(s)
    q.put((rank, fid.fid(), ometrics.frechet_distance(out, ori), l1.avg(), ref_l1.avg()))
    dist.destroy_process_group()


def test_metric_reduction_over_two_ranks_gloo():
    """Each rank holds the statistics of its shard; after the all-reduce every rank has the FID / L1div of the whole test set."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for _rank, fid, fid_ref, l1, l1_ref in res:
        assert abs(fid - fid_ref) < 1e-6 * max(1.0, abs(fid_ref))
        assert abs(l1 - l1_ref) < 1e-6 * l1_ref


@pytest.mark.gpu
def test_native_accumulators_vs_oracle():
    torch.set_grad_enabled(False)
    dev = torch.device("cuda")
    out, ori = _data(30, 1500, 240, 0.1), _data(31, 1400, 240)
    fid = evaltail.FIDAccumulator(240, dev)
    for i in range(0, 1500, 500):                                  # three "sequences" per side
        fid.update(torch.from_numpy(out[i:i + 500]).to(dev), torch.from_numpy(ori[i:i + 500][:max(0, min(500, 1400 - i))]).to(dev))
    mu, cov = fid.out.mean_cov()
    assert np.allclose(mu, np.mean(out, 0), atol=1e-6) and np.allclose(cov, np.cov(out, rowvar=False), atol=1e-5)
    ref = ometrics.frechet_distance(out, ori)
    assert abs(fid.fid() - ref) < 1e-4 * max(1.0, abs(ref))
    l1, ref_l1 = evaltail.L1divAccumulator(dev), ometrics.L1div()
    for i in range(4):
        s = _data(40 + i, 90 + 11 * i, 165)
        l1.run(torch.from_numpy(s).to(dev)); ref_l1.run(s)
    assert abs(l1.avg() - ref_l1.avg()) < 1e-5 * ref_l1.avg()
    # 330-d -> 165-d axis-angle of the result files
    g = torch.Generator().manual_seed(5)
    pose = torch.randn(3, 40, 330, generator=g)
    aa = evaltail.poses_aa165(pose.to(dev)).cpu()
    aa_ref = ometrics.poses_aa165(pose)
    d = (aa - aa_ref).abs()
    assert float(d.max()) < 1e-3 and float((d > 1e-5).float().mean()) < 2e-2      # the axis-angle vector is ill-conditioned near pi (the reference's own fp32 result is 1.8e-4 from its fp64 one here) ...
    from oracle import pose as opose
    R, R_ref = opose.axis_angle_to_matrix(aa.reshape(-1, 3)), opose.axis_angle_to_matrix(aa_ref.reshape(-1, 3))
    dR = (R - R_ref).abs()
    print(f"aa165: max-abs {float(d.max()):.2e}, rotation max-abs {float(dR.max()):.2e}, share of rotation entries over 1e-5: {float((dR > 1e-5).float().mean()):.2e}")
    assert float(dR.max()) < 1e-3 and float((dR > 1e-5).float().mean()) < 2e-2    # the old pytorch3d matrix_to_quaternion is itself fp32-noisy near pi
    # one sequence through the whole tail with stand-in callables
    enc = lambda p: p.reshape(p.shape[0], -1, 8 * 330)[..., :240]
    fid2, l12 = evaltail.FIDAccumulator(240, dev), evaltail.L1divAccumulator(dev)
    poses = evaltail.eval_tail_step(pose.to(dev), pose.flip(0).to(dev), fid2, l12, enc)
    assert poses.shape == (120, 165) and float(fid2.out.state[0]) == 3 * 4 and float(l12.state[1]) == 120
