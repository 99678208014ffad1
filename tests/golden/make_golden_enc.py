"""Golden fixture for the RVQ-VAE encoder side (`RVQVAE.map2latent`, models/vq/model.py:95-100) from the REAL
reference module (build container only; reuses the import shims of make_golden.py).

    python tests/golden/make_golden_enc.py        # needs /root/reference; writes tests/golden/rvq_enc.npz
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (applies the shims, chdirs into the reference)

from syntalker_b200 import synth  # noqa: E402


def main():
    g = torch.Generator().manual_seed(11)
    out = {}
    for d in synth.PART_DIMS_BEATX:
        vq = mg.build_vq(d)
        pose = torch.randn(2, 128, d, generator=g)              # normalised 6d pose features of one window (trainer:290-294)
        out[f"pose{d}"] = pose.clone()
        out[f"lat{d}"] = vq.map2latent(pose)                     # [2, 32, 512]
    mg.save("rvq_enc", **out)
    print("done")


if __name__ == "__main__":
    main()
