"""Acceptance rate of the Metropolis-Hastings chain at the bench shape (C2) along the EM iterations, and the
number of distinct kept samples per frame (= slots with non-zero multiplicity).  Run on the B200 box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "guided-vae-nmf_b200")):
    sys.path.insert(0, p_)
from gvn import engine as E  # noqa: E402
from gvn.pipeline import McemConfig, Enhancer  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vae = bench.build_model()
cfg = McemConfig(model="M2", niter=100, nmf_rank=10, precision="f16")
enh = Enhancer(vae, cfg, "cuda:0")
x, s, nz, labels = bench.make_inputs(B, 0)
b = enh.prepare(list(x), labels, seed=0)
scratch = E.MstepScratch(b, 100)
valid = b.frame_utt >= 0
for it in range(100):
    tr = E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=it, trace=(it % 10 == 0 or it == 99))
    if tr is not None:
        acc, dec_, cnt, zs = tr
        d = dec_[:, valid].float()
        used = (b.Vs_w[:10][:, valid] > 0).float().sum(0)
        t8 = (b.Vs_w[:10] > 0).float().sum(0).reshape(-1, 8).max(1).values
        print("iter %3d: acceptance %.3f (burn-in %.3f, kept %.3f); slots in use per frame: mean %.2f, max over 8-frame tile: mean %.2f"
              % (it, d.mean().item(), d[:30].mean().item(), d[30:].mean().item(), used.mean().item(), t8.mean().item()))
    E.mstep(b, 10, scratch, it, 1)
