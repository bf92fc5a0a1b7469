"""Chain launch time with the noise read from a precomputed tape (eps, u in HBM) instead of the in-kernel Philox draw
(run on the B200 box): python tools/estep_tape_time.py [batch]"""
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
cfg = McemConfig(model="M2", niter=2, nmf_rank=10, precision="f16")
enh = Enhancer(vae, cfg, "cuda:0")
x, s, nz, labels = bench.make_inputs(B, 0)
b = enh.prepare(list(x), labels, seed=0)
eps = torch.randn(40, b.L, b.NP, device="cuda")
u = torch.rand(40, b.NP, device="cuda").clamp_(1e-6, 1.0)
for mode in ("philox", "tape", "philox", "tape"):
    kw = dict(eps=eps, u=u) if mode == "tape" else {}
    for it in range(3):
        E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=it, **kw)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    for i in range(10):
        ev[i].record()
        E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=10 + i, **kw)
    ev[10].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
    print("%-7s chain (30+10 steps, %d frames) median %.4f ms  min %.4f ms" % (mode, b.NP, ts[5], ts[0]))
