"""Times the E-step chain launch alone at the benchmark shape (run on the B200 box):
python tools/estep_time.py [batch] [precision]"""
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
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
vae = bench.build_model()
cfg = McemConfig(model="M2", niter=2, nmf_rank=10, precision=prec)
enh = Enhancer(vae, cfg, "cuda:0")
x, s, nz, labels = bench.make_inputs(B, 0)
b = enh.prepare(list(x), labels, seed=0)
for it in range(3):
    E.estep(b, enh.dec, 30, 10, 0.01, prec, seed=1, chain=it)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
for i in range(10):
    ev[i].record()
    E.estep(b, enh.dec, 30, 10, 0.01, prec, seed=1, chain=10 + i)
ev[10].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
print("chain (30+10 steps, %d frames) median %.4f ms  min %.4f ms; finite=%s" %
      (b.NP, ts[5], ts[0], bool(torch.isfinite(b.Vs).all())))
