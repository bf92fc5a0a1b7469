"""Times the M-step launches alone at the benchmark shape (run on the B200 box): python tools/mstep_time.py [batch]"""
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
scratch = E.MstepScratch(b, 4)
E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=0)
state = [t.clone() for t in (b.W, b.H, b.g, b.Vb)]
for it in range(3):
    E.mstep(b, 10, scratch, 0, 1)
torch.cuda.synchronize()
ts = []
for i in range(10):
    for t, s0 in zip((b.W, b.H, b.g, b.Vb), state):
        t.copy_(s0)
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    E.mstep(b, 10, scratch, 0, 1)
    c.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(c))
ts.sort()
print("diag %s: M-step (%d frames) median %.4f ms  min %.4f ms" % (os.environ.get("GVN_MSTEP_DIAG", "0"), b.NP, ts[5], ts[0]))
