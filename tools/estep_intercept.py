"""Fixed cost of one chain launch (launch + prologue + tail) against its per-step cost: times gvn_estep at the benchmark shape
for several chain lengths and fits T = T0 + steps * t_step (run on the B200 box): python tools/estep_intercept.py [batch]"""
import os
import sys

import numpy as np
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
E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=0)               # packs XV; every later call vouches for it
for R in (1, 10):
    pts = []
    for burnin in (0, 10, 30, 70):
        for it in range(2):
            E.estep(b, enh.dec, burnin, R, 0.01, "f16", seed=1, chain=it, xv_current=True)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        for i in range(10):
            ev[i].record()
            E.estep(b, enh.dec, burnin, R, 0.01, "f16", seed=1, chain=10 + i, xv_current=True)
        ev[10].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
        pts.append((burnin + R, ts[5]))
        print("R=%d burnin=%d: %d steps, median %.4f ms, min %.4f ms" % (R, burnin, burnin + R, ts[5], ts[0]))
    st, t = np.array(pts).T
    slope, icpt = np.polyfit(st, t, 1)
    print("R=%d: %.2f us per step, intercept %.1f us (= launch + prologue + tail + the initial decode%s)" %
          (R, 1e3 * slope, 1e3 * icpt, " + slot-0 decode + store extras" if R else ""))
