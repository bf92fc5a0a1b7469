"""E-step / M-step launch times at BASELINE config 4 (M1, 8 utterances x 30 s, K=32, R_E=10); run on the B200 box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "guided-vae-nmf_b200")):
    sys.path.insert(0, p_)
from gvn import engine as E  # noqa: E402
from gvn.pipeline import McemConfig, Enhancer  # noqa: E402
from gvn.synth import synth_batch  # noqa: E402
from python.models.models import VariationalAutoencoder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(0)
vae = VariationalAutoencoder([513, 16, [128, 128]]).eval()
cfg = McemConfig(model="M1", niter=2, nmf_rank=K, burnin_E_step=10, precision="f16")
enh = Enhancer(vae, cfg, "cuda:0")
x, s, n = synth_batch(B, seed=0, T=480000)
b = enh.prepare(list(x), None, seed=0)
scratch = E.MstepScratch(b, 4)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    return sorted(ts)[len(ts) // 2]


te = timed(lambda: E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=0))
state = [t.clone() for t in (b.W, b.H, b.g, b.Vb)]


def m():
    for t, s0 in zip((b.W, b.H, b.g, b.Vb), state):
        t.copy_(s0)
    E.mstep(b, 10, scratch, 0, 1)


def restore_only():
    for t, s0 in zip((b.W, b.H, b.g, b.Vb), state):
        t.copy_(s0)


tm = timed(m) - timed(restore_only)
nbytes = 2 * 11 * 513 * sum(b.n_frames_host) * 4
print("C4 shape B=%d K=%d NP=%d: E-step chain (30+10) %.3f ms | M-step %.3f ms = %.0f GB/s algorithmic (%.1f %% of 6556)"
      % (B, K, b.NP, te, tm, nbytes / tm / 1e6, 100 * nbytes / tm / 1e6 / 6556))
