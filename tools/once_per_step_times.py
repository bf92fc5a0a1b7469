"""CUDA-event times of the kernels that run once per enhancement at the benchmark shape (Wiener filter, STFT, ISTFT, encoder init),
in steady state, back to back (run on the B200 box): python tools/once_per_step_times.py [batch]"""
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
up = enh.upload(list(x), labels)
b = enh.prepare(None, None, seed=0, uploaded=up)
cost, S, Nn, _, _ = E.run_mcem(b, enh.dec, cfg.niter, *cfg.chains(), cfg.var_RW, "f16", seed=0)
(R_E, b_E), (R_W, b_W) = cfg.chains()


def timed(name, fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    for i in range(reps):
        ev[i].record()
        fn()
    ev[reps].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    print("%-28s median %.4f ms  min %.4f ms" % (name, ts[reps // 2], ts[0]))


timed("wiener (R=%d slots)" % R_W, lambda: E.wiener(b, R_W))
timed("istft x2", lambda: (E.istft_from(b, S, b.T, b.T_stride, b.nfft, b.hop), E.istft_from(b, Nn, b.T, b.T_stride, b.nfft, b.hop)))
timed("prepare (stft, init, encoder)", lambda: enh.prepare(None, None, seed=1, uploaded=up))
timed("encode_init", lambda: E.encode_init(b, vae))
