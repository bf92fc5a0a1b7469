"""Where the end-to-end step goes: host time stamps + CUDA events around the phases of
Enhancer.enhance_many at the bench shape (C2).  Usage: python tools/e2e_trace.py [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "guided-vae-nmf_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from gvn.pipeline import McemConfig, Enhancer  # noqa: E402
from gvn import engine as E  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = McemConfig(model="M2", niter=100, nmf_rank=10, precision="f16", mstep_variant=1)
enh = Enhancer(bench.build_model(), cfg, dev)
x, s, nz, labels = bench.make_inputs(64, first=0)
wavs = list(x)

log = []
t_origin = [0.0]


def wrap(obj, name, label=None):
    fn = getattr(obj, name)

    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        log.append((label or name, (t0 - t_origin[0]) * 1e3, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(obj, name, w)


wrap(enh, "upload")
wrap(enh, "prepare")
wrap(enh, "run")
wrap(E, "download")
wrap(E, "energy_ratios")


def batches(k):
    for _ in range(k):
        yield dict(wavs=wavs, labels=labels, refs=(s, nz))


for out in enh.enhance_many(batches(2), seed=1):
    pass
torch.cuda.synchronize()
log.clear()
t_origin[0] = time.perf_counter()
marks = []
for out in enh.enhance_many(batches(steps), seed=5):
    marks.append((time.perf_counter() - t_origin[0]) * 1e3)
torch.cuda.synchronize()
total = (time.perf_counter() - t_origin[0]) * 1e3
print("total %.1f ms for %d steps = %.1f ms/step (%.0f utt/s)" % (total, steps, total / steps, 64 * steps / total * 1e3))
print("yield times:", ["%.1f" % m for m in marks])
for name, t0, dt in log:
    print("%-14s start %8.2f ms  host %7.2f ms" % (name, t0, dt))

# device-only step for comparison
up = enh.upload(wavs, labels, refs=(s, nz), slot=2)
torch.cuda.synchronize()
for rep in range(2):
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    a.record()
    bt = enh.prepare(None, None, seed=3, uploaded=up)
    h1 = time.perf_counter()
    r = enh.run(bt, seed=3)
    h2 = time.perf_counter()
    b_.record()
    torch.cuda.synchronize()
    print("device step %.2f ms; host: prepare %.2f ms, run (launch queueing) %.2f ms" % (a.elapsed_time(b_), (h1 - h0) * 1e3, (h2 - h1) * 1e3))
