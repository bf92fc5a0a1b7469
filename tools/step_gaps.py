"""Where does a step's time go?  Device time of run_mcem (events around the whole loop) with and
without the per-kernel event brackets, and the host time needed to enqueue it."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "guided-vae-nmf_b200")):
    sys.path.insert(0, p_)
from gvn import engine as E  # noqa: E402
from gvn.pipeline import McemConfig, Enhancer  # noqa: E402
import bench  # noqa: E402
vae = bench.build_model()
cfg = McemConfig(model="M2", niter=100, nmf_rank=10, precision="f16")
enh = Enhancer(vae, cfg, "cuda:0")
x, s, nz, labels = bench.make_inputs(64, 0)
up = enh.upload(list(x), labels)
for use_timers in (False, True, False):
    b = enh.prepare(None, None, seed=1, uploaded=up)
    torch.cuda.synchronize()
    tm = E.KernelTimers() if use_timers else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record()
    out = E.run_mcem(b, enh.dec, cfg.niter, *cfg.chains(), cfg.var_RW, "f16", seed=3, timers=tm)
    e1.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    h2 = time.perf_counter()
    msg = "timers=%s: device %.2f ms, host enqueue %.2f ms, host until done %.2f ms" % (use_timers, e0.elapsed_time(e1), (h1 - h0) * 1e3, (h2 - h0) * 1e3)
    if tm is not None:
        msg += " | estep %.2f ms, mstep %.2f ms" % (tm.total_ms("estep"), tm.total_ms("mstep"))
    print(msg)
