"""Wall time of the phases of one bench step (synchronising after each phase)."""
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
wavs = list(x)
def t():
    torch.cuda.synchronize()
    return time.perf_counter()
for rep in range(3):
    t0 = t(); up = enh.upload(wavs, labels)
    t1 = t(); b = enh.prepare(None, None, seed=rep, uploaded=up)
    t2 = t(); cost, S, Nn, _, _ = E.run_mcem(b, enh.dec, cfg.niter, *cfg.chains(), cfg.var_RW, "f16", seed=rep)
    t3 = t(); s_hat = E.istft_from(b, S, b.T, b.T_stride, b.nfft, b.hop); n_hat = E.istft_from(b, Nn, b.T, b.T_stride, b.nfft, b.hop)
    t4 = t(); out = (s_hat.cpu(), n_hat.cpu(), cost[-1].cpu())
    t5 = t()
    print("rep %d: upload %.2f | prepare %.2f | mcem %.2f | istft %.2f | d2h %.2f ms" % (rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); b = enh.prepare(None, None, seed=9, uploaded=up); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
