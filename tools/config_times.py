"""Timings of the other BASELINE configs on one B200 (not bench lines): C1 = one 4 s utterance, M1;
C3 shape = M2-VAD with classifier labels, 64 utterances; C4 = M1, 30 s, K=32, 10 kept samples, 8 utterances."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "guided-vae-nmf_b200")):
    sys.path.insert(0, p_)
from gvn.pipeline import McemConfig, Enhancer  # noqa: E402
from gvn.synth import synth_batch  # noqa: E402
from python.models.models import DeepGenerativeModel, VariationalAutoencoder, Classifier  # noqa: E402

def timeit(enh, wavs, labels=None, reps=3):
    for _ in range(2):
        enh.enhance(wavs, labels, seed=1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(reps):
        enh.enhance(wavs, labels, seed=2 + i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

torch.manual_seed(0)
m1 = VariationalAutoencoder([513, 16, [128, 128]]).eval()
x, s, n = synth_batch(1, seed=0, T=64000)
for prec in ("f16", "fp32"):
    cfg = McemConfig(model="M1", niter=100, precision=prec)          # M1 quirk: R_E = 30, burn-in 30; Wiener R = 75
    t = timeit(Enhancer(m1, cfg, "cuda:0"), list(x))
    print("C1  M1, 1 utterance x 4 s, K=10, niter=100, %s: %.1f ms per utterance (%.1f utt/s)" % (prec, t * 1e3, 1 / t))
x, s, n = synth_batch(64, seed=0, T=64000)
vad = DeepGenerativeModel([513, 1, 16, [128, 128]], None).eval()
clf = Classifier([513, [128, 128], 1]).eval()
cfg = McemConfig(model="M2", niter=100, precision="f16")
t = timeit(Enhancer(vad, cfg, "cuda:0", classifier=clf, mean=np.zeros((513, 1), np.float32), std=np.ones((513, 1), np.float32)), list(x))
print("C3  M2-VAD, labels from the classifier on the device, 64 utterances x 4 s per GPU, f16: %.1f ms per batch (%.1f utt/s)" % (t * 1e3, 64 / t))
x, s, n = synth_batch(8, seed=0, T=480000)
cfg = McemConfig(model="M1", niter=100, nmf_rank=32, burnin_E_step=10, precision="f16")
t = timeit(Enhancer(m1, cfg, "cuda:0"), list(x), reps=2)
print("C4  M1, 8 utterances x 30 s (N=1876), K=32, R_E=10, niter=100, f16 (generic column sweep for K=32): %.1f ms per batch (%.2f utt/s)" % (t * 1e3, 8 / t))
