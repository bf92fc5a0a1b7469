"""Smallest end-to-end case for compute-sanitizer: one short utterance, both chain kernels, both M-step
variants, Wiener, STFT/ISTFT, metrics.  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "guided-vae-nmf_b200")):
    sys.path.insert(0, p_)
from gvn import engine as E  # noqa: E402
from gvn.pipeline import McemConfig, Enhancer  # noqa: E402
from gvn.synth import synth_batch  # noqa: E402
from python.models.models import DeepGenerativeModel  # noqa: E402

torch.manual_seed(0)
F, L = 513, 16
vae = DeepGenerativeModel([F, 1, L, [128, 128]], None).eval()
x, s, n = synth_batch(2, seed=1, T=9000)            # 2 utterances x 37 frames (ragged padding inside a 128-frame tile)
rs = np.random.RandomState(0)
for prec in ("f16", "fp32"):
    for variant in (1, 0):
        cfg = McemConfig(model="M2", niter=2, nsamples_E_step=10, burnin_E_step=3, nsamples_WF=4, burnin_WF=3, nmf_rank=10,
                         precision=prec, mstep_variant=variant)
        enh = Enhancer(vae, cfg, "cuda:0")
        labels = [(rs.rand(1, 37) > 0.4).astype(np.uint8) for _ in range(2)]
        for out in enh.enhance_many([dict(wavs=list(x), labels=labels, refs=(s, n))] * 2, seed=3):
            assert np.isfinite(out["s_hat"].numpy()).all() and np.isfinite(out["metrics"].numpy()).all()
        print(prec, variant, "ok", out["metrics"].numpy()[:, 0])
