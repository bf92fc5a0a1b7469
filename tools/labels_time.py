"""Times the device label path (clean-speech STFT + gvn_speech_labels) at the benchmark shape: python tools/labels_time.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "guided-vae-nmf_b200")):
    sys.path.insert(0, p_)
from gvn import engine as E  # noqa: E402
from gvn.pipeline import McemConfig, Enhancer  # noqa: E402
from gvn.synth import synth_batch  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vae = bench.build_model()
enh = Enhancer(vae, McemConfig(model="M2", niter=1, precision="f16"), "cuda:0", label_source="oracle_ibm")
x, s, nz = synth_batch(B, seed=0, T=64000)
up = enh.upload(list(x), clean=list(s))
b = enh.prepare(None, None, seed=0, uploaded=up)
S = torch.zeros(b.F, b.NP, 2, device="cuda")
P2 = torch.zeros(b.F, b.NP, device="cuda")
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for vad in (False, True):
    for it in range(3):
        t0 = ev()
        E.stft_to(b, up["clean"], up["T"], up["T_stride"], 1024, 256, [g[2] for g in up["geo"]], S, P2)
        t1 = ev()
        y = E.speech_labels(b, S, vad, 0.999, 0.999)
        t2 = ev()
        torch.cuda.synchronize()
    print("vad=%d  clean STFT %.3f ms   labels %.3f ms   (%d utterances, %d frames)" % (vad, t0.elapsed_time(t1), t1.elapsed_time(t2), B, b.NP))
