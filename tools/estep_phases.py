"""Phase breakdown of the tensor-core chain kernel from its in-kernel cycle counters (run on the
B200 box): python tools/estep_phases.py [batch]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "guided-vae-nmf_b200")):
    sys.path.insert(0, p_)
from gvn import _lib, engine as E  # noqa: E402
from gvn.pipeline import McemConfig, Enhancer  # noqa: E402
from gvn.synth import synth_batch  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _lib.load()
vae = bench.build_model()
cfg = McemConfig(model="M2", niter=2, nmf_rank=10, precision="f16")
enh = Enhancer(vae, cfg, "cuda:0")
x, s, nz, labels = bench.make_inputs(B, 0)
b = enh.prepare(list(x), labels, seed=0)
tiles = (b.NP + 127) // 128
buf = torch.zeros(tiles, 10, 16, dtype=torch.int64, device="cuda")    # [tile][warp][counter], as the kernel indexes it
for it in range(3):
    E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=it)
torch.cuda.synchronize()
lib.gvn_debug_profile_buffer(C.c_void_p(buf.data_ptr()))
E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=5)             # first launch of the counting instantiation (attribute, load)
torch.cuda.synchronize()
buf.zero_()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
E.estep(b, enh.dec, 30, 10, 0.01, "f16", seed=1, chain=7)
t1.record()
torch.cuda.synchronize()
lib.gvn_debug_profile_buffer(C.c_void_p(0))
ms = t0.elapsed_time(t1)
a = buf.cpu().numpy().astype(np.float64)[:, :8, :12]          # epilogue warps only
names = ["pre-hidden (put_z, noise, accept tail)", "wait L1 acc", "hidden0 math", "wait L2 acc", "hidden1 math", "wait chunk acc",
         "wait ring", "stage math", "energy barrier", "accept/other", "-", "-"]
tot = a.sum(-1).mean()
print("launch %.3f ms, %d tiles; mean cycles per epilogue warp %.0f (%.3f ms at 1.965 GHz)" % (ms, tiles, tot, tot / 1.965e6))
for half, nm in ((slice(0, 4), "owner warps 0-3"), (slice(4, 8), "noise warps 4-7")):
    m = a[:, half, :].mean((0, 1))
    print(nm)
    for i, nme in enumerate(names[:10]):
        print("   %-42s %9.0f cyc  %5.1f%%" % (nme, m[i], 100 * m[i] / m.sum()))
