"""Hardware probe of the tcgen05 conventions (run on the B200 box): prints the error of
gvn_selftest_umma for every convention variant and a few shapes."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "guided-vae-nmf_b200"))
from gvn import _lib  # noqa: E402

lib = _lib.load()
torch.manual_seed(0)
for (N, K) in [(128, 16), (128, 128), (176, 128), (16, 128), (256, 64)]:
    A = torch.randn(128, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * 0.1
    ref16 = (A.half().float() @ W.half().float().T)
    ref32 = A.double() @ W.double().T
    for variant in range(8):
        D = torch.full((128, N), float("nan"), device="cuda")
        rc = lib.gvn_selftest_umma(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), N, K, variant,
                                   C.c_void_p(D.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        try:
            torch.cuda.synchronize()
            e16 = float((D - ref16).abs().max())
            e32 = float((D.double() - ref32).abs().max())
            print("N=%3d K=%3d variant=%d rc=%d  max|D-ref_f16|=%.3e  max|D-ref_f64|=%.3e  scale=%.2f"
                  % (N, K, variant, rc, e16, e32, float(ref32.abs().max())), flush=True)
        except Exception as ex:  # a trap poisons the context: stop here
            print("N=%d K=%d variant=%d FAILED: %s" % (N, K, variant, str(ex)[:200]), flush=True)
            sys.exit(1)
