"""Shared helpers for the tests: golden loading and oracle drivers (test-side only)."""
import os

import numpy as np
import torch

from oracle.mcem_oracle import McemNoNmfOracle, McemOracle, NoiseTape, split_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN, "mcem_%s.npz" % tag), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_state_dict(g):
    return {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd_")}


def golden_tape(g):
    """Rebuilds the draw list in consumption order: rand(F,K), rand(K,N), then per MH
    step randn(L,N), rand(N)  (SURVEY.md section 8a row R0)."""
    draws = [("rand", torch.from_numpy(g["rand_W"])), ("rand", torch.from_numpy(g["rand_H"]))]
    for e, u in zip(g["tape_eps"], g["tape_u"]):
        draws.append(("randn", torch.from_numpy(e)))
        draws.append(("rand", torch.from_numpy(u)))
    return draws


def oracle_from_golden(g, dtype=torch.float32):
    nE, bE, nW, bW = [int(v) for v in g["chain"]]
    o = McemOracle(int(g["niter"]), nE, bE, nW, bW, float(g["var_RW"]), model=str(g["model"]), dtype=dtype)
    sd = golden_state_dict(g)
    y = None if g["y"].shape[1] == 0 else torch.from_numpy(g["y"])
    o.init_parameters(g["X"], y, split_state_dict(sd, "decoder"), split_state_dict(sd, "encoder"),
                      int(g["K"]), float(g["eps"]), NoiseTape(draws=golden_tape(g), dtype=dtype))
    return o


def nonmf_tape(g):
    draws = []
    for e, u in zip(g["tape_eps"], g["tape_u"]):
        draws.append(("randn", torch.from_numpy(e)))
        draws.append(("rand", torch.from_numpy(u)))
    return draws


def nonmf_oracle_from_golden(g, dtype=torch.float32):
    """MCEM_M2_noNMF restatement fed with the golden inputs (tests/golden/mcem_M2_noNMF.npz)."""
    nE, bE, nW, bW = [int(v) for v in g["chain"]]
    sd = golden_state_dict(g)
    return McemNoNmfOracle(g["X"], g["Vb"], torch.from_numpy(g["g0"]), torch.from_numpy(g["Z0"]), torch.from_numpy(g["y"]),
                           split_state_dict(sd, "decoder"), int(g["niter"]), NoiseTape(draws=nonmf_tape(g), dtype=dtype),
                           nE, bE, nW, bW, float(g["var_RW"]), dtype=dtype)
