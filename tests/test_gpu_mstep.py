"""M-step variant 1 (column tiles staged once in shared memory, pair-wise reciprocal/log) against
variant 0 (the straightforward schedule: IEEE divisions, logf) on ragged batches at the benchmark shape,
including tiles that lie entirely in padding and slots of multiplicity zero.  Both variants are compared
with the ORACLE in tests/test_gpu_parity.py (golden shapes, variant 1) and tests/test_gpu_fullshape.py
(benchmark shapes, variants 0 and 1); this file checks that they agree with each other on many more shapes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _random_batch(n_frames, F, K, L, R, seed):
    from gvn import engine as E
    g = torch.Generator(device="cuda").manual_seed(seed)
    b = E.Batch(n_frames, F, K, L, R, "cuda:0", with_complex=False)
    b.X2.copy_(torch.rand(F, b.NP, generator=g, device="cuda") * 3 + 1e-3)
    E.init_nmf(b, 1e-8, generator=g)                       # also builds the column-tile copy of X2
    vs = torch.exp(torch.randn(R, F, b.NP, generator=g, device="cuda") * 2 - 1)
    b.g.copy_(torch.rand(b.NP, generator=g, device="cuda") + 0.5)
    b.H.copy_(torch.rand(K, b.NP, generator=g, device="cuda") + 1e-3)
    b.W.copy_(torch.rand(b.B, F, K, generator=g, device="cuda") + 1e-3)
    for i in range(b.B):                                   # Vb is the product W @ H, as after any M-step
        b.Vb[:, b.cols(i)] = b.W[i] @ b.H[:, b.cols(i)]
    # slot multiplicities (include/gvn.h): slot 0 always live, some slots dead (rejected proposals)
    w = torch.randint(0, 3, (R, b.NP), generator=g, device="cuda").float()
    w[0] += 1
    b.Vs_w[:R].copy_(w)
    # poison the padding columns of Vs: they must never be read into a result
    pad = b.frame_utt < 0
    vs[:, :, pad] = float("nan")
    b.set_samples(vs, b.Vs_w[:R].clone())
    return b


@pytest.mark.parametrize("n_frames,F,K,R", [([251, 100, 37, 256], 513, 10, 10), ([64], 513, 12, 9), ([40, 33], 257, 4, 3),
                                            ([251] * 6, 513, 10, 10), ([70], 129, 1, 2),
                                            # shapes of the generic column sweep: K = 32 (config 4), R = 30 (MCEM_M1's E chain), odd sizes
                                            ([251, 90], 513, 32, 10), ([251], 513, 10, 30), ([100, 37], 257, 20, 5), ([64], 513, 16, 10),
                                            # few long utterances: the W sweep splits the frame axis across CTAs (ragged split included)
                                            ([1876], 513, 32, 10), ([1000, 640], 513, 10, 10), ([251], 513, 10, 10), ([700, 90, 333], 129, 7, 4)])
def test_mstep_v1_matches_v0(n_frames, F, K, R):
    from gvn import engine as E
    out = {}
    for variant in (0, 1):
        b = _random_batch(n_frames, F, K, 16, R, seed=3)
        sc = E.MstepScratch(b, 2)
        E.mstep(b, R, sc, 0, variant)
        E.mstep(b, R, sc, 1, variant)                 # second iteration uses the normalised W and new H, g, Vb
        cost = E.cost_reduce(b, R, sc, 2)
        torch.cuda.synchronize()
        valid = (b.frame_utt >= 0).cpu().numpy()
        out[variant] = dict(W=b.W.cpu().numpy(), H=b.H.cpu().numpy()[:, valid], g=b.g.cpu().numpy()[valid],
                            Vb=b.Vb.cpu().numpy()[:, valid], cost=cost.cpu().numpy())
    for k in ("W", "H", "g", "Vb", "cost"):
        assert np.all(np.isfinite(out[1][k])), k
        np.testing.assert_allclose(out[1][k], out[0][k], rtol=3e-5, err_msg=k)
    np.testing.assert_allclose(np.abs(out[1]["W"]).sum(1), 1.0, rtol=1e-5)          # mcem.py:128-131


def test_mstep_v1_falls_back_when_nothing_fits():
    from gvn import engine as E
    b = _random_batch([64], 513, 10, 16, 120, seed=1)                                # R=120: no ring stage fits, variant 0 runs
    sc = E.MstepScratch(b, 1)
    E.mstep(b, 120, sc, 0, 1)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(b.W).all())
