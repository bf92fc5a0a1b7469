"""gvn_spp_mask ("timo" guide labels) against the golden mask of the reference's timo_mask_estimation and the
numpy restatement."""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN
from oracle import spp_oracle

pytestmark = pytest.mark.gpu


def test_dropin_matches_reference_golden():
    from python.models.spp_estimation import timo_mask_estimation
    z = np.load(os.path.join(GOLDEN, "spp_mask.npz"))
    m = timo_mask_estimation(z["power"])
    assert m.shape == z["mask"].shape and m.dtype == np.float32
    np.testing.assert_allclose(m, z["mask"], rtol=1e-5, atol=1e-7)
    away = np.abs(z["mask"] - 0.5) > 1e-5
    assert np.array_equal((m > 0.5)[away], (z["mask"] > 0.5)[away])
    assert np.all(m[:, :10] == 0)                                 # spp_estimation.py:106 -- SPP is 0 while initialising


def test_ragged_batch_equals_oracle_per_utterance():
    from gvn import engine as E
    rs = np.random.RandomState(0)
    Ns, F = [7, 40, 33, 100], 70                                  # shorter than the init phase, chunk edges, ragged F
    P = [(rs.rand(F, n) ** 4 * 10).astype(np.float32) for n in Ns]
    b = E.Batch(Ns, F, 1, 1, 1, "cuda:0", with_complex=False)
    b.scatter_cols(b.X2, [torch.from_numpy(p) for p in P])
    soft, hard = E.spp_mask(b)
    for i, p in enumerate(P):
        ref = spp_oracle.timo_mask(p)
        got = soft[:, b.cols(i)].cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-7)
        h = hard[:, b.cols(i)].cpu().numpy()
        away = np.abs(ref - 0.5) > 1e-5
        assert np.array_equal(h[away], (ref > 0.5)[away].astype(np.float32))
    # padding frames stay zero
    pad = (b.frame_utt < 0).cpu().numpy()
    assert float(hard[:, torch.from_numpy(pad).cuda()].abs().sum()) == 0.0


def test_enhancer_with_timo_labels():
    from gvn.pipeline import Enhancer, McemConfig
    from gvn.synth import synth_utterance
    from python.models.models import DeepGenerativeModel
    torch.manual_seed(0)
    vae = DeepGenerativeModel([513, 513, 16, [128, 128]], None).eval()
    cfg = McemConfig(model="M2", niter=1, nsamples_E_step=2, burnin_E_step=2, nsamples_WF=2, burnin_WF=2, precision="f16")
    enh = Enhancer(vae, cfg, "cuda:0", label_source="timo")
    x = [synth_utterance(i, seed=1, T=8000)[0] for i in range(2)]
    s_hat, n_hat, cost = enh.enhance(x, None, seed=0)
    assert np.isfinite(cost).all() and all(np.isfinite(s).all() for s in s_hat)
    np.testing.assert_allclose(s_hat[0] + n_hat[0], x[0], atol=1e-4)
