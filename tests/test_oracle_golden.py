"""The oracle (oracle/*.py) against the golden vectors produced by the reference itself
(oracle/make_golden.py) and against the reference's own STFT fixture."""
import os

import numpy as np
import pytest
import torch

from oracle import stft_oracle
from _util import GOLDEN, load_golden, nonmf_oracle_from_golden, oracle_from_golden

STFT_KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, dtype="complex64")


def test_stft_matches_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "stft_wsj0_slice.npz"))
    S = stft_oracle.stft(z["x"], **STFT_KW)
    assert S.dtype == np.complex64 and S.shape[0] == 513
    k = int(z["n_keep"])
    P = np.abs(S[:, :k]) ** 2
    # the fixture is float32 |.|^2 of a complex64 STFT: allow a few ulp of the frame maximum
    assert np.max(np.abs(P - z["power"])) <= 2e-6 * np.max(z["power"])


@pytest.mark.parametrize("T", [64000, 6144, 20000, 777])
def test_stft_istft_roundtrip(T):
    # property of the reference's tests/processing/test_stft.py:10-50 (6 decimals)
    x = np.random.RandomState(T).randn(T) * 0.1
    S = stft_oracle.stft(x, **STFT_KW)
    nfr = 1 + (T + (256 if stft_oracle.needs_end_pad(T, 16000, 64e-3, 0.25) else 0)) // 256
    assert S.shape == (513, nfr)
    xh = stft_oracle.istft(S, fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, max_len=T)
    assert xh.dtype == np.float32 and xh.shape == (T,)
    np.testing.assert_array_almost_equal(x, xh, decimal=6)


def test_stft_rejects_fractional_window():
    with pytest.raises(ValueError, match="not an integer"):
        stft_oracle.stft(np.zeros(100), fs=16000, wlen_sec=50.01e-3)


@pytest.mark.parametrize("tag", ["M1", "M2_ibm", "M2_vad"])
def test_oracle_reproduces_reference_run(tag):
    g = load_golden(tag)
    o = oracle_from_golden(g)
    for k in ("W", "H", "g", "Z", "Vb"):
        np.testing.assert_array_equal(getattr(o, k).numpy(), g["init_" + k], err_msg="init " + k)
    seen = {}

    def hook(oo, n):
        seen[n] = {k: getattr(oo, k).numpy().copy() for k in ("W", "H", "g", "Z", "Vb")}
        if n == 0:
            seen["Vs0"] = oo.Vs.numpy().copy()
    o.iter_hook = hook
    cost = o.run()
    # same torch ops in the same order on the same draws: bit-for-bit in fp32
    np.testing.assert_array_equal(cost, g["cost"])
    for n in range(int(g["niter"])):
        for k in ("W", "H", "g", "Z", "Vb"):
            np.testing.assert_array_equal(seen[n][k], g["M%d_%s" % (n, k)], err_msg="iter %d %s" % (n, k))
    np.testing.assert_array_equal(seen["Vs0"], g["Vs_E0"])
    np.testing.assert_array_equal(o.S_hat, g["S_hat"])
    np.testing.assert_array_equal(o.N_hat, g["N_hat"])


def test_m1_quirk_chain_lengths():
    # mcem.py:461-462 / :477-478 shift the positional arguments
    g = load_golden("M1")
    o = oracle_from_golden(g)
    (RE, bE), (RW, bW) = o.chain_lengths()
    nE, bEc, nW, bWc = [int(v) for v in g["chain"]]
    assert (RE, bE, RW, bW) == (bEc, 30, bWc, 30)
    assert g["Vs_E0"].shape[0] == bEc
    assert len(g["tape_u"]) == int(g["niter"]) * (bEc + 30) + (bWc + 30)


def test_oracle_fp64_close_to_fp32():
    g = load_golden("M2_ibm")
    o = oracle_from_golden(g, dtype=torch.float64)
    cost = o.run()
    np.testing.assert_allclose(cost, g["cost"], rtol=2e-3)


def test_nonmf_oracle_reproduces_reference_run():
    # MCEM_M2_noNMF (mcem.py:609-760): fixed Vb, gain-only M-step
    g = load_golden("M2_noNMF")
    o = nonmf_oracle_from_golden(g)
    seen = {}

    def hook(oo, n):
        seen[n] = dict(g=oo.g.numpy().copy(), Z=oo.Z.numpy().copy())
        if n == 0:
            seen["Vs0"] = oo.Vs.numpy().copy()
    o.iter_hook = hook
    cost = o.run()
    np.testing.assert_array_equal(cost, g["cost"])
    for n in range(int(g["niter"])):
        np.testing.assert_array_equal(seen[n]["g"], g["M%d_g" % n])
        np.testing.assert_array_equal(seen[n]["Z"], g["E%d_Z" % n])
    np.testing.assert_array_equal(seen["Vs0"], g["Vs_E0"])
    np.testing.assert_array_equal(o.S_hat, g["S_hat"])
    np.testing.assert_array_equal(o.N_hat, g["N_hat"])


def test_spp_oracle_reproduces_reference_mask():
    # timo_mask_estimation (python/models/spp_estimation.py:198-218) run by oracle/make_golden.py
    from oracle import spp_oracle
    z = np.load(os.path.join(GOLDEN, "spp_mask.npz"))
    m = spp_oracle.timo_mask(z["power"])
    assert m.dtype == np.float32
    np.testing.assert_array_equal(m, z["mask"])
    assert 0.01 < (m > 0.5).mean() < 0.5


def test_oracle_reproduces_reference_on_real_wsj0_slice():
    # mixture = speech + noise from the reference's own fixture, enhanced by the unmodified reference (make_golden.golden_real)
    from oracle.mcem_oracle import energy_ratios
    g = load_golden("M2_vad_wsj0")
    np.testing.assert_allclose(g["x"], g["s"] + g["n"], atol=1e-12)
    o = oracle_from_golden(g)
    cost = o.run()
    np.testing.assert_array_equal(cost, g["cost"])
    np.testing.assert_array_equal(o.S_hat, g["S_hat"])
    np.testing.assert_array_equal(o.N_hat, g["N_hat"])
    s_hat = stft_oracle.istft(o.S_hat, fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, max_len=len(g["x"]))
    r = np.array(energy_ratios(s_hat.astype(np.float64), g["s"], g["n"]))
    np.testing.assert_allclose(r, g["ratios"], rtol=0, atol=1e-9)            # python/metrics.py of the reference
    assert r[0] > energy_ratios(g["x"], g["s"], g["n"])[0] + 2.0              # the enhancement does enhance (+3.1 dB SI-SDR)


@pytest.mark.parametrize("q", [0.999, 0.98])
def test_label_oracle_reproduces_reference_masks(q):
    """tests/golden/labels.npz: clean_speech_IBM / clean_speech_VAD of the reference's python/processing/target.py:7-50 run
    unmodified (oracle/make_golden.py golden_labels); discrete outputs, so exact.  Also the reference's own assertions
    (tests/processing/test_target.py:49-50: float32, values {0, 1})."""
    from oracle.mcem_oracle import clean_speech_IBM, clean_speech_VAD
    z = np.load(os.path.join(GOLDEN, "labels.npz"))
    S = z["S"]
    assert S.dtype == np.complex64
    ibm, vad = clean_speech_IBM(S, q, 0.999), clean_speech_VAD(S, q, 0.999)
    assert ibm.dtype == np.float32 and np.unique(ibm).tolist() == [0.0, 1.0]
    assert vad.dtype == np.float32 and vad.shape == (1, S.shape[1])
    np.testing.assert_array_equal(ibm, z["ibm_%d" % round(q * 1000)])
    np.testing.assert_array_equal(vad, z["vad_%d" % round(q * 1000)])


def test_label_oracle_reproduces_the_references_own_label_fixture():
    """tests/golden/labels_wsj0.npz: the reference's COMMITTED label pickles (data/subset/pickle/CSR-1-WSJ-0/si_dt_05_labels.p and
    si_dt_05_vad_labels.p, written by its tests/dataset/test_csr1_wjs0_dataset.py with clean_speech_IBM / clean_speech_VAD at
    quantile 0.98) next to the power spectrograms they were made from (si_dt_05_frames.p): three WSJ0 utterances, 513 x 976 bins.
    The threshold rule of the oracle reproduces every label, per utterance."""
    from oracle.mcem_oracle import lorenz_mask
    z = np.load(os.path.join(GOLDEN, "labels_wsj0.npz"))
    P, bounds = z["power"], z["bounds"]
    ibm = np.unpackbits(z["ibm"], axis=1)[:, :P.shape[1]].astype(np.float32)
    vad = z["vad"].astype(np.float32)
    assert P.shape == (513, 976) and bounds.tolist() == [0, 341, 640, 976] and 0.05 < ibm.mean() < 0.15
    for a, b in zip(bounds[:-1], bounds[1:]):
        np.testing.assert_array_equal(lorenz_mask(P[:, a:b], 0.98, 0.999), ibm[:, a:b])
        np.testing.assert_array_equal(lorenz_mask(P[:, a:b].sum(axis=0), 0.98, 0.999)[None], vad[:, a:b])
    # the rule is per utterance: over the concatenation it gives different labels
    assert not np.array_equal(lorenz_mask(P, 0.98, 0.999), ibm)
