"""STFT / ISTFT / dense-layer kernels against the oracle, the reference's own fixture and
torch fp32 (dense is a floating-point GEMM kernel)."""
import os

import numpy as np
import pytest
import torch

from oracle import stft_oracle
from _util import GOLDEN

pytestmark = pytest.mark.gpu
KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25)


def test_stft_matches_reference_fixture():
    from python.processing.stft import stft
    z = np.load(os.path.join(GOLDEN, "stft_wsj0_slice.npz"))
    S = stft(z["x"], dtype="complex64", **KW)
    assert S.dtype == np.complex64 and S.shape[0] == 513
    k = int(z["n_keep"])
    P = np.abs(S[:, :k]) ** 2
    # fp32 FFT: error relative to the strongest bin (the fixture came from a float64 FFT)
    assert np.max(np.abs(P - z["power"])) <= 2e-5 * np.max(z["power"])


@pytest.mark.parametrize("T", [64000, 6144, 20000, 9001, 777])
def test_stft_istft_against_oracle_and_roundtrip(T):
    from python.processing.stft import stft, istft
    x = np.random.RandomState(T).randn(T) * 0.1
    S = stft(x, dtype="complex64", **KW)
    So = stft_oracle.stft(x, dtype="complex64", **KW)
    assert S.shape == So.shape                                   # incl. the end-pad rule (stft.py:48-53)
    assert np.max(np.abs(S - So)) <= 2e-5 * np.max(np.abs(So))
    xh = istft(S, max_len=T, **KW)
    xo = stft_oracle.istft(So, max_len=T, **KW)
    assert xh.dtype == np.float32 and xh.shape == (T,)
    assert np.max(np.abs(xh - xo)) <= 2e-5
    # property of the reference's tests/processing/test_stft.py:10-50 (within fp32 FFT accuracy)
    np.testing.assert_array_almost_equal(x, xh, decimal=5)


def test_stft_errors_match_reference():
    from python.processing.stft import stft, istft
    with pytest.raises(ValueError, match="not an integer"):
        stft(np.zeros(4000), fs=16000, wlen_sec=50.01e-3)
    with pytest.raises(ValueError, match="not an integer"):
        istft(np.zeros((513, 4), np.complex64), fs=16000, wlen_sec=50.01e-3)
    from gvn._lib import GvnError
    with pytest.raises(GvnError):                                # 80 ms -> 1280 samples: not a power of two
        stft(np.zeros(16000), fs=16000, wlen_sec=80e-3)


def test_stft_ragged_batch():
    from gvn import engine as E
    Ts = [6144, 9001, 4000]
    rs = np.random.RandomState(3)
    wavs = [rs.randn(T) * 0.1 for T in Ts]
    geo = [E.stft_geometry(T, 16000, 64e-3, 0.25) for T in Ts]
    b = E.Batch([g[3] for g in geo], 513, 1, 1, 1, "cuda:0")
    wav, T, Ts_ = E.upload_waveforms(wavs, b.device)
    E.stft_into(b, wav, T, Ts_, 1024, 256, [g[2] for g in geo])
    for i, w in enumerate(wavs):
        So = stft_oracle.stft(w, dtype="complex64", **KW)
        S = b.Xc[:, b.cols(i), :].cpu().numpy()
        S = S[..., 0] + 1j * S[..., 1]
        assert np.max(np.abs(S - So)) <= 2e-5 * np.max(np.abs(So))
        X2 = b.X2[:, b.cols(i)].cpu().numpy()
        np.testing.assert_allclose(X2, np.abs(So) ** 2, rtol=1e-3, atol=2e-5 * np.max(np.abs(So)) ** 2)
    out = E.istft_from(b, b.Xc, Ts, Ts_, 1024, 256).cpu().numpy()
    for i, w in enumerate(wavs):
        np.testing.assert_array_almost_equal(out[i, :Ts[i]], w, decimal=5)
        assert np.all(out[i, Ts[i]:] == 0)


@pytest.mark.parametrize("act", ["none", "tanh", "relu", "sigmoid"])
def test_dense_matches_torch_fp32(act):
    from gvn import engine as E
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    D0, D1, Dout, NP = 513, 7, 128, 96
    W = torch.randn(Dout, D0 + D1, device=dev) * 0.05
    bias = torch.randn(Dout, device=dev)
    a, c = torch.randn(D0, NP, device=dev), torch.randn(D1, NP, device=dev)
    mean, std = torch.randn(D0, device=dev), torch.rand(D0, device=dev) + 0.5
    out = E.dense(W, bias, a, c, act, NP, mean, std, 1e-3)
    x = torch.cat([(a - mean[:, None]) / (std[:, None] + 1e-3), c], 0)
    ref = W @ x + bias[:, None]
    ref = {"none": lambda v: v, "tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[act](ref)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_classifier_hard_labels():
    from gvn import engine as E
    from python.models.models import Classifier
    torch.manual_seed(0)
    clf = Classifier([513, [128, 128], 1]).eval()
    b = E.Batch([40], 513, 1, 1, 1, "cuda:0")
    b.X2[:, b.cols(0)] = torch.rand(513, 40, device=b.device) * 3
    y = E.classify(b, clf, torch.zeros(513, 1), torch.ones(513, 1), 1e-8)
    with torch.no_grad():
        ref = (clf(b.X2[:, b.cols(0)].T.cpu() / (1 + 1e-8)) > 0.5).float().T
    assert y.shape == (1, b.NP)
    np.testing.assert_array_equal(y[:, b.cols(0)].cpu().numpy(), ref.numpy())


def test_energy_ratios_match_reference_metrics():
    """gvn_energy_ratios against python/metrics.py:12-60 (float64 numpy) on ragged synthetic signals."""
    from gvn import engine as E
    from gvn.synth import synth_utterance
    from python.metrics import energy_ratios
    Ts = [64000, 50001, 777]
    stride = 64000
    rs = np.random.RandomState(3)
    est = np.zeros((3, stride), np.float32); s = np.zeros_like(est); n = np.zeros_like(est)
    for i, T in enumerate(Ts):
        x, sp, no = synth_utterance(i, seed=9, T=T)
        s[i, :T], n[i, :T] = sp, no
        est[i, :T] = 0.7 * sp + 0.2 * no + 0.05 * rs.randn(T)
    out = E.energy_ratios(torch.from_numpy(est).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(n).cuda(), Ts).cpu().numpy()
    for i, T in enumerate(Ts):
        ref = energy_ratios(est[i, :T].astype(np.float64), s[i, :T].astype(np.float64), n[i, :T].astype(np.float64))
        np.testing.assert_allclose(out[i], ref, rtol=1e-9, atol=1e-9)
