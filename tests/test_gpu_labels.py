"""Oracle guide labels on the device (gvn_speech_labels, csrc/labels.cu) against the numpy restatement of
clean_speech_IBM / clean_speech_VAD (reference python/processing/target.py:7-50).  The labels are discrete: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import mcem_oracle as O
from oracle import stft_oracle

pytestmark = pytest.mark.gpu
KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25)


def _speech_like(seed, T):
    from gvn.synth import synth_utterance
    _, s, _ = synth_utterance(seed, seed=11, T=T)
    return s


@pytest.mark.parametrize("vad", [False, True])
@pytest.mark.parametrize("qf", [0.999, 0.98])
def test_ragged_batch_from_power_is_bit_exact(vad, qf):
    """The power array numpy itself computed goes in: sort, pairwise sum, sequential cumsum, division and threshold of the
    device path must reproduce numpy's float32 results exactly, utterance by utterance."""
    from gvn import engine as E
    Ts = [64000, 20000, 33123]
    specs = [stft_oracle.stft(_speech_like(i, T), dtype="complex64", **KW) for i, T in enumerate(Ts)]
    b = E.Batch([s.shape[1] for s in specs], 513, 1, 1, 1, "cuda:0", with_complex=False)
    P = torch.full((513, b.NP), float("nan"), device="cuda")          # padding frames must never be read
    for i, s in enumerate(specs):
        P[:, b.cols(i)] = torch.from_numpy(np.abs(s * s.conj())).cuda()
    y = E.speech_labels(b, P, vad, qf, 0.999, from_power=True)
    torch.cuda.synchronize()
    fn = O.clean_speech_VAD if vad else O.clean_speech_IBM
    for i, s in enumerate(specs):
        ref = fn(s, qf, 0.999)
        got = y[:, b.cols(i)].cpu().numpy()
        assert got.dtype == np.float32 and set(np.unique(got)) <= {0.0, 1.0}     # reference tests/processing/test_target.py:49-50
        np.testing.assert_array_equal(got, ref)
        assert 0 < ref.mean() < 1
    pad = np.ones(b.NP, bool)
    for i in range(len(specs)):
        pad[b.cols(i)] = False
    assert not y[:, torch.from_numpy(pad).cuda()].any()


@pytest.mark.parametrize("T", [64000, 480000])
def test_mirror_functions_match_oracle(T):
    """python.processing.target (the drop-in module): complex64 STFT in, numpy mask out; 4 s and 30 s utterances."""
    from python.processing.target import clean_speech_IBM, clean_speech_VAD
    S = stft_oracle.stft(_speech_like(3, T), dtype="complex64", **KW)
    a, r = clean_speech_IBM(S, 0.999, 0.999), O.clean_speech_IBM(S, 0.999, 0.999)
    assert a.shape == r.shape == S.shape and a.dtype == np.float32
    np.testing.assert_array_equal(a, r)
    v, rv = clean_speech_VAD(S, 0.999, 0.999), O.clean_speech_VAD(S, 0.999, 0.999)
    assert v.shape == rv.shape == (1, S.shape[1])
    np.testing.assert_array_equal(v, rv)
    rs = np.random.RandomState(0)                                      # and a non-speech input with many near-ties
    S2 = ((rs.randn(65, 40) + 1j * rs.randn(65, 40)) * rs.rand(65, 1)).astype(np.complex64)
    np.testing.assert_array_equal(clean_speech_IBM(S2, 0.999, 0.999), O.clean_speech_IBM(S2, 0.999, 0.999))
    np.testing.assert_array_equal(clean_speech_VAD(S2, 0.98), O.clean_speech_VAD(S2, 0.98))


@pytest.mark.parametrize("q", [0.999, 0.98])
def test_mirror_functions_match_reference_golden(q):
    """The drop-in module against masks the reference's own target.py produced (tests/golden/labels.npz)."""
    import os
    from python.processing.target import clean_speech_IBM, clean_speech_VAD
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "labels.npz"))
    np.testing.assert_array_equal(clean_speech_IBM(z["S"], q, 0.999), z["ibm_%d" % round(q * 1000)])
    np.testing.assert_array_equal(clean_speech_VAD(z["S"], q, 0.999), z["vad_%d" % round(q * 1000)])


def test_enhancer_oracle_labels_on_device():
    """Enhancer(label_source='oracle_ibm'): clean speech in, STFT and labels on the device; equal to the oracle labels of
    the device's own clean-speech STFT, and the enhancement equals the one run with host-made labels."""
    import bench
    from gvn import engine as E
    from gvn.pipeline import McemConfig, Enhancer
    from gvn.synth import synth_batch
    from python.processing.stft import stft
    vae = bench.build_model()
    cfg = McemConfig(model="M2", niter=2, nmf_rank=10, precision="f16")
    x, s, nz = synth_batch(3, seed=0, T=16000)
    enh = Enhancer(vae, cfg, "cuda:0", label_source="oracle_ibm")
    up = enh.upload(list(x), clean=list(s))
    b = enh.prepare(None, None, seed=4, uploaded=up)
    torch.cuda.synchronize()
    labels = []
    for i in range(3):
        ref = O.clean_speech_IBM(stft(s[i], dtype="complex64", **KW), 0.999, 0.999)
        np.testing.assert_array_equal(b.y[:, b.cols(i)].cpu().numpy(), ref)
        labels.append(ref.astype(np.uint8))
    s1, n1, c1 = enh.run(b, seed=4)
    c1, s1 = c1.cpu().numpy(), s1.cpu().numpy()
    enh2 = Enhancer(vae, cfg, "cuda:0")
    b2 = enh2.prepare(list(x), labels, seed=4)
    s2, n2, c2 = enh2.run(b2, seed=4)
    np.testing.assert_array_equal(c1, c2.cpu().numpy())
    np.testing.assert_array_equal(s1, s2.cpu().numpy())


def test_argument_checks():
    from gvn import _lib, engine as E
    b = E.Batch([40], 513, 1, 1, 1, "cuda:0", with_complex=False)
    S = torch.zeros(513, b.NP, 2, device="cuda")
    for qf, qw in ((0.0, 0.999), (1.5, 0.999), (0.98, 0.0), (0.98, 1.5)):
        with pytest.raises(_lib.GvnError):
            E.speech_labels(b, S, False, qf, qw)


@pytest.mark.parametrize("vad", [False, True])
def test_device_labels_reproduce_the_references_own_label_fixture(vad):
    """gvn_speech_labels on the reference's committed power spectrograms (three WSJ0 utterances as one ragged batch) against the
    reference's committed label pickles (tests/golden/labels_wsj0.npz, see tests/test_oracle_golden.py): every label equal."""
    import os
    from gvn import engine as E
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "labels_wsj0.npz"))
    P, bounds = z["power"], z["bounds"]
    ref = z["vad"].astype(np.float32) if vad else np.unpackbits(z["ibm"], axis=1)[:, :P.shape[1]].astype(np.float32)
    b = E.Batch([int(n) for n in np.diff(bounds)], 513, 1, 1, 1, "cuda:0", with_complex=False)
    Pd = torch.full((513, b.NP), float("nan"), device="cuda")
    for i, (a, c) in enumerate(zip(bounds[:-1], bounds[1:])):
        Pd[:, b.cols(i)] = torch.from_numpy(P[:, a:c]).cuda()
    y = E.speech_labels(b, Pd, vad, 0.98, 0.999, from_power=True)
    torch.cuda.synchronize()
    for i, (a, c) in enumerate(zip(bounds[:-1], bounds[1:])):
        np.testing.assert_array_equal(y[:, b.cols(i)].cpu().numpy(), ref[:, a:c])
