"""Generates tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run from the repo root (the reference tree is read-only, so no bytecode is written):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

* ``stft_wsj0_slice.npz``  -- a slice of the reference's own fixture pair
  ``data/subset/pickle/CSR-1-WSJ-0/si_et_05_speech-505.p`` -> ``si_et_05_frames.p``
  (produced by the reference's ``tests/dataset/test_csr1_wjs0_dataset.py:17-83``): the first
  20 480 samples of utterance 0 after peak normalisation, and the frames of the stored power
  spectrogram whose analysis windows lie inside that slice.
* ``mcem_{M1,M2_ibm,M2_vad}.npz`` -- ``/root/reference/python/models/{mcem,models}.py``
  imported as they are; their module-level ``torch`` name is replaced by a proxy whose
  ``randn``/``rand`` read a :class:`oracle.mcem_oracle.NoiseTape`, so every random draw is
  recorded in consumption order.  The reference is driven
  through its own ``init_parameters`` / ``E_step`` / ``M_step`` /
  ``compute_expected_neg_log_like`` / ``compute_WF`` (the body of ``EM.run``,
  ``mcem.py:155-178``), the state after every step is stored, and a second instance is run
  through ``run()`` itself to check that the stepwise drive is the same computation.

* ``mcem_M2_vad_wsj0.npz`` -- a slice of the reference's own WSJ0 fixture (mixture = speech + noise) enhanced by the
  unmodified ``MCEM_M2`` with ``clean_speech_VAD`` labels, scored by the reference's ``python/metrics.py``.
* ``spp_mask.npz`` -- ``timo_mask_estimation`` of ``python/models/spp_estimation.py`` on a synthetic mixture.
* ``mcem_M2_noNMF.npz`` -- the same drive for ``MCEM_M2_noNMF`` (``mcem.py:609-760``).
* ``labels.npz`` -- ``clean_speech_IBM`` / ``clean_speech_VAD`` of ``python/processing/target.py`` (imported unmodified) on
  the complex64 STFT of a synthetic speech signal, at the scripts' quantile 0.999 and the functions' default 0.98.

* ``labels_wsj0.npz`` -- the reference's OWN committed label fixtures (``data/subset/pickle/CSR-1-WSJ-0/si_dt_05_labels.p``,
  ``si_dt_05_vad_labels.p``: ``clean_speech_IBM`` / ``clean_speech_VAD`` at quantile 0.98 per utterance, written by
  ``tests/dataset/test_csr1_wjs0_dataset.py``) next to the power spectrograms they were made from (``si_dt_05_frames.p``) and
  the utterance boundaries, which the fixture does not store: they are recovered here as the one segmentation under which the
  restated rule reproduces every one of the 513 x 976 labels.

/root/reference does not exist on the GPU box; only the committed .npz files travel.
"""
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "guided-vae-nmf_b200"))
sys.dont_write_bytecode = True

from oracle import stft_oracle                      # noqa: E402
from oracle.mcem_oracle import NoiseTape, clean_speech_IBM  # noqa: E402
from gvn.synth import synth_utterance               # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STFT_KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, dtype="complex64")


class _NumpyOnly(pickle.Unpickler):
    """The pickles are untrusted input: only numpy reconstruction is allowed."""

    def find_class(self, module, name):
        if module.split(".")[0] == "numpy":
            return super().find_class(module, name)
        raise pickle.UnpicklingError("blocked %s.%s" % (module, name))


def _load(name):
    with open(os.path.join(REF, "data/subset/pickle/CSR-1-WSJ-0", name), "rb") as f:
        return _NumpyOnly(f).load()


def golden_stft():
    speech = _load("si_et_05_speech-505.p")
    frames = _load("si_et_05_frames.p")
    s = speech[0] / np.max(np.abs(speech[0]))        # test_csr1_wjs0_dataset.py:47
    full = np.abs(stft_oracle.stft(s, **STFT_KW)) ** 2
    n0 = full.shape[1]
    err = np.max(np.abs(full - frames[:, :n0])) / np.max(frames[:, :n0])
    print("stft restatement vs reference fixture (utt 0, %d frames): max err / max = %.2e" % (n0, err))
    assert err < 1e-6
    T = 20480
    keep = (T - 512) // 256 - 1                       # frames untouched by the truncation
    np.savez_compressed(os.path.join(OUT, "stft_wsj0_slice.npz"),
                        x=s[:T], power=frames[:, :keep], n_keep=keep,
                        source="si_et_05_speech-505.p[0]/max -> si_et_05_frames.p[:, :%d]" % keep)
    # second pin: all three utterances, stored as a per-frame checksum (sum over frequency)
    col = []
    for u in speech:
        u = u / np.max(np.abs(u))
        col.append((np.abs(stft_oracle.stft(u, **STFT_KW)) ** 2).sum(0))
    col = np.concatenate(col)
    ref_col = frames.astype(np.float64).sum(0)
    print("frame-energy checksum, 3 utterances: max rel err %.2e" % np.max(np.abs(col - ref_col) / ref_col))
    assert np.max(np.abs(col - ref_col) / ref_col) < 1e-5


class _TorchProxy:
    def __init__(self, tape):
        self._tape = tape

    def __getattr__(self, name):
        return getattr(torch, name)

    def randn(self, *shape, **kw):
        return self._tape.randn(*shape)

    def rand(self, *shape, **kw):
        return self._tape.rand(*shape)


def _reference_modules():
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "python" or k.startswith("python.")]:
        del sys.modules[k]
    from python.models import mcem as ref_mcem, models as ref_models
    sys.path.remove(REF)
    return ref_mcem, ref_models


def _snap(m):
    return dict(W=m.W.numpy().copy(), H=m.H.numpy().copy(), g=m.g.numpy().copy(),
                Z=m.Z.numpy().copy(), Vb=m.Vb.numpy().copy())


def golden_mcem(tag, model, y_kind, L, K, niter, chain, T=6144, seed=0):
    ref_mcem, ref_models = _reference_modules()
    nE, bE, nW, bW = chain
    x, s, _ = synth_utterance(0, seed=seed, T=T)
    X = stft_oracle.stft(x, **STFT_KW).T                       # (N,F) as process_utt passes it
    N, F = X.shape
    torch.manual_seed(0)
    if model == "M1":
        vae = ref_models.VariationalAutoencoder([F, L, [128, 128]])
        y = None
    else:
        y_dim = F if y_kind == "ibm" else 1
        vae = ref_models.DeepGenerativeModel([F, y_dim, L, [128, 128]], None)
        if y_kind == "ibm":
            y = torch.from_numpy(clean_speech_IBM(stft_oracle.stft(s, **STFT_KW), 0.999, 0.999).T.copy())
        else:
            g = torch.Generator().manual_seed(5)
            y = (torch.rand(N, 1, generator=g) > 0.4).float()
    vae.eval()
    for p in vae.parameters():
        p.requires_grad = False
    # the default Xavier init gives near-flat spectra; perturb the biases so the decoder
    # output spans a few decades like a trained model would
    with torch.no_grad():
        vae.decoder.reconstruction.bias.copy_(torch.linspace(-6.0, -1.0, F))
        vae.decoder.hidden[0].bias.normal_(0, 0.3)
        vae.decoder.hidden[1].bias.normal_(0, 0.3)

    def make(tape):
        ref_mcem.torch = _TorchProxy(tape)
        cls = ref_mcem.MCEM_M1 if model == "M1" else ref_mcem.MCEM_M2
        m = cls(niter=niter, nsamples_E_step=nE, burnin_E_step=bE, nsamples_WF=nW,
                burnin_WF=bW, var_RW=0.01)
        if model == "M1":
            m.init_parameters(X=X, vae=vae, nmf_rank=K, eps=1e-8, device="cpu")
        else:
            m.init_parameters(X=X, y=y, vae=vae, nmf_rank=K, eps=1e-8, device="cpu")
        return m

    tape = NoiseTape(seed=1234 + seed)
    m = make(tape)
    out = {"init": _snap(m)}
    cost = []
    Vs_first = None
    for n in range(niter):
        m.E_step()
        out["E%d" % n] = dict(Z=m.Z.numpy().copy())
        if n == 0:
            Vs_first = m.Vs.numpy().copy()
        m.M_step()
        cost.append(float(m.compute_expected_neg_log_like()))
        out["M%d" % n] = _snap(m)
    WFs, WFn = m.compute_WF(sample=True)
    S_hat = WFs.numpy() * m.X
    N_hat = WFn.numpy() * m.X
    # the same thing through run() itself
    m2 = make(NoiseTape(draws=tape.draws))
    cost2 = m2.run()
    assert np.array_equal(cost2, np.array(cost)), (cost2, cost)
    assert np.array_equal(m2.S_hat, S_hat) and np.array_equal(m2.N_hat, N_hat)
    ref_mcem.torch = torch

    draws = tape.draws
    assert draws[0][0] == "rand" and draws[1][0] == "rand"
    eps_ = np.stack([t.numpy() for k, t in draws[2:] if k == "randn"])
    u_ = np.stack([t.numpy() for k, t in draws[2:] if k == "rand"])
    assert len(eps_) == len(u_)
    flat = {}
    for k, d in out.items():
        for kk, v in d.items():
            flat["%s_%s" % (k, kk)] = v
    sd = {"sd_" + k: v.numpy() for k, v in vae.state_dict().items()}
    np.savez_compressed(
        os.path.join(OUT, "mcem_%s.npz" % tag),
        model=model, y_kind=str(y_kind), L=L, K=K, niter=niter, chain=np.array(chain),
        var_RW=0.01, eps=1e-8, X=X, y=(np.zeros((N, 0), np.float32) if y is None else y.numpy()),
        rand_W=draws[0][1].numpy(), rand_H=draws[1][1].numpy(), tape_eps=eps_, tape_u=u_,
        cost=np.array(cost), S_hat=S_hat, N_hat=N_hat, WFs=WFs.numpy(), WFn=WFn.numpy(),
        Vs_E0=Vs_first, **flat, **sd)
    print("golden %-7s N=%d L=%d K=%d niter=%d steps=%d cost=%s" % (tag, N, L, K, niter, len(u_), np.round(cost, 4)))


def golden_nonmf(tag="M2_noNMF", L=16, niter=3, chain=(3, 4, 3, 5), T=6144, seed=0):
    """MCEM_M2_noNMF (mcem.py:609-760): fixed noise variance, gain-only M-step.  Same drive as golden_mcem."""
    ref_mcem, ref_models = _reference_modules()
    nE, bE, nW, bW = chain
    x, s, _ = synth_utterance(0, seed=seed, T=T)
    X = stft_oracle.stft(x, **STFT_KW).T
    N, F = X.shape
    torch.manual_seed(0)
    vae = ref_models.DeepGenerativeModel([F, 1, L, [128, 128]], None)
    vae.eval()
    for p in vae.parameters():
        p.requires_grad = False
    with torch.no_grad():
        vae.decoder.reconstruction.bias.copy_(torch.linspace(-6.0, -1.0, F))
        vae.decoder.hidden[0].bias.normal_(0, 0.3)
        vae.decoder.hidden[1].bias.normal_(0, 0.3)
    rs = np.random.RandomState(11)
    P = np.abs(X) ** 2
    Vb = (0.5 * P.mean(0, keepdims=True) * (0.5 + rs.rand(N, F))).astype(np.float32)       # (N,F) as the constructor takes it
    g0 = torch.from_numpy((0.5 + rs.rand(N)).astype(np.float32))
    Z0 = torch.from_numpy((0.3 * rs.randn(N, L)).astype(np.float32))
    y = (torch.from_numpy(rs.rand(N, 1).astype(np.float32)) > 0.4).float()

    def make(tape):
        ref_mcem.torch = _TorchProxy(tape)
        return ref_mcem.MCEM_M2_noNMF(X=X, Vb=Vb, g=g0.clone(), Z=Z0.clone(), y=y, vae=vae, niter=niter, device="cpu",
                                      nsamples_E_step=nE, burnin_E_step=bE, nsamples_WF=nW, burnin_WF=bW, var_RW=0.01)

    tape = NoiseTape(seed=4321 + seed)
    m = make(tape)
    flat, cost, Vs_first = {}, [], None
    for n in range(niter):
        m.E_step()
        flat["E%d_Z" % n] = m.Z.numpy().copy()
        if n == 0:
            Vs_first = m.Vs.numpy().copy()
        m.M_step()
        cost.append(float(m.compute_expected_neg_log_like()))
        flat["M%d_g" % n] = m.g.numpy().copy()
    WFs, WFn = m.compute_WF(sample=True)
    S_hat, N_hat = WFs.numpy() * m.X, WFn.numpy() * m.X
    m2 = make(NoiseTape(draws=tape.draws))
    cost2 = m2.run()
    assert np.array_equal(cost2, np.array(cost)), (cost2, cost)
    assert np.array_equal(m2.S_hat, S_hat) and np.array_equal(m2.N_hat, N_hat)
    ref_mcem.torch = torch
    eps_ = np.stack([t.numpy() for k, t in tape.draws if k == "randn"])
    u_ = np.stack([t.numpy() for k, t in tape.draws if k == "rand"])
    sd = {"sd_" + k: v.numpy() for k, v in vae.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "mcem_%s.npz" % tag), L=L, niter=niter, chain=np.array(chain), var_RW=0.01,
                        X=X, Vb=Vb, g0=g0.numpy(), Z0=Z0.numpy(), y=y.numpy(), tape_eps=eps_, tape_u=u_, cost=np.array(cost),
                        S_hat=S_hat, N_hat=N_hat, WFs=WFs.numpy(), WFn=WFn.numpy(), Vs_E0=Vs_first, **flat, **sd)
    print("golden %-8s N=%d L=%d niter=%d steps=%d cost=%s" % (tag, N, L, niter, len(u_), np.round(cost, 4)))


def golden_real(tag="M2_vad_wsj0", L=16, K=8, niter=3, chain=(4, 6, 5, 8), start=24000, T=12288):
    """A slice of the reference's own WSJ0 fixture (data/subset/pickle/CSR-1-WSJ-0/si_et_05_{mixture,speech,noise}-505.p,
    x = s + n) enhanced by the UNMODIFIED reference: MCEM_M2 guided by clean_speech_VAD labels (target.py:29-50), then
    SI-SDR / SI-SIR / SI-SAR by the reference's python/metrics.py.  Real speech, real noise, reference code end to end
    (only the STFT/ISTFT are the restatement: librosa is not installed)."""
    ref_mcem, ref_models = _reference_modules()
    sys.path.insert(0, REF)
    from python.processing.target import clean_speech_VAD
    from python.metrics import energy_ratios
    sys.path.remove(REF)
    x = _load("si_et_05_mixture-505.p")[0][start:start + T].astype(np.float64)
    sp = _load("si_et_05_speech-505.p")[0][start:start + T].astype(np.float64)
    no = _load("si_et_05_noise-505.p")[0][start:start + T].astype(np.float64)
    assert np.max(np.abs(x - sp - no)) < 1e-12
    X = stft_oracle.stft(x, **STFT_KW).T
    N, F = X.shape
    y_np = clean_speech_VAD(stft_oracle.stft(sp, **STFT_KW), quantile_fraction=0.999, quantile_weight=0.999)   # (1, N)
    y = torch.from_numpy(y_np.T.copy())
    torch.manual_seed(3)
    vae = ref_models.DeepGenerativeModel([F, 1, L, [128, 128]], None).eval()
    for p_ in vae.parameters():
        p_.requires_grad = False
    with torch.no_grad():
        # no trained weights ship with the reference: give the random decoder the long-term spectrum of the clean speech
        # as its output bias, so that the Wiener filter has a speech model of the right scale
        S_pow = np.abs(stft_oracle.stft(sp, **STFT_KW)) ** 2
        vae.decoder.reconstruction.bias.copy_(torch.from_numpy(np.log(S_pow.mean(1) + 1e-8).astype(np.float32)))
        vae.decoder.hidden[0].bias.normal_(0, 0.3)
        vae.decoder.hidden[1].bias.normal_(0, 0.3)
    tape = NoiseTape(seed=99)
    ref_mcem.torch = _TorchProxy(tape)
    nE, bE, nW, bW = chain
    m = ref_mcem.MCEM_M2(niter=niter, nsamples_E_step=nE, burnin_E_step=bE, nsamples_WF=nW, burnin_WF=bW, var_RW=0.01)
    m.init_parameters(X=X, y=y, vae=vae, nmf_rank=K, eps=1e-8, device="cpu")
    cost = m.run()
    ref_mcem.torch = torch
    s_hat = stft_oracle.istft(m.S_hat, fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, max_len=T)
    ratios = np.array(energy_ratios(s_hat.astype(np.float64), sp, no))
    draws = tape.draws
    eps_ = np.stack([t.numpy() for k, t in draws[2:] if k == "randn"])
    u_ = np.stack([t.numpy() for k, t in draws[2:] if k == "rand"])
    sd = {"sd_" + k: v.numpy() for k, v in vae.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "mcem_%s.npz" % tag), model="M2", y_kind="vad", L=L, K=K, niter=niter, chain=np.array(chain),
                        var_RW=0.01, eps=1e-8, x=x, s=sp, n=no, X=X, y=y.numpy(), rand_W=draws[0][1].numpy(), rand_H=draws[1][1].numpy(),
                        tape_eps=eps_, tape_u=u_, cost=cost, S_hat=m.S_hat, N_hat=m.N_hat, ratios=ratios, **sd)
    print("golden %s: N=%d, VAD active %.2f, cost %s, SI-SDR/SIR/SAR %s (input SI-SDR %.2f dB)"
          % (tag, N, float(y_np.mean()), np.round(cost, 4), np.round(ratios, 3), energy_ratios(x, sp, no)[0]))


def golden_spp():
    """timo_mask_estimation (python/models/spp_estimation.py:198-218) of the reference on a synthetic mixture."""
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "python" or k.startswith("python.")]:
        del sys.modules[k]
    from python.models.spp_estimation import timo_mask_estimation
    sys.path.remove(REF)
    from oracle import spp_oracle
    x, _, _ = synth_utterance(3, seed=2, T=16000)
    P = (np.abs(stft_oracle.stft(x, **STFT_KW)) ** 2).astype(np.float32)        # evaluate_M2_ibm.py:137 (float32)
    mask = timo_mask_estimation(P)
    assert mask.dtype == np.float32 and np.array_equal(spp_oracle.timo_mask(P), mask)
    np.savez_compressed(os.path.join(OUT, "spp_mask.npz"), x=x, power=P, mask=mask)
    print("golden spp: %s frames, speech share %.3f" % (P.shape, float((mask > 0.5).mean())))


def golden_labels(T=24000):
    """clean_speech_IBM / clean_speech_VAD (python/processing/target.py:7-50) of the reference on a synthetic clean-speech
    STFT (complex64, as evaluate_M2_ibm.py:115-123 feeds them).  The masks are discrete; the oracle must equal them exactly."""
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "python" or k.startswith("python.")]:
        del sys.modules[k]
    from python.processing import target as ref_target
    sys.path.remove(REF)
    from oracle import mcem_oracle
    _, s, _ = synth_utterance(5, seed=11, T=T)
    S = stft_oracle.stft(s, **STFT_KW)
    out = dict(S=S)
    for q in (0.999, 0.98):
        np.random.seed(0)                                   # target.py:17 draws (and discards) np.random.rand
        ibm = ref_target.clean_speech_IBM(S, quantile_fraction=q, quantile_weight=0.999)
        vad = ref_target.clean_speech_VAD(S, quantile_fraction=q, quantile_weight=0.999)
        assert ibm.dtype == np.float32 and vad.dtype == np.float32 and vad.shape == (1, S.shape[1])
        assert np.array_equal(mcem_oracle.clean_speech_IBM(S, q, 0.999), ibm)
        assert np.array_equal(mcem_oracle.clean_speech_VAD(S, q, 0.999), vad)
        out["ibm_%d" % round(q * 1000)] = ibm.astype(np.uint8)
        out["vad_%d" % round(q * 1000)] = vad.astype(np.uint8)
        print("golden labels q=%.3f: %s, speech share IBM %.4f VAD %.4f" % (q, S.shape, ibm.mean(), vad.mean()))
    np.savez_compressed(os.path.join(OUT, "labels.npz"), **out)


def golden_label_fixture(name="si_dt_05"):
    """The reference's committed label pickles against its committed power spectrograms (see the module docstring)."""
    import itertools
    import pickle
    from oracle.mcem_oracle import lorenz_mask

    class NumpyOnly(pickle.Unpickler):                      # the pickles are untrusted input: numpy reconstruction only
        def find_class(self, module, cls):
            if module.split(".")[0] == "numpy":
                return super().find_class(module, cls)
            raise pickle.UnpicklingError("blocked %s.%s" % (module, cls))

    base = os.path.join(REF, "data", "subset", "pickle", "CSR-1-WSJ-0")
    load = lambda suffix: NumpyOnly(open(os.path.join(base, "%s_%s.p" % (name, suffix)), "rb")).load()
    P, ibm, vad = load("frames"), load("labels"), load("vad_labels")
    assert P.dtype == np.float32 and P.shape == ibm.shape and vad.shape == (1, P.shape[1])
    N = P.shape[1]
    # three utterances per subset: the first boundary is where a prefix reproduces its labels, the second where a suffix does
    first = [b for b in range(50, N - 50) if np.array_equal(lorenz_mask(P[:, :b], 0.98), ibm[:, :b])]
    last = [a for a in range(50, N - 50) if np.array_equal(lorenz_mask(P[:, a:], 0.98), ibm[:, a:])]
    found = [(b1, b2) for b1, b2 in itertools.product(first, last) if b1 < b2 and
             np.array_equal(lorenz_mask(P[:, b1:b2], 0.98), ibm[:, b1:b2])]
    assert len(found) == 1, found
    bounds = np.array([0, found[0][0], found[0][1], N], np.int32)
    for a, b in zip(bounds[:-1], bounds[1:]):
        assert np.array_equal(lorenz_mask(P[:, a:b].sum(axis=0), 0.98)[None], vad[:, a:b])
    np.savez_compressed(os.path.join(OUT, "labels_wsj0.npz"), power=P, ibm=np.packbits(ibm.astype(np.uint8), axis=1),
                        vad=vad.astype(np.uint8), bounds=bounds)
    print("golden label fixture %s: %s, utterances at %s, speech share IBM %.4f VAD %.4f" % (name, P.shape, bounds, ibm.mean(), vad.mean()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["labels"]:
        golden_labels()
        golden_label_fixture()
        sys.exit(0)
    golden_stft()
    golden_mcem("M1", "M1", None, L=16, K=4, niter=2, chain=(3, 4, 2, 5))
    golden_mcem("M2_ibm", "M2", "ibm", L=16, K=10, niter=3, chain=(3, 5, 4, 6))
    golden_mcem("M2_vad", "M2", "vad", L=32, K=10, niter=2, chain=(2, 3, 2, 3))
    golden_nonmf()
    golden_spp()
    golden_real()
    golden_labels()
    golden_label_fixture()
