"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances (north star): W, H, g, posterior speech variance within rtol 1e-4 in fp32 mode;
decisions of the Metropolis-Hastings chain are discrete, so the chain is compared (a) on the
continuous log acceptance ratio, (b) on decisions away from ties, (c) in forced-decision
replay (the oracle's accept stream), where all state must agree to rtol 1e-4.
"""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, load_golden, golden_state_dict, oracle_from_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-4
STFT_KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25)


def _vae_from_golden(g):
    from python.models.models import DeepGenerativeModel, VariationalAutoencoder
    F, L = g["X"].shape[1], int(g["L"])
    if str(g["model"]) == "M1":
        vae = VariationalAutoencoder([F, L, [128, 128]])
    else:
        vae = DeepGenerativeModel([F, g["y"].shape[1], L, [128, 128]], None)
    vae.load_state_dict(golden_state_dict(g))
    return vae.eval()


def _mcem_from_golden(g, forced=None, precision="fp32"):
    from python.models.mcem import MCEM_M1, MCEM_M2
    nE, bE, nW, bW = [int(v) for v in g["chain"]]
    cls = MCEM_M1 if str(g["model"]) == "M1" else MCEM_M2
    m = cls(int(g["niter"]), nE, bE, nW, bW, float(g["var_RW"]))
    m.precision = precision
    m.replay = dict(rand_W=g["rand_W"], rand_H=g["rand_H"], eps=g["tape_eps"], u=g["tape_u"], forced=forced)
    vae = _vae_from_golden(g)
    if str(g["model"]) == "M1":
        m.init_parameters(X=g["X"], vae=vae, nmf_rank=int(g["K"]), eps=float(g["eps"]), device="cuda:0")
    else:
        m.init_parameters(X=g["X"], y=torch.from_numpy(g["y"]).cuda(), vae=vae, nmf_rank=int(g["K"]),
                          eps=float(g["eps"]), device="cuda:0")
    return m


def _oracle_trace(g):
    o = oracle_from_golden(g)
    o.trace = []
    snaps = {}

    def hook(oo, n):
        snaps[n] = {k: getattr(oo, k).numpy().copy() for k in ("W", "H", "g", "Z", "Vb")}
        snaps[n]["Vs"] = oo.Vs.numpy().copy()
    o.iter_hook = hook
    cost = o.run()
    acc = np.stack([t[0].numpy() for t in o.trace])
    dec = np.stack([t[1].numpy() for t in o.trace])
    logu = np.stack([t[2].numpy() for t in o.trace])
    return o, cost, acc, dec, logu, snaps


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["M1", "M2_ibm", "M2_vad"])
def test_init_matches_golden(tag):
    g = load_golden(tag)
    m = _mcem_from_golden(g)
    np.testing.assert_allclose(m.W.cpu().numpy(), g["init_W"], rtol=0, atol=0)
    np.testing.assert_allclose(m.H.cpu().numpy(), g["init_H"], rtol=0, atol=0)
    np.testing.assert_allclose(m.Vb.cpu().numpy(), g["init_Vb"], rtol=1e-6)
    # encoder mean (models.py:90-104) through gvn_dense
    np.testing.assert_allclose(m.Z.cpu().numpy(), g["init_Z"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("tag", ["M1", "M2_ibm", "M2_vad"])
def test_forced_decision_run_matches_golden(tag):
    """Whole run() with the oracle's accept stream: every iterate within rtol 1e-4."""
    g = load_golden(tag)
    o, cost_o, acc_o, dec_o, logu_o, snaps = _oracle_trace(g)
    np.testing.assert_array_equal(cost_o, g["cost"])                       # oracle == reference
    m = _mcem_from_golden(g, forced=dec_o)
    seen = {}
    for n in range(int(g["niter"])):
        m.E_step()
        np.testing.assert_allclose(m.Vs.cpu().numpy(), snaps[n]["Vs"], rtol=RTOL, err_msg="Vs iter %d" % n)
        m.M_step()
        for k in ("W", "H", "g", "Vb", "Z"):
            # Z is a sum of O(1) increments: its entries near zero carry an absolute error
            atol = 2e-6 if k == "Z" else 1e-7
            np.testing.assert_allclose(getattr(m, k).cpu().numpy(), g["M%d_%s" % (n, k)], rtol=RTOL, atol=atol,
                                       err_msg="%s iter %d" % (k, n))
        c = float(m.compute_expected_neg_log_like())
        assert abs(c - g["cost"][n]) <= RTOL * abs(g["cost"][n])
    WFs, WFn = m.compute_WF(sample=True)
    np.testing.assert_allclose(WFs.cpu().numpy(), g["WFs"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(WFn.cpu().numpy(), g["WFn"], rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("tag", ["M1", "M2_ibm", "M2_vad"])
def test_free_running_run_matches_golden(tag):
    """run() on the recorded noise without forcing: the log acceptance ratios are continuous
    and must agree; decisions may only differ at ties; then everything matches rtol 1e-4."""
    g = load_golden(tag)
    o, cost_o, acc_o, dec_o, logu_o, snaps = _oracle_trace(g)
    m = _mcem_from_golden(g)
    # first chain with trace: compare acc_prob and decisions
    (R, burnin), _ = m.chain_lengths()
    m._run_chain(R, burnin, trace=True)
    acc = m.last_trace["acc_prob"].cpu().numpy()
    dec = m.last_trace["accepted"].cpu().numpy().astype(bool)
    n1 = R + burnin
    margin = np.abs(acc_o[:n1] - logu_o[:n1])
    np.testing.assert_allclose(acc, acc_o[:n1], rtol=1e-3, atol=2e-3)
    assert np.array_equal(dec[margin > 5e-3], dec_o[:n1][margin > 5e-3])
    # full free run
    m = _mcem_from_golden(g)
    cost = m.run()
    if np.array_equal(dec, dec_o[:n1]):
        np.testing.assert_allclose(cost, g["cost"], rtol=RTOL)
        np.testing.assert_allclose(m.S_hat, g["S_hat"], rtol=1e-3, atol=1e-5 * np.max(np.abs(g["S_hat"])))
    else:                                                    # a tie flipped: statistical agreement only
        np.testing.assert_allclose(cost, g["cost"], rtol=5e-2)
    assert m.S_hat.dtype == np.complex64 and m.S_hat.shape == g["S_hat"].shape
    # S_hat + N_hat == X  (WFs + WFn == 1)
    np.testing.assert_allclose(m.S_hat + m.N_hat, g["X"].T, rtol=1e-4, atol=1e-6 * np.max(np.abs(g["X"])))


def test_sample_posterior_api_and_shapes():
    g = load_golden("M2_ibm")
    m = _mcem_from_golden(g)
    Z0 = m.Z.clone()
    zs, zsy = m.sample_posterior(Z0, m.y, nsamples=3, burnin=2)
    N, L = g["X"].shape[0], int(g["L"])
    assert zs.shape == (N, 3, L) and zsy.shape == (N, 3, L + g["y"].shape[1])
    assert m.Vs.shape == (3, g["X"].shape[1], N)
    # the last kept sample is the new state (mcem.py:319)
    np.testing.assert_array_equal(zs[:, -1, :].T.cpu().numpy(), m.Z.cpu().numpy())
    # Vs of the kept samples equals the decoder applied to them (compute_Vs, mcem.py:297-307)
    from oracle.mcem_oracle import decode, split_state_dict
    dec = split_state_dict(golden_state_dict(g), "decoder")
    ref = decode(dec, zsy.cpu()).permute(1, 2, 0).numpy()
    np.testing.assert_allclose(m.Vs.cpu().numpy(), ref, rtol=RTOL)


def test_batch_equals_single_utterances():
    """Ragged batch of three utterances (different N) == each run alone, on replayed noise."""
    from gvn import engine as E
    from gvn.pipeline import McemConfig, Enhancer
    from gvn.synth import synth_utterance
    from python.models.models import DeepGenerativeModel
    torch.manual_seed(1)
    F, L, K = 513, 16, 5
    vae = DeepGenerativeModel([F, 1, L, [128, 128]], None).eval()
    with torch.no_grad():
        vae.decoder.reconstruction.bias.copy_(torch.linspace(-6.0, -1.0, F))
    cfg = McemConfig(model="M2", niter=2, nsamples_E_step=2, burnin_E_step=3, nsamples_WF=2, burnin_WF=3, nmf_rank=K)
    Ts = [6144, 9000, 4000]                                   # 9000 and 4000 trigger the end-pad rule
    wavs = [synth_utterance(i, seed=2, T=T)[0] for i, T in enumerate(Ts)]
    enh = Enhancer(vae, cfg, "cuda:0")
    rs = np.random.RandomState(0)
    geo = [E.stft_geometry(T, cfg.fs, cfg.wlen_sec, cfg.hop_percent) for T in Ts]
    Ns = [gq[3] for gq in geo]
    labels = [(rs.rand(1, n) > 0.4).astype(np.float32) for n in Ns]
    steps = cfg.niter * 5 + 5
    eps = [rs.randn(steps, L, n).astype(np.float32) for n in Ns]
    us = [rs.rand(steps, n).astype(np.float32) for n in Ns]
    rW = [rs.rand(F, K).astype(np.float32) for _ in Ns]
    rH = [rs.rand(K, n).astype(np.float32) for n in Ns]

    def run(idx):
        b = enh.prepare([wavs[i] for i in idx], [labels[i] for i in idx], rand=([rW[i] for i in idx], [rH[i] for i in idx]))
        nz = E.ReplayNoise.from_utterance_tapes(b, [torch.from_numpy(eps[i]) for i in idx],
                                                [torch.from_numpy(us[i]) for i in idx], [5] * cfg.niter + [5])
        s, n, c = enh.run(b, noise=nz)
        torch.cuda.synchronize()
        return s.cpu().numpy(), c.cpu().numpy(), b
    s_all, c_all, b_all = run([0, 1, 2])
    assert b_all.NP == sum((n + 31) // 32 * 32 for n in Ns)
    for i in range(3):
        s_i, c_i, _ = run([i])
        np.testing.assert_allclose(c_all[:, i], c_i[:, 0], rtol=1e-6)
        np.testing.assert_allclose(s_all[i, :Ts[i]], s_i[0, :Ts[i]], rtol=1e-4, atol=1e-6)
    assert np.all(np.isfinite(s_all))


def test_philox_run_properties_full_shape():
    """Throughput mode (in-kernel Philox) at the benchmark shape (4 s, F=513, K=10, L=16):
    size-independent properties of the algorithm."""
    from gvn.pipeline import McemConfig, Enhancer
    from gvn.synth import synth_batch
    from python.models.models import DeepGenerativeModel
    from python.processing.target import clean_speech_IBM
    from oracle import stft_oracle
    torch.manual_seed(0)
    vae = DeepGenerativeModel([513, 513, 16, [128, 128]], None).eval()
    cfg = McemConfig(model="M2", niter=6, nmf_rank=10)
    x, s, n = synth_batch(3, seed=0, T=64000)
    labels = [clean_speech_IBM(stft_oracle.stft(si, dtype="complex64", **STFT_KW), 0.999, 0.999) for si in s]
    enh = Enhancer(vae, cfg, "cuda:0")
    b = enh.prepare(list(x), labels, seed=5)
    assert b.NP == 3 * 256 and b.n_frames_host == [251] * 3
    from gvn import engine as E
    cost, S, Nn, WFs, WFn = E.run_mcem(b, enh.dec, cfg.niter, *cfg.chains(), cfg.var_RW, "fp32", seed=5, want_masks=True)
    torch.cuda.synchronize()
    cost = cost.cpu().numpy()
    assert np.all(np.isfinite(cost)) and np.all(cost[-1] < cost[0])          # EM decreases the cost
    for i in range(3):
        c = b.cols(i)
        np.testing.assert_allclose(b.W[i].abs().sum(0).cpu().numpy(), 1.0, rtol=1e-5)   # mcem.py:128-131
        w = (WFs[:, c] + WFn[:, c]).cpu().numpy()
        np.testing.assert_allclose(w, 1.0, rtol=1e-5)
        assert float(WFs[:, c].min()) >= 0 and float(WFs[:, c].max()) <= 1
    # determinism: same seed -> identical result; different seed -> different chain
    b2 = enh.prepare(list(x), labels, seed=5)
    cost2 = E.run_mcem(b2, enh.dec, cfg.niter, *cfg.chains(), cfg.var_RW, "fp32", seed=5)[0].cpu().numpy()
    np.testing.assert_array_equal(cost, cost2)
    b3 = enh.prepare(list(x), labels, seed=5)
    cost3 = E.run_mcem(b3, enh.dec, cfg.niter, *cfg.chains(), cfg.var_RW, "fp32", seed=6)[0].cpu().numpy()
    assert not np.array_equal(cost, cost3)
    np.testing.assert_allclose(cost3, cost, rtol=2e-2)


def test_acceptance_rate_is_sane():
    from python.models.mcem import MCEM_M2
    g = load_golden("M2_ibm")
    m = MCEM_M2(1)
    m.seed = 11
    m.init_parameters(X=g["X"], y=torch.from_numpy(g["y"]).cuda(), vae=_vae_from_golden(g), nmf_rank=10, eps=1e-8,
                      device="cuda:0")
    m._run_chain(10, 30, trace=True)
    rate = float(m.last_trace["n_accepted"].float().mean()) / 40
    assert 0.3 < rate < 0.99, rate
