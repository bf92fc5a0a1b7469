"""End-to-end agreement of the enhanced signal (north star: SI-SDR within 0.05 dB): one synthetic
4 s utterance at the benchmark shape (F=513, K=10, L=16, M2 with oracle IBM labels) enhanced by
(a) the oracle = the reference's torch-CPU algorithm on a recorded noise tape, (b) the CUDA path in
fp32 mode replaying the same tape, (c) the tensor-core (f16) mode replaying the same tape,
(d) the tensor-core mode on its own Philox stream.  (a)-(b) is the parity claim; (c) and (d) differ
from (a) by Monte-Carlo noise only (a near-tie decision that flips restarts an independent chain)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25)


def _si_sdr(est, s, n):
    from python.metrics import energy_ratios
    return energy_ratios(est, s, n)[0]


def test_si_sdr_agreement_with_the_oracle():
    from gvn.synth import synth_utterance
    from oracle import stft_oracle
    from oracle.mcem_oracle import McemOracle, NoiseTape, split_state_dict
    from python.models.mcem import MCEM_M2
    from python.models.models import DeepGenerativeModel
    from python.processing.stft import stft, istft
    from python.processing.target import clean_speech_IBM

    x, s, n = synth_utterance(0, seed=11, T=64000)
    X = stft_oracle.stft(x, dtype="complex64", **KW)                        # (F, N)
    y = torch.from_numpy(clean_speech_IBM(stft_oracle.stft(s, dtype="complex64", **KW), 0.999, 0.999).T.copy())
    F, N, L, K, niter = X.shape[0], X.shape[1], 16, 10, 12
    torch.manual_seed(0)
    vae = DeepGenerativeModel([F, F, L, [128, 128]], None).eval()
    chain = (10, 30, 25, 75)

    # (a) oracle with a recorded tape
    o = McemOracle(niter, *chain, 0.01, model="M2")
    o.trace = []
    tape = NoiseTape(seed=5)
    sd = vae.state_dict()
    o.init_parameters(X.T, y, split_state_dict(sd, "decoder"), split_state_dict(sd, "encoder"), K, 1e-8, tape)
    o.run()
    d = tape.draws
    eps = np.stack([t.numpy() for k, t in d[2:] if k == "randn"])
    u = np.stack([t.numpy() for k, t in d[2:] if k == "rand"])
    s_o = stft_oracle.istft(o.S_hat, max_len=len(x), **KW)
    sdr_o = _si_sdr(s_o, s, n)

    def run_gpu(precision, replay):
        m = MCEM_M2(niter, *chain, 0.01)
        m.precision = precision
        m.seed = 123
        if replay:
            m.replay = dict(rand_W=d[0][1], rand_H=d[1][1], eps=eps, u=u)
        m.init_parameters(X=X.T, y=y.cuda(), vae=vae, nmf_rank=K, eps=1e-8, device="cuda:0")
        m.run()
        return _si_sdr(istft(m.S_hat, max_len=len(x), **KW), s, n)

    sdr_fp32 = run_gpu("fp32", True)
    sdr_tc = run_gpu("f16", True)
    sdr_tc_philox = run_gpu("f16", False)
    print("SI-SDR [dB]: oracle %.4f | fp32 replay %.4f | f16 replay %.4f | f16 philox %.4f" % (sdr_o, sdr_fp32, sdr_tc, sdr_tc_philox))
    assert abs(sdr_fp32 - sdr_o) < 0.05                    # parity (same noise, same decisions up to ties)
    assert abs(sdr_tc - sdr_o) < 0.05                      # measured 0.0001 dB (same tape) and 0.012 dB (own stream)
    assert abs(sdr_tc_philox - sdr_o) < 0.05


@pytest.mark.parametrize("precision", ["f16", "fp32"])
def test_config4_shape_properties(precision):
    """BASELINE config 4 (NMF stress): M1, 30 s utterance (N = 1876 frames), K = 32, 10 kept samples per
    frame (MCEM_M1(burnin_E_step=10)).  Size-independent properties at the full shape."""
    from gvn.synth import synth_utterance
    from python.models.mcem import MCEM_M1
    from python.models.models import VariationalAutoencoder
    from python.processing.stft import stft
    x, s, n = synth_utterance(1, seed=4, T=480000)
    X = stft(x, dtype="complex64", **KW)
    assert X.shape == (513, 1876)
    torch.manual_seed(0)
    vae = VariationalAutoencoder([513, 16, [128, 128]]).eval()
    m = MCEM_M1(niter=3, nsamples_E_step=10, burnin_E_step=10, nsamples_WF=25, burnin_WF=12, var_RW=0.01)
    m.precision = precision
    m.seed = 7
    m.init_parameters(X=X.T, vae=vae, nmf_rank=32, eps=1e-8, device="cuda:0")
    cost = m.run()
    assert np.all(np.isfinite(cost)) and cost[-1] < cost[0]
    assert m.Vs.shape == (12, 513, 1876)                                   # Wiener chain keeps burnin_WF samples (quirk)
    np.testing.assert_allclose(m.W.abs().sum(0).cpu().numpy(), 1.0, rtol=1e-5)
    np.testing.assert_allclose(m.S_hat + m.N_hat, X, rtol=1e-4, atol=1e-6 * np.max(np.abs(X)))
    assert m.S_hat.shape == (513, 1876) and m.S_hat.dtype == np.complex64


def test_enhance_many_equals_enhance():
    """The pipelined end-to-end API (upload of batch i+1 overlapping batch i, byte labels, reused staging
    buffers and batch state) returns exactly what one-batch-at-a-time enhance() returns."""
    from gvn import engine as E
    from gvn.pipeline import McemConfig, Enhancer
    from gvn.synth import synth_batch
    from python.metrics import energy_ratios
    from python.models.models import DeepGenerativeModel
    torch.manual_seed(0)
    vae = DeepGenerativeModel([513, 1, 16, [128, 128]], None).eval()
    cfg = McemConfig(model="M2", niter=3, nsamples_E_step=4, burnin_E_step=4, nsamples_WF=4, burnin_WF=4, nmf_rank=6, precision="f16")
    enh = Enhancer(vae, cfg, "cuda:0")
    rs = np.random.RandomState(1)
    items = []
    for k, (B, T) in enumerate([(3, 9000), (2, 12000), (3, 9000)]):
        x, s, n = synth_batch(B, seed=20 + k, T=T)
        N = E.stft_geometry(T, 16000, 64e-3, 0.25)[3]
        labels = [(rs.rand(1, N) > 0.5).astype(np.uint8) for _ in range(B)]
        items.append(dict(wavs=list(x), labels=labels, refs=(s, n)))
    outs = []
    for out in enh.enhance_many(items, seed=40):
        outs.append((out["s_hat"].numpy().copy(), out["cost"].numpy().copy(), out["metrics"].numpy().copy(), list(out["T"])))
    assert len(outs) == 3
    for k, it in enumerate(items):
        s_ref, n_ref, cost_ref = enh.enhance(it["wavs"], [l.astype(np.float32) for l in it["labels"]], seed=40 + k)
        s_hat, cost, metrics, T = outs[k]
        np.testing.assert_array_equal(cost, cost_ref)
        for i in range(len(it["wavs"])):
            np.testing.assert_array_equal(s_hat[i, :T[i]], s_ref[i])
            ref = energy_ratios(s_ref[i].astype(np.float64), it["refs"][0][i], it["refs"][1][i])
            np.testing.assert_allclose(metrics[i], ref, rtol=1e-6, atol=1e-6)


def test_enhance_many_over_a_ragged_list_with_device_labels():
    """A file-list-like job: nine batches, every one with its own geometry (batch size AND lengths differ; more shapes than
    the enhancer keeps batch state for), oracle labels made on the upload stream, two passes over the list.  The upload
    thread allocates and drops memory while the compute stream is still working on the previous batches, so batch state
    dropped by one stream must not be handed to the other early (this crashed with an illegal address when the label upload
    shared full batch objects with the compute stream).  Results must equal the one-batch-at-a-time path exactly."""
    from gvn.pipeline import McemConfig, Enhancer
    from gvn.synth import synth_utterance
    from python.models.models import DeepGenerativeModel
    torch.manual_seed(0)
    vae = DeepGenerativeModel([513, 513, 16, [128, 128]], None).eval()
    cfg = McemConfig(model="M2", niter=2, nsamples_E_step=3, burnin_E_step=3, nsamples_WF=3, burnin_WF=3, nmf_rank=10, precision="f16")
    enh = Enhancer(vae, cfg, "cuda:0", label_source="oracle_ibm")
    rs = np.random.RandomState(3)
    items = []
    for k in range(9):
        B = int(rs.randint(2, 9))
        xs, ss, ns = zip(*[synth_utterance(10 * k + i, seed=6, T=256 * int(rs.randint(40, 140))) for i in range(B)])
        items.append(dict(wavs=list(xs), refs=(list(ss), list(ns))))
    assert len({tuple(len(w) for w in it["wavs"]) for it in items}) == 9
    outs = []
    for out in enh.enhance_many(items + items, seed=70):
        outs.append((out["s_hat"].numpy().copy(), out["cost"].numpy().copy(), out["metrics"].numpy().copy(), list(out["T"])))
    assert len(outs) == 18
    for k, it in enumerate(items + items):
        up = enh.upload(it["wavs"], refs=it["refs"])
        b = enh.prepare(None, None, seed=70 + k, uploaded=up)
        s_ref, _, cost_ref = enh.run(b, seed=70 + k)
        torch.cuda.synchronize()
        s_hat, cost, metrics, T = outs[k]
        assert np.all(np.isfinite(cost)) and np.all(np.isfinite(metrics))
        np.testing.assert_array_equal(cost, cost_ref.cpu().numpy())
        s_ref = s_ref.cpu().numpy()
        for i in range(len(it["wavs"])):
            np.testing.assert_array_equal(s_hat[i, :T[i]], s_ref[i, :T[i]])


@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_real_wsj0_slice_matches_reference_golden(precision):
    """Real speech + noise from the reference's own fixture, enhanced by the unmodified reference (tests/golden/
    mcem_M2_vad_wsj0.npz): the CUDA path on the same tape with the same decisions gives the same cost and the same
    SI-SDR / SI-SIR / SI-SAR (north star: within 0.05 dB)."""
    from _util import load_golden, golden_state_dict, oracle_from_golden
    from gvn import engine as E
    from python.models.mcem import MCEM_M2
    from python.models.models import DeepGenerativeModel
    from python.processing.stft import istft
    g = load_golden("M2_vad_wsj0")
    o = oracle_from_golden(g)
    o.trace = []
    o.run()
    forced = np.stack([t[1].numpy() for t in o.trace])
    F, L = g["X"].shape[1], int(g["L"])
    vae = DeepGenerativeModel([F, 1, L, [128, 128]], None)
    vae.load_state_dict(golden_state_dict(g))
    nE, bE, nW, bW = [int(v) for v in g["chain"]]
    m = MCEM_M2(int(g["niter"]), nE, bE, nW, bW, float(g["var_RW"]))
    m.precision = precision
    m.replay = dict(rand_W=g["rand_W"], rand_H=g["rand_H"], eps=g["tape_eps"], u=g["tape_u"], forced=forced)
    m.init_parameters(X=g["X"], y=torch.from_numpy(g["y"]).cuda(), vae=vae.eval(), nmf_rank=int(g["K"]), eps=float(g["eps"]), device="cuda:0")
    cost = m.run()
    np.testing.assert_allclose(cost, g["cost"], rtol=1e-4 if precision == "fp32" else 2e-3)
    T = len(g["x"])
    s_hat = istft(m.S_hat, max_len=T, **KW)
    from python.metrics import energy_ratios
    r = np.array(energy_ratios(s_hat.astype(np.float64), g["s"], g["n"]))
    assert np.max(np.abs(r - g["ratios"])) < 0.05, (r, g["ratios"])
    # the same three numbers from the device kernel (gvn_energy_ratios)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda().reshape(1, -1)
    rd = E.energy_ratios(dev(s_hat), dev(g["s"]), dev(g["n"]), [T]).cpu().numpy()[0]
    assert np.max(np.abs(rd - g["ratios"])) < 0.05, (rd, g["ratios"])
