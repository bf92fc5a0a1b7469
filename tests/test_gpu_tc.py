"""Tensor-core (tcgen05/TMEM) E-step: hardware self test of the UMMA plumbing and parity of the
f16 chain against the fp32 CUDA-core chain and the oracle.  f16 operands carry 11 mantissa bits
(= TF32), so this is the north star's looser-tolerance mode: Vs within 5e-3.  The chain reads X2
and Vb as bf16 (a fixed 2^-9 perturbation of the target density, identical on both sides of the
ratio), so the log acceptance ratio -- a difference of two sums of 513 terms -- is compared within
0.25 absolute, and 99% of the steps within 0.05."""
import ctypes as C

import numpy as np
import pytest
import torch

from _util import load_golden
from test_gpu_parity import _mcem_from_golden, _oracle_trace

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 16), (128, 32), (128, 128), (16, 128), (176, 128), (256, 64)])
def test_umma_selftest(N, K):
    from gvn import _lib
    lib = _lib.load()
    torch.manual_seed(N * 1000 + K)
    A = torch.randn(128, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * 0.1
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref16 = A.half().float() @ W.half().float().T
    ref64 = (A.double() @ W.double().T)
    for variant, ref, tol in ((0, ref16, 2e-5), (4, ref64.float(), 5e-5)):
        D = torch.full((128, N), float("nan"), device="cuda")
        _lib.check(lib.gvn_selftest_umma(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), N, K, variant,
                                         C.c_void_p(D.data_ptr()), st))
        torch.cuda.synchronize()
        assert float((D - ref).abs().max()) < tol * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("tag", ["M2_ibm", "M1", "M2_vad"])
def test_tc_chain_matches_fp32_chain(tag):
    """Same state, same noise, same (forced) decisions: f16 tensor-core chain vs fp32 chain."""
    g = load_golden(tag)
    o, cost_o, acc_o, dec_o, logu_o, snaps = _oracle_trace(g)
    out = {}
    for prec in ("fp32", "f16"):
        m = _mcem_from_golden(g, forced=dec_o, precision=prec)
        (R, burnin), _ = m.chain_lengths()
        m._run_chain(R, burnin, trace=True)
        torch.cuda.synchronize()
        out[prec] = dict(acc=m.last_trace["acc_prob"].cpu().numpy(), Vs=m.Vs.cpu().numpy(), Z=m.Z.cpu().numpy(),
                         zs=m.last_trace["z_samples"].cpu().numpy(), cnt=m.last_trace["n_accepted"].cpu().numpy())
    a, b = out["fp32"], out["f16"]
    np.testing.assert_array_equal(a["cnt"], b["cnt"])
    np.testing.assert_allclose(b["Z"], a["Z"], rtol=1e-6, atol=1e-6)            # forced decisions: same path
    np.testing.assert_allclose(b["zs"], a["zs"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(b["Vs"], a["Vs"], rtol=5e-3)
    np.testing.assert_allclose(b["acc"], a["acc"], atol=0.25, rtol=2e-2)
    assert np.quantile(np.abs(b["acc"] - a["acc"]) / (1 + np.abs(a["acc"])), 0.99) < 0.05
    # and against the oracle itself
    n1 = R + burnin
    np.testing.assert_allclose(b["acc"], acc_o[:n1], atol=0.25, rtol=2e-2)
    np.testing.assert_allclose(b["Vs"], snaps[0]["Vs"], rtol=5e-3)


@pytest.mark.parametrize("tag", ["M2_ibm", "M1"])
def test_tc_full_run_forced(tag):
    g = load_golden(tag)
    o, cost_o, acc_o, dec_o, logu_o, snaps = _oracle_trace(g)
    m = _mcem_from_golden(g, forced=dec_o, precision="f16")
    cost = m.run()
    np.testing.assert_allclose(cost, g["cost"], rtol=2e-3)
    np.testing.assert_allclose(m.W.cpu().numpy(), g["M%d_W" % (int(g["niter"]) - 1)], rtol=2e-2, atol=1e-6)
    np.testing.assert_allclose(m.S_hat, g["S_hat"], rtol=2e-2, atol=2e-3 * np.max(np.abs(g["S_hat"])))


def test_tc_philox_matches_fp32_statistics():
    """Free-running Philox chains at the benchmark shape: the f16 chain and the fp32 chain use the
    same random stream, so costs agree closely and the final Wiener masks are statistically equal."""
    from gvn import engine as E
    from gvn.pipeline import McemConfig, Enhancer
    from gvn.synth import synth_batch
    from python.models.models import DeepGenerativeModel
    from python.processing.target import clean_speech_IBM
    from oracle import stft_oracle
    kw = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25)
    torch.manual_seed(0)
    vae = DeepGenerativeModel([513, 513, 16, [128, 128]], None).eval()
    x, s, n = synth_batch(3, seed=0, T=64000)
    labels = [clean_speech_IBM(stft_oracle.stft(si, dtype="complex64", **kw), 0.999, 0.999) for si in s]
    res = {}
    for prec in ("fp32", "f16"):
        cfg = McemConfig(model="M2", niter=5, nmf_rank=10, precision=prec)
        enh = Enhancer(vae, cfg, "cuda:0")
        b = enh.prepare(list(x), labels, seed=5)
        cost, S, Nn, WFs, WFn = E.run_mcem(b, enh.dec, cfg.niter, *cfg.chains(), cfg.var_RW, prec, seed=5, want_masks=True)
        torch.cuda.synchronize()
        res[prec] = (cost.cpu().numpy(), WFs.cpu().numpy())
    assert np.all(np.isfinite(res["f16"][0]))
    np.testing.assert_allclose(res["f16"][0], res["fp32"][0], rtol=1e-2)
    assert np.mean(np.abs(res["f16"][1] - res["fp32"][1])) < 0.02


@pytest.mark.parametrize("L,y_dim", [(32, 513), (32, 0), (8, 1)])
def test_tc_chain_other_latent_dims(L, y_dim):
    """z_dim = 32 is what the evaluate scripts themselves use (scripts/evaluate_M2_ibm.py:33-38); the
    tensor-core chain pads L to a multiple of 16.  Same noise, same forced decisions: f16 chain vs fp32 chain."""
    from gvn import engine as E
    from python.models.models import DeepGenerativeModel, VariationalAutoencoder
    torch.manual_seed(3)
    F, N, K, R, burnin = 513, 70, 6, 4, 5
    vae = (VariationalAutoencoder([F, L, [128, 128]]) if y_dim == 0 else DeepGenerativeModel([F, y_dim, L, [128, 128]], None)).eval()
    with torch.no_grad():
        vae.decoder.reconstruction.bias.copy_(torch.linspace(-6.0, -1.0, F))
    dec = E.PackedDecoder(vae, "cuda:0")
    g = torch.Generator(device="cuda").manual_seed(5)
    steps = R + burnin
    eps = torch.randn(steps, L, 96, generator=g, device="cuda")
    u = torch.rand(steps, 96, generator=g, device="cuda").clamp_(1e-6, 1.0)
    forced = (torch.rand(steps, 96, generator=g, device="cuda") < 0.7).to(torch.uint8)
    y = None if y_dim == 0 else (torch.rand(y_dim, 96, generator=g, device="cuda") > 0.5).float()
    z0 = torch.randn(L, 96, generator=g, device="cuda") * 0.5
    out = {}
    for prec in ("fp32", "f16"):
        b = E.Batch([N], F, K, L, R, "cuda:0", with_complex=False)
        b.X2.copy_(torch.rand(F, b.NP, generator=torch.Generator(device="cuda").manual_seed(9), device="cuda") * 2 + 1e-3)
        E.init_nmf(b, 1e-8, generator=torch.Generator(device="cuda").manual_seed(11))
        E.set_labels(b, dec, y)
        b.Z.copy_(z0)
        acc, dec_, cnt, zs = E.estep(b, dec, burnin, R, 0.01, prec, eps=eps, u=u, forced=forced, trace=True)
        torch.cuda.synchronize()
        c = b.cols(0)
        out[prec] = dict(acc=acc[:, c].cpu().numpy(), Z=b.Z[:, c].cpu().numpy(), Vs=b.expand_samples(R, 0).cpu().numpy(),
                         w=b.Vs_w[:R, c].cpu().numpy(), zs=zs[..., c].cpu().numpy())
    a, t = out["fp32"], out["f16"]
    np.testing.assert_array_equal(t["w"], a["w"])
    assert np.all(a["w"].sum(0) == R)                                       # multiplicities add up to R for every frame
    np.testing.assert_allclose(t["Z"], a["Z"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(t["zs"], a["zs"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(t["Vs"], a["Vs"], rtol=5e-3)
    np.testing.assert_allclose(t["acc"], a["acc"], atol=0.25, rtol=2e-2)


@pytest.mark.parametrize("K,variant", [(10, 1), (20, 1), (10, 0)])
def test_xv_kept_current_by_mstep(K, variant):
    """gvn_mstep rewrites XV next to every Vb it writes (staged sweep, generic sweep, variant 0), so a chain
    launched with GVN_PREC_XV_CURRENT sees the words the packing pass would have produced: bit-identical runs."""
    from gvn import engine as E
    from gvn.pipeline import McemConfig, Enhancer
    import bench
    vae = bench.build_model()
    cfg = McemConfig(model="M2", niter=3, nmf_rank=K, precision="f16", mstep_variant=variant)
    enh = Enhancer(vae, cfg, "cuda:0")
    x, s, nz, labels = bench.make_inputs(3, 0, T=12000)
    res = []
    for hook in (None, lambda b, n: None):                  # a hook disables the shortcut: every chain packs XV itself
        b = enh.prepare(list(x), labels, seed=4)
        cost, S, Nn, _, _ = E.run_mcem(b, enh.dec, 3, (10, 30), (25, 75), 0.01, "f16", 7, None, variant, iter_hook=hook)
        torch.cuda.synchronize()
        res.append((cost.cpu().numpy(), S.cpu().numpy(), b.XV.cpu().numpy().copy(), b.Vb.cpu().numpy().copy()))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][1], res[1][1])
    # and the words themselves: low half = bf16(Vb), whole word as f32 within 2^-8 of X2
    xv, vb = res[0][2].view(np.uint32), res[0][3]
    valid = (b.frame_utt >= 0).cpu().numpy()
    lo = (xv & 0xffff).astype(np.uint32) << 16
    np.testing.assert_allclose(lo.view(np.float32)[:, valid], vb[:, valid], rtol=2.0 ** -8)
    x2 = b.X2.cpu().numpy()
    np.testing.assert_allclose(xv.view(np.float32)[:, valid], x2[:, valid], rtol=2.0 ** -8, atol=1e-30)


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_stored_speech_variance_is_clamped_to_a_finite_maximum(prec):
    """A proposal whose decoder output overflows is rejected (acc_prob = -inf), but its slot (multiplicity 0) is still
    multiplied into the weighted sums of the M-step: gvn_estep therefore never stores inf (GVN_VS_MAX, include/gvn.h)."""
    from gvn import engine as E
    from python.models.models import VariationalAutoencoder
    torch.manual_seed(1)
    F, L, N, R, burnin = 513, 16, 40, 3, 2
    vae = VariationalAutoencoder([F, L, [128, 128]]).eval()
    with torch.no_grad():
        vae.decoder.reconstruction.bias.fill_(-3.0)
        vae.decoder.reconstruction.bias[7] = 100.0          # exp(100) overflows fp32
    dec = E.PackedDecoder(vae, "cuda:0")
    b = E.Batch([N], F, 4, L, R, "cuda:0", with_complex=False)
    b.X2.copy_(torch.rand(F, b.NP, device="cuda") + 1e-3)
    E.init_nmf(b, 1e-8, generator=torch.Generator(device="cuda").manual_seed(2))
    E.set_labels(b, dec, None)
    E.estep(b, dec, burnin, R, 0.01, prec, seed=3)
    torch.cuda.synchronize()
    vs = b.Vs[:R].permute(0, 2, 1, 3).reshape(R, F, b.NP)[:, :, b.cols(0)]
    assert bool(torch.isfinite(vs).all())
    assert float(vs[:, 7].min()) == float(vs[:, 7].max()) == float(np.float32(1e18))
    assert float(vs[:, :7].max()) < 1.0
