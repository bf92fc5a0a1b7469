"""MCEM_M2_noNMF (reference mcem.py:609-760: fixed noise variance, gain-only M-step) through gvn_mstep_gain,
against the oracle restatement and the golden run of the reference itself."""
import numpy as np
import pytest
import torch

from _util import load_golden, golden_state_dict, nonmf_oracle_from_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _model(g, forced=None, precision="fp32"):
    from python.models.mcem import MCEM_M2_noNMF
    from python.models.models import DeepGenerativeModel
    F, L = g["X"].shape[1], int(g["L"])
    vae = DeepGenerativeModel([F, 1, L, [128, 128]], None)
    vae.load_state_dict(golden_state_dict(g))
    nE, bE, nW, bW = [int(v) for v in g["chain"]]
    m = MCEM_M2_noNMF(X=g["X"], Vb=g["Vb"], g=torch.from_numpy(g["g0"]), Z=torch.from_numpy(g["Z0"]), y=torch.from_numpy(g["y"]),
                      vae=vae.eval(), niter=int(g["niter"]), device="cuda:0", nsamples_E_step=nE, burnin_E_step=bE,
                      nsamples_WF=nW, burnin_WF=bW, var_RW=float(g["var_RW"]))
    m.precision = precision
    m.replay = dict(eps=g["tape_eps"], u=g["tape_u"], forced=forced)
    return m


def _oracle_run(g):
    o = nonmf_oracle_from_golden(g)
    o.trace = []
    snaps = {}
    o.iter_hook = lambda oo, n: snaps.__setitem__(n, dict(g=oo.g.numpy().copy(), Z=oo.Z.numpy().copy(), Vs=oo.Vs.numpy().copy()))
    cost = o.run()
    return o, cost, np.stack([t[1].numpy() for t in o.trace]), snaps


def test_nonmf_forced_run_matches_golden():
    g = load_golden("M2_noNMF")
    o, cost_o, dec_o, snaps = _oracle_run(g)
    np.testing.assert_array_equal(cost_o, g["cost"])
    m = _model(g, forced=dec_o)
    np.testing.assert_allclose(m.Vb.cpu().numpy(), g["Vb"].T, rtol=0, atol=0)          # the noise variance is an input
    for n in range(int(g["niter"])):
        m.E_step()
        np.testing.assert_allclose(m.Vs.cpu().numpy(), snaps[n]["Vs"], rtol=RTOL)
        np.testing.assert_allclose(m.Z.cpu().numpy(), g["E%d_Z" % n], rtol=RTOL, atol=2e-6)
        m.M_step()
        np.testing.assert_allclose(m.g.cpu().numpy(), g["M%d_g" % n], rtol=RTOL)
        c = float(m.compute_expected_neg_log_like())
        assert abs(c - g["cost"][n]) <= RTOL * abs(g["cost"][n])
        np.testing.assert_allclose(m.Vb.cpu().numpy(), g["Vb"].T, rtol=0, atol=0)      # ... and never changes
    WFs, WFn = m.compute_WF(sample=True)
    np.testing.assert_allclose(WFs.cpu().numpy(), g["WFs"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(WFn.cpu().numpy(), g["WFn"], rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_nonmf_run_api(precision):
    g = load_golden("M2_noNMF")
    o, cost_o, dec_o, snaps = _oracle_run(g)
    m = _model(g, forced=dec_o, precision=precision)
    cost = m.run()
    assert cost.shape == (int(g["niter"]),)
    np.testing.assert_allclose(cost, g["cost"], rtol=RTOL if precision == "fp32" else 2e-3)
    assert m.S_hat.dtype == np.complex64 and m.S_hat.shape == g["S_hat"].shape
    tol = 1e-3 if precision == "fp32" else 2e-2
    np.testing.assert_allclose(m.S_hat, g["S_hat"], rtol=tol, atol=tol * 1e-2 * np.max(np.abs(g["S_hat"])))
    np.testing.assert_allclose(m.S_hat + m.N_hat, g["X"].T, rtol=1e-4, atol=1e-6 * np.max(np.abs(g["X"])))
    with pytest.raises(NameError):
        class RVAE:                                                          # mcem.py:616-617
            pass
        from python.models.mcem import MCEM_M2_noNMF
        MCEM_M2_noNMF(g["X"], g["Vb"], g["g0"], g["Z0"], g["y"], RVAE(), 1, "cuda:0")
