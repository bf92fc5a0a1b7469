"""CPU oracle for the MCEM-NMF hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product path (guided-vae-nmf_b200/) never does and fails
loudly when its CUDA library is missing.

What it restates (reference files relative to /root/reference)
--------------------------------------------------------------
* ``python/models/models.py:107-121``   Decoder.forward          -> :func:`decode`
* ``python/models/models.py:90-104``    Encoder.forward (mean)   -> :func:`encode_mean`
* ``python/models/models.py:41-62``     Classifier.forward       -> :func:`classify`
* ``python/models/mcem.py:36-57``       EM.init_parameters       -> :meth:`McemOracle.init_parameters`
* ``python/models/mcem.py:218-294``     MCEM_M2.sample_posterior -> :meth:`McemOracle.sample_posterior`
* ``python/models/mcem.py:371-441``     MCEM_M1.sample_posterior -> same, ``y is None``
* ``python/models/mcem.py:297-307``     compute_Vs               -> :meth:`McemOracle.compute_Vs`
* ``python/models/mcem.py:309-325 / 456-471`` E_step (incl. the M1 positional-argument
  quirk of ``:461-462`` / ``:477-478``)                          -> :meth:`McemOracle.E_step`
* ``python/models/mcem.py:90-152``      EM.M_step                -> :meth:`McemOracle.M_step`
* ``python/models/mcem.py:68-70``       cost                     -> :meth:`McemOracle.cost`
* ``python/models/mcem.py:327-345 / 473-490`` compute_WF         -> :meth:`McemOracle.compute_WF`
* ``python/models/mcem.py:155-178``     EM.run                   -> :meth:`McemOracle.run`
* ``python/models/mcem.py:493-607``     EM_noNMF (fixed noise variance, gain-only M-step)
  and ``:609-760`` MCEM_M2_noNMF                                  -> :class:`McemNoNmfOracle`

It is written with the same torch CPU ops in the same order as the reference (so that in
fp32 it reproduces the reference bit-for-bit when both consume the same noise tape), but
it is a restatement, not a copy: all random draws come from an explicit :class:`NoiseTape`
(SURVEY.md section 8a row R0 lists the consumption order), the state is explicit, and
the chain can return a per-step trace (``acc_prob``, ``is_acc``) or be driven with forced
accept decisions -- the hooks the parity tests need.

Pin status: the reference holds NO test or golden vector for this arithmetic (SURVEY.md
section 8c).  This restatement is therefore pinned against outputs of the reference itself
run in the build container: ``oracle/make_golden.py`` imports ``/root/reference/python/
models/{mcem,models}.py`` unmodified, feeds it the same tape, and commits inputs+outputs as
``tests/golden/mcem_*.npz``; ``tests/test_oracle_golden.py`` checks the oracle against them.
"""
import numpy as np
import torch


# ----------------------------------------------------------------------------------------
# noise tape (row R0 of SURVEY.md section 8a)
# ----------------------------------------------------------------------------------------
class NoiseTape:
    """Replays / records the random draws of one utterance in consumption order.

    ``randn(shape)`` and ``rand(shape)`` mirror the two calls the chain makes per MH step
    (``mcem.py:257`` then ``:271``).  If constructed with ``draws`` it replays them (and
    checks shapes); otherwise it draws from ``generator`` and records.
    """

    def __init__(self, draws=None, seed=0, dtype=torch.float32):
        self.replay = draws is not None
        self.draws = list(draws) if draws is not None else []
        self.pos = 0
        self.dtype = dtype
        self.gen = torch.Generator().manual_seed(seed)

    def _next(self, kind, shape):
        shape = tuple(int(s) for s in shape)
        if self.replay:
            k, t = self.draws[self.pos]
            assert k == kind and tuple(t.shape) == shape, (k, kind, tuple(t.shape), shape)
            self.pos += 1
            return torch.as_tensor(t).to(self.dtype)
        t = (torch.randn(shape, generator=self.gen) if kind == "randn"
             else torch.rand(shape, generator=self.gen))
        self.draws.append((kind, t))
        self.pos += 1
        return t.to(self.dtype)

    def randn(self, *shape):
        return self._next("randn", shape)

    def rand(self, *shape):
        return self._next("rand", shape)


# ----------------------------------------------------------------------------------------
# the three MLPs (weights passed as plain dicts of tensors, keys = reference state-dict keys)
# ----------------------------------------------------------------------------------------
def _lin(x, w, b):
    return torch.nn.functional.linear(x, w, b)


def decode(dec, zin):
    """models.py:118-121 -- tanh hidden layers, exp output.  ``zin`` is (..., D_in)."""
    x = zin
    i = 0
    while ("hidden.%d.weight" % i) in dec:
        x = torch.tanh(_lin(x, dec["hidden.%d.weight" % i], dec["hidden.%d.bias" % i]))
        i += 1
    return torch.exp(_lin(x, dec["reconstruction.weight"], dec["reconstruction.bias"]))


def encode_mean(enc, xin):
    """models.py:101-104 + :32-38 -- returns only the mean head (what MCEM keeps)."""
    x = xin
    i = 0
    while ("hidden.%d.weight" % i) in enc:
        x = torch.tanh(_lin(x, enc["hidden.%d.weight" % i], enc["hidden.%d.bias" % i]))
        i += 1
    return _lin(x, enc["sample.mu.weight"], enc["sample.mu.bias"])


def classify(clf, xin):
    """models.py:57-62 -- ReLU hidden layers, sigmoid output."""
    x = xin
    i = 0
    while ("hidden.%d.weight" % i) in clf:
        x = torch.relu(_lin(x, clf["hidden.%d.weight" % i], clf["hidden.%d.bias" % i]))
        i += 1
    return torch.sigmoid(_lin(x, clf["output_layer.weight"], clf["output_layer.bias"]))


def split_state_dict(sd, prefix):
    """{'decoder.hidden.0.weight': t, ...} -> {'hidden.0.weight': t, ...} for one sub-module."""
    p = prefix + "."
    return {k[len(p):]: torch.as_tensor(v) for k, v in sd.items() if k.startswith(p)}


# ----------------------------------------------------------------------------------------
# the algorithm
# ----------------------------------------------------------------------------------------
class McemOracle:
    """One utterance of MCEM (M1 when ``y is None``, M2 otherwise)."""

    def __init__(self, niter, nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25,
                 burnin_WF=75, var_RW=0.01, model="M2", dtype=torch.float32):
        assert model in ("M1", "M2")
        self.niter = niter
        self.nsamples_E_step = nsamples_E_step
        self.burnin_E_step = burnin_E_step
        self.nsamples_WF = nsamples_WF
        self.burnin_WF = burnin_WF
        self.var_RW = var_RW
        self.model = model
        self.dtype = dtype
        self.trace = None           # set to [] to collect per-step (acc_prob, is_acc)
        self.forced_accept = None   # iterator of bool (N,) tensors overriding the decision
        self.iter_hook = None       # callable(oracle, n) after each EM iteration

    # effective (R, burnin) of the two chains, including the M1 quirk
    def chain_lengths(self):
        if self.model == "M2":
            return (self.nsamples_E_step, self.burnin_E_step), (self.nsamples_WF, self.burnin_WF)
        # mcem.py:461-462 / :477-478: sample_posterior(Z, y=nsamples, nsamples=burnin, burnin=30)
        return (self.burnin_E_step, 30), (self.burnin_WF, 30)

    # mcem.py:36-57 (+ :207-216 / :361-369)
    def init_parameters(self, X, y, dec, enc, nmf_rank, eps, tape, W0=None, H0=None):
        """X: (N,F) complex; y: (N,y_dim) float tensor or None; dec/enc: weight dicts."""
        dt = self.dtype
        N, F = X.shape
        self.tape = tape
        if W0 is None:
            W0 = torch.max(tape.rand(F, nmf_rank), eps * torch.ones(F, nmf_rank, dtype=dt))
            H0 = torch.max(tape.rand(nmf_rank, N), eps * torch.ones(nmf_rank, N, dtype=dt))
        self.X = np.asarray(X).T                                            # (F,N) complex
        self.X_abs_2 = torch.tensor(np.abs(np.asarray(X).T) ** 2).to(dt)    # (F,N)
        self.W = torch.as_tensor(W0).to(dt).clone()
        self.H = torch.as_tensor(H0).to(dt).clone()
        self.Vb = self.W @ self.H
        self.g = torch.ones(N, dtype=dt)
        self.dec = {k: v.to(dt) for k, v in dec.items()}
        self.y = None if y is None else torch.t(torch.as_tensor(y).to(dt))  # (y_dim,N)
        if enc is not None:
            enc = {k: v.to(dt) for k, v in enc.items()}
            xin = self.X_abs_2 if self.y is None else torch.cat([self.X_abs_2, self.y], dim=0)
            self.Z = torch.t(encode_mean(enc, torch.t(xin)))                # (L,N)
        self.Vs = self.Vs_scaled = self.Vx = None

    def _decode_cols(self, Z):
        """decoder applied to the columns of Z (L,N), M2 appends y (mcem.py:242 / :392)."""
        zin = Z if self.y is None else torch.cat([Z, self.y], dim=0)
        return torch.t(decode(self.dec, torch.t(zin)))                       # (F,N)

    # mcem.py:218-294 / :371-441
    def sample_posterior(self, Z, nsamples, burnin):
        L, N = Z.shape
        dt = self.dtype
        sd = torch.sqrt(torch.tensor(np.float32(self.var_RW))).to(dt)
        Zs = torch.zeros(N, nsamples, L, dtype=dt)
        Z_t = Z.clone()
        Vs_t = self._decode_cols(Z_t)
        g_t, Vb_t = self.g.clone(), self.Vb.clone()
        Vx_t = g_t * Vs_t + Vb_t
        cpt = 0
        for m in range(nsamples + burnin):
            Zp = Z_t + sd * self.tape.randn(L, N)
            Vsp = self._decode_cols(Zp)
            Vxp = g_t * Vsp + Vb_t
            acc_prob = (torch.sum(torch.log(Vx_t) - torch.log(Vxp)
                                  + (1 / Vx_t - 1 / Vxp) * self.X_abs_2, 0)
                        + .5 * torch.sum(Z_t.pow(2) - Zp.pow(2), 0))
            logu = torch.log(self.tape.rand(N))
            is_acc = logu < acc_prob
            if self.forced_accept is not None:
                is_acc = torch.as_tensor(next(self.forced_accept)).bool()
            if self.trace is not None:
                self.trace.append((acc_prob.clone(), is_acc.clone(), logu.clone()))
            Z_t[:, is_acc] = Zp[:, is_acc]
            Vs_t = self._decode_cols(Z_t)
            Vx_t = g_t * Vs_t + Vb_t
            if m > burnin - 1:
                Zs[:, cpt, :] = torch.t(Z_t)
                cpt += 1
        return Zs

    # mcem.py:297-307 / :444-454
    def compute_Vs(self, Zs):
        N, R, L = Zs.shape
        if self.y is not None:
            yy = torch.t(self.y).unsqueeze(1).expand(N, R, self.y.shape[0])
            Zs = torch.cat([Zs, yy], dim=2)
        Vs_t = decode(self.dec, Zs)                       # (N,R,F)
        self.Vs = Vs_t.permute(1, 2, 0)                   # (R,F,N)

    def _refresh(self):
        self.Vs_scaled = self.g * self.Vs
        self.Vx = self.Vs_scaled + self.Vb

    # mcem.py:309-325 / :456-471
    def E_step(self):
        (R, burnin), _ = self.chain_lengths()
        Zs = self.sample_posterior(self.Z, R, burnin)
        self.Z = torch.t(Zs[:, -1, :]).clone()
        self.compute_Vs(Zs)
        self._refresh()

    # mcem.py:90-152
    def M_step(self):
        X2 = self.X_abs_2
        num = (X2 * torch.sum(self.Vx ** -2, axis=0)) @ self.H.T
        den = torch.sum(self.Vx ** -1, axis=0) @ self.H.T
        self.W = self.W * (num / den) ** .5
        self.Vb = self.W @ self.H
        self.Vx = self.Vs_scaled + self.Vb
        num = self.W.T @ (X2 * torch.sum(self.Vx ** -2, axis=0))
        den = self.W.T @ torch.sum(self.Vx ** -1, axis=0)
        self.H = self.H * (num / den) ** .5
        self.Vb = self.W @ self.H
        self.Vx = self.Vs_scaled + self.Vb
        c = torch.sum(torch.abs(self.W), axis=0)
        self.W = self.W / c.unsqueeze(0)
        self.H = self.H * c.unsqueeze(1)
        num = torch.sum(X2 * torch.sum(self.Vs * (self.Vx ** -2), axis=0), axis=0)
        den = torch.sum(torch.sum(self.Vs * (self.Vx ** -1), axis=0), axis=0)
        self.g = self.g * (num / den) ** .5
        self._refresh()

    # mcem.py:68-70
    def cost(self):
        return float(torch.mean(torch.log(self.Vx) + self.X_abs_2 / self.Vx))

    # mcem.py:327-345 / :473-490
    def compute_WF(self, sample=True):
        if sample:
            _, (R, burnin) = self.chain_lengths()
            Zs = self.sample_posterior(self.Z, R, burnin)
            self.compute_Vs(Zs)
            self._refresh()
        WFs = torch.mean(self.Vs_scaled / self.Vx, axis=0)
        WFn = torch.mean(self.Vb / self.Vx, axis=0)
        return WFs, WFn

    # mcem.py:155-178
    def run(self):
        cost = np.zeros(self.niter)
        for n in range(self.niter):
            self.E_step()
            self.M_step()
            cost[n] = self.cost()
            if self.iter_hook is not None:
                self.iter_hook(self, n)
        WFs, WFn = self.compute_WF(sample=True)
        self.WFs, self.WFn = WFs, WFn
        self.S_hat = WFs.to(torch.float32).numpy() * self.X
        self.N_hat = WFn.to(torch.float32).numpy() * self.X
        return cost


class McemNoNmfOracle(McemOracle):
    """MCEM_M2_noNMF (mcem.py:609-760) on EM_noNMF (mcem.py:493-607): the noise variance Vb is given
    and stays fixed, the M-step updates the gain only.  The chain, compute_Vs, the cost and the Wiener
    filter are the M2 ones (the reference repeats their code verbatim)."""

    def __init__(self, X, Vb, g, Z, y, dec, niter, tape, nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25,
                 burnin_WF=75, var_RW=0.01, dtype=torch.float32):
        super().__init__(niter, nsamples_E_step, burnin_E_step, nsamples_WF, burnin_WF, var_RW, model="M2", dtype=dtype)
        self.tape = tape
        self.X = np.asarray(X).T                                              # mcem.py:504
        self.X_abs_2 = torch.from_numpy((np.abs(np.asarray(X).T) ** 2).astype(np.float32)).to(dtype)   # :505
        self.Vb = torch.tensor(np.asarray(Vb).T).to(dtype)                    # :508
        self.g = torch.as_tensor(g).to(dtype)                                 # :509
        self.Z = torch.t(torch.as_tensor(Z).to(dtype))                        # :623
        self.y = torch.t(torch.as_tensor(y).to(dtype))                        # :624
        self.dec = {k: v.to(dtype) for k, v in dec.items()}
        self.Vs = self.Vs_scaled = self.Vx = None

    # mcem.py:551-588
    def M_step(self):
        self.Vx = self.Vs_scaled + self.Vb
        num = torch.sum(self.X_abs_2 * torch.sum(self.Vs * (self.Vx ** -2), axis=0), axis=0)
        den = torch.sum(torch.sum(self.Vs * (self.Vx ** -1), axis=0), axis=0)
        self.g = self.g * (num / den) ** .5
        self._refresh()


# ----------------------------------------------------------------------------------------
# labels and metric (small, so kept with the oracle)
# ----------------------------------------------------------------------------------------
def lorenz_mask(power, quantile_fraction=0.98, quantile_weight=0.999):
    """The threshold rule both label functions of processing/target.py share (:18-26, :41-49), on a power array of any shape:
    sort descending, cumulative share (Lorenz curve), threshold = last value whose share is below `quantile_fraction`,
    mask = power > threshold, softened by `quantile_weight` and rounded back to {0, 1}."""
    srt = np.sort(power, axis=None)[::-1]
    lorenz = np.cumsum(srt) / np.sum(srt)
    thr = srt[lorenz < quantile_fraction][-1]
    return np.float32(np.round(0.5 + quantile_weight * ((power > thr) - 0.5)))


def clean_speech_IBM(S, quantile_fraction=0.98, quantile_weight=0.999):
    """processing/target.py:7-27 -- Lorenz-curve threshold on the clean-speech power."""
    return lorenz_mask(np.abs(S * S.conj()), quantile_fraction, quantile_weight)


def clean_speech_VAD(S, quantile_fraction=0.98, quantile_weight=0.999):
    """processing/target.py:29-50 -- same on the per-frame summed power; returns (1,N)."""
    return lorenz_mask(np.abs(S * S.conj()).sum(axis=0), quantile_fraction, quantile_weight)[None]


def energy_ratios(s_hat, s, n):
    """metrics.py:12-60 -- SI-SDR / SI-SIR / SI-SAR with the 3-way projection."""
    a_s = np.dot(s_hat, s) / np.linalg.norm(s) ** 2
    a_n = np.dot(s_hat, n) / np.linalg.norm(n) ** 2
    s_t, e_n = a_s * s, a_n * n
    e_a = s_hat - s_t - e_n
    e = np.linalg.norm(s_t) ** 2
    return (10 * np.log10(e / np.linalg.norm(e_n + e_a) ** 2),
            10 * np.log10(e / np.linalg.norm(e_n) ** 2),
            10 * np.log10(e / np.linalg.norm(e_a) ** 2))
