"""Drop-in for the reference's ``python/models/mcem.py`` -- same class names, constructor and
method signatures and public attributes -- backed by the sm_100a kernels of libgvn.so.

``EM_noNMF`` / ``MCEM_M2_noNMF`` (mcem.py:493-760, fixed noise variance, gain-only M-step) follow at the end.

``MCEM_M1`` / ``MCEM_M2`` keep the single-utterance API of the reference
(``init_parameters`` -> ``run`` -> ``S_hat`` / ``N_hat``; mcem.py:185-216, :350-369, :155-178);
underneath it is a batch of one on the engine in ``gvn.engine``.  The batched entry point
that the throughput numbers use is ``gvn.pipeline.enhance_batch``.

Kept from the reference on purpose
* the M1 positional-argument quirk (mcem.py:461-462, :477-478): ``MCEM_M1`` runs its E-step
  chain with ``R = burnin_E_step`` kept samples after 30 burn-in steps, and its Wiener chain
  with ``R = burnin_WF`` after 30;
* ``NameError('MCEM algorithm only valid for FFNN VAE')`` for a model class named ``RVAE``;
* objects are picklable until ``init_parameters`` (no CUDA handle in ``__init__``), because
  the evaluate scripts ship one object to every worker process.

Different on purpose: there is no CPU path.  ``device`` must be a CUDA device; a missing
libgvn.so raises ImportError.
"""
import numpy as np
import torch

from gvn import engine as _E

_chain_counter = [0]


class EM:
    """State and the EM skeleton shared by M1 and M2 (reference mcem.py:8-178)."""

    def __init__(self, niter=100):
        self.niter = niter
        self._R = None
        self.precision = "fp32"        # "fp32" | "f16"  (decoder arithmetic)
        self.mstep_variant = 1
        self.seed = None               # Philox seed; None -> derived from torch's seed
        self.replay = None             # dict(rand_W, rand_H, eps, u[, forced]) for parity runs
        self._batch = None

    # ---- reference attribute surface (views of the batch state, (F,N) orientation) ----
    def _cols(self, t):
        return t[..., self._batch.cols(0)]

    W = property(lambda self: self._batch.W[0])
    H = property(lambda self: self._cols(self._batch.H))
    g = property(lambda self: self._cols(self._batch.g))
    Z = property(lambda self: self._cols(self._batch.Z))
    Vb = property(lambda self: self._cols(self._batch.Vb))
    X_abs_2 = property(lambda self: self._cols(self._batch.X2))

    @property
    def Vs(self):
        """(R,F,N) speech variance of the kept samples (mcem.py:307), expanded from the slot form."""
        return None if self._R is None or self._batch is None else self._batch.expand_samples(self._R, 0)

    @property
    def Vs_scaled(self):
        return None if self.Vs is None else self.g * self.Vs

    @property
    def Vx(self):
        return None if self.Vs is None else self.g * self.Vs + self.Vb

    def chain_lengths(self):
        return ((self.nsamples_E_step, self.burnin_E_step), (self.nsamples_WF, self.burnin_WF))

    # ---- mcem.py:36-57 + :207-216 / :361-369 ----
    def _init_common(self, X, y, vae, nmf_rank, eps, device):
        if type(vae).__name__ == "RVAE":
            raise NameError("MCEM algorithm only valid for FFNN VAE")
        dev = _E._require_cuda(device)
        self.device = device
        X = np.asarray(X)
        N, F = X.shape
        cached = getattr(self, "_dec_cache", None)
        if cached is None or cached[0] is not vae or cached[1].device != dev:
            self._dec_cache = (vae, _E.PackedDecoder(vae, dev))
        dec = self._dec_cache[1]
        (R_E, _), (R_W, _) = self.chain_lengths()
        with torch.cuda.device(dev):
            b = _E.Batch([N], F, nmf_rank, dec.L, max(R_E, R_W), dev)
            self._batch, self._dec, self.vae = b, dec, vae
            self.X = X.T                                                      # (F,N) complex, host
            Xt = np.ascontiguousarray(X.T.astype(np.complex64))
            b.Xc[:, b.cols(0), :] = torch.from_numpy(Xt.view(np.float32).reshape(F, N, 2)).to(dev)
            b.scatter_cols(b.X2, [torch.from_numpy(np.abs(Xt) ** 2)])
            rp = self.replay
            if rp is not None:
                _E.init_nmf(b, eps, [rp["rand_W"]], [rp["rand_H"]])
            else:
                _E.init_nmf(b, eps)
            yd = None
            if y is not None:
                yd = torch.zeros(dec.y_dim, b.NP, dtype=torch.float32, device=dev)
                b.scatter_cols(yd, [torch.t(torch.as_tensor(y)).float()])
                self.y = yd[:, b.cols(0)]
            _E.set_labels(b, dec, yd)
            _E.encode_init(b, vae)
        self.X_abs_2_t = self.X_abs_2
        self._R = None
        self._iter = 0
        self._scratch = None
        self._chain = 0
        self._replay_pos = 0
        if self.seed is None:
            _chain_counter[0] += 1
            self._seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _chain_counter[0]) & (2 ** 63 - 1)
        else:
            self._seed = int(self.seed)

    def _run_chain(self, R, burnin, Z=None, trace=False):
        b = self._batch
        if Z is not None:
            b.scatter_cols(b.Z, [Z])
        eps = u = forced = None
        if self.replay is not None:
            steps, p = R + burnin, self._replay_pos
            f32 = dict(dtype=torch.float32, device=b.device)
            eps = torch.zeros(steps, b.L, b.NP, **f32)
            u = torch.full((steps, b.NP), 0.5, **f32)
            b.scatter_cols(eps, [torch.as_tensor(self.replay["eps"][p:p + steps])])
            b.scatter_cols(u, [torch.as_tensor(self.replay["u"][p:p + steps])])
            if self.replay.get("forced") is not None:
                forced = torch.zeros(steps, b.NP, dtype=torch.uint8, device=b.device)
                b.scatter_cols(forced, [torch.as_tensor(self.replay["forced"][p:p + steps]).to(torch.uint8)])
            self._replay_pos += steps
        with torch.cuda.device(b.device):
            out = _E.estep(b, self._dec, burnin, R, self.var_RW, self.precision, self._seed, self._chain,
                           eps, u, forced, trace)
        self._chain += 1
        self._R = R
        if trace:
            acc, dec_, cnt, zs = out
            self.last_trace = dict(acc_prob=acc[:, b.cols(0)], accepted=dec_[:, b.cols(0)],
                                   n_accepted=cnt[b.cols(0)], z_samples=zs[..., b.cols(0)])
        return out

    # ---- mcem.py:309-325 / :456-471 ----
    def E_step(self):
        (R, burnin), _ = self.chain_lengths()
        self._run_chain(R, burnin)

    # ---- mcem.py:90-152 (+ the cost of :68-70, which the same kernels produce) ----
    def M_step(self):
        b = self._batch
        if self._scratch is None:
            self._scratch = _E.MstepScratch(b, max(1, self.niter))
        with torch.cuda.device(b.device):
            _E.mstep(b, self._R, self._scratch, self._iter % max(1, self.niter), self.mstep_variant)
        self._last_iter = self._iter % max(1, self.niter)
        self._iter += 1

    def compute_expected_neg_log_like(self):
        """Cost of the last M-step as a 0-dim float64 device tensor (mcem.py:68-70)."""
        b = self._batch
        with torch.cuda.device(b.device):
            c = _E.cost_reduce(b, self._R, self._scratch, max(1, self.niter))
        return c[self._last_iter, 0]

    # ---- mcem.py:327-345 / :473-490 ----
    def compute_WF(self, sample=False):
        b = self._batch
        if sample:
            _, (R, burnin) = self.chain_lengths()
            self._run_chain(R, burnin)
        with torch.cuda.device(b.device):
            S, Nn, WFs, WFn = _E.wiener(b, self._R, want_masks=True)
        self._S, self._N = S, Nn
        return WFs[:, b.cols(0)], WFn[:, b.cols(0)]

    # ---- mcem.py:155-178 ----
    def run(self):
        b = self._batch
        self._iter = 0
        self._scratch = _E.MstepScratch(b, max(1, self.niter))
        for _ in range(self.niter):
            self.E_step()
            self.M_step()
        self.compute_WF(sample=True)
        with torch.cuda.device(b.device):
            cost = _E.cost_reduce(b, self.chain_lengths()[0][0], self._scratch, self.niter) if self.niter else None
        c = b.cols(0)
        to_c = lambda t: np.ascontiguousarray(t[:, c].cpu().numpy()).view(np.complex64)[..., 0]
        self.S_hat = to_c(self._S)
        self.N_hat = to_c(self._N)
        return cost[:, 0].cpu().numpy() if self.niter else np.zeros(0)


class MCEM_M2(EM):
    """Label-conditioned model (reference mcem.py:181-345)."""

    def __init__(self, niter, nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25, burnin_WF=75, var_RW=0.01):
        super().__init__(niter=niter)
        self.nsamples_E_step = nsamples_E_step
        self.burnin_E_step = burnin_E_step
        self.nsamples_WF = nsamples_WF
        self.burnin_WF = burnin_WF
        self.var_RW = var_RW

    def init_parameters(self, X, y, vae, nmf_rank, eps, device):
        self._init_common(X, y, vae, nmf_rank, eps, device)

    def sample_posterior(self, Z, y, nsamples=10, burnin=30):
        """Returns (Z_sampled (N,R,L), Z_sampled_y (N,R,L+y_dim)) like mcem.py:218-294."""
        out = self._run_chain(nsamples, burnin, Z=Z, trace=True)
        zs = self.last_trace["z_samples"].permute(2, 0, 1).contiguous()        # (N,R,L)
        yy = torch.t(self.y).unsqueeze(1).expand(zs.shape[0], zs.shape[1], self.y.shape[0])
        return zs, torch.cat([zs, yy], dim=2)


class MCEM_M1(EM):
    """Unconditioned VAE (reference mcem.py:348-490), including its chain-length quirk."""

    def __init__(self, niter, nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25, burnin_WF=75, var_RW=0.01):
        super().__init__(niter=niter)
        self.nsamples_E_step = nsamples_E_step
        self.burnin_E_step = burnin_E_step
        self.nsamples_WF = nsamples_WF
        self.burnin_WF = burnin_WF
        self.var_RW = var_RW

    def chain_lengths(self):
        # sample_posterior(self.Z, self.nsamples_*, self.burnin_*) binds (y, nsamples) and leaves
        # burnin at its default 30 -- mcem.py:461-462, :477-478
        return ((self.burnin_E_step, 30), (self.burnin_WF, 30))

    def init_parameters(self, X, vae, nmf_rank, eps, device):
        self._init_common(X, None, vae, nmf_rank, eps, device)

    def sample_posterior(self, Z, y, nsamples=10, burnin=30):
        """Same signature as mcem.py:371 (``y`` is unused there too).  Returns Z_sampled (N,R,L)."""
        self._run_chain(nsamples, burnin, Z=Z, trace=True)
        return self.last_trace["z_samples"].permute(2, 0, 1).contiguous()


class EM_noNMF(EM):
    """Reference mcem.py:493-607: the noise variance ``Vb`` is given and fixed, the M-step updates the gain
    only (``gvn_mstep_gain``).  Same constructor as the reference (everything arrives in ``__init__``; X and Vb
    are (N,F))."""

    def __init__(self, X, Vb, g, vae, niter=100, device="cpu"):
        EM.__init__(self, niter=niter)
        self._ctor = dict(X=X, Vb=Vb, g=g, vae=vae, device=device)

    def _init_fixed_noise(self, Z, y):
        c = self._ctor
        vae, device = c["vae"], c["device"]
        if type(vae).__name__ == "RVAE":
            raise NameError("MCEM algorithm only valid for FFNN VAE")
        dev = _E._require_cuda(device)
        self.device = device
        X = np.asarray(c["X"])
        N, F = X.shape
        dec = _E.PackedDecoder(vae, dev)
        (R_E, _), (R_W, _) = self.chain_lengths()
        f32 = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            b = _E.Batch([N], F, 1, dec.L, max(R_E, R_W), dev)          # K = 1: W / H are allocated but never used
            self._batch, self._dec, self.vae = b, dec, vae
            self.X = X.T
            Xt = np.ascontiguousarray(X.T.astype(np.complex64))
            b.Xc[:, b.cols(0), :] = torch.from_numpy(Xt.view(np.float32).reshape(F, N, 2)).to(dev)
            b.scatter_cols(b.X2, [torch.from_numpy((np.abs(X.T) ** 2).astype(np.float32))])       # mcem.py:505
            b.X2t.copy_(b.X2.reshape(F, b.NP // b.X2t.shape[2], b.X2t.shape[2]).permute(1, 0, 2))
            b.scatter_cols(b.Vb, [torch.as_tensor(np.asarray(c["Vb"]).T).to(**f32)])              # :508
            b.scatter_cols(b.g, [torch.as_tensor(c["g"]).to(**f32).reshape(-1)])                  # :509
            b.scatter_cols(b.Z, [torch.t(torch.as_tensor(Z).to(**f32))])                          # :623
            yd = torch.zeros(dec.y_dim, b.NP, **f32)
            b.scatter_cols(yd, [torch.t(torch.as_tensor(y).to(**f32))])                           # :624
            self.y = yd[:, b.cols(0)]
            _E.set_labels(b, dec, yd)
        self.X_abs_2_t = self.X_abs_2
        self._R = None
        self._iter = 0
        self._scratch = None
        self._chain = 0
        self._replay_pos = 0
        if self.seed is None:
            _chain_counter[0] += 1
            self._seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _chain_counter[0]) & (2 ** 63 - 1)
        else:
            self._seed = int(self.seed)

    W = property(lambda self: None)
    H = property(lambda self: None)

    # ---- mcem.py:551-588 (+ the cost of :530-532) ----
    def M_step(self):
        b = self._batch
        if self._scratch is None:
            self._scratch = _E.MstepScratch(b, max(1, self.niter))
        with torch.cuda.device(b.device):
            _E.mstep_gain(b, self._R, self._scratch, self._iter % max(1, self.niter))
        self._last_iter = self._iter % max(1, self.niter)
        self._iter += 1


class MCEM_M2_noNMF(EM_noNMF):
    """Reference mcem.py:609-760.  Like the reference's constructor (mcem.py:505-509, :623-624) this one places the
    state on the device right away.  ``replay`` (dict(eps, u[, forced])) and ``precision`` may be set on the object
    before the first chain, as for the other classes."""

    def __init__(self, X, Vb, g, Z, y, vae, niter, device, nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25,
                 burnin_WF=75, var_RW=0.01):
        super().__init__(X=X, Vb=Vb, g=g, vae=vae, niter=niter, device=device)
        if type(vae).__name__ == "RVAE":
            raise NameError("MCEM algorithm only valid for FFNN VAE")
        self.nsamples_E_step = nsamples_E_step
        self.burnin_E_step = burnin_E_step
        self.nsamples_WF = nsamples_WF
        self.burnin_WF = burnin_WF
        self.var_RW = var_RW
        self._init_fixed_noise(Z, y)

    def sample_posterior(self, Z, y, nsamples=10, burnin=30):
        self._run_chain(nsamples, burnin, Z=Z, trace=True)
        zs = self.last_trace["z_samples"].permute(2, 0, 1).contiguous()
        yy = torch.t(self.y).unsqueeze(1).expand(zs.shape[0], zs.shape[1], self.y.shape[0])
        return zs, torch.cat([zs, yy], dim=2)
