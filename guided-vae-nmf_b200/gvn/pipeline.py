"""Batched whole-utterance enhancement: what ``process_utt`` of the reference's evaluate
scripts does for one file (scripts/evaluate_M2_ibm.py:95-171, evaluate_M1.py:111-166), done
for a batch of utterances per call:

    waveforms -> STFT -> [classifier label] -> init (NMF, encoder) -> niter x (E, M) ->
    Wiener chain -> ISTFT x2 -> waveforms

Every arrow is a libgvn.so kernel; torch only carries the buffers and the copies.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import engine as E


@dataclass
class McemConfig:
    """Constants block of the evaluate scripts (scripts/evaluate_M2_ibm.py:33-38, :71-80)."""
    model: str = "M2"               # "M1" | "M2"
    niter: int = 100
    nsamples_E_step: int = 10
    burnin_E_step: int = 30
    nsamples_WF: int = 25
    burnin_WF: int = 75
    var_RW: float = 0.01
    nmf_rank: int = 10
    eps: float = 1e-8
    fs: int = 16000
    wlen_sec: float = 64e-3
    hop_percent: float = 0.25
    precision: str = "fp32"
    mstep_variant: int = 1

    def chains(self):
        """(R, burnin) of the E-step and Wiener chains, incl. the M1 quirk (mcem.py:461-462)."""
        if self.model == "M1":
            return (self.burnin_E_step, 30), (self.burnin_WF, 30)
        return (self.nsamples_E_step, self.burnin_E_step), (self.nsamples_WF, self.burnin_WF)


class Enhancer:
    """Holds the packed model on one device and enhances batches of utterances."""

    def __init__(self, vae, cfg, device, classifier=None, mean=None, std=None):
        self.device = E._require_cuda(device)
        self.cfg = cfg
        self.vae = vae
        self.classifier = classifier
        self.mean, self.std = mean, std
        with torch.cuda.device(self.device):
            self.dec = E.PackedDecoder(vae, self.device)
        self._batches = {}          # batch state in HBM is allocated once per shape and reused (stream-ordered)

    def upload(self, wavs, labels=None):
        """Host -> device copy of one batch of inputs (pinned staging, async on the stream):
        the waveforms and, for oracle-label M2, the (y_dim, N_b) label arrays."""
        cfg, dev = self.cfg, self.device
        geo = [E.stft_geometry(len(w), cfg.fs, cfg.wlen_sec, cfg.hop_percent) for w in wavs]
        with torch.cuda.device(dev):
            wav, T, T_stride = E.upload_waveforms(wavs, dev)
            y, nbytes = None, wav.numel() * 4
            if cfg.model == "M2" and labels is not None:
                A = E.GVN_FRAME_ALIGN
                off = np.cumsum([0] + [(g[3] + A - 1) // A * A for g in geo])
                host = E.pinned_buffer("labels", (self.dec.y_dim, int(off[-1])))
                hn = host.numpy()
                hn[:] = 0
                for i, l in enumerate(labels):
                    hn[:, off[i]:off[i] + geo[i][3]] = np.asarray(l, dtype=np.float32)
                y = host.to(dev, non_blocking=True)
                nbytes += y.numel() * 4
        return dict(wav=wav, T=T, T_stride=T_stride, geo=geo, y=y, h2d_bytes=nbytes)

    def prepare(self, wavs, labels=None, seed=0, rand=None, uploaded=None):
        """STFT + initialisation of a batch.  ``labels``: None (M1 / classifier) or a list of
        (y_dim, N_b) arrays; ``uploaded``: the result of :meth:`upload` when the inputs are
        already in HBM.  Returns the batch; everything is queued on the current stream."""
        cfg, dev = self.cfg, self.device
        up = uploaded if uploaded is not None else self.upload(wavs, labels)
        geo = up["geo"]
        nfft, hop = geo[0][0], geo[0][1]
        (R_E, _), (R_W, _) = cfg.chains()
        with torch.cuda.device(dev):
            key = (tuple(g[3] for g in geo), nfft // 2 + 1, cfg.nmf_rank, self.dec.L, max(R_E, R_W))
            b = self._batches.get(key)
            if b is None:
                if len(self._batches) >= 4:
                    self._batches.clear()
                b = self._batches[key] = E.Batch(list(key[0]), key[1], key[2], key[3], key[4], dev)
            E.stft_into(b, up["wav"], up["T"], up["T_stride"], nfft, hop, [g[2] for g in geo])
            if rand is None:
                E.init_nmf(b, cfg.eps, generator=torch.Generator(device=dev).manual_seed(int(seed)))
            else:
                E.init_nmf(b, cfg.eps, rand[0], rand[1])
            y = None
            if cfg.model == "M2":
                y = up["y"] if up["y"] is not None else E.classify(b, self.classifier, self.mean, self.std, cfg.eps)
            E.set_labels(b, self.dec, y)
            E.encode_init(b, self.vae)
        b.T, b.T_stride, b.nfft, b.hop = up["T"], up["T_stride"], nfft, hop
        return b

    def run(self, b, seed=0, noise=None, timers=None):
        """The MCEM loop + ISTFT on a prepared batch.  Returns (s_hat, n_hat, cost) device
        tensors: (B, T_stride) f32 x2 and (niter, B) f64."""
        cfg = self.cfg
        cE, cW = cfg.chains()
        with torch.cuda.device(self.device):
            cost, S, Nn, _, _ = E.run_mcem(b, self.dec, cfg.niter, cE, cW, cfg.var_RW, cfg.precision, seed, noise,
                                           cfg.mstep_variant, timers=timers)
            s_hat = E.istft_from(b, S, b.T, b.T_stride, b.nfft, b.hop)
            n_hat = E.istft_from(b, Nn, b.T, b.T_stride, b.nfft, b.hop)
        return s_hat, n_hat, cost

    def enhance(self, wavs, labels=None, seed=0):
        """Host waveforms in, host waveforms out (the end-to-end call)."""
        b = self.prepare(wavs, labels, seed)
        s_hat, n_hat, cost = self.run(b, seed)
        s_hat, n_hat, cost = E.download(s_hat, "s_hat"), E.download(n_hat, "n_hat"), E.download(cost, "cost")
        torch.cuda.current_stream(self.device).synchronize()
        s_hat, n_hat, cost = s_hat.numpy().copy(), n_hat.numpy().copy(), cost.numpy().copy()
        return ([s_hat[i, :b.T[i]] for i in range(b.B)], [n_hat[i, :b.T[i]] for i in range(b.B)], cost)
