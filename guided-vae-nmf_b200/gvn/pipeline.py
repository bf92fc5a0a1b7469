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

    def __init__(self, vae, cfg, device, classifier=None, mean=None, std=None, label_source=None,
                 quantile_fraction=0.999, quantile_weight=0.999):
        """``label_source`` for M2 when no labels are passed in: None = the classifier (evaluate_M2_ibm.py:121-131),
        "timo" = the speech-presence-probability mask (:136-141; needs y_dim == F), "oracle_ibm" / "oracle_vad" = the
        oracle labels of the clean speech uploaded with the batch (``clean=`` of :meth:`upload`, or the clean-speech
        metric reference; :132-134, python/processing/target.py:7-50 with the scripts' quantiles 0.999 / 0.999)."""
        self.quantile_fraction, self.quantile_weight = quantile_fraction, quantile_weight
        self.device = E._require_cuda(device)
        self.label_source = label_source
        self.cfg = cfg
        self.vae = vae
        self.classifier = classifier
        # standardisation constants of the classifier input live on the device from the start (a host array
        # here would mean a blocking pageable copy in every classify call)
        to_dev = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32).reshape(-1).to(self.device)
        self.mean, self.std = to_dev(mean), to_dev(std)
        with torch.cuda.device(self.device):
            self.dec = E.PackedDecoder(vae, self.device)
        self._batches = {}          # batch state in HBM is allocated once per shape and reused (stream-ordered)
        import threading
        self._batches_lock = threading.Lock()   # upload() may run on the packing thread of enhance_many

    def _batch_for(self, geo):
        """Device state for a batch of this geometry, allocated once per shape and reused (stream-ordered)."""
        cfg = self.cfg
        (R_E, _), (R_W, _) = cfg.chains()
        key = (tuple(g[3] for g in geo), geo[0][0] // 2 + 1, cfg.nmf_rank, self.dec.L, max(R_E, R_W))
        with self._batches_lock:
            b = self._batches.pop(key, None)
            if b is None:
                while len(self._batches) >= 4:              # every batch of a real file list has its own geometry
                    self._batches.pop(next(iter(self._batches)))     # the least recently used one goes
                b = E.Batch(list(key[0]), key[1], key[2], key[3], key[4], self.device)
            self._batches[key] = b                          # most recently used last
        b.use_on(torch.cuda.current_stream(self.device))    # memory is reclaimed in the order of EVERY stream that used it
        return b

    def upload(self, wavs, labels=None, refs=None, slot=0, clean=None):
        """Host -> device copy of one batch of inputs (pinned staging, async on the current stream): the
        waveforms and, for oracle-label M2, the (y_dim, N_b) label arrays.  Binary labels travel as one
        byte per bin and are widened on the device.  ``refs`` = (clean speech, noise) arrays (B, T) for the
        quality metrics.  ``slot`` names the staging buffers (two slots allow an upload to overlap the
        enhancement of the previous batch, see :meth:`enhance_many`)."""
        cfg, dev = self.cfg, self.device
        geo = [E.stft_geometry(len(w), cfg.fs, cfg.wlen_sec, cfg.hop_percent) for w in wavs]
        with torch.cuda.device(dev):
            wav, T, T_stride = E.upload_waveforms(wavs, dev, tag="wav%d" % slot)
            y, nbytes = None, wav.numel() * 4
            if cfg.model == "M2" and labels is not None:
                A = E.GVN_FRAME_ALIGN
                off = np.cumsum([0] + [(g[3] + A - 1) // A * A for g in geo])
                binary = all(getattr(l, "dtype", None) == np.uint8 for l in labels)    # 0/1 masks may be passed as bytes
                dt = torch.uint8 if binary else torch.float32
                host = E.pinned_buffer("labels%d" % slot, (self.dec.y_dim, int(off[-1])), dt)
                hn = host.numpy()
                for i, l in enumerate(labels):
                    hn[:, off[i]:off[i] + geo[i][3]] = l
                    hn[:, off[i] + geo[i][3]:off[i + 1]] = 0
                yd = E.device_buffer("labels%d" % slot, host.shape, dt, dev)
                yd.copy_(host, non_blocking=True)
                y = yd.float() if binary else yd
                nbytes += host.numel() * host.element_size()
            out = dict(wav=wav, T=T, T_stride=T_stride, geo=geo, y=y, h2d_bytes=nbytes)
            if clean is not None:                           # clean speech for the oracle labels (label_source "oracle_*")
                cw, _, cs = E.upload_waveforms(clean, dev, tag="clean%d" % slot)
                assert cs == T_stride, "clean speech and mixture must have the same lengths"
                out["clean"] = cw
                out["h2d_bytes"] += cw.numel() * 4
            if refs is not None:
                for name, arr in zip(("ref_s", "ref_n"), refs):
                    host = E.pinned_buffer("%s%d" % (name, slot), (len(wavs), T_stride))
                    if isinstance(arr, np.ndarray) and arr.ndim == 2:            # equal lengths: one vectorised copy (f64 -> f32 on the way)
                        host[:, :arr.shape[1]].copy_(torch.from_numpy(np.ascontiguousarray(arr)))
                    else:                                                        # ragged list (numpy assignments: see upload_waveforms)
                        hn = host.numpy()
                        for i, a in enumerate(arr):
                            hn[i, :len(a)] = a
                    d = E.device_buffer("%s%d" % (name, slot), host.shape, torch.float32, dev)
                    d.copy_(host, non_blocking=True)
                    out[name] = d
                    out["h2d_bytes"] += host.numel() * 4
            # Oracle labels are made HERE, on the stream of the upload (in enhance_many: the copy stream, which idles while
            # the compute stream runs the EM loop of the previous batch), from the clean speech that has just arrived:
            # STFT, ranking and threshold per utterance on the device (gvn_speech_labels).  They travel on as up["y"].
            if cfg.model == "M2" and y is None and self.label_source in ("oracle_ibm", "oracle_vad"):
                src = out.get("clean", out.get("ref_s"))
                if src is None:
                    raise ValueError("label_source=%r needs the clean speech of the batch (clean= or refs=)" % self.label_source)
                # only the frame index arrays are read: a batch of its own, allocated, used and dropped on this stream
                # (the full batch state is allocated by prepare() on the compute stream)
                b = E.Batch([g[3] for g in geo], geo[0][0] // 2 + 1, 1, 1, 1, dev, index_only=True)
                Sc = E.device_buffer("clean_Xc", (b.F, b.NP, 2), torch.float32, dev)
                P2 = E.device_buffer("clean_X2", (b.F, b.NP), torch.float32, dev)
                E.stft_to(b, src, T, T_stride, geo[0][0], geo[0][1], [g[2] for g in geo], Sc, P2)
                out["y"] = E.speech_labels(b, Sc, self.label_source == "oracle_vad", self.quantile_fraction, self.quantile_weight)
        return out

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
            b = self._batch_for(geo)
            E.stft_into(b, up["wav"], up["T"], up["T_stride"], nfft, hop, [g[2] for g in geo])
            if rand is None:
                E.init_nmf(b, cfg.eps, generator=torch.Generator(device=dev).manual_seed(int(seed)))
            else:
                E.init_nmf(b, cfg.eps, rand[0], rand[1])
            y = None
            if cfg.model == "M2":
                if up["y"] is not None:
                    y = up["y"]
                elif self.label_source == "timo":
                    b.y_soft, y = E.spp_mask(b)
                else:
                    y = E.classify(b, self.classifier, self.mean, self.std, cfg.eps)
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

    def enhance_many(self, batches, seed=0, device_hook=None):
        """Pipelined end-to-end enhancement of a sequence of batches.  ``batches`` yields dicts with keys
        ``wavs``, optionally ``labels`` and ``refs`` (clean speech, noise: adds the quality metrics).
        Three things overlap: the GPU enhances batch i while the host packs batch i+1 into pinned memory and a
        second stream copies it to the device, and the kernels of batch i+1 are queued behind those of batch i
        before the host waits for the results of batch i -- the GPU never drains between batches.  Nothing in
        the loop orders the host behind the compute stream except the wait for a finished batch.  Results come
        back in pinned host buffers (two sets, valid until the next iteration).  Yields dicts with ``s_hat``,
        ``n_hat`` (B, T_stride) f32, ``cost`` (niter, B) f64, ``metrics`` (B, 3) f64 or None, ``T``, byte counts.
        ``device_hook(i, cost_dev, metrics_dev)`` is called right after the kernels of batch i have been
        queued (e.g. to queue a collective on the result rows, in stream order, without touching the host copies)."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(dev)
        consumed = [None, None]                             # main-stream events: device inputs of slot p have been read
        copied = [None, None]                               # copy-stream events: pinned staging of slot p has been read

        def start_upload(item, slot):
            if copied[slot] is not None:
                copied[slot].synchronize()                  # the staging buffers are about to be overwritten by the host
            with torch.cuda.stream(copy):
                if consumed[slot] is not None:
                    copy.wait_event(consumed[slot])
                up = self.upload(item["wavs"], item.get("labels"), item.get("refs"), slot=slot)
                ev = torch.cuda.Event()
                ev.record(copy)
            copied[slot] = ev
            return up, ev

        def launch(up, ev, i):
            slot = i & 1
            main.wait_event(ev)
            if up["y"] is not None:
                up["y"].record_stream(main)                 # allocated on the copy stream, read on this one
            b = self.prepare(None, None, seed=seed + i, uploaded=up)
            s_hat, n_hat, cost = self.run(b, seed=seed + i)
            metrics = metrics_dev = None
            if "ref_s" in up:
                metrics_dev = E.energy_ratios(s_hat, up["ref_s"], up["ref_n"], b.T)
                metrics = E.download(metrics_dev, "metrics%d" % slot)
            if device_hook is not None:
                device_hook(i, cost, metrics_dev)
            consumed[slot] = torch.cuda.Event()
            consumed[slot].record(main)
            out = dict(s_hat=E.download(s_hat, "s_hat%d" % slot), n_hat=E.download(n_hat, "n_hat%d" % slot),
                       cost=E.download(cost, "cost%d" % slot), metrics=metrics, T=b.T, h2d_bytes=up["h2d_bytes"],
                       d2h_bytes=(s_hat.numel() + n_hat.numel()) * 4 + cost.numel() * 8 + (0 if metrics is None else metrics.numel() * 8))
            done = torch.cuda.Event()
            done.record(main)
            return out, done

        # Packing (float64 -> float32 into pinned memory, a few tens of milliseconds per batch and far more when eight
        # ranks share the host's cores) runs on a worker thread, so that it overlaps not only the GPU but also this
        # thread's own work: queueing the ~1300 kernel launches of the batch before it.  torch copies and the ctypes
        # calls into libgvn.so release the interpreter lock.
        from concurrent.futures import ThreadPoolExecutor
        pool = getattr(self, "_pack_pool", None)
        if pool is None:
            pool = self._pack_pool = ThreadPoolExecutor(1, thread_name_prefix="gvn-pack")

        def upload_job(item, slot):
            with torch.cuda.device(dev):
                return start_upload(item, slot)

        it = iter(batches)
        nxt = next(it, None)
        pending = pool.submit(upload_job, nxt, 0) if nxt is not None else None
        in_flight = None                                    # (results, event) of the batch queued before the current one
        i = 0
        while pending is not None:
            up, ev = pending.result()
            nxt = next(it, None)                            # pack + copy the next batch while this thread queues the kernels of this one
            pending = pool.submit(upload_job, nxt, (i + 1) & 1) if nxt is not None else None
            cur = launch(up, ev, i)                         # queued behind the batch in flight
            if in_flight is not None:
                in_flight[1].synchronize()
                yield in_flight[0]
            in_flight = cur
            i += 1
        if in_flight is not None:
            in_flight[1].synchronize()
            yield in_flight[0]

    def enhance(self, wavs, labels=None, seed=0):
        """Host waveforms in, host waveforms out (the end-to-end call)."""
        b = self.prepare(wavs, labels, seed)
        s_hat, n_hat, cost = self.run(b, seed)
        s_hat, n_hat, cost = E.download(s_hat, "s_hat"), E.download(n_hat, "n_hat"), E.download(cost, "cost")
        torch.cuda.current_stream(self.device).synchronize()
        s_hat, n_hat, cost = s_hat.numpy().copy(), n_hat.numpy().copy(), cost.numpy().copy()
        return ([s_hat[i, :b.T[i]] for i in range(b.B)], [n_hat[i, :b.T[i]] for i in range(b.B)], cost)
