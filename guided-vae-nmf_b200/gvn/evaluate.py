"""The on-disk contract of the reference's evaluate scripts (scripts/evaluate_M1.py:111-166,
evaluate_M2_ibm.py:95-171, evaluate_M2_vad.py:96-174), for a whole file list and batches of utterances:

    <processed_dir>/<stem>_x.wav  (mixture)  [+ <stem>_s.wav (clean speech) for the oracle labels]
        -> <output_dir>/<stem>_s_est.wav, <stem>_n_est.wav            16-bit PCM, as sf.write's default
        -> <output_dir>/<stem> _ibm_soft_est.pt, <stem>_ibm_hard_est.pt   (M2 only; torch.save of the label: hard =
           (N, y_dim) tensor; soft = the classifier's sigmoid output (N, y_dim), or the (y_dim, N) numpy mask for
           oracle labels -- the blank in the first name is the reference's, evaluate_M2_ibm.py:170)

The file list is split across ranks like ``np.array_split(file_paths, nb_devices)`` (evaluate_M1.py:203-207); inside
a rank, consecutive files form batches.  Reading (threads, one per file) of batch i+1 and writing of batch i-1
overlap the enhancement of batch i; the waveforms travel through the pinned staging buffers of
``Enhancer.upload``.  Everything numeric is a libgvn.so kernel (``gvn.pipeline.Enhancer``); this module is file
plumbing only.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import engine as E
from . import wavio
from .shard import shard_list

LABEL_SOURCES = (None, "classifier", "timo", "oracle_ibm", "oracle_vad")


def _stem(file_path):
    return os.path.splitext(file_path)[0]


def _read_item(processed_dir, file_path, need_clean, fs):
    x, fs_x = wavio.read(os.path.join(processed_dir, _stem(file_path) + "_x.wav"))
    if fs_x != fs:
        raise ValueError("%s: sampling rate %d, the model expects %d" % (file_path, fs_x, fs))
    s = None
    if need_clean:
        s, _ = wavio.read(os.path.join(processed_dir, _stem(file_path) + "_s.wav"))
    return x, s


def _write_item(output_dir, file_path, s_hat, n_hat, fs, y_soft, y_hard):
    out = os.path.join(output_dir, _stem(file_path))
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    wavio.write(out + "_s_est.wav", s_hat, fs)
    wavio.write(out + "_n_est.wav", n_hat, fs)
    if y_hard is not None:
        torch.save(y_soft, out + " _ibm_soft_est.pt")
        torch.save(y_hard, out + "_ibm_hard_est.pt")


def evaluate_file_list(enhancer, file_paths, processed_dir, output_dir, label_source=None, batch_size=64, seed=0,
                       quantile_fraction=0.999, quantile_weight=0.999, world=1, rank=0, io_threads=8, progress=None):
    """Enhances this rank's shard of ``file_paths``.  ``label_source``: None (M1), "classifier" (the enhancer's
    supervised classifier), "timo" (speech presence probability of the mixture itself, spp_estimation.py:198-218),
    "oracle_ibm" / "oracle_vad" (from ``<stem>_s.wav``, target.py:7-50).
    Returns the list of (file_path, cost (niter,) float64) of the shard."""
    if label_source not in LABEL_SOURCES:
        raise ValueError("label_source must be one of %r" % (LABEL_SOURCES,))
    cfg = enhancer.cfg
    if (cfg.model == "M2") != (label_source is not None):
        raise ValueError("model %s and label_source %r do not go together" % (cfg.model, label_source))
    if label_source == "classifier" and enhancer.classifier is None:
        raise ValueError("label_source='classifier' needs an Enhancer built with a classifier")
    if label_source in ("timo", "oracle_ibm", "oracle_vad"):
        enhancer.label_source = label_source
        enhancer.quantile_fraction, enhancer.quantile_weight = quantile_fraction, quantile_weight
    files = shard_list(list(file_paths), world, rank)
    groups = [files[i:i + batch_size] for i in range(0, len(files), batch_size)]
    oracle = label_source in ("oracle_ibm", "oracle_vad")
    results = []
    with ThreadPoolExecutor(io_threads) as readers, ThreadPoolExecutor(max(2, io_threads // 2)) as writers:
        def fetch(group):
            return [readers.submit(_read_item, processed_dir, fp, oracle, cfg.fs) for fp in group]

        pending_w = []
        nxt = fetch(groups[0]) if groups else None
        for gi, group in enumerate(groups):
            items = [f.result() for f in nxt]
            nxt = fetch(groups[gi + 1]) if gi + 1 < len(groups) else None        # disk reads of the next batch start now
            wavs = [it[0] for it in items]
            with torch.cuda.device(enhancer.device):
                # oracle labels: the clean speech travels with the batch; STFT, ranking and threshold run on the device
                up = enhancer.upload(wavs, clean=[it[1] for it in items] if oracle else None)
                b = enhancer.prepare(None, None, seed=seed + gi, uploaded=up)
                y_soft = y_hard = None
                if cfg.model == "M2":
                    y_hard = b.y
                    if label_source == "classifier":
                        y_soft = E.classify(b, enhancer.classifier, enhancer.mean, enhancer.std, cfg.eps, hard=False)
                    else:
                        y_soft = b.y_soft if label_source == "timo" else b.y
                s_hat, n_hat, cost = enhancer.run(b, seed=seed + gi)
                tag = str(gi & 1)                                                # two sets of pinned result buffers
                s_h, n_h, c_h = E.download(s_hat, "ev_s" + tag), E.download(n_hat, "ev_n" + tag), E.download(cost, "ev_c" + tag)
                ys = None if y_soft is None else y_soft.to("cpu", non_blocking=False)
                yh = None if y_hard is None else y_hard.to("cpu", non_blocking=False)
                torch.cuda.current_stream(enhancer.device).synchronize()
            for f in pending_w:                                                  # writes of batch i-1 ran during batch i
                f.result()
            pending_w = []
            c_np = c_h.numpy().copy()
            for i, fp in enumerate(group):
                T = b.T[i]
                cols = b.cols(i)
                ysi = None if ys is None else torch.t(ys[:, cols]).contiguous()   # (N, y_dim) as the scripts save them
                if oracle and ysi is not None:
                    ysi = ys[:, cols].numpy().astype(np.float32)                  # oracle: the (y_dim, N) numpy mask itself (evaluate_M2_ibm.py:133, :170)
                yhi = None if yh is None else torch.t(yh[:, cols]).contiguous()
                pending_w.append(writers.submit(_write_item, output_dir, fp, s_h[i, :T].numpy().copy(), n_h[i, :T].numpy().copy(),
                                                cfg.fs, ysi, yhi))
                results.append((fp, c_np[:, i]))
            if progress is not None:
                progress(gi + 1, len(groups))
        for f in pending_w:
            f.result()
    return results
