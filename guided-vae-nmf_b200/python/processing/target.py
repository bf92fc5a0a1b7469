"""Drop-in for the oracle guide labels of the reference's ``python/processing/target.py`` (:7-50): same
``clean_speech_IBM`` / ``clean_speech_VAD`` signatures and return types (numpy float32 masks in {0, 1}), computed on
the current CUDA device by ``gvn_speech_labels`` (csrc/labels.cu): power, descending sort, Lorenz share and threshold
per utterance, in numpy's own float32 summation order, so the labels equal the reference's bit for bit.

Only the two functions the evaluate scripts call (scripts/evaluate_M2_ibm.py:132-134) exist here; the training-set
label variants of the reference file are out of scope.  The reference's ``clean_speech_IBM`` also draws
``np.random.rand(N)`` into a variable it overwrites on the next line (target.py:17); that dead draw (it only advances
numpy's global generator, which nothing on the MCEM path reads) is not reproduced.
"""
import numpy as np
import torch

from gvn import engine as _E


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("gvn: the label targets run on a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _labels(observations, quantile_fraction, quantile_weight, vad):
    S = np.ascontiguousarray(np.asarray(observations).astype(np.complex64))
    F, N = S.shape
    dev = _device()
    with torch.cuda.device(dev):
        b = _E.Batch([N], F, 1, 1, 1, dev, with_complex=False)
        Sd = torch.zeros(F, b.NP, 2, dtype=torch.float32, device=dev)
        Sd[:, b.cols(0), :] = torch.from_numpy(S.view(np.float32).reshape(F, N, 2)).to(dev)
        y = _E.speech_labels(b, Sd, vad, quantile_fraction, quantile_weight)
        return y[:, b.cols(0)].cpu().numpy()


def clean_speech_IBM(observations, quantile_fraction=0.98, quantile_weight=0.999):
    """(F,N) complex STFT -> (F,N) float32 mask in {0,1}."""
    return _labels(observations, quantile_fraction, quantile_weight, False)


def clean_speech_VAD(observations, quantile_fraction=0.98, quantile_weight=0.999):
    """(F,N) complex STFT -> (1,N) float32 voice-activity flags."""
    return _labels(observations, quantile_fraction, quantile_weight, True)
