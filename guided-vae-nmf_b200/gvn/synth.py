"""Synthetic noisy utterances for benchmarks and tests.

Utterance ``i`` of a run with seed ``s`` uses ``RandomState(s*100003 + i)``: an AR(1)
"speech-like" signal with a 4 Hz syllabic envelope plus white noise at 0 dB SNR, mixed
and normalised with the recipe the reference uses to build its test set
(``scripts/create_test_set.py:83-103``: gain from the power ratio, then divide by the max
over ``[speech, noise, speech+noise]``).  Host-side numpy only.
"""
import numpy as np
from scipy.signal import lfilter


def synth_utterance(i, seed=0, T=64000, fs=16000, snr_db=0.0):
    """Returns (mixture, speech, noise) as float64 arrays of length T."""
    rng = np.random.RandomState(seed * 100003 + i)
    t = np.arange(T) / fs
    sp = lfilter([1.0], [1.0, -0.95], rng.randn(T)) * (0.1 + np.abs(np.sin(2 * np.pi * 4 * t)))
    sp = sp / np.max(np.abs(sp))
    no = rng.randn(T)
    k = np.sum(sp ** 2) * 10 ** (-snr_db / 10) / np.sum(no ** 2)
    no = no * np.sqrt(k)
    norm = np.max(np.abs(np.concatenate([sp, no, sp + no])))
    return (sp + no) / norm, sp / norm, no / norm


def synth_batch(n, seed=0, T=64000, fs=16000, first=0):
    """Stacks ``n`` utterances: three (n, T) float64 arrays (mixture, speech, noise)."""
    xs, ss, ns = zip(*(synth_utterance(first + i, seed, T, fs) for i in range(n)))
    return np.stack(xs), np.stack(ss), np.stack(ns)
