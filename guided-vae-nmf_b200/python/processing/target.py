"""Oracle guide labels from the clean speech (reference python/processing/target.py:7-50).

Host-side numpy: these labels are an *input* of the MCEM path in the oracle-label
configuration (scripts/evaluate_M2_ibm.py:132-134), computed once per utterance.  Both
functions rank the time-frequency (or per-frame) power, walk down the ranking until the
requested share of the total energy is covered, and flag everything above that level.
"""
import numpy as np


def _energy_share_mask(power, quantile_fraction, quantile_weight):
    ranked = np.sort(power, axis=None)[::-1]
    share = np.cumsum(ranked) / np.sum(ranked)
    level = ranked[share < quantile_fraction][-1]
    soft = 0.5 + quantile_weight * ((power > level) - 0.5)
    return np.round(soft).astype(np.float32)


def clean_speech_IBM(observations, quantile_fraction=0.98, quantile_weight=0.999):
    """(F,N) complex STFT -> (F,N) float32 mask in {0,1}."""
    return _energy_share_mask(np.abs(observations * observations.conj()), quantile_fraction, quantile_weight)


def clean_speech_VAD(observations, quantile_fraction=0.98, quantile_weight=0.999):
    """(F,N) complex STFT -> (1,N) float32 voice-activity flags."""
    power = np.abs(observations * observations.conj()).sum(axis=0)
    return _energy_share_mask(power, quantile_fraction, quantile_weight)[None]
