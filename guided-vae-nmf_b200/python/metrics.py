"""Scale-invariant SDR / SIR / SAR (reference python/metrics.py:12-60): project the estimate
on the clean speech and on the noise, the remainder is the artefact."""
import numpy as np


def si_sdr_components(s_hat, s, n):
    s_target = (np.dot(s_hat, s) / np.dot(s, s)) * s
    e_noise = (np.dot(s_hat, n) / np.dot(n, n)) * n
    return s_target, e_noise, s_hat - s_target - e_noise


def energy_ratios(s_hat, s, n):
    s_target, e_noise, e_art = si_sdr_components(s_hat, s, n)
    p = np.sum(s_target ** 2)
    db = lambda e: 10 * np.log10(p / np.sum(e ** 2))
    return db(e_noise + e_art), db(e_noise), db(e_art)
