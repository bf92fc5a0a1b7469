"""CPU oracle for the "timo" guide labels.  TEST INFRASTRUCTURE ONLY (see oracle/mcem_oracle.py).

Restates ``SPPNoiseEstimator.update`` (no-``v_spp_in`` branch) and ``timo_mask_estimation`` of the reference's
``python/models/spp_estimation.py:84-141, :198-218`` as one loop over frames on whole-spectrum vectors, in numpy
with the reference's dtypes (float32 periodogram, float64 state).  Pinned by ``tests/golden/spp_mask.npz``, which
``oracle/make_golden.py`` produces by importing the reference module unmodified.
"""
import numpy as np


def timo_mask(spectrogram, fixed_smooth=0.8, prob_smooth=0.9, prior=0.5, snr_opt_db=15, n_init=10):
    """spectrogram: (F, N) noisy power |Y|^2 -> (F, N) speech presence probability, dtype of the input."""
    F, N = spectrogram.shape
    snr = 10.0 ** (snr_opt_db / 10.0)
    glr_factor = (1 - prior) / prior * (1.0 + snr)               # :80
    glr_exp = snr / (1.0 + snr)                                  # :81
    old_psd = np.zeros(F)
    smooth = np.zeros(F)
    mask = np.zeros_like(spectrogram)
    for i in range(N):
        per = spectrogram[:, i]
        if i < n_init:                                           # :98-108
            old_psd = old_psd + per / n_init
            spp = np.zeros(F)
        else:                                                    # :110-136
            inv_glr = glr_factor * np.exp(-per / (old_psd + 1e-8) * glr_exp)
            spp = 1.0 / (1.0 + inv_glr)
            smooth = (1 - prob_smooth) * spp + prob_smooth * smooth
            stuck = smooth > 0.99
            spp[stuck] = np.minimum(spp[stuck], 0.99)
            noise_per = (1.0 - spp) * per + spp * old_psd
            old_psd = (1.0 - fixed_smooth) * noise_per + fixed_smooth * old_psd
        mask[:, i] = spp
    return mask
