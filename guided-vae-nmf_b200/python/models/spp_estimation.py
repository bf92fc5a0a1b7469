"""Drop-in for the part of the reference's ``python/models/spp_estimation.py`` that the evaluate scripts call
(``timo_mask_estimation``, scripts/evaluate_M2_ibm.py:136-141; reconstruct_timo_classif.py:96): the speech presence
probability mask from the SPP-based noise tracker, computed by ``gvn_spp_mask`` of libgvn.so on the current CUDA
device.  Same constants as the reference (spp_estimation.py:10-14)."""
import numpy as np
import torch

from gvn import engine as _E

SPP_FIX_SMOOTH = 0.8
SPP_PROB_SMOOTH = 0.9
SPP_PRIOR = 0.5
SPP_SNR_OPT_DB = 15
SPP_NUM_FRAMES_INIT = 10


def timo_mask_estimation(spectrogram):
    """(freq_bins, frames) noisy power spectrogram |Y|^2 -> (freq_bins, frames) mask, dtype of the input."""
    if not torch.cuda.is_available():
        raise RuntimeError("gvn: timo_mask_estimation runs on a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    P = np.asarray(spectrogram)
    F, N = P.shape
    with torch.cuda.device(dev):
        b = _E.Batch([N], F, 1, 1, 1, dev, with_complex=False)
        b.scatter_cols(b.X2, [torch.from_numpy(np.ascontiguousarray(P, dtype=np.float32))])
        soft, _ = _E.spp_mask(b, SPP_FIX_SMOOTH, SPP_PROB_SMOOTH, SPP_PRIOR, SPP_SNR_OPT_DB, SPP_NUM_FRAMES_INIT)
        out = soft[:, b.cols(0)].cpu().numpy()
    return out.astype(P.dtype, copy=False)
