"""Drop-in for the reference's ``python/processing/stft.py``: same ``stft`` / ``istft``
signatures and return types (numpy, ``(freq_bins, frames)`` complex64 / float32 time
signal), computed by the fused framing+window+FFT kernels of libgvn.so on the current CUDA
device.  The librosa semantics the reference selects (stft.py:55-62, :92-98) are described
in csrc/stft.cu.  n_fft must be a power of two (1024 in every evaluate script)."""
import numpy as np
import torch

from gvn import engine as _E


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("gvn: stft/istft run on a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stft(x, fs=16e3, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, pad_mode='reflect',
         pad_at_end=True, dtype='complex64'):
    """Returns Sxx of shape (n_fft/2+1, n_frames); the null frequency is included."""
    if wlen_sec * fs != int(wlen_sec * fs):
        raise ValueError("wlen_sample of STFT is not an integer.")
    if win != 'hann' or not center or pad_mode != 'reflect':
        raise NotImplementedError("only the configuration the evaluate scripts use is implemented "
                                  "(win='hann', center=True, pad_mode='reflect')")
    dev = _device()
    x = np.asarray(x)
    nfft, hop, end_pad, n_frames = _E.stft_geometry(len(x), fs, wlen_sec, hop_percent)
    if not pad_at_end:
        end_pad, n_frames = False, 1 + len(x) // hop
    with torch.cuda.device(dev):
        b = _E.Batch([n_frames], nfft // 2 + 1, 1, 1, 1, dev)
        wav, T, T_stride = _E.upload_waveforms([x], dev)
        _E.stft_into(b, wav, T, T_stride, nfft, hop, [end_pad])
        out = b.Xc[:, b.cols(0), :].cpu().numpy()
    return np.ascontiguousarray(out).view(np.complex64)[..., 0].astype(dtype, copy=False)


def istft(Sxx, fs=16000, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, dtype='float32',
          max_len=None):
    """Inverse STFT of a (freq_bins, frames) spectrogram; pads / trims to ``max_len`` samples."""
    if wlen_sec * fs != int(wlen_sec * fs):
        raise ValueError("wlen_sample of iSTFT is not an integer.")
    if win != 'hann' or not center:
        raise NotImplementedError("only win='hann', center=True is implemented")
    dev = _device()
    nfft = int(wlen_sec * fs)
    hop = int(hop_percent * nfft)
    Sxx = np.ascontiguousarray(np.asarray(Sxx).astype(np.complex64))
    F, N = Sxx.shape
    out_len = int(max_len) if max_len else hop * (N - 1)
    with torch.cuda.device(dev):
        b = _E.Batch([N], F, 1, 1, 1, dev)
        S = torch.zeros(F, b.NP, 2, dtype=torch.float32, device=dev)
        S[:, b.cols(0), :] = torch.from_numpy(Sxx.view(np.float32).reshape(F, N, 2)).to(dev)
        T_stride = (out_len + 3) // 4 * 4
        y = _E.istft_from(b, S, [out_len], T_stride, nfft, hop)[0, :out_len].cpu().numpy()
    y = y.astype(dtype, copy=False)
    if max_len:
        y = y[:int(max_len * fs)]
    return y
