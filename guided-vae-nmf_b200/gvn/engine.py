"""Host side of the hot path: batch state in HBM and the calls into libgvn.so.

PyTorch is used for device memory, streams and host<->device copies only; every arithmetic
step of the path (STFT, encoder/classifier layers, MH chains, NMF updates, Wiener filter,
ISTFT) is a kernel of libgvn.so reached through the C ABI in include/gvn.h.

The unit of work is a *batch* of utterances laid side by side on one padded global frame
axis (see include/gvn.h).  The reference API (one utterance per ``init_parameters``/``run``,
python/models/mcem.py:155-178, :207-216) is the B=1 view provided by ``python/models/mcem.py``
in this package.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import GvnBatch, GvnNoise, GvnTrace, GVN_FRAME_ALIGN, GVN_HIDDEN, PRECISIONS, check


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("gvn: the MCEM hot path runs on CUDA (sm_100a) only; got device %r. "
                           "There is no CPU fallback." % (device,))
    return device


# --------------------------------------------------------------------------------------
# decoder / encoder / classifier weights
# --------------------------------------------------------------------------------------
def latent_dim(vae):
    """mcem.py:224-227 -- `latent_dim` first, then `z_dim`."""
    if hasattr(vae, "latent_dim"):
        return int(vae.latent_dim)
    return int(vae.z_dim)


class PackedDecoder:
    """Decoder weights (models.py:107-121) repacked once per model for the kernels."""

    def __init__(self, vae, device):
        if type(vae).__name__ == "RVAE":                        # mcem.py:208-209 / :362-363
            raise NameError("MCEM algorithm only valid for FFNN VAE")
        device = _require_cuda(device)
        lib = _lib.load()
        dec = vae.decoder
        if len(dec.hidden) != 2:
            raise _lib.GvnError(_lib.E_UNSUPPORTED_SHAPE, "decoder must have two hidden layers")
        f32 = dict(device=device, dtype=torch.float32)
        w = [dec.hidden[0].weight, dec.hidden[0].bias, dec.hidden[1].weight, dec.hidden[1].bias,
             dec.reconstruction.weight, dec.reconstruction.bias]
        w = [t.detach().to(**f32).contiguous() for t in w]
        self.L = latent_dim(vae)
        self.hidden = w[0].shape[0]
        self.y_dim = w[0].shape[1] - self.L
        self.F = w[4].shape[0]
        self.device = device
        nbytes = lib.gvn_decoder_packed_bytes(self.L, self.y_dim, self.F, self.hidden)
        if nbytes == 0:
            raise _lib.GvnError(_lib.E_UNSUPPORTED_SHAPE,
                                "unsupported decoder shape L=%d y_dim=%d F=%d hidden=%d (hidden must be %d, L<=%d)"
                                % (self.L, self.y_dim, self.F, self.hidden, GVN_HIDDEN, _lib.GVN_MAX_L))
        self.packed = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            check(lib.gvn_pack_decoder(*[_ptr(t) for t in w], self.L, self.y_dim, self.F, self.hidden,
                                       _ptr(self.packed), _stream()))
        self._keep = w


_PARAMS = {}


def _device_param(p, device):
    """Device f32 copy of a module parameter, made once per (parameter, in-place version): a pageable
    host-to-device copy per call would order the host behind the whole stream (see const_i32)."""
    if p.device == torch.device(device) and p.dtype == torch.float32:
        return p.detach().contiguous()
    key = (id(p), str(device))
    hit = _PARAMS.get(key)
    if hit is None or hit[0] is not p or hit[1] != p._version:
        if len(_PARAMS) >= 256:
            _PARAMS.clear()
        hit = _PARAMS[key] = (p, p._version, p.detach().to(device=device, dtype=torch.float32).contiguous())
    return hit[2]


def _linear_params(layer, device):
    return _device_param(layer.weight, device), _device_param(layer.bias, device)


def dense(W, b, in0, in1, act, NP, mean=None, std=None, eps=0.0):
    """out[j][n] = act(b[j] + sum_i W[j][i] [in0;in1][i][n]) on feature-major activations."""
    lib = _lib.load()
    D0 = 0 if in0 is None else in0.shape[0]
    D1 = 0 if in1 is None else in1.shape[0]
    assert W.shape[1] == D0 + D1, (W.shape, D0, D1)
    out = torch.empty(W.shape[0], NP, dtype=torch.float32, device=W.device)
    acts = {"none": 0, "tanh": 1, "relu": 2, "sigmoid": 3, "hard": 4}
    check(lib.gvn_dense(_ptr(W), _ptr(b), _ptr(in0), D0, _ptr(in1), D1, _ptr(mean), _ptr(std), float(eps),
                        W.shape[0], NP, acts[act], _ptr(out), _stream()))
    return out


# --------------------------------------------------------------------------------------
# batch state
# --------------------------------------------------------------------------------------
class Batch:
    """Device state of B utterances (layout documented in include/gvn.h)."""

    def __init__(self, n_frames, F, K, L, R_cap, device, with_complex=True, index_only=False):
        """``index_only``: only the frame index arrays (frame_off, n_frames, frame_utt) are allocated -- enough for the
        entry points that take their arrays as separate arguments (gvn_stft_power through stft_to, gvn_speech_labels)."""
        device = _require_cuda(device)
        self.device = device
        n_frames = [int(n) for n in n_frames]
        assert len(n_frames) > 0 and min(n_frames) > 0
        self.B, self.F, self.K, self.L, self.R_cap = len(n_frames), int(F), int(K), int(L), int(R_cap)
        A = GVN_FRAME_ALIGN
        off = [0]
        for n in n_frames:
            off.append(off[-1] + (n + A - 1) // A * A)
        self.NP = off[-1]
        self.n_frames_host = n_frames
        self.frame_off_host = off
        utt = np.full(self.NP, -1, np.int32)
        for b, n in enumerate(n_frames):
            utt[off[b]:off[b] + n] = b
        # One pinned block, one asynchronous copy: a pageable host-to-device copy (torch.tensor(list, device=...)) is
        # ordered by the runtime behind ALL work queued on the stream and blocks the host until then -- with a batch in
        # flight that is the whole enhancement of the previous batch (see const_i32).
        nb = len(n_frames)
        seg = (nb + 1 + 63) // 64 * 64                      # every array starts on a 256-byte boundary
        host = torch.zeros(2 * seg + self.NP, dtype=torch.int32).pin_memory()
        hn = host.numpy()
        hn[:nb + 1], hn[seg:seg + nb], hn[2 * seg:] = off, n_frames, utt
        idx = host.to(device, non_blocking=True)
        self.frame_off, self.n_frames, self.frame_utt = idx[:nb + 1], idx[seg:seg + nb], idx[2 * seg:]
        self._alloc_stream = torch.cuda.current_stream(device)
        self._streams = {self._alloc_stream}
        self._struct = None
        self.y = None
        if index_only:
            for name in ("X2", "Xc", "W", "Wun", "H", "g", "Vb", "Z", "Vs", "X2t", "Vs_w", "XV", "yproj"):
                setattr(self, name, None)
            return
        f32 = dict(dtype=torch.float32, device=device)
        NP = self.NP
        self.X2 = torch.ones(F, NP, **f32)
        self.Xc = torch.zeros(F, NP, 2, **f32) if with_complex else None
        self.W = torch.empty(self.B, F, K, **f32)
        self.Wun = torch.empty(self.B, F, K, **f32)
        self.H = torch.ones(K, NP, **f32)
        self.g = torch.ones(NP, **f32)
        self.Vb = torch.ones(F, NP, **f32)
        self.Z = torch.zeros(L, NP, **f32)
        VT = _lib.GVN_VS_TILE
        self.Vs = torch.zeros(R_cap, NP // VT, F, VT, **f32)  # slot form in column-tile order, see include/gvn.h
        self.X2t = torch.ones(NP // VT, F, VT, **f32)
        self.Vs_w = torch.zeros(R_cap, NP, **f32)
        self.XV = torch.empty(F, NP, dtype=torch.int32, device=device)
        self.yproj = torch.zeros(GVN_HIDDEN, NP, **f32)

    def use_on(self, stream):
        """Declares that kernels reading or writing this batch are being queued on `stream`.  The tensors belong to the
        caching allocator's pool of the stream they were allocated on; without this, dropping the batch while another
        stream still works on it hands its memory to the next allocation on the first stream (an upload thread building
        the next batch, say) -- which overwrites index arrays a running kernel is about to read."""
        if stream in self._streams:
            return
        self._streams.add(stream)
        for t in vars(self).values():
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(stream)

    def struct(self):
        if self._struct is not None:            # pointers and dims never change after construction
            return C.byref(self._struct)
        s = GvnBatch()
        s.B, s.F, s.K, s.L, s.NP, s.R_cap = self.B, self.F, self.K, self.L, self.NP, self.R_cap
        for name in ("frame_off", "n_frames", "frame_utt", "X2", "Xc", "W", "Wun", "H", "g", "Vb", "Z", "Vs", "yproj", "Vs_w", "XV", "X2t"):
            t = getattr(self, name)
            setattr(s, name, 0 if t is None else t.data_ptr())
        self._struct = s
        return C.byref(s)

    # ---- scatter / gather between per-utterance arrays and the global frame axis ----
    def cols(self, b):
        o = self.frame_off_host[b]
        return slice(o, o + self.n_frames_host[b])

    def scatter_cols(self, dst, per_utt):
        """dst[..., NP]  <-  list of [..., N_b] tensors (host or device)."""
        for b, t in enumerate(per_utt):
            dst[..., self.cols(b)] = torch.as_tensor(t).to(device=dst.device, dtype=dst.dtype)

    def gather_cols(self, src, b):
        return src[..., self.cols(b)]

    def expand_samples(self, R, b=None):
        """The reference's Vs (R,F,N) (mcem.py:307) from the slot form: sample r is the last slot
        <= r with a non-zero multiplicity.  For tests and the `Vs` attribute of the drop-in classes."""
        w = self.Vs_w[:R]
        idx = torch.arange(R, device=w.device).unsqueeze(1) * (w > 0)
        src = torch.cummax(idx, dim=0).values                              # (R, NP)
        vs = self.Vs[:R].permute(0, 2, 1, 3).reshape(R, self.F, self.NP)    # column tiles -> (R,F,NP)
        out = torch.gather(vs, 0, src.unsqueeze(1).expand(R, self.F, self.NP))
        return out if b is None else out[..., self.cols(b)]

    def set_samples(self, vs, w=None):
        """(R,F,NP) sample tensor -> the column-tile slot form (tests)."""
        R = vs.shape[0]
        VT = _lib.GVN_VS_TILE
        self.Vs[:R].copy_(vs.reshape(R, self.F, self.NP // VT, VT).permute(0, 2, 1, 3))
        self.Vs_w[:R].copy_(torch.ones(R, self.NP, device=vs.device) if w is None else w)


# --------------------------------------------------------------------------------------
# STFT / ISTFT (python/processing/stft.py)
# --------------------------------------------------------------------------------------
def stft_geometry(T, fs, wlen_sec, hop_percent):
    """Window, hop, end-pad flag and frame count for a signal of T samples (stft.py:37-53)."""
    if wlen_sec * fs != int(wlen_sec * fs):
        raise ValueError("wlen_sample of STFT is not an integer.")
    nfft = int(wlen_sec * fs)
    hop = int(hop_percent * nfft)
    q = (T / fs) / wlen_sec / hop_percent
    end_pad = math.ceil(q) != int(q)
    Tx = T + (hop if end_pad else 0)
    return nfft, hop, end_pad, 1 + Tx // hop


_PINNED = {}


def _grow_view(cache, key, shape, make):
    """View of `shape` into a flat buffer kept per key; the buffer only grows (by at least a quarter, so that a list of
    files with slowly growing lengths does not reallocate every batch).  Batches of a real file list all differ in
    T_stride and NP: one buffer per distinct shape would pin memory in proportion to the length of the list."""
    n = 1
    for d in shape:
        n *= int(d)
    flat = cache.get(key)
    if flat is None or flat.numel() < n:
        flat = cache[key] = make(max(n, 0 if flat is None else flat.numel() * 5 // 4))
    return flat[:n].view(*shape)


def pinned_buffer(tag, shape, dtype=torch.float32):
    """Page-locked staging buffer, one per (tag, dtype), grown on demand and handed out as a view: cudaHostAlloc of
    tens of megabytes per call costs more than the copy it serves, and keeping one buffer per shape leaks."""
    return _grow_view(_PINNED, (tag, dtype), tuple(shape), lambda n: torch.zeros(n, dtype=dtype).pin_memory())


_DEVBUF = {}


def device_buffer(tag, shape, dtype, device):
    """Device-side landing buffer of an upload, one per (tag, dtype, device), grown on demand (see pinned_buffer)."""
    return _grow_view(_DEVBUF, (tag, dtype, str(device)), tuple(shape), lambda n: torch.empty(n, dtype=dtype, device=device))


_CONST_I32 = {}


def const_i32(values, device):
    """Small read-only int32 vector on the device (lengths, paddings), created once per distinct content.
    A fresh ``torch.tensor(list, device=...)`` is a pageable host-to-device copy, which the CUDA runtime
    orders behind ALL work already queued on the stream: placed after the EM loop it stalls the host for
    the whole enhancement and nothing can be prepared for the next batch meanwhile -- so the copy goes through
    pinned memory, asynchronously.  The vectors are shared between streams (an upload stream makes them, the
    compute stream reads them again): every use is recorded, so that dropping old entries cannot hand their
    memory to another stream's allocation while a queued kernel still reads it."""
    key = (tuple(int(v) for v in values), str(device))
    t = _CONST_I32.get(key)
    if t is None:
        while len(_CONST_I32) >= 256:
            _CONST_I32.pop(next(iter(_CONST_I32)))
        host = torch.tensor(list(key[0]), dtype=torch.int32).pin_memory()
        t = _CONST_I32[key] = host.to(device, non_blocking=True)
        t._gvn_streams = set()
    st = torch.cuda.current_stream(t.device)
    if st not in t._gvn_streams:
        t._gvn_streams.add(st)
        t.record_stream(st)
    return t


def upload_waveforms(wavs, device, pinned=None, tag="wav"):
    """Packs B waveforms into one zero-padded (B, T_stride) f32 tensor on the device (pinned staging
    buffer and device buffer named by `tag`, so that two uploads can be in flight with two tags)."""
    B = len(wavs)
    T = [len(w) for w in wavs]
    T_stride = (max(T) + 3) // 4 * 4
    host = pinned if pinned is not None else pinned_buffer(tag, (B, T_stride))
    hn = host.numpy()
    # numpy assignments (float64 -> float32 on the way): a torch copy_ per utterance costs ~1 ms of dispatch each -- 59 ms
    # per 64-utterance batch on two threads against 5 ms -- which is what held the 8-rank end-to-end number back
    if isinstance(wavs, np.ndarray) and wavs.ndim == 2:
        hn[:, :wavs.shape[1]] = wavs
        if wavs.shape[1] < T_stride:
            hn[:, wavs.shape[1]:] = 0
    else:
        for b, w in enumerate(wavs):
            hn[b, :T[b]] = w
            if T[b] < T_stride:
                hn[b, T[b]:] = 0
    dev = device_buffer(tag, (B, T_stride), torch.float32, device)
    dev.copy_(host, non_blocking=True)
    return dev, T, T_stride


def download(t, tag):
    """Device tensor -> reused pinned host buffer (asynchronous on the current stream)."""
    host = pinned_buffer(tag, t.shape, t.dtype)
    host.copy_(t, non_blocking=True)
    return host


def stft_into(batch, wav_dev, T, T_stride, n_fft, hop, end_pad):
    lib = _lib.load()
    T_d = const_i32(T, batch.device)
    ep_d = const_i32(end_pad, batch.device)
    if min(T) <= n_fft // 2:
        raise _lib.GvnError(_lib.E_INVALID, "signal shorter than n_fft/2 cannot be reflect-padded")
    check(lib.gvn_stft_power(batch.struct(), _ptr(wav_dev), T_stride, _ptr(T_d), _ptr(ep_d), n_fft, hop, _stream()))
    return T_d, ep_d


def stft_to(batch, wav_dev, T, T_stride, n_fft, hop, end_pad, Xc_out, X2_out):
    """STFT of B waveforms with the geometry of `batch` into caller-supplied planes ([F][NP][2] and [F][NP] f32) instead
    of the batch's own Xc / X2: the clean-speech spectrogram the oracle labels are made from."""
    lib = _lib.load()
    batch.struct()
    alias = GvnBatch.from_buffer_copy(batch._struct)
    alias.Xc, alias.X2 = Xc_out.data_ptr(), X2_out.data_ptr()
    check(lib.gvn_stft_power(C.byref(alias), _ptr(wav_dev), T_stride, _ptr(const_i32(T, batch.device)),
                             _ptr(const_i32(end_pad, batch.device)), n_fft, hop, _stream()))


def istft_from(batch, S, out_len, T_stride, n_fft, hop):
    """S: [F][NP][2] f32 on device -> (B, T_stride) f32 waveforms on device."""
    lib = _lib.load()
    ol = const_i32(out_len, batch.device)
    out = torch.empty(batch.B, T_stride, dtype=torch.float32, device=batch.device)
    ws = torch.empty(lib.gvn_istft_workspace_bytes(batch.struct(), n_fft), dtype=torch.uint8, device=batch.device)
    check(lib.gvn_istft(batch.struct(), _ptr(S), n_fft, hop, _ptr(ol), _ptr(out), T_stride, _ptr(ws), _stream()))
    return out


# --------------------------------------------------------------------------------------
# per-utterance initialisation (mcem.py:36-57, :207-216, :361-369)
# --------------------------------------------------------------------------------------
def init_nmf(batch, eps, rand_W=None, rand_H=None, generator=None):
    """W = max(rand(F,K), eps), H = max(rand(K,N), eps), g = 1, Vb = W@H  (mcem.py:41-51)."""
    lib = _lib.load()
    f32 = dict(dtype=torch.float32, device=batch.device)
    if rand_W is None:
        rand_W = torch.rand(batch.B, batch.F, batch.K, generator=generator, **f32)
        rand_H = torch.rand(batch.K, batch.NP, generator=generator, **f32)
    else:
        rw = torch.stack([torch.as_tensor(w) for w in rand_W]).to(**f32)
        rh = torch.ones(batch.K, batch.NP, **f32)
        batch.scatter_cols(rh, rand_H)
        rand_W, rand_H = rw.contiguous(), rh
    check(lib.gvn_init_nmf(batch.struct(), _ptr(rand_W), _ptr(rand_H), float(eps), _stream()))


def set_labels(batch, dec, y):
    """y: None (M1) or [y_dim][NP] device tensor.  Computes yproj (label half of layer 1)."""
    lib = _lib.load()
    batch.y = y
    check(lib.gvn_label_projection(_ptr(dec.packed), _ptr(y), dec.L, dec.y_dim, dec.F, batch.NP,
                                   _ptr(batch.yproj), _stream()))


def encode_init(batch, vae):
    """Z <- encoder mean of [X2; y]  (mcem.py:214-215 / :367-368; models.py:90-104)."""
    enc = vae.encoder
    x, y = batch.X2, batch.y
    h = None
    for i, layer in enumerate(enc.hidden):
        W, b = _linear_params(layer, batch.device)
        h = dense(W, b, x, y, "tanh", batch.NP) if i == 0 else dense(W, b, h, None, "tanh", batch.NP)
    W, b = _linear_params(enc.sample.mu, batch.device)
    batch.Z.copy_(dense(W, b, h, None, "none", batch.NP))


def classify(batch, classifier, mean=None, std=None, eps=0.0, hard=True):
    """Guide label from the supervised classifier (models.py:41-62 with the standardisation
    of scripts/evaluate_M2_ibm.py:121-130).  Returns [y_dim][NP] on the device."""
    h = None
    for i, layer in enumerate(classifier.hidden):
        W, b = _linear_params(layer, batch.device)
        if i == 0:
            m = None if mean is None else torch.as_tensor(mean, dtype=torch.float32, device=batch.device).reshape(-1).contiguous()
            s = None if std is None else torch.as_tensor(std, dtype=torch.float32, device=batch.device).reshape(-1).contiguous()
            h = dense(W, b, batch.X2, None, "relu", batch.NP, m, s, eps)
        else:
            h = dense(W, b, h, None, "relu", batch.NP)
    W, b = _linear_params(classifier.output_layer, batch.device)
    return dense(W, b, h, None, "hard" if hard else "sigmoid", batch.NP)


def speech_labels(batch, S, vad=False, quantile_fraction=0.98, quantile_weight=0.999, from_power=False):
    """Oracle guide labels of the batch from the clean-speech STFT S ([F][NP][2] f32, or the power [F][NP] with
    ``from_power``): clean_speech_IBM / clean_speech_VAD (python/processing/target.py:7-50) per utterance.
    Returns [F][NP] (IBM) or [1][NP] (VAD) f32 on the device."""
    lib = _lib.load()
    y = torch.empty(1 if vad else batch.F, batch.NP, dtype=torch.float32, device=batch.device)
    ws = torch.empty(lib.gvn_speech_labels_workspace_bytes(batch.struct()), dtype=torch.uint8, device=batch.device)
    check(lib.gvn_speech_labels(batch.struct(), _ptr(S), int(bool(from_power)), int(bool(vad)), float(quantile_fraction),
                                float(quantile_weight), _ptr(y), _ptr(ws), _stream()))
    return y


def spp_mask(batch, fixed_smooth=0.8, prob_smooth=0.9, prior=0.5, snr_opt_db=15.0, n_init=10, want_soft=True):
    """"timo" guide labels (python/models/spp_estimation.py:198-218): speech presence probability of every bin of
    batch.X2 and its threshold at 0.5.  Returns (soft or None, hard), both [F][NP] on the device."""
    lib = _lib.load()
    f32 = dict(dtype=torch.float32, device=batch.device)
    soft = torch.zeros(batch.F, batch.NP, **f32) if want_soft else None
    hard = torch.zeros(batch.F, batch.NP, **f32)
    check(lib.gvn_spp_mask(batch.struct(), float(fixed_smooth), float(prob_smooth), float(prior), float(snr_opt_db),
                           int(n_init), _ptr(soft), _ptr(hard), _stream()))
    return soft, hard


# --------------------------------------------------------------------------------------
# the MCEM loop (mcem.py:155-178)
# --------------------------------------------------------------------------------------
class ReplayNoise:
    """Recorded draws for parity runs: a list over chains of (eps [steps][L][NP], u [steps][NP])."""

    def __init__(self, chains, forced=None):
        self.chains = chains
        self.forced = forced

    @staticmethod
    def from_utterance_tapes(batch, tapes_eps, tapes_u, chain_steps):
        """tapes_eps[b]: (total_steps, L, N_b), tapes_u[b]: (total_steps, N_b); split into chains."""
        chains, s0 = [], 0
        f32 = dict(dtype=torch.float32, device=batch.device)
        for steps in chain_steps:
            e = torch.zeros(steps, batch.L, batch.NP, **f32)
            u = torch.full((steps, batch.NP), 0.5, **f32)
            batch.scatter_cols(e, [t[s0:s0 + steps] for t in tapes_eps])
            batch.scatter_cols(u, [t[s0:s0 + steps] for t in tapes_u])
            chains.append((e.contiguous(), u.contiguous()))
            s0 += steps
        return ReplayNoise(chains)


def estep(batch, dec, burnin, R, var_RW, precision="fp32", seed=0, chain=0, eps=None, u=None, forced=None,
          trace=False, xv_current=False):
    """One MH chain over all frames (gvn_estep).  Returns the trace tensors if requested.
    ``xv_current``: the caller vouches that gvn_mstep was the last writer of Vb (it keeps batch.XV, the packed
    per-bin constants of the tensor-core chain, in step), so the packing pass is skipped."""
    lib = _lib.load()
    nz = GvnNoise()
    nz.eps, nz.u = (0 if eps is None else eps.data_ptr()), (0 if u is None else u.data_ptr())
    nz.forced_accept = 0 if forced is None else forced.data_ptr()
    nz.seed, nz.chain = int(seed) & (2 ** 64 - 1), int(chain)
    tr, out = None, None
    if trace:
        steps = burnin + R
        acc = torch.zeros(steps, batch.NP, dtype=torch.float32, device=batch.device)
        dec_ = torch.zeros(steps, batch.NP, dtype=torch.uint8, device=batch.device)
        cnt = torch.zeros(batch.NP, dtype=torch.int32, device=batch.device)
        zs = torch.zeros(R, batch.L, batch.NP, dtype=torch.float32, device=batch.device)
        tr = GvnTrace()
        tr.acc_prob, tr.accepted, tr.n_accepted = acc.data_ptr(), dec_.data_ptr(), cnt.data_ptr()
        tr.z_samples = zs.data_ptr()
        out = (acc, dec_, cnt, zs)
    check(lib.gvn_estep(batch.struct(), _ptr(dec.packed), int(burnin), int(R), float(np.float32(var_RW)),
                        C.byref(nz), C.byref(tr) if tr is not None else None,
                        PRECISIONS[precision] | (_lib.GVN_PREC_XV_CURRENT if xv_current else 0), _stream()))
    return out


class MstepScratch:
    def __init__(self, batch, niter):
        lib = _lib.load()
        self.ws = torch.empty(max(16, lib.gvn_mstep_workspace_bytes(batch.struct())), dtype=torch.uint8, device=batch.device)
        self.ntiles = batch.NP // _lib.GVN_COST_TILE
        self.cost_part = torch.zeros(niter, self.ntiles, dtype=torch.float32, device=batch.device)


def mstep(batch, R, scratch, it, variant=1):
    lib = _lib.load()
    check(lib.gvn_mstep(batch.struct(), int(R), _ptr(scratch.cost_part[it]), _ptr(scratch.ws), int(variant), _stream()))


def mstep_gain(batch, R, scratch, it):
    """Gain-only M-step with a fixed noise variance (EM_noNMF.M_step, mcem.py:551-588) + cost."""
    lib = _lib.load()
    check(lib.gvn_mstep_gain(batch.struct(), int(R), _ptr(scratch.cost_part[it]), _stream()))


def cost_reduce(batch, R, scratch, niter):
    lib = _lib.load()
    cost = torch.empty(niter, batch.B, dtype=torch.float64, device=batch.device)
    check(lib.gvn_cost_reduce(batch.struct(), int(R), int(niter), _ptr(scratch.cost_part), _ptr(cost), _stream()))
    return cost


def wiener(batch, R, want_masks=False):
    lib = _lib.load()
    f32 = dict(dtype=torch.float32, device=batch.device)
    S = torch.empty(batch.F, batch.NP, 2, **f32)
    Nn = torch.empty(batch.F, batch.NP, 2, **f32)
    WFs = torch.empty(batch.F, batch.NP, **f32) if want_masks else None
    WFn = torch.empty(batch.F, batch.NP, **f32) if want_masks else None
    check(lib.gvn_wiener(batch.struct(), int(R), _ptr(S), _ptr(Nn), _ptr(WFs), _ptr(WFn), _stream()))
    return S, Nn, WFs, WFn


def energy_ratios(est, s_ref, n_ref, T):
    """(SI-SDR, SI-SIR, SI-SAR) in dB of B signals (python/metrics.py:12-60) -> (B, 3) f64 device tensor.
    est, s_ref, n_ref: (B, T_stride) f32 device tensors; T: per-utterance lengths (list or device i32 tensor)."""
    lib = _lib.load()
    B, T_stride = est.shape
    assert s_ref.shape == est.shape and n_ref.shape == est.shape
    Td = T if torch.is_tensor(T) else const_i32(T, est.device)
    out = torch.empty(B, 3, dtype=torch.float64, device=est.device)
    check(lib.gvn_energy_ratios(_ptr(est.contiguous()), _ptr(s_ref.contiguous()), _ptr(n_ref.contiguous()), B, T_stride,
                                _ptr(Td), _ptr(out), _stream()))
    return out


class KernelTimers:
    """CUDA-event brackets around the E-step and M-step launches of a run (for the roofline
    figures of bench.py).  Events are recorded on the launching stream; read after a sync."""

    def __init__(self):
        self.spans = {"estep": [], "mstep": []}
        self.launches = 0

    def bracket(self, kind):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.spans[kind].append((a, b))
        return a, b

    def total_ms(self, kind):
        return sum(a.elapsed_time(b) for a, b in self.spans[kind])

    def count(self, kind):
        return len(self.spans[kind])


def run_mcem(batch, dec, niter, chain_E, chain_WF, var_RW, precision="fp32", seed=0, noise=None,
             mstep_variant=1, want_masks=False, iter_hook=None, timers=None):
    """EM.run (mcem.py:155-178) for the whole batch: niter x (E-step chain, M-step), then the
    Wiener chain.  chain_E / chain_WF are (R, burnin).  Returns (cost[niter][B] f64 device
    tensor, S_hat, N_hat, WFs, WFn).  Nothing here synchronises the host."""
    (R_E, b_E), (R_W, b_W) = chain_E, chain_WF
    assert max(R_E, R_W) <= batch.R_cap
    scratch = MstepScratch(batch, niter)
    def timed(kind, fn, *a):
        if timers is None:
            return fn(*a)
        t0, t1 = timers.bracket(kind)
        t0.record()
        fn(*a)
        t1.record()

    for n in range(niter):
        e, u = (noise.chains[n] if noise is not None else (None, None))
        # after the first M-step XV is kept current by gvn_mstep (unless a hook may have touched Vb)
        timed("estep", estep, batch, dec, b_E, R_E, var_RW, precision, seed, n, e, u, None, False, n > 0 and iter_hook is None)
        timed("mstep", mstep, batch, R_E, scratch, n, mstep_variant)
        if iter_hook is not None:
            iter_hook(batch, n)
    e, u = (noise.chains[niter] if noise is not None else (None, None))
    timed("estep", estep, batch, dec, b_W, R_W, var_RW, precision, seed, niter, e, u, None, False, niter > 0 and iter_hook is None)
    cost = cost_reduce(batch, R_E, scratch, niter)
    S, Nn, WFs, WFn = wiener(batch, R_W, want_masks)
    return cost, S, Nn, WFs, WFn
