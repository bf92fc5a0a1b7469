"""WAV files as the evaluate scripts use them (``sf.read`` / ``sf.write`` of the ``soundfile`` package:
scripts/evaluate_M2_ibm.py:98, :166-167; evaluate_M1.py:114, :158-159).  ``soundfile`` (libsndfile) is not part
of this image, so the two calls are provided here with libsndfile's conventions:

* ``read(path)`` returns ``(data float64, fs)``; integer PCM is scaled by ``1 / 2**(bits-1)`` (16-bit: 1/32768),
  float files come back as stored; multi-channel files give ``(frames, channels)``;
* ``write(path, data, fs)`` writes 16-bit PCM by default (libsndfile's default subtype for WAV): the sample is
  ``lrint(x * 0x7FFF)`` -- round half to even -- with saturation outside [-1, 1] (libsndfile wraps there unless
  clipping is switched on; saturation is the safe choice and identical for in-range signals).  ``subtype='FLOAT'``
  writes IEEE float32.

Plain RIFF/WAVE, little endian, formats 1 (PCM), 3 (IEEE float) and 0xFFFE (extensible with those sub-formats).
"""
import struct

import numpy as np

_PCM, _FLOAT, _EXT = 1, 3, 0xFFFE


def _chunks(buf):
    pos = 12
    while pos + 8 <= len(buf):
        cid, size = buf[pos:pos + 4], struct.unpack_from("<I", buf, pos + 4)[0]
        yield cid, buf[pos + 8:pos + 8 + size]
        pos += 8 + size + (size & 1)


def read(path, dtype="float64"):
    """-> (data, samplerate).  data: (frames,) or (frames, channels) in ``dtype``."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 12 or buf[:4] != b"RIFF" or buf[8:12] != b"WAVE":
        raise ValueError("%s: not a RIFF/WAVE file" % path)
    fmt = data = None
    for cid, body in _chunks(buf):
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            data = body
            break
    if fmt is None or data is None or len(fmt) < 16:
        raise ValueError("%s: missing fmt or data chunk" % path)
    tag, ch, fs, _, align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == _EXT and len(fmt) >= 26:
        tag = struct.unpack_from("<H", fmt, 24)[0]
    nbytes = bits // 8
    n = len(data) // (nbytes * ch) * ch
    if tag == _PCM:
        if bits == 16:
            x = np.frombuffer(data, "<i2", n).astype(np.float64) / 32768.0
        elif bits == 32:
            x = np.frombuffer(data, "<i4", n).astype(np.float64) / 2147483648.0
        elif bits == 24:
            b = np.frombuffer(data, np.uint8, n * 3).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = (v - ((v & 0x800000) << 1)).astype(np.float64) / 8388608.0
        elif bits == 8:
            x = (np.frombuffer(data, np.uint8, n).astype(np.float64) - 128.0) / 128.0
        else:
            raise ValueError("%s: unsupported PCM width %d" % (path, bits))
    elif tag == _FLOAT and bits in (32, 64):
        x = np.frombuffer(data, "<f4" if bits == 32 else "<f8", n).astype(np.float64)
    else:
        raise ValueError("%s: unsupported WAV format tag %d / %d bits" % (path, tag, bits))
    x = x.astype(dtype, copy=False)
    return (x.reshape(-1, ch) if ch > 1 else x), int(fs)


def pcm16(x):
    """float -> int16 the way libsndfile converts on write (scale 0x7FFF, round half to even), saturating."""
    return np.clip(np.rint(np.asarray(x, np.float64) * 32767.0), -32768, 32767).astype("<i2")


def write(path, data, samplerate, subtype="PCM_16"):
    data = np.asarray(data)
    ch = 1 if data.ndim == 1 else data.shape[1]
    if subtype == "PCM_16":
        body, tag, bits = pcm16(data).tobytes(), _PCM, 16
    elif subtype == "FLOAT":
        body, tag, bits = np.asarray(data, "<f4").tobytes(), _FLOAT, 32
    else:
        raise ValueError("subtype %r not supported (PCM_16, FLOAT)" % (subtype,))
    align = ch * bits // 8
    fmt = struct.pack("<HHIIHH", tag, ch, int(samplerate), int(samplerate) * align, align, bits)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(body) + (len(body) & 1)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<I", len(fmt)) + fmt)
        f.write(b"data" + struct.pack("<I", len(body)) + body + (b"\x00" if len(body) & 1 else b""))
