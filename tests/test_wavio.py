"""gvn.wavio against Python's own ``wave`` module (an independent RIFF implementation) and libsndfile's
documented conversions (PCM16 read = int/32768, write = lrint(x*0x7FFF))."""
import os
import struct
import wave

import numpy as np
import pytest

from gvn import wavio


def test_write_pcm16_is_readable_by_stdlib_and_matches_libsndfile_rounding(tmp_path):
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 1.5 / 32767, 2.5 / 32767, -1.5 / 32767, 1.2, -1.3, 1e-6])
    p = str(tmp_path / "a.wav")
    wavio.write(p, x, 16000)
    with wave.open(p, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, len(x))
        raw = np.frombuffer(w.readframes(len(x)), "<i2")
    # lrint: round half to even; saturation outside [-1, 1]
    np.testing.assert_array_equal(raw, [0, 32767, -32767, 16384, -16384, 2, 2, -2, 32767, -32768, 0])
    y, fs = wavio.read(p)
    assert fs == 16000 and y.dtype == np.float64
    np.testing.assert_array_equal(y, raw / 32768.0)


def test_read_file_written_by_stdlib(tmp_path):
    rs = np.random.RandomState(0)
    pcm = rs.randint(-32768, 32768, size=(1000, 2)).astype("<i2")
    p = str(tmp_path / "b.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(8000)
        w.writeframes(pcm.tobytes())
    y, fs = wavio.read(p)
    assert fs == 8000 and y.shape == (1000, 2)
    np.testing.assert_array_equal(y, pcm / 32768.0)


@pytest.mark.parametrize("T", [1, 7, 64000])
def test_float_roundtrip_and_odd_sizes(tmp_path, T):
    x = np.random.RandomState(T).randn(T).astype(np.float32) * 0.3
    p = str(tmp_path / "c.wav")
    wavio.write(p, x, 16000, subtype="FLOAT")
    y, fs = wavio.read(p, dtype="float32")
    np.testing.assert_array_equal(x, y)
    assert os.path.getsize(p) % 2 == 0


def test_pcm24_and_extensible_header(tmp_path):
    v = np.array([0, 1, -1, 8388607, -8388608, 123456], np.int32)
    body = b"".join(struct.pack("<i", int(t))[:3] for t in v)
    fmt = struct.pack("<HHIIHH", 0xFFFE, 1, 16000, 48000, 3, 24) + struct.pack("<HHI", 22, 24, 4) + struct.pack("<H", 1) + b"\x00" * 14
    p = str(tmp_path / "d.wav")
    with open(p, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + 4 + 8 + len(body)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<I", len(fmt)) + fmt)
        f.write(b"LIST" + struct.pack("<I", 4) + b"abcd")                   # an unrelated chunk before the data
        f.write(b"data" + struct.pack("<I", len(body)) + body)
    y, fs = wavio.read(p)
    np.testing.assert_array_equal(y, v / 8388608.0)


def test_rejects_garbage(tmp_path):
    p = str(tmp_path / "e.wav")
    open(p, "wb").write(b"not a wav file at all")
    with pytest.raises(ValueError):
        wavio.read(p)


@pytest.mark.parametrize("subtype", ["PCM_16", "FLOAT"])
def test_cross_check_with_scipy_wavfile(tmp_path, subtype):
    """A second, independent RIFF implementation (scipy.io.wavfile): files written here are read by scipy sample for sample,
    and files written by scipy (int16, int32, float32, stereo) come back from `read` with libsndfile's scaling (int / 2^(bits-1))."""
    wavfile = pytest.importorskip("scipy.io.wavfile")
    rs = np.random.RandomState(4)
    x = np.clip(0.3 * rs.randn(4001), -1.0, 1.0)
    p = str(tmp_path / "a.wav")
    wavio.write(p, x, 16000, subtype=subtype)
    fs, y = wavfile.read(p)
    assert fs == 16000 and y.shape == (4001,)
    if subtype == "PCM_16":
        assert y.dtype == np.int16
        np.testing.assert_array_equal(y, wavio.pcm16(x))
    else:
        assert y.dtype == np.float32
        np.testing.assert_array_equal(y, x.astype(np.float32))
    for arr, scale in ((rs.randint(-32768, 32768, size=(3000, 2)).astype(np.int16), 32768.0),
                       (rs.randint(-2 ** 31, 2 ** 31 - 1, size=2999).astype(np.int32), 2147483648.0),
                       (rs.randn(1234).astype(np.float32), 1.0)):
        q = str(tmp_path / "b.wav")
        wavfile.write(q, 8000, arr)
        z, fs = wavio.read(q)
        assert fs == 8000 and z.shape == arr.shape and z.dtype == np.float64
        np.testing.assert_array_equal(z, arr.astype(np.float64) / scale)
