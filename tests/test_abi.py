"""CPU-side checks of the C-ABI library and the host logic (no compute calls)."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest
import torch

from gvn import _lib, engine
from gvn.pipeline import McemConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gvn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gvn_[a-z_0-9]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "libgvn.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding lacks %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.gvn_version() == 100


def test_struct_layouts_match_header():
    # 6 int32 + 16 pointers; noise: 3 pointers + 2 u64; trace: 4 pointers
    assert ctypes.sizeof(_lib.GvnBatch) == 6 * 4 + 16 * 8
    assert ctypes.sizeof(_lib.GvnNoise) == 5 * 8
    assert ctypes.sizeof(_lib.GvnTrace) == 4 * 8


def test_packed_size_and_shape_limits():
    lib = _lib.load()
    assert lib.gvn_decoder_packed_bytes(16, 513, 513, 128) > 4 * (128 * 16 + 128 * 128 + 513 * 128)
    assert lib.gvn_decoder_packed_bytes(16, 0, 513, 64) == 0          # hidden must be 128
    assert lib.gvn_decoder_packed_bytes(65, 0, 513, 128) == 0         # L limit


def test_argument_validation_reports_errors_without_a_gpu():
    lib = _lib.load()
    rc = lib.gvn_pack_decoder(None, None, None, None, None, None, 16, 0, 513, 64, None, None)
    assert rc == _lib.E_UNSUPPORTED_SHAPE and b"hidden" in lib.gvn_last_error()
    rc = lib.gvn_estep(None, None, 1, 1, 0.01, None, None, 0, None)
    assert rc == _lib.E_INVALID
    with pytest.raises(_lib.GvnError):
        _lib.check(rc)


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.Batch([10], 513, 4, 16, 2, "cpu")


def test_stft_geometry_matches_reference_rule():
    # stft.py:48-53: hop zeros are appended unless T/fs/wlen/hop% is integral
    assert engine.stft_geometry(64000, 16000, 64e-3, 0.25) == (1024, 256, False, 251)
    assert engine.stft_geometry(480000, 16000, 64e-3, 0.25) == (1024, 256, False, 1876)
    assert engine.stft_geometry(136983, 16000, 64e-3, 0.25) == (1024, 256, True, 537)
    with pytest.raises(ValueError, match="not an integer"):
        engine.stft_geometry(100, 16000, 50.01e-3, 0.25)


def test_m1_quirk_and_pickle():
    from python.models.mcem import MCEM_M1, MCEM_M2
    m1 = MCEM_M1(niter=100)
    assert m1.chain_lengths() == ((30, 30), (75, 30))                 # mcem.py:461-462, :477-478
    m2 = MCEM_M2(niter=100)
    assert m2.chain_lengths() == ((10, 30), (25, 75))
    assert McemConfig(model="M1").chains() == ((30, 30), (75, 30))
    m = pickle.loads(pickle.dumps(m2))                                # shipped to workers before init
    assert m.niter == 100 and m.var_RW == 0.01


def test_rvae_is_rejected_like_the_reference():
    from python.models.mcem import MCEM_M1

    class RVAE:
        pass
    with pytest.raises(NameError, match="only valid for FFNN VAE"):
        MCEM_M1(1).init_parameters(np.zeros((4, 513), np.complex64), RVAE(), 2, 1e-8, "cuda:0")


def test_models_state_dict_keys_and_seeded_init():
    from python.models.models import DeepGenerativeModel, Classifier
    from _util import load_golden
    g = load_golden("M2_ibm")
    torch.manual_seed(0)
    vae = DeepGenerativeModel([513, 513, 16, [128, 128]], None)
    sd = vae.state_dict()
    assert set(sd) == {k[3:] for k in g if k.startswith("sd_")}
    # weights of the golden model came from the reference's constructor under the same seed
    for k in ("encoder.hidden.0.weight", "decoder.hidden.0.weight", "decoder.reconstruction.weight"):
        np.testing.assert_array_equal(sd[k].numpy(), g["sd_" + k])
    assert set(Classifier([513, [128, 128], 1]).state_dict()) == {
        "hidden.0.weight", "hidden.0.bias", "hidden.1.weight", "hidden.1.bias", "output_layer.weight", "output_layer.bias"}


def test_metric_matches_oracle():
    from python.metrics import energy_ratios
    from oracle import mcem_oracle as O
    rs = np.random.RandomState(0)
    s, n = rs.randn(1000), rs.randn(1000)
    e = 0.7 * s + 0.2 * n + 0.05 * np.random.RandomState(1).randn(1000)
    np.testing.assert_allclose(energy_ratios(e, s, n), O.energy_ratios(e, s, n), rtol=1e-12)


def test_staging_buffers_grow_instead_of_accumulating():
    """Staging buffers are kept per tag, not per shape: batches of a real file list all differ in T_stride / NP, and one
    page-locked buffer per distinct shape would pin memory in proportion to the length of the list."""
    import torch
    cache = {}
    make = lambda n: torch.zeros(n)
    a = engine._grow_view(cache, "wav", (3, 5), make)
    b = engine._grow_view(cache, "wav", (2, 4), make)
    assert a.data_ptr() == b.data_ptr() and tuple(b.shape) == (2, 4) and b.is_contiguous()
    c = engine._grow_view(cache, "wav", (10, 10), make)
    assert tuple(c.shape) == (10, 10) and len(cache) == 1 and cache["wav"].numel() >= 100
    for k in range(200):                                   # 200 distinct shapes: still one buffer
        engine._grow_view(cache, "wav", (7, 100 + k), make)
    assert len(cache) == 1 and cache["wav"].numel() < 2 * 7 * 300


def test_bench_reports_ncu_traffic_only_for_the_same_build(tmp_path):
    """bench.py copies the DRAM traffic of the committed ncu capture into `roofline.traffic` only when the capture was
    made on a build of the same sources (binary or source sha256), for the same workload and chain arithmetic."""
    import json
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    h = bench.src_hash()
    assert h == bench.src_hash() and len(h) == 64
    p = tmp_path / "traffic.json"
    rec = dict(libgvn_sha256="0" * 64, libgvn_src_sha256=h, config="C2", precision="f16", estep_bytes_per_launch=1, mstep_bytes_per_launch=2)
    p.write_text(json.dumps(rec))
    t, note = bench.ncu_traffic("C2", "f16", str(p))
    assert t["estep_bytes_per_launch"] == 1 and "same source sha256" in note
    for cfg, prec in (("C4", "f16"), ("C2", "fp32")):
        t, note = bench.ncu_traffic(cfg, prec, str(p))
        assert t == {} and "not reported" in note
    rec["libgvn_src_sha256"] = "1" * 64
    p.write_text(json.dumps(rec))
    t, note = bench.ncu_traffic("C2", "f16", str(p))
    assert t == {} and "not reported" in note
    assert bench.ncu_traffic("C2", "f16", str(tmp_path / "missing.json"))[0] == {}
