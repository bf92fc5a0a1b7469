"""The file-level driver (gvn.evaluate: the on-disk contract of scripts/evaluate_*.py) on synthetic WAV files."""
import os

import numpy as np
import pytest
import torch

from gvn import wavio
from gvn.evaluate import evaluate_file_list
from gvn.pipeline import Enhancer, McemConfig
from gvn.synth import synth_utterance

pytestmark = pytest.mark.gpu


def _corpus(tmp_path, n, lengths):
    proc = tmp_path / "processed"
    files = []
    for i in range(n):
        x, s, _ = synth_utterance(i, seed=3, T=lengths[i % len(lengths)])
        rel = "spk%d/utt%03d.wav" % (i % 2, i)
        os.makedirs(proc / ("spk%d" % (i % 2)), exist_ok=True)
        wavio.write(str(proc / (os.path.splitext(rel)[0] + "_x.wav")), x, 16000)
        wavio.write(str(proc / (os.path.splitext(rel)[0] + "_s.wav")), s, 16000)
        files.append(rel)
    return str(proc) + "/", files


def _vae(y_dim):
    from python.models.models import DeepGenerativeModel, VariationalAutoencoder
    torch.manual_seed(0)
    m = VariationalAutoencoder([513, 16, [128, 128]]) if y_dim == 0 else DeepGenerativeModel([513, y_dim, 16, [128, 128]], None)
    return m.eval()


def test_m1_files_equal_direct_api(tmp_path):
    proc, files = _corpus(tmp_path, 5, [8000, 6144, 7000])
    cfg = McemConfig(model="M1", niter=2, nsamples_E_step=2, burnin_E_step=3, nsamples_WF=2, burnin_WF=3, precision="f16")
    enh = Enhancer(_vae(0), cfg, "cuda:0")
    out = str(tmp_path / "out") + "/"
    res = evaluate_file_list(enh, files, proc, out, label_source=None, batch_size=3, seed=5)
    assert [r[0] for r in res] == files and all(r[1].shape == (2,) and np.isfinite(r[1]).all() for r in res)
    # batch 0 = files[0:3] with seed 5: the same call through the in-memory API, quantised like sf.write does
    wavs = [wavio.read(proc + os.path.splitext(f)[0] + "_x.wav")[0] for f in files[:3]]
    s_hat, n_hat, _ = enh.enhance(wavs, None, seed=5)
    for i, f in enumerate(files[:3]):
        y, fs = wavio.read(out + os.path.splitext(f)[0] + "_s_est.wav")
        assert fs == 16000 and len(y) == len(wavs[i])
        np.testing.assert_array_equal(y, wavio.pcm16(s_hat[i]) / 32768.0)
        yn, _ = wavio.read(out + os.path.splitext(f)[0] + "_n_est.wav")
        np.testing.assert_array_equal(yn, wavio.pcm16(n_hat[i]) / 32768.0)
    assert not os.path.exists(out + os.path.splitext(files[0])[0] + "_ibm_hard_est.pt")


@pytest.mark.parametrize("source,y_dim", [("oracle_ibm", 513), ("oracle_vad", 1), ("classifier", 1)])
def test_m2_label_sources_and_sharding(tmp_path, source, y_dim):
    from python.models.models import Classifier
    proc, files = _corpus(tmp_path, 4, [6144, 8000])
    cfg = McemConfig(model="M2", niter=1, nsamples_E_step=2, burnin_E_step=2, nsamples_WF=2, burnin_WF=2, precision="f16")
    torch.manual_seed(1)
    clf = Classifier([513, [128, 128], 1]).eval() if source == "classifier" else None
    enh = Enhancer(_vae(y_dim), cfg, "cuda:0", classifier=clf, mean=np.zeros((513, 1), np.float32), std=np.ones((513, 1), np.float32))
    out = str(tmp_path / "out") + "/"
    done = []
    for rank in range(2):                                                   # two ranks, one after the other: disjoint shards
        done += [r[0] for r in evaluate_file_list(enh, files, proc, out, label_source=source, batch_size=2, world=2, rank=rank)]
    assert done == files
    for f in files:
        st = out + os.path.splitext(f)[0]
        x, _ = wavio.read(proc + os.path.splitext(f)[0] + "_x.wav")
        s, _ = wavio.read(st + "_s_est.wav")
        n, _ = wavio.read(st + "_n_est.wav")
        assert len(s) == len(x) == len(n)
        np.testing.assert_allclose(s + n, x, atol=3.0 / 32768)               # WFs + WFn = 1, three 16-bit roundings
        hard = torch.load(st + "_ibm_hard_est.pt", weights_only=False)
        soft = torch.load(st + " _ibm_soft_est.pt", weights_only=False)
        n_frames = 1 + (len(x) + (256 if (len(x) / 256) % 1 else 0)) // 256
        assert tuple(hard.shape) == (n_frames, y_dim) and set(np.unique(hard.numpy())) <= {0.0, 1.0}
        if source == "classifier":
            assert tuple(soft.shape) == (n_frames, 1) and torch.equal((soft > 0.5).float(), hard)
        else:
            assert soft.shape == (y_dim, n_frames) and np.array_equal(soft.T, hard.numpy())


def test_rejects_mismatched_configuration(tmp_path):
    cfg = McemConfig(model="M1", niter=1)
    enh = Enhancer(_vae(0), cfg, "cuda:0")
    with pytest.raises(ValueError):
        evaluate_file_list(enh, [], str(tmp_path), str(tmp_path), label_source="oracle_ibm")
    with pytest.raises(ValueError):
        evaluate_file_list(enh, [], str(tmp_path), str(tmp_path), label_source="nonsense")
