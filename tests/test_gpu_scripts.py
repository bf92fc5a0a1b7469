"""The drop-in proof: the reference's OWN evaluate scripts -- scripts/evaluate_M1.py,
evaluate_M2_ibm.py, evaluate_M2_vad.py, byte-for-byte copies under oracle/_ref/scripts (oracle/make_ref.py; git-ignored) --
are imported UNMODIFIED and their process_utt() is called with this repository's modules standing where the reference's
stood: `python.processing.stft`, `python.processing.target`, `python.models.mcem`, `python.models.models` resolve to
guided-vae-nmf_b200/python (the CUDA path), `soundfile` to gvn.wavio (libsndfile is not in this image).  What a user of
the reference does to switch is exactly this: put guided-vae-nmf_b200 first on PYTHONPATH.
The outputs on disk are checked against the batched driver of this repository (gvn.evaluate) on the same files."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(ROOT, "oracle", "_ref", "scripts")


def _load_script(name, monkeypatch):
    path = os.path.join(SCRIPTS, name + ".py")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/scripts is absent (python oracle/make_ref.py needs /root/reference)")
    from gvn import wavio
    sf = types.ModuleType("soundfile")
    sf.read, sf.write = wavio.read, wavio.write
    monkeypatch.setitem(sys.modules, "soundfile", sf)
    # modules the scripts import at the top but do not need for process_utt (training-set listing, a parameter counter)
    ds = types.ModuleType("python.dataset.csr1_wjs0_dataset")
    ds.speech_list = lambda **kw: []
    pkg = types.ModuleType("python.dataset")
    pkg.__path__ = []
    ut = types.ModuleType("python.utils")
    ut.count_parameters = lambda model: sum(p.numel() for p in model.parameters() if p.requires_grad)
    for k, v in (("python.dataset", pkg), ("python.dataset.csr1_wjs0_dataset", ds), ("python.utils", ut)):
        monkeypatch.setitem(sys.modules, k, v)
    spec = importlib.util.spec_from_file_location("ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                                  # module level = the constants block; main() is not run
    import python.models.mcem as mirror
    assert mod.stft.__module__ == "python.processing.stft" and sys.modules["python.models.mcem"] is mirror
    assert "guided-vae-nmf_b200" in mirror.__file__              # the scripts got the drop-in, not the reference
    return mod


def _write_inputs(tmp, n=2, T=20000):
    from gvn import wavio
    from gvn.synth import synth_utterance
    files = []
    for i in range(n):
        x, s, nz = synth_utterance(i, seed=8, T=T + 512 * i)
        stem = os.path.join("spk%d" % i, "utt%d" % i)
        os.makedirs(os.path.join(tmp, "processed", "spk%d" % i), exist_ok=True)
        for tag, sig in (("_x", x), ("_s", s), ("_n", nz)):
            wavio.write(os.path.join(tmp, "processed", stem + tag + ".wav"), sig, 16000)
        files.append(stem + ".wav")
    return files


def _check_outputs(tmp, out_dir, files):
    from gvn import wavio
    for fp in files:
        stem = os.path.splitext(fp)[0]
        x, fs = wavio.read(os.path.join(tmp, "processed", stem + "_x.wav"))
        s_est, fs1 = wavio.read(os.path.join(out_dir, stem + "_s_est.wav"))
        n_est, _ = wavio.read(os.path.join(out_dir, stem + "_n_est.wav"))
        assert fs == fs1 == 16000 and len(s_est) == len(n_est) == len(x)          # istft(max_len=T_orig), evaluate_M2_ibm.py:156-157
        assert np.all(np.isfinite(s_est)) and np.std(s_est) > 1e-4
        # WFs + WFn = 1 (mcem.py:341-343): the two estimates add up to the mixture, up to the STFT round trip and the
        # 16-bit quantisation of the three files
        np.testing.assert_allclose(s_est + n_est, x, atol=5e-4)


def test_reference_process_utt_M2_ibm_runs_on_the_dropin(tmp_path, monkeypatch):
    mod = _load_script("evaluate_M2_ibm", monkeypatch)
    from python.models.mcem import MCEM_M2
    from python.models.models import DeepGenerativeModel
    from oracle import mcem_oracle as O
    from oracle import stft_oracle
    from gvn import wavio
    tmp = str(tmp_path)
    files = _write_inputs(tmp)
    mod.processed_data_dir = os.path.join(tmp, "processed") + os.sep
    mod.output_data_dir = os.path.join(tmp, "out") + os.sep
    mod.classif_type = "oracle"                                   # the oracle-label branch (:132-134)
    torch.manual_seed(0)
    model = DeepGenerativeModel([513, 513, mod.z_dim, mod.h_dim], None).eval()       # the script's own dims (z_dim = 32)
    mcem = MCEM_M2(niter=3, nsamples_E_step=mod.nsamples_E_step, burnin_E_step=mod.burnin_E_step, nsamples_WF=mod.nsamples_WF,
                   burnin_WF=mod.burnin_WF, var_RW=mod.var_RW)
    for fp in files:
        mod.process_utt(mcem, model, None, None, None, fp, "cuda:0")
    _check_outputs(tmp, os.path.join(tmp, "out"), files)
    for fp in files:                                              # the label files the script saves (:170-171), against the oracle labels
        stem = os.path.join(tmp, "out", os.path.splitext(fp)[0])
        soft = torch.load(stem + " _ibm_soft_est.pt", weights_only=False)
        hard = torch.load(stem + "_ibm_hard_est.pt", weights_only=False)
        s, _ = wavio.read(os.path.join(tmp, "processed", os.path.splitext(fp)[0] + "_s.wav"))
        S = mod.stft(s, fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, dtype="complex64")
        np.testing.assert_array_equal(soft, O.clean_speech_IBM(S, 0.999, 0.999))
        np.testing.assert_array_equal(hard.cpu().numpy(), soft.T)


def test_reference_process_utt_M1_and_vad_run_on_the_dropin(tmp_path, monkeypatch):
    tmp = str(tmp_path)
    files = _write_inputs(tmp, n=1)
    from python.models.mcem import MCEM_M1, MCEM_M2
    from python.models.models import Classifier, DeepGenerativeModel, VariationalAutoencoder
    # M1 (evaluate_M1.py:111-166)
    m1 = _load_script("evaluate_M1", monkeypatch)
    m1.processed_data_dir = os.path.join(tmp, "processed") + os.sep
    m1.output_data_dir = os.path.join(tmp, "out_m1") + os.sep
    torch.manual_seed(0)
    vae = VariationalAutoencoder([513, m1.z_dim, m1.h_dim]).eval()
    mcem = MCEM_M1(niter=2, nsamples_E_step=10, burnin_E_step=5, nsamples_WF=25, burnin_WF=6, var_RW=0.01)
    m1.process_utt(mcem, vae, files[0], "cuda:0")
    _check_outputs(tmp, os.path.join(tmp, "out_m1"), files)
    # M2 with VAD labels from the classifier (evaluate_M2_vad.py:96-174, classif_type 'dnn')
    mv = _load_script("evaluate_M2_vad", monkeypatch)
    mv.processed_data_dir = os.path.join(tmp, "processed") + os.sep
    mv.output_data_dir = os.path.join(tmp, "out_vad") + os.sep
    torch.manual_seed(1)
    dgm = DeepGenerativeModel([513, 1, mv.z_dim, mv.h_dim], None).eval().cuda()
    clf = Classifier([513, mv.h_dim_cl, 1]).eval().cuda()
    mean = torch.zeros(513, 1, device="cuda")
    std = torch.ones(513, 1, device="cuda")
    mcem = MCEM_M2(niter=2, nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25, burnin_WF=75, var_RW=0.01)
    with torch.no_grad():
        mv.process_utt(mcem, dgm, clf, mean, std, files[0], "cuda:0")
    _check_outputs(tmp, os.path.join(tmp, "out_vad"), files)


def test_batched_driver_writes_the_same_files(tmp_path, monkeypatch):
    """gvn.evaluate (batches, threads, pinned staging) against the reference script's one-utterance-at-a-time process_utt
    on the same inputs: same file names, same lengths, same label files; the waveforms agree as two Monte-Carlo runs do."""
    mod = _load_script("evaluate_M2_ibm", monkeypatch)
    from python.models.mcem import MCEM_M2
    from python.models.models import DeepGenerativeModel
    from gvn import wavio
    from gvn.evaluate import evaluate_file_list
    from gvn.pipeline import McemConfig, Enhancer
    tmp = str(tmp_path)
    files = _write_inputs(tmp, n=3)
    mod.processed_data_dir = os.path.join(tmp, "processed") + os.sep
    mod.output_data_dir = os.path.join(tmp, "out_ref") + os.sep
    mod.classif_type = "oracle"
    torch.manual_seed(0)
    model = DeepGenerativeModel([513, 513, 16, [128, 128]], None).eval()
    mcem = MCEM_M2(niter=20, nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25, burnin_WF=75, var_RW=0.01)
    for fp in files:
        mod.process_utt(mcem, model, None, None, None, fp, "cuda:0")
    enh = Enhancer(model, McemConfig(model="M2", niter=20, nmf_rank=10, precision="fp32"), "cuda:0")
    evaluate_file_list(enh, files, os.path.join(tmp, "processed"), os.path.join(tmp, "out_gvn"), label_source="oracle_ibm", batch_size=2)
    for fp in files:
        stem = os.path.splitext(fp)[0]
        a, _ = wavio.read(os.path.join(tmp, "out_ref", stem + "_s_est.wav"))
        b, _ = wavio.read(os.path.join(tmp, "out_gvn", stem + "_s_est.wav"))
        assert len(a) == len(b)
        assert np.corrcoef(a, b)[0, 1] > 0.98, np.corrcoef(a, b)[0, 1]
        ha = torch.load(os.path.join(tmp, "out_ref", stem + "_ibm_hard_est.pt"), weights_only=False).cpu().numpy()
        hb = torch.load(os.path.join(tmp, "out_gvn", stem + "_ibm_hard_est.pt"), weights_only=False).cpu().numpy()
        np.testing.assert_array_equal(ha, hb)
