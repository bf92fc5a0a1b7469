#!/usr/bin/env python
"""Benchmark of the MCEM-NMF enhancement hot path (BASELINE.json metric: utterances/s at a
fixed iteration count; %tensor and %HBM roofline of the two hot kernels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C1|C2|C3|C4|C5] [--impl reference]

One *step* = one whole enhancement (STFT -> labels -> init -> niter EM iterations -> Wiener chain ->
ISTFT x2 -> quality metrics) of one batch of synthetic utterances on every rank.  The default workload is
BASELINE.json configs[1] ("C2"): M2 guided VAE with oracle IBM labels (made on the device from the clean
speech), 64 utterances of 4 s per GPU, n_fft=1024 (F=513), K=10, z_dim=16, 100 EM iterations, chains
(10,30)/(25,75).  The other configs of BASELINE.json are selectable (they are parity cases first, bench
lines second): C1 = M1, one utterance (the reference's CPU-runnable case); C3 = M2 with VAD labels from the
supervised classifier (on the device), 64 utterances per GPU; C4 = M1, 30 s utterances, K = 32, 10 MH
samples per frame (the NMF stress case), 8 utterances per GPU; C5 = a fixed list of --utterances ragged
utterances (537-748 frames, the range of the reference's own WSJ0 fixture) of the C2 model, sorted by length,
dealt to the ranks and enhanced in batches: STRONG scaling (the list does not grow with the ranks).
Ranks own disjoint utterance shards (no data-path collective); the per-utterance result rows are brought
together with one NCCL all-gather: per step in the device-timed loop, once at the end of the job in the
end-to-end loop.

`value`  : utterances/s with the inputs of the path already in HBM when the timed region starts: the waveforms,
           the clean-speech / noise references of the quality metrics and the guide labels made from the clean
           speech (CUDA events, max over ranks).
`e2e`    : the same through Enhancer.enhance_many with HOST buffers, everything inside the timed region: packing,
           pinned H2D of the waveforms and references, the oracle labels (clean-speech STFT, ranking, threshold: on
           the device, queued on the copy stream behind the upload), the enhancement, D2H of both enhanced waveforms
           and of the result rows.
`roofline`: the dominant kernel (the fused decoder + Metropolis-Hastings chain, gvn_estep)
           against the measured bf16 tensor peak; `roofline_nmf`: the NMF M-step kernels
           against the measured HBM copy bandwidth.
`cpu_baseline` / `--impl reference`: the reference's own MCEM classes (oracle/_ref, kind "reference";
           the oracle port when that copy is absent, kind "port") timed on this box's host cores:
           one process with all threads, and one single-thread process per core (the reference's own
           parallelism is process-level, scripts/evaluate_M1.py:215-216); `torch_cuda_baseline`: the same
           reference code on device='cuda', one utterance at a time, as scripts/evaluate_M2_ibm.py:196-197.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "guided-vae-nmf_b200")
REF_COPY = os.path.join(ROOT, "oracle", "_ref")

METRIC = "utterances/sec (MCEM-NMF, fixed iters)"
STFT_KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25)

# model, label source, utterances per GPU, samples, NMF rank, MCEM constructor arguments (BASELINE.json configs)
CONFIGS = {
    "C1": dict(model="M1", y_dim=0, labels=None, batch=1, T=64000, K=10, mcem={},
               text="C1: M1 VAE, one utterance x 4 s @16 kHz, F=513, K=10, z_dim=16 (chains 30+30 / 75+30 by the reference's positional-argument quirk)"),
    "C2": dict(model="M2", y_dim=513, labels="oracle_ibm", batch=64, T=64000, K=10, mcem={},
               text="C2: M2 guided VAE, oracle IBM labels, {B} utt/GPU x 4 s @16 kHz, F=513, K={K}, z_dim=16, chains (10,30)/(25,75)"),
    "C3": dict(model="M2", y_dim=1, labels="classifier", batch=64, T=64000, K=10, mcem={},
               text="C3: M2 guided VAE, VAD labels from the supervised classifier (on the device), {B} utt/GPU x 4 s @16 kHz, F=513, K={K}, z_dim=16, chains (10,30)/(25,75)"),
    "C4": dict(model="M1", y_dim=0, labels=None, batch=8, T=480000, K=32, mcem=dict(burnin_E_step=10),
               text="C4: M1 VAE, {B} utt/GPU x 30 s @16 kHz (N=1876), F=513, K={K}, z_dim=16, 10 MH samples per frame (NMF stress)"),
    "C5": dict(model="M2", y_dim=513, labels="oracle_ibm", batch=64, T=None, K=10, mcem={},
               text="C5: M2 guided VAE, oracle IBM labels, a fixed list of {U} ragged utterances (537-748 frames), length-sorted, {BT}, F=513, K={K}, z_dim=16"),
}


def product_paths():
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)


def workload(args):
    c = CONFIGS[args.config]
    bt = "batches of %d" % args.batch if args.waves <= 0 else "batches sized to %d full waves of 128-frame chain tiles" % args.waves
    d = dict(workload=c["text"].format(B=args.batch, K=args.rank_k, U=args.utterances, BT=bt) + ", niter=%d" % args.niter,
             utterances_per_gpu=args.batch, niter=args.niter, precision=args.precision,
             l2="working set (Vs alone > 300 MB/GPU) exceeds the 126 MB L2; no explicit flush",
             parallelism="utterance shards, 1 process/GPU")
    if args.config == "C5":
        d["utterances_total"] = args.utterances
        if args.waves > 0:
            del d["utterances_per_gpu"]                     # the batch size follows from the lengths (wave_batches)
    return d


def build_models(cfgname):
    """Random-init models of the reference's architecture (no trained weights ship with it): SURVEY.md 8d."""
    import torch
    from python.models.models import Classifier, DeepGenerativeModel, VariationalAutoencoder
    c = CONFIGS[cfgname]
    torch.manual_seed(0)
    if c["model"] == "M1":
        vae = VariationalAutoencoder([513, 16, [128, 128]]).eval()
    else:
        vae = DeepGenerativeModel([513, c["y_dim"], 16, [128, 128]], None).eval()
    clf = Classifier([513, [128, 128], 1]).eval() if c["labels"] == "classifier" else None
    for m in (vae, clf):
        if m is not None:
            for p_ in m.parameters():
                p_.requires_grad = False
    return vae, clf


def build_model(F=513, y_dim=513, L=16):
    """The C2 model (kept for the tests and tools that import it)."""
    import torch
    from python.models.models import DeepGenerativeModel
    torch.manual_seed(0)
    vae = DeepGenerativeModel([F, y_dim, L, [128, 128]], None).eval()
    for p_ in vae.parameters():
        p_.requires_grad = False
    return vae


def make_inputs(n, first, T=64000, cpu_only=False):
    """Synthetic utterances + oracle IBM labels made on the HOST (clean_speech_IBM of the clean
    speech STFT, scripts/evaluate_M2_ibm.py:133-134): for the tests and tools; the benchmark itself makes the labels on
    the device."""
    import numpy as np
    from gvn.synth import synth_batch
    from oracle import stft_oracle
    from oracle.mcem_oracle import clean_speech_IBM
    x, s, nz = synth_batch(n, seed=0, T=T, first=first)
    labels = [clean_speech_IBM(stft_oracle.stft(si, dtype="complex64", **STFT_KW), 0.999, 0.999).astype(np.uint8) for si in s]
    return x, s, nz, labels                                  # binary masks as bytes (0/1), waveforms float64


def ragged_lengths(U):
    """Frame counts of the C5 list: uniform in [537, 748], the range of the reference's WSJ0 fixture; T = 256 (N-1)."""
    import numpy as np
    return np.random.RandomState(5).randint(537, 749, size=U)


def decoder_flops_per_frame(L=16, F=513):
    return 2 * (L * 128 + 128 * 128 + 128 * F)


def lib_hash():
    try:
        return hashlib.sha256(open(os.path.join(PKG, "libgvn.so"), "rb").read()).hexdigest()
    except OSError:
        return None


def src_hash():
    """sha256 over the sources libgvn.so is built from (csrc/*.cu, *.cuh, the Makefile, include/gvn.h), in name order.
    nvcc names the internal-linkage symbols of every compilation differently, so two builds of the same sources are not
    byte-identical: the ncu capture is tied to this hash, the library's own sha256 is reported next to it."""
    h = hashlib.sha256()
    csrc = os.path.join(PKG, "csrc")
    files = sorted(f for f in os.listdir(csrc) if f.endswith((".cu", ".cuh")) or f == "Makefile")
    for f in [os.path.join(csrc, f) for f in files] + [os.path.join(ROOT, "include", "gvn.h")]:
        h.update(os.path.basename(f).encode() + b"\0")
        h.update(open(f, "rb").read())
    return h.hexdigest()


def ncu_traffic(config, precision, path=None):
    """DRAM bytes per launch of the two hot kernels from the committed `ncu --set full` capture -- reported only when the
    capture was made on a build of THESE sources (binary or source sha256), for this workload and this chain arithmetic.
    Returns (dict or {}, note)."""
    path = path or os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        tj = json.load(open(path))
    except Exception:
        return {}, "no ncu capture of this build"
    same_bin = tj.get("libgvn_sha256") is not None and tj.get("libgvn_sha256") == lib_hash()
    same_src = tj.get("libgvn_src_sha256") is not None and tj.get("libgvn_src_sha256") == src_hash()
    if (same_bin or same_src) and tj.get("config") == config and tj.get("precision", "f16") == precision:
        return tj, "%s (capture of a build of the same sources: %s)" % (
            os.path.relpath(path, ROOT), "same libgvn.so sha256" if same_bin else "same source sha256")
    return {}, "%s is from another build or workload (sha256 / config / precision mismatch): not reported" % os.path.relpath(path, ROOT)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        import numpy as np
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [v.strip() for v in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU (or stock-torch CUDA) implementation of the path
# ------------------------------------------------------------------------------------------------
def reference_runner(args, device):
    """Returns (kind, fn(x, s_clean, niter)) enhancing ONE utterance like process_utt of the evaluate scripts
    (scripts/evaluate_M2_ibm.py:95-171): stft -> label -> init_parameters -> run() -> istft x2.  kind "reference" =
    the reference's own classes from oracle/_ref; "port" = the oracle restatement (oracle/mcem_oracle.py).
    STFT/ISTFT are the numpy restatement in both cases (the reference's stft.py needs librosa, absent from this image)."""
    import numpy as np
    import torch
    c = CONFIGS[args.config]
    sys.path.insert(0, ROOT)
    from oracle import stft_oracle
    have_ref = os.path.exists(os.path.join(REF_COPY, "python", "models", "mcem.py"))
    if have_ref:
        sys.path.insert(0, REF_COPY)                        # `python` now is the reference's package, not the mirror
        from python.models.mcem import MCEM_M1, MCEM_M2
        from python.models.models import Classifier, DeepGenerativeModel, VariationalAutoencoder
        from python.processing.target import clean_speech_IBM
        torch.manual_seed(0)
        if c["model"] == "M1":
            vae = VariationalAutoencoder([513, 16, [128, 128]])
        else:
            vae = DeepGenerativeModel([513, c["y_dim"], 16, [128, 128]], None)
        clf = Classifier([513, [128, 128], 1]) if c["labels"] == "classifier" else None
        for m in (vae, clf):
            if m is not None:
                m.to(device).eval()
                for p_ in m.parameters():
                    p_.requires_grad = False

        def run(x, s_clean, niter):
            x_tf = stft_oracle.stft(x, dtype="complex64", **STFT_KW).T          # (N, F), evaluate_M2_ibm.py:100-108
            kw = dict(niter=niter, nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25, burnin_WF=75, var_RW=0.01)
            kw.update(c["mcem"])
            with torch.no_grad():
                if c["model"] == "M1":
                    m = MCEM_M1(**kw)
                    m.init_parameters(X=x_tf, vae=vae, nmf_rank=args.rank_k, eps=1e-8, device=device)
                else:
                    if c["labels"] == "classifier":                                  # evaluate_M2_vad.py:121-131 (mean 0, std 1)
                        xin = torch.tensor(np.abs(x_tf) ** 2, device=device)
                        y = (clf(xin) > 0.5).float()
                    else:
                        S = stft_oracle.stft(s_clean, dtype="complex64", **STFT_KW)
                        y = torch.from_numpy(clean_speech_IBM(S, 0.999, 0.999).T.copy()).to(device)
                    m = MCEM_M2(**kw)
                    m.init_parameters(X=x_tf, y=y, vae=vae, nmf_rank=args.rank_k, eps=1e-8, device=device)
                m.run()
            stft_oracle.istft(m.S_hat, max_len=len(x), **STFT_KW)
            stft_oracle.istft(m.N_hat, max_len=len(x), **STFT_KW)
        return "reference", run

    # fall-back: the oracle port (CPU only)
    product_paths()
    from oracle.mcem_oracle import McemOracle, NoiseTape, split_state_dict, clean_speech_IBM, classify
    vae, clf = build_models(args.config)
    sd = vae.state_dict()

    def run(x, s_clean, niter):
        X = stft_oracle.stft(x, dtype="complex64", **STFT_KW).T
        y = None
        if c["model"] == "M2":
            if c["labels"] == "classifier":
                y = (classify({k: v for k, v in clf.state_dict().items()}, torch.from_numpy(np.abs(X) ** 2)) > 0.5).float()
            else:
                y = torch.from_numpy(clean_speech_IBM(stft_oracle.stft(s_clean, dtype="complex64", **STFT_KW), 0.999, 0.999).T.copy())
        kw = dict(nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25, burnin_WF=75)
        kw.update(c["mcem"])
        o = McemOracle(niter, kw["nsamples_E_step"], kw["burnin_E_step"], kw["nsamples_WF"], kw["burnin_WF"], 0.01, model=c["model"])
        o.init_parameters(X, y, split_state_dict(sd, "decoder"), split_state_dict(sd, "encoder"), args.rank_k, 1e-8, NoiseTape(seed=1))
        o.run()
        stft_oracle.istft(o.S_hat, max_len=len(x), **STFT_KW)
        stft_oracle.istft(o.N_hat, max_len=len(x), **STFT_KW)
    return "port", run


def reference_utterance(args, i=0):
    """One synthetic utterance of the workload for the reference arm (numpy only: no product import)."""
    sys.path.append(PKG)
    from gvn.synth import synth_utterance                   # pure numpy / scipy
    T = CONFIGS[args.config]["T"] or 256 * (int(ragged_lengths(max(1, args.utterances))[i % max(1, args.utterances)]) - 1)
    x, s, _ = synth_utterance(i, seed=0, T=T)
    return x, s


def run_reference(args):
    """--impl reference: the reference's implementation of the path on the host cores (or, with --ref-device cuda, its
    stock-torch path on the GPU).  Each step enhances ONE utterance of the workload (same shape, same niter).
    --ref-threads 1 --start-at T: a worker of the utterance-parallel measurement."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    threads = args.ref_threads or cores
    torch.set_num_threads(threads)
    dev = args.ref_device
    kind, run = reference_runner(args, dev)
    if dev != "cpu" and kind != "reference":
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is absent: the port runs on the CPU only"}))
        return
    x, s = reference_utterance(args)
    sync = (lambda: torch.cuda.synchronize()) if dev != "cpu" else (lambda: None)
    niter_t = args.ref_niter or args.niter
    for _ in range(args.warmup):
        run(x, s, max(1, args.niter // 20))
    sync()
    if args.start_at:
        time.sleep(max(0.0, args.start_at - time.time()))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(x, s, niter_t)
    sync()
    dt = (time.perf_counter() - t0) / args.steps
    extrap = None
    if niter_t != args.niter:
        # time(niter) = (niter + c) * t_iter, c = the final Wiener chain in EM-iteration equivalents of the E-step chain
        (rE, bE), (rW, bW) = chains_of(args.config)
        cW = (rW + bW) / float(rE + bE)
        dt, extrap = dt * (args.niter + cW) / (niter_t + cW), "measured at niter=%d, scaled by (niter + %.2f) / (%d + %.2f)" % (niter_t, cW, niter_t, cW)
    v = 1.0 / dt
    sample = "1 utterance per step (niter=%d%s), warm-up steps at niter=%d; STFT/ISTFT = numpy restatement (librosa absent)" % (
        niter_t, "" if extrap is None else "; " + extrap, max(1, args.niter // 20))
    cfg = workload(args)
    cfg["precision"] = "fp32"                               # the reference computes in fp32 throughout
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "utt/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": "utt/s", "cores": threads, "kind": kind, "device": dev, "sample": sample},
        "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def chains_of(cfgname):
    c = CONFIGS[cfgname]
    kw = dict(nsamples_E_step=10, burnin_E_step=30, nsamples_WF=25, burnin_WF=75)
    kw.update(c["mcem"])
    if c["model"] == "M1":                                  # mcem.py:461-462, :477-478
        return (kw["burnin_E_step"], 30), (kw["burnin_WF"], 30)
    return (kw["nsamples_E_step"], kw["burnin_E_step"]), (kw["nsamples_WF"], kw["burnin_WF"])


def child_reference(args, extra, timeout=900):
    """Runs `bench.py --impl reference ...` in a child process (the reference's `python` package and the mirror's cannot
    live in one interpreter) and returns its JSON line, or None."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config, "--niter", str(args.niter),
           "--rank-k", str(args.rank_k), "--utterances", str(args.utterances)] + extra
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=timeout, env=env).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception:
        return None


def host_baselines(args):
    """cpu_baseline (+ torch_cuda_baseline) of the gvn arm: bounded samples of the same workload, each in child processes."""
    cores = os.cpu_count()
    heavy = args.config in ("C4",)                          # ~90 s per utterance at niter=100: measure fewer iterations and scale
    ref_niter = ["--ref-niter", str(max(2, args.niter // 10))] if heavy else []
    out = {}
    one = child_reference(args, ["--steps", "1" if heavy or args.config == "C1" else "3", "--warmup", "1"] + ref_niter)
    if one and "cpu_baseline" in one:
        out["cpu_baseline"] = one["cpu_baseline"]
    # utterance-parallel: one single-thread process per core, started together (the reference's own parallelism)
    start = time.time() + 25.0
    par_niter = max(2, args.niter // (20 if heavy else 5))
    cmd = ["--steps", "1", "--warmup", "1", "--ref-threads", "1", "--ref-niter", str(par_niter), "--start-at", "%.3f" % start]
    procs = []
    for _ in range(cores):
        c = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config, "--niter", str(args.niter),
             "--rank-k", str(args.rank_k), "--utterances", str(args.utterances)] + cmd
        env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
            env.pop(k, None)
        procs.append(subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env))
    vals = []
    for p_ in procs:
        try:
            o, _ = p_.communicate(timeout=900)
            vals.append(json.loads(o.strip().splitlines()[-1])["value"])
        except Exception:
            p_.kill()
    if vals and "cpu_baseline" in out:
        out["cpu_baseline"]["utterance_parallel"] = {
            "value": len(vals) * min(vals), "unit": "utt/s", "processes": len(vals), "threads_each": 1,
            "sample": "%d single-thread processes started together, one utterance each at niter=%d, scaled to niter=%d like the "
                      "Wiener-chain-aware formula of --ref-niter; value = processes x the slowest process's rate" % (len(vals), par_niter, args.niter)}
    cu = child_reference(args, ["--steps", "1", "--warmup", "1", "--ref-device", "cuda", "--ref-niter", str(max(2, args.niter // 5))])
    if cu and "cpu_baseline" in cu:
        out["torch_cuda_baseline"] = {"value": cu["value"], "unit": "utt/s", "kind": cu["cpu_baseline"]["kind"],
                                      "what": "the reference's stock-torch path on device='cuda' of this B200, one utterance at a time "
                                              "(scripts/evaluate_M2_ibm.py:196-197)", "sample": cu["cpu_baseline"]["sample"]}
    return out


# ------------------------------------------------------------------------------------------------
# the gvn arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gvn")
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU and step (default: the config's)")
    ap.add_argument("--utterances", type=int, default=1024, help="C5: length of the fixed utterance list")
    ap.add_argument("--waves", type=int, default=2,
                    help="C5: a batch takes as many utterances as fit this many full waves of 128-frame chain tiles "
                         "(one tile per SM); 0 = fixed batches of --batch utterances")
    ap.add_argument("--niter", type=int, default=100)
    ap.add_argument("--rank-k", type=int, default=None)
    ap.add_argument("--precision", default=os.environ.get("GVN_PRECISION", "f16"),
                    help="decoder arithmetic inside the chain: f16 = tcgen05 tensor cores, f16 operands (11-bit mantissa = "
                         "TF32) with fp32 accumulate -- the north star's tensor-core mode; fp32 = CUDA-core parity mode")
    ap.add_argument("--mstep-variant", type=int, default=int(os.environ.get("GVN_MSTEP_VARIANT", "1")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", default="cpu")
    ap.add_argument("--ref-threads", type=int, default=0)
    ap.add_argument("--ref-niter", type=int, default=0)
    ap.add_argument("--start-at", type=float, default=0.0)
    args = ap.parse_args()
    c = CONFIGS[args.config]
    if args.batch is None:
        args.batch = c["batch"]
    if args.rank_k is None:
        args.rank_k = c["K"]
    if args.impl == "reference":
        return run_reference(args)

    product_paths()
    import numpy as np
    import torch
    import torch.distributed as dist
    from gvn.pipeline import McemConfig, Enhancer
    from gvn.shard import length_sorted_shard, wave_batches
    from gvn.synth import synth_batch, synth_utterance
    from gvn import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # host side of a rank = packing waveforms into pinned memory: share the cores between the ranks instead of
        # letting every process start one intra-op thread per core
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

    cfg = McemConfig(model=c["model"], niter=args.niter, nmf_rank=args.rank_k, precision=args.precision,
                     mstep_variant=args.mstep_variant, **c["mcem"])
    vae, clf = build_models(args.config)
    enh = Enhancer(vae, cfg, dev, classifier=clf, mean=None if clf is None else np.zeros(513, np.float32),
                   std=None if clf is None else np.ones(513, np.float32),
                   label_source=c["labels"] if c["labels"] in ("oracle_ibm", "oracle_vad") else None)
    B = args.batch

    # ---- this rank's utterances: a list of batches (one batch per step, except C5: the rank's share of the fixed list) ----
    if args.config == "C5":
        lens = ragged_lengths(args.utterances)
        mine = length_sorted_shard(lens, world, rank)       # longest first, dealt to the ranks in a snake
        pool = {}                                           # distinct synthetic signals: one per length class, reused down the list
        def utt(i):
            T = 256 * (int(lens[i]) - 1)
            key = (i % 64, T)
            if key not in pool:
                pool[key] = synth_utterance(i % 64, seed=0, T=T)
            return pool[key]
        if args.waves > 0:                                  # batches sized to whole waves of chain tiles (gvn.shard.wave_batches)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            groups = [mine[a:b_] for a, b_ in wave_batches([lens[i] for i in mine], sms, args.waves, max_batch=4 * B)]
        else:
            groups = [mine[k:k + B] for k in range(0, len(mine), B)]
        batches = []
        for g in groups:
            xs, ss, ns = zip(*[utt(i) for i in g])
            batches.append(dict(wavs=list(xs), refs=(list(ss), list(ns)), ids=g))
        n_utt_rank, n_utt_total, scaling = len(mine), args.utterances, "strong"
    else:
        x, s, nz = synth_batch(B, seed=0, T=c["T"], first=rank * B)
        batches = [dict(wavs=list(x), refs=(s, nz), ids=list(range(rank * B, rank * B + B)))]
        n_utt_rank, n_utt_total, scaling = B, world * B, "weak"

    gather_rows_buf = None
    def finish_step(s_hat, cost, refs, T, ids):
        # per-utterance result rows [utt_id, SI-SDR, SI-SIR, SI-SAR, final cost] (python/metrics.py:12-60 on the
        # device); bringing them together is the only collective of the path
        q = E.energy_ratios(s_hat[:, :refs[0].shape[1]].contiguous(), refs[0], refs[1], T)
        rows = torch.cat([torch.tensor(ids, device=dev, dtype=torch.float64)[:, None], q, cost[-1][:, None]], 1)
        return rows

    def gather(rows_list):
        nonlocal gather_rows_buf
        rows = torch.cat(rows_list, 0)
        if world > 1:
            cap = (n_utt_total + world - 1) // world + 1    # shards differ by at most one utterance
            pad = torch.full((cap, 5), -1.0, dtype=torch.float64, device=dev)
            pad[:rows.shape[0]] = rows
            if gather_rows_buf is None:
                gather_rows_buf = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(gather_rows_buf, pad)
        return rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------
    ups = [enh.upload(bt["wavs"], None, refs=bt["refs"], slot=2 + k) for k, bt in enumerate(batches)]

    def device_step(timers=None, seed=0):
        rows, last = [], None
        for k, (bt, up) in enumerate(zip(batches, ups)):
            b = enh.prepare(None, None, seed=seed + k, uploaded=up)
            s_hat, n_hat, cost = enh.run(b, seed=seed + k, timers=timers)
            rows.append(finish_step(s_hat, cost, (up["ref_s"], up["ref_n"]), b.T, bt["ids"]))
            last = (s_hat, cost, b)
        return gather(rows), last

    for i in range(args.warmup):
        device_step(seed=1000 * i)
    from gvn import _lib
    lib = _lib.load()
    timers = E.KernelTimers()
    clocks = ClockSampler(local)
    barrier()
    launches0 = int(lib.gvn_launch_count())
    if rank == 0:
        clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()                             # `ncu --profile-from-start off` sees the timed region only
    t0.record()
    for i in range(args.steps):
        rows, (s_hat, cost, b) = device_step(timers=timers, seed=100000 + 1000 * i)
    t1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    barrier()
    launches = int(lib.gvn_launch_count()) - launches0      # kernels of libgvn.so launched inside the timed region
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    value = n_utt_total / (ms_step * 1e-3)
    assert bool(torch.isfinite(cost).all()) and bool(torch.isfinite(s_hat).all()), "non-finite result"

    # ---- end-to-end timing (host buffers in, host buffers out) --------------------------
    # Enhancer.enhance_many is the public batched API: every step packs this step's waveforms and references
    # into pinned memory, copies them to the device, makes the labels there, enhances, and reads both enhanced
    # waveforms + the result rows back into pinned host memory; the upload of batch i+1 overlaps batch i and its
    # kernels are queued behind those of batch i before the host waits for batch i.
    def stream_batches(k):
        for _ in range(k):
            for bt in batches:
                yield dict(wavs=bt["wavs"], refs=bt["refs"])
    ids_dev = [torch.tensor(bt["ids"], device=dev, dtype=torch.float64) for bt in batches]
    pending_rows = []
    def collect_hook(i, cost_d, metrics_d):                # result rows stay on the device, in stream order
        pending_rows.append(torch.cat([ids_dev[i % len(batches)][:, None], metrics_d, cost_d[-1][:, None]], 1))
    def gather_job():
        # The path's one collective: the result rows of the whole job, once, at its end --
        # not once per step, which would tie eight host processes together at every batch (any rank's host jitter
        # then stalls the other seven GPUs; measured: 4 % of the 8-GPU end-to-end throughput).
        rows = torch.cat(pending_rows, 0)
        pending_rows.clear()
        if world > 1:
            cap = torch.tensor([rows.shape[0]], device=dev)
            dist.all_reduce(cap, op=dist.ReduceOp.MAX)
            pad = torch.full((int(cap), 5), -1.0, dtype=torch.float64, device=dev)
            pad[:rows.shape[0]] = rows
            out_ = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(out_, pad)
    for out in enh.enhance_many(stream_batches(1 if args.config == "C5" else 2), seed=150, device_hook=collect_hook):
        pass
    gather_job()
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h2d = d2h = 0
    for out in enh.enhance_many(stream_batches(args.steps), seed=200, device_hook=collect_hook):
        h2d += out["h2d_bytes"]; d2h += out["d2h_bytes"]
    gather_job()
    e1.record()
    barrier()
    e_ms = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n_utt_total / (float(e_ms) / args.steps * 1e-3)
    assert bool(np.isfinite(out["s_hat"].numpy()).all())

    # ---- roofline of the two hot kernels (events recorded inside the timed region) ------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    frames = sum(g[3] for up in ups for g in up["geo"])     # frames of this rank per step
    (R_E, b_E), (R_W, b_W) = cfg.chains()
    sweeps = args.niter * (R_E + b_E) + (R_W + b_W)                 # proposals per step per frame
    e_ms_tot = timers.total_ms("estep")
    flops = args.steps * sweeps * frames * decoder_flops_per_frame()
    e_tflops = flops / (e_ms_tot * 1e-3) / 1e12
    m_ms_tot = timers.total_ms("mstep")
    nmf_bytes = args.steps * args.niter * 2 * (R_E + 1) * 513 * frames * 4
    m_gbs = nmf_bytes / (m_ms_tot * 1e-3) / 1e9
    # DRAM traffic per launch of the two hot kernels from the committed `ncu --set full` capture (profiles/): only
    # reported when the capture was made with THIS build of libgvn.so and this workload
    traffic, tr_note = ncu_traffic(args.config, args.precision)
    n_e = timers.count("estep")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16 operands, f32 accumulate",
            "data": "synthetic" + (" (64 distinct signals per length, reused down the list)" if args.config == "C5" else ""),
            "config": workload(args), "clocks": clk,
            "quality": {"si_sdr_db_mean": float(rows[:, 1].mean()), "note": "random-init decoder (no trained weights ship with the "
                        "reference): the number checks plumbing, not enhancement quality"},
            "e2e": {"value": e2e_value, "unit": "utt/s", "h2d_bytes_per_step": int(h2d // args.steps), "d2h_bytes_per_step": int(d2h // args.steps)},
            "gpu_launches": launches, "libgvn_sha256": lib_hash(), "libgvn_src_sha256": src_hash(),
            "roofline": {"kernel": "gvn_estep (decoder MLP + MH chain, %s)" % args.precision, "bound": "tensor",
                         "achieved": e_tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": e_tflops / tf_peak,
                         "traffic": traffic.get("estep_bytes_per_launch"), "traffic_source": tr_note,
                         "peak_source": peak_src + " bf16_tflops_sustained",
                         "avg_launch_ms": e_ms_tot / n_e, "launches": n_e, "share_of_step": e_ms_tot / (ms_step * args.steps)},
            "roofline_nmf": {"kernel": "gvn_mstep (k_tile_meta + W sweep + column sweep)", "bound": "hbm",
                             "achieved": m_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": m_gbs / hbm_peak,
                             "traffic": traffic.get("mstep_bytes_per_launch"), "traffic_source": tr_note, "peak_source": peak_src + " hbm_gbs",
                             "avg_launch_ms": m_ms_tot / timers.count("mstep"),
                             "share_of_step": m_ms_tot / (ms_step * args.steps)},
        }
        if world == 1 and not args.no_cpu_baseline:
            line.update(host_baselines(args))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
