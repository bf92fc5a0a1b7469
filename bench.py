#!/usr/bin/env python
"""Benchmark of the MCEM-NMF enhancement hot path (BASELINE.json metric: utterances/s at a
fixed iteration count; %tensor and %HBM roofline of the two hot kernels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One *step* = one whole enhancement (STFT -> init -> niter EM iterations -> Wiener chain ->
ISTFT x2) of one batch of synthetic utterances on every rank.  Workload = BASELINE.json
configs[1] ("C2"): M2 guided VAE with oracle IBM labels, 64 utterances of 4 s per GPU,
n_fft=1024 (F=513), K=10, z_dim=16, 100 EM iterations, chains (10,30)/(25,75).  Ranks own
disjoint utterance shards (weak scaling, no data-path collective); one NCCL all-gather of the
per-utterance result rows closes each step.

`value`  : utterances/s with the waveforms and labels already in HBM (CUDA events, max over
           ranks).
`e2e`    : the same through Enhancer.enhance-style calls with HOST buffers: pinned H2D of the
           waveforms+labels and D2H of both enhanced waveforms inside the timed region.
`roofline`: the dominant kernel (the fused decoder + Metropolis-Hastings chain, gvn_estep)
           against the measured bf16 tensor peak; `roofline_nmf`: the NMF M-step kernels
           against the measured HBM copy bandwidth.
`cpu_baseline` / `--impl reference`: the oracle port of the reference's torch-CPU path
           (oracle/mcem_oracle.py + oracle/stft_oracle.py) timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "guided-vae-nmf_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "utterances/sec (MCEM-NMF, fixed iters)"
STFT_KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25)


def workload(args):
    return dict(workload="C2: M2 guided VAE, oracle IBM labels, %d utt/GPU x 4 s @16 kHz, F=513, K=%d, z_dim=16, "
                         "niter=%d, chains (10,30)/(25,75)" % (args.batch, args.rank_k, args.niter),
                utterances_per_gpu=args.batch, niter=args.niter, precision=args.precision,
                l2="working set (Vs 330 MB/GPU) exceeds the 126 MB L2; no explicit flush",
                parallelism="utterance shards, 1 process/GPU")


def build_model(F=513, y_dim=513, L=16):
    from python.models.models import DeepGenerativeModel
    torch.manual_seed(0)
    vae = DeepGenerativeModel([F, y_dim, L, [128, 128]], None).eval()
    for p_ in vae.parameters():
        p_.requires_grad = False
    return vae


def make_inputs(n, first, T=64000, cpu_only=False):
    """Synthetic utterances + oracle IBM labels (clean_speech_IBM of the clean
    speech STFT, scripts/evaluate_M2_ibm.py:133-134).  The product arm computes that STFT with the
    library's own kernel; only the CPU reference arm (cpu_only) uses the oracle's numpy STFT."""
    from gvn.synth import synth_batch
    from python.processing.target import clean_speech_IBM
    if cpu_only:
        from oracle import stft_oracle
        stft = stft_oracle.stft
    else:
        from python.processing.stft import stft
    x, s, nz = synth_batch(n, seed=0, T=T, first=first)
    labels = [clean_speech_IBM(stft(si, dtype="complex64", **STFT_KW), 0.999, 0.999).astype(np.uint8) for si in s]
    return x, s, nz, labels                                  # binary masks as bytes (0/1), waveforms float64


def decoder_flops_per_frame(L=16, F=513):
    return 2 * (L * 128 + 128 * 128 + 128 * F)


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


def cpu_reference_utterance(x, s_clean, vae, args, niter):
    """The reference's CPU path for ONE utterance through the oracle port: stft -> labels are
    given -> init_parameters -> run() -> istft x2."""
    from oracle import stft_oracle
    from oracle.mcem_oracle import McemOracle, NoiseTape, split_state_dict, clean_speech_IBM
    sd = vae.state_dict()
    X = stft_oracle.stft(x, dtype="complex64", **STFT_KW).T
    y = torch.from_numpy(clean_speech_IBM(stft_oracle.stft(s_clean, dtype="complex64", **STFT_KW), 0.999, 0.999).T.copy())
    o = McemOracle(niter, 10, 30, 25, 75, 0.01, model="M2")
    o.init_parameters(X, y, split_state_dict(sd, "decoder"), split_state_dict(sd, "encoder"), args.rank_k, 1e-8,
                      NoiseTape(seed=1))
    o.run()
    stft_oracle.istft(o.S_hat, max_len=len(x), **STFT_KW)
    stft_oracle.istft(o.N_hat, max_len=len(x), **STFT_KW)


def run_reference(args):
    """--impl reference: the oracle port of the reference's torch-CPU path on the host cores.
    Each step enhances ONE utterance of the workload (same shape, same niter)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    vae = build_model()
    x, s, _, _ = make_inputs(1, 0, cpu_only=True)
    for _ in range(args.warmup):
        cpu_reference_utterance(x[0], s[0], vae, args, max(1, args.niter // 20))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_utterance(x[0], s[0], vae, args, args.niter)
    dt = (time.perf_counter() - t0) / args.steps
    v = 1.0 / dt
    sample = "1 utterance per step (full niter=%d), warm-up steps at niter=%d" % (args.niter, max(1, args.niter // 20))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "utt/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(args),
        "cpu_baseline": {"value": v, "unit": "utt/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gvn")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--niter", type=int, default=100)
    ap.add_argument("--rank-k", type=int, default=10)
    ap.add_argument("--precision", default=os.environ.get("GVN_PRECISION", "f16"),
                    help="decoder arithmetic inside the chain: f16 = tcgen05 tensor cores, f16 operands (11-bit mantissa = "
                         "TF32) with fp32 accumulate -- the north star's tensor-core mode; fp32 = CUDA-core parity mode")
    ap.add_argument("--mstep-variant", type=int, default=int(os.environ.get("GVN_MSTEP_VARIANT", "1")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from gvn.pipeline import McemConfig, Enhancer
    from gvn import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # host side of a rank = packing waveforms/labels into pinned memory: share the cores between the ranks instead of
        # letting every process start one intra-op thread per core
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

    cfg = McemConfig(model="M2", niter=args.niter, nmf_rank=args.rank_k, precision=args.precision,
                     mstep_variant=args.mstep_variant)
    vae = build_model()
    enh = Enhancer(vae, cfg, dev)
    B = args.batch
    x, s, nz, labels = make_inputs(B, first=rank * B)       # this rank's shard of the utterance list
    wavs = list(x)

    gather_buf = [torch.zeros(B, 5, dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None
    utt_ids = torch.arange(rank * B, rank * B + B, device=dev, dtype=torch.float64)


    def finish_step(s_hat, cost, refs, T):
        # per-utterance result rows [utt_id, SI-SDR, SI-SIR, SI-SAR, final cost] (python/metrics.py:12-60 on the
        # device); bringing them together is the only collective of the path
        q = E.energy_ratios(s_hat[:, :refs[0].shape[1]].contiguous(), refs[0], refs[1], T)
        rows = torch.cat([utt_ids[:, None], q, cost[-1][:, None]], 1)
        if world > 1:
            dist.all_gather(gather_buf, rows)
        return rows

    def device_step(up, refs, timers=None, seed=0):
        b = enh.prepare(None, None, seed=seed, uploaded=up)
        s_hat, n_hat, cost = enh.run(b, seed=seed, timers=timers)
        rows = finish_step(s_hat, cost, refs, b.T)
        return s_hat, n_hat, cost, b, rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------
    up = enh.upload(wavs, labels, refs=(s, nz), slot=2)
    refs = (up["ref_s"], up["ref_n"])
    for i in range(args.warmup):
        device_step(up, refs, seed=i)
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
        s_hat, n_hat, cost, b, rows = device_step(up, refs, timers=timers, seed=100 + i)
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
    value = world * B / (ms_step * 1e-3)
    assert bool(torch.isfinite(cost).all()) and bool(torch.isfinite(s_hat).all()), "non-finite result"

    # ---- end-to-end timing (host buffers in, host buffers out) --------------------------
    # Enhancer.enhance_many is the public batched API: every step packs this step's waveforms, labels and
    # metric references into pinned memory, copies them to the device, enhances, and reads both enhanced
    # waveforms + the result rows back into pinned host memory; the upload of step i+1 overlaps step i and its
    # kernels are queued behind those of step i before the host waits for step i.
    def batches(k):
        for _ in range(k):
            yield dict(wavs=wavs, labels=labels, refs=(s, nz))
    def gather_rows(i, cost_d, metrics_d):                 # the step's only collective, queued in stream order
        if world > 1:
            dist.all_gather(gather_buf, torch.cat([utt_ids[:, None], metrics_d, cost_d[-1][:, None]], 1))
    for out in enh.enhance_many(batches(2), seed=150, device_hook=gather_rows):
        pass
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for out in enh.enhance_many(batches(args.steps), seed=200, device_hook=gather_rows):
        h2d, d2h = out["h2d_bytes"], out["d2h_bytes"]
    e1.record()
    barrier()
    e_ms = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (float(e_ms) / args.steps * 1e-3)
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
    frames = sum(b.n_frames_host)
    (R_E, b_E), (R_W, b_W) = cfg.chains()
    sweeps = args.niter * (R_E + b_E) + (R_W + b_W)                 # proposals per step per frame
    e_ms_tot = timers.total_ms("estep")
    flops = args.steps * sweeps * frames * decoder_flops_per_frame()
    e_tflops = flops / (e_ms_tot * 1e-3) / 1e12
    m_ms_tot = timers.total_ms("mstep")
    nmf_bytes = args.steps * args.niter * 2 * (R_E + 1) * b.F * frames * 4
    m_gbs = nmf_bytes / (m_ms_tot * 1e-3) / 1e9
    # DRAM traffic per launch of the two hot kernels from the committed ncu --set full capture (profiles/)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
    except Exception:
        pass
    n_e = timers.count("estep")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16 operands, f32 accumulate",
            "data": "synthetic", "config": workload(args), "clocks": clk,
            "quality": {"si_sdr_db_mean": float(rows[:, 1].mean()), "note": "random-init decoder (no trained weights ship with the "
                        "reference): the number checks plumbing, not enhancement quality"},
            "e2e": {"value": e2e_value, "unit": "utt/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches,
            "roofline": {"kernel": "gvn_estep (decoder MLP + MH chain, %s)" % args.precision, "bound": "tensor",
                         "achieved": e_tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": e_tflops / tf_peak,
                         "traffic": traffic.get("estep_bytes_per_launch"), "peak_source": peak_src + " bf16_tflops_sustained",
                         "avg_launch_ms": e_ms_tot / n_e, "launches": n_e, "share_of_step": e_ms_tot / (ms_step * args.steps)},
            "roofline_nmf": {"kernel": "gvn_mstep (k_tile_meta + k_w_v2 + k_cols_v1)", "bound": "hbm",
                             "achieved": m_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": m_gbs / hbm_peak,
                             "traffic": traffic.get("mstep_bytes_per_launch"), "peak_source": peak_src + " hbm_gbs",
                             "avg_launch_ms": m_ms_tot / timers.count("mstep"),
                             "share_of_step": m_ms_tot / (ms_step * args.steps)},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            cpu_reference_utterance(x[0], s[0], vae, args, 2)          # warm the thread pool and the allocator
            n_cpu = 3                                                  # a bounded sample: ~12 s of host work at C2
            c0 = time.perf_counter()
            for i in range(n_cpu):
                cpu_reference_utterance(x[i % len(x)], s[i % len(s)], vae, args, args.niter)
            t_cpu = time.perf_counter() - c0
            line["cpu_baseline"] = {"value": n_cpu / t_cpu, "unit": "utt/s", "cores": cores, "kind": "port",
                                    "sample": "oracle port, %d utterances of the workload at the full niter=%d, one after "
                                              "the other with all host threads (%.1f s)" % (n_cpu, args.niter, t_cpu)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
