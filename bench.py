#!/usr/bin/env python
"""bench.py -- spectrogram bins/s of batch_lws (100 iterations) on B200, with the reference's CPU
path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg1|cfg2|cfg3|cfg4|cfg5] [--thresholds default|zero] [--cpu-seconds S]

A *step* is one full call of the workload's hot path on one batch of synthetic magnitude spectrograms
(BASELINE.json configs[1] by default: 64 utterances x 10 s at 16 kHz, 1024-pt STFT, hop 256,
batch_lws with 100 iterations, default thresholds; cfg3 = configs[2]: nofuture_lws + online_lws in RTISI-LA mode;
cfg4 = configs[3]: the full run_lws chain, 32 utterances per GPU = 256 over 8; cfg5 = configs[4]: 2048-pt frames,
200 iterations, 4 utterances per GPU = 32 over 8).  One JSON line on stdout (rank 0):

* ``value``   bins/s with the magnitudes already resident in HBM: lwsb_load (extend, |.|, mean)
              + the sweep kernel + lwsb_store (crop) per step, CUDA events on the launching
              stream, max over ranks.  N > 1: every rank runs the same-sized batch on its own
              GPU (utterances are independent: no collective on the data path) -> weak scaling.
* ``e2e``     the same metric through the public API ``lws_b200.lws(...).batch_lws(A, out=Y)``
              with pinned HOST buffers, H2D and D2H inside the timed region.
* ``roofline`` algorithmic bytes (40 B per bin-iteration) of the sweep kernel
              over its CUDA-event duration, against MEASURED_PEAKS.json's copy bandwidth.
* ``cpu_baseline`` the compiled reference (oracle/_ref, kind "reference") or the oracle port
              (kind "port"), one core, on a bounded sample of the same utterances.

``--impl reference`` times the reference's own CPU implementation on all host cores (one
process per core over utterances; the reference itself is single-threaded).
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config index in BASELINE.json, utterances per GPU, samples, fsize, hop, batch iterations)
    "cfg1": (0, 1, 32000, 512, 128, 100),
    "cfg2": (1, 64, 160000, 1024, 256, 100),
    "cfg3": (2, 64, 160000, 1024, 256, 0),     # mode='music' without the batch stage: nofuture 1 sweep + online 10 iterations, LA 3
    "cfg4": (3, 32, 160000, 1024, 256, 100),   # mode='music': nofuture -> online -> batch; 256 utterances over 8 GPUs
    "cfg5": (4, 4, 1440000, 2048, 256, 200),   # 32 utterances over 8 GPUs
}
MUSIC = ("cfg3", "cfg4")                         # workloads run with lws.lws(..., mode='music') (lws.pyx:432-437)
STAGES = {"cfg3": "nofuture_lws + online_lws", "cfg4": "run_lws: nofuture -> online -> batch"}
ALGO_BYTES_PER_BIN_ITER = 40.0  # 16 B read + 16 B write of the complex128 state + 8 B amplitude
FLOP_PER_ACTIVE_BIN = {2: 100.0, 4: 270.0, 8: 902.0}  # separately rounded fp64 operations per updated bin (batch sweeps, default windows):
                                                       # Q = 4 folded: 256 adds / multiplies + projection; Q = 8: 74 terms x 12 + projection


def metric_name(name):
    it = WORKLOADS[name][5]
    if name in MUSIC:
        return "spectrogram bins/sec (%s)" % STAGES[name]
    return "spectrogram bins/sec (batch_lws, %d iters)" % it


def make_plugin(mod, name, **kw):
    """the reference-API object of the workload: lws.lws(fsize, hop[, mode='music'][, batch_iterations=0])"""
    idx, B, n, fs, hop, it = WORKLOADS[name]
    if name in MUSIC:
        return mod.lws(fs, hop, mode="music", batch_iterations=it, **kw)
    return mod.lws(fs, hop, **kw)


def hot_path(p, name, A, thr, **kw):
    """one call of the workload's hot path through the reference API"""
    if name in MUSIC:
        return p.run_lws(A, **kw)       # nofuture -> online (-> batch when batch_iterations > 0), lws.pyx:495-499
    return p.batch_lws(A, thresholds=thr, **kw)


def workload_desc(name, thresholds):
    idx, B, n, fs, hop, it = WORKLOADS[name]
    what = ("batch_lws %d iters, %s thresholds" % (it, thresholds)) if name not in MUSIC else (
        "mode='music' (nofuture 1 iter, online 10 iters, look-ahead 3)" + (", batch %d iters" % it if it else ", no batch stage") + ": " + STAGES[name])
    return {"workload": "BASELINE.json configs[%d]: %d utterances x %d samples per GPU, %d-pt STFT hop %d, %s" % (idx, B, n, fs, hop, what),
            "utterances_per_gpu": B, "fsize": fs, "hop": hop, "iterations": it, "thresholds": thresholds,
            "l2": "state per step (extended spectrogram + amplitude, fp64) is larger than the 126 MB L2"
                  if name != "cfg1" else "L2 flushed between steps"}


def signals(name, rank=0):
    idx, B, n, fs, hop, it = WORKLOADS[name]
    c = idx + 1  # utterance b of config c is seeded 1000*c + b
    return np.stack([np.random.default_rng(1000 * c + b + 100000 * rank).standard_normal(n) for b in range(B)])


def thresholds_for(name, mode):
    it = WORKLOADS[name][5]
    if mode == "zero":
        return np.zeros(it)
    return 100.0 * np.exp(-0.1 * np.arange(it))  # lws.pyx:203-206 with the class defaults (lws.pyx:382)


# ---------------------------------------------------------------------------------- CPU side
def load_reference():
    """(module, kind): the compiled reference from oracle/_ref if present, else the oracle port."""
    d = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(d):
        for f in sorted(os.listdir(d)):
            if f.startswith("lws_ref") and f.endswith(".so"):
                try:
                    spec = importlib.util.spec_from_file_location("lws_ref", os.path.join(d, f))
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    return mod, "reference"
                except Exception:
                    break
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lws_oracle
    lws_oracle.lib()
    return lws_oracle, "port"


_CPU_INPUTS = {}  # utterance index -> magnitude spectrogram; filled before the pool forks


def _cpu_worker(args):
    name, thr_mode, b = args
    mod, _ = load_reference()
    idx, B, n, fs, hop, it = WORKLOADS[name]
    A = _CPU_INPUTS[b]
    p = make_plugin(mod, name)
    thr = thresholds_for(name, thr_mode)
    t0 = time.perf_counter()
    hot_path(p, name, A, thr)
    return time.perf_counter() - t0, A.size


def cpu_baseline_one_core(name, thr_mode, budget_s):
    """bins/s of the reference on ONE core over a bounded sample of the workload's utterances."""
    mod, kind = load_reference()
    idx, B, n, fs, hop, it = WORKLOADS[name]
    p = make_plugin(mod, name)
    thr = thresholds_for(name, thr_mode)
    bins, secs, used = 0, 0.0, 0
    for b in range(B):
        x = np.random.default_rng(1000 * (idx + 1) + b).standard_normal(n)
        A = np.abs(p.stft(x))
        if name == "cfg5":  # 158 s per utterance: time a 1/16 slice of the frames instead
            A = A[: A.shape[0] // 16]
        t0 = time.perf_counter()
        hot_path(p, name, A, thr)
        secs += time.perf_counter() - t0
        bins += A.size
        used += 1
        if secs >= budget_s:
            break
    return {"value": bins / secs, "unit": "bins/s", "cores": 1, "kind": kind,
            "sample": "%d of %d utterances of the workload, sequential, %.1f s" % (used, B, secs)}


def run_reference_arm(args):
    """--impl reference: all host cores, one process per core (the reference is single-threaded)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    idx, B, n, fs, hop, it = WORKLOADS[name]
    mod, kind = load_reference()
    cores = os.cpu_count() or 1
    per_step = min(cores, B)
    p = make_plugin(mod, name)
    for b in range(per_step):  # untimed set-up: the step times the hot path, not input synthesis
        A = np.abs(p.stft(np.random.default_rng(1000 * (idx + 1) + b).standard_normal(n)))
        _CPU_INPUTS[b] = A[: A.shape[0] // 16] if name == "cfg5" else A
    ctx = mp.get_context("fork")
    times = []
    bins_step = 0
    with ctx.Pool(min(cores, per_step)) as pool:
        for s in range(args.warmup + args.steps):
            jobs = [(name, args.thresholds, j) for j in range(per_step)]
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
                bins_step = sum(r[1] for r in res)
            if sum(times) > 240:  # keep the whole run within a few minutes
                break
    steps = len(times)
    tot = sum(times)
    value = bins_step * steps / tot
    sample = "%d utterances of the workload per step on %d processes%s" % (
        per_step, min(cores, per_step), " (first 1/16 of the frames of each)" if name == "cfg5" else "")
    line = {"impl": "reference", "metric": metric_name(name), "value": value,
            "unit": "bins/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_desc(name, args.thresholds),
            "cpu_baseline": {"value": value, "unit": "bins/s", "cores": min(cores, per_step), "kind": kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": "bins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cpus": cores}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------- GPU side
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class Ranks(object):
    """One process per GPU (torchrun): utterances are independent, so the only traffic between
    ranks is the barrier around the timed region and a MAX over the per-rank device times."""

    def __init__(self, backend):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            # (NCCL_DEBUG is left as the launcher set it: main() points fd 1 at stderr while the arm runs, so
            # NCCL's banner cannot reach the JSON line, and the driver can read the rank check from stderr)
            if backend == "nccl":
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            else:
                dist.init_process_group(backend)

    def barrier(self):
        if self.backend == "nccl":
            self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        if self.backend == "nccl":
            self.torch.cuda.synchronize()

    def max(self, x):
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def throughput(units_per_rank_step, steps, world_units_factor, seconds_max):
    """whole-job units per second: every rank did `units_per_rank_step` per step; time = max over ranks"""
    return units_per_rank_step * world_units_factor * steps / seconds_max


def active_bin_iters(A, thr):
    """exact number of bin updates a batch_lws call performs: bins with |S| > thresholds[i] * mean|S| (lwslib.cpp:295-296),
    summed over sweeps and utterances"""
    total = 0
    for b in range(A.shape[0]):
        a = np.sort(A[b].ravel())
        lim = thr * np.mean(A[b])
        total += int((a.size - np.searchsorted(a, lim, side="right")).sum())
    return total


def online_chain_len(T, it, LA):
    """row updates of TF_RTISI_LA (lwslib.cpp:1432-1491): sum over frames of 1 + it * (min(LA, m) + 1)"""
    return sum(1 + it * (min(LA, m) + 1) for m in range(T))


def run_b200_arm(args):
    import torch
    import lws_b200
    from lws_b200 import _native

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    ranks = Ranks("nccl")
    world, rank, local = ranks.world, ranks.rank, ranks.local
    name = args.workload
    idx, B, n, fs, hop, it = WORKLOADS[name]
    music = name in MUSIC
    thr = thresholds_for(name, args.thresholds)
    p = make_plugin(lws_b200, name, device=local)
    nf_thr = lws_b200.get_thresholds(p.nofuture_iterations, p.nofuture_alpha, p.nofuture_beta, p.nofuture_gamma)
    on_thr = lws_b200.get_thresholds(p.online_iterations, p.online_alpha, p.online_beta, p.online_gamma)
    if music:
        thr = lws_b200.get_thresholds(it, p.batch_alpha, p.batch_beta, p.batch_gamma)  # what run_lws uses (lws.pyx:487-499)

    # synthetic magnitudes (untimed set-up): |STFT| of white noise, computed by the library's own stft
    x = signals(name, rank)
    A_host = np.abs(p.stft(x))                       # (B, T, Nreal) float64
    Bn, T, Nreal = A_host.shape
    bins_rank = Bn * T * Nreal

    # ---- value: data resident in HBM, CUDA events on the stream the kernels run on
    stream = torch.cuda.current_stream()
    ctx = _native.Context(local, stream.cuda_stream)
    ctx.set_weights(_native.W, p.W)
    if music:
        ctx.set_weights(_native.W_AI, p.W_ai)
        ctx.set_weights(_native.W_AF, p.W_af)
    A_dev = torch.from_numpy(A_host).cuda()
    Y_dev = torch.empty((Bn, T, Nreal), dtype=torch.complex128, device="cuda")
    in_ptrs = [A_dev[b].data_ptr() for b in range(Bn)]
    out_ptrs = [Y_dev[b].data_ptr() for b in range(Bn)]
    Ts = [T] * Bn
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if name == "cfg1" else None

    def step_device():
        if music:
            ctx.run_lws_device(in_ptrs, out_ptrs, Ts, Nreal, _native.F64, nf_thr, on_thr, p.look_ahead, thr)
        else:
            ctx.load_device(in_ptrs, Ts, Nreal, _native.F64)
            ctx.batch(thr)
            ctx.store_device(out_ptrs)

    barrier = ranks.barrier

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()   # every rank watches its own GPU: the slowest rank sets the time
    barrier()
    l0 = ctx.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    stage_ms = []
    t_wall0 = time.perf_counter()
    ev[0].record(stream)
    dev_ms = 0.0
    for _ in range(args.steps):
        if flush is not None:
            flush.zero_()  # a memset of 2 x the L2 size: the working set of configs[0] would otherwise stay in L2 from step to step
        step_device()
        stage_ms.append(ctx.last_stage_ms())
    ev[1].record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    dev_ms = ev[0].elapsed_time(ev[1])
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    dev_ms_max = ranks.max(dev_ms)
    per_rank = [None] * world
    mine = {"rank": rank, "device_ms_per_step": dev_ms / args.steps, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")}
    if world > 1:
        ranks.dist.all_gather_object(per_rank, mine)
    else:
        per_rank = [mine]
    bins_all = ranks.sum(bins_rank)  # every rank holds its own utterances (weak scaling)
    value = throughput(bins_all / world, args.steps, world, dev_ms_max * 1e-3)
    work = ctx.last_batch_work() if it > 0 else None
    plan = ctx.last_batch_plan() if it > 0 else None
    cycles = ctx.last_batch_cycles() if it > 0 else None

    # ---- e2e: public API, pinned host in/out, H2D + D2H inside the timed region
    A_pin = torch.from_numpy(A_host).pin_memory()
    Y_pin = torch.empty((Bn, T, Nreal), dtype=torch.complex128).pin_memory()
    A_np, Y_np = A_pin.numpy(), Y_pin.numpy()
    e2e_warm = max(1, min(args.warmup, 2))
    for _ in range(e2e_warm):
        hot_path(p, name, A_np, thr, out=Y_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hot_path(p, name, A_np, thr, out=Y_np)
        chk = float(np.abs(Y_np[0, 0, 0]))  # the step's result is read on the host
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_value = throughput(bins_all / world, args.steps, world, ranks.max(e2e_s))

    # the plain drop-in call: pageable numpy in, fresh numpy out (what a user of the reference writes)
    A_page = np.array(A_host)
    for _ in range(2):  # a caller's loop `Y = p.batch_lws(A)`: the previous result is alive during the call
        Y_plain = hot_path(p, name, A_page, thr)
    barrier()
    t0 = time.perf_counter()
    nplain = max(1, min(args.steps, 3))
    for _ in range(nplain):
        Y_plain = hot_path(p, name, A_page, thr)
    plain_s = (time.perf_counter() - t0) / nplain
    plain_s = ranks.max(plain_s)

    # sanity: device-resident and host paths agree bit for bit, magnitudes preserved
    Yd = Y_dev.cpu().numpy()
    assert np.array_equal(Yd, Y_np), "device-resident and host-API results differ"
    assert np.array_equal(np.asarray(Y_plain), Y_np), "plain-call and pinned-buffer results differ"
    assert np.allclose(np.abs(Y_np[0]), A_host[0], rtol=1e-10, atol=1e-12 * A_host.max())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        stages = {k: statistics.mean([m[k] for m in stage_ms]) for k in ("nofuture", "online", "batch") if stage_ms[0][k] is not None}
        # algorithmic bytes per launch of each stage: 40 B per bin per row update
        rows = {"nofuture": Bn * T * len(nf_thr), "online": Bn * online_chain_len(T, len(on_thr), p.look_ahead) if len(on_thr) else 0,
                "batch": Bn * T * it}
        dom = max(stages, key=lambda k: stages[k])
        k_ms = stages[dom]
        algo_bytes = ALGO_BYTES_PER_BIN_ITER * rows[dom] * Nreal
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)",
                "unit": "GB/s", "frac": achieved / peak,
                "kernel": {"batch": "batch sweep kernel (k_batch_strips)", "online": "online chain kernel (%s)" % {0: "k_online_generic", 1: "k_online_ring", 2: "k_online_ring2", 3: "k_online_duo",
                                                                               4: "k_online_flow", 5: "k_online_rail"}.get(ctx.last_online_kernel(), "?"),
                           "nofuture": "no-future sweep kernel"}[dom],
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": algo_bytes,
                "algorithmic_bytes_definition": "40 B x bins x row updates asked for, independent of thresholding (SURVEY.md section 8d)",
                "stage_ms": stages,
                "stage_frac": {k: ALGO_BYTES_PER_BIN_ITER * rows[k] * Nreal / (stages[k] * 1e-3) / 1e9 / peak for k in stages},
                "note": "bit-exact fp64 Gauss-Seidel stencil: bounded by shared-memory bandwidth and fp64 issue, not by HBM "
                        "(DESIGN.md section 5); fp64_frac is the fraction of the non-FMA fp64 issue peak"}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % name)
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch")
            roof["traffic_source"] = "%s (ncu --set full, builder run: %s)" % (os.path.relpath(tp, ROOT), tj.get("source", "see profiles/README.md"))
        roof["traffic"] = traffic
        if it > 0 and "batch" in stages:
            Q = fs // hop
            sm_mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
            sms = ctx.device_info()["sm_count"]
            fp64_peak = sms * 64 * sm_mhz * 1e6          # separately rounded fp64 operations per second (no FMA: the reference has none)
            act = active_bin_iters(A_host, thr)
            flops = FLOP_PER_ACTIVE_BIN.get(Q, 0.0) * act
            roof["fp64_frac"] = flops / (stages["batch"] * 1e-3) / fp64_peak
            roof["fp64_peak_ops"] = fp64_peak
            roof["fp64_ops_per_active_bin"] = FLOP_PER_ACTIVE_BIN.get(Q)
            roof["bin_iters"] = {"asked": int(bins_rank) * it, "in_sweeps_executed": work["bin_iters_executed"] if work else None,
                                 "active_bins_updated": act}
            if work and work["bin_iters_executed"]:
                roof["achieved_executed_sweeps"] = ALGO_BYTES_PER_BIN_ITER * work["bin_iters_executed"] / (stages["batch"] * 1e-3) / 1e9
                roof["frac_executed_sweeps"] = roof["achieved_executed_sweeps"] / peak
        cpu = cpu_baseline_one_core(name, args.thresholds, args.cpu_seconds)
        line = {
            "metric": metric_name(name), "value": value, "unit": "bins/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_desc(name, args.thresholds),
            "bin_iters_per_s": value * max(it, 1),
            "e2e": {"value": e2e_value, "unit": "bins/s", "h2d_bytes_per_step": int(A_host.nbytes),
                    "d2h_bytes_per_step": int(Y_np.nbytes), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "lws_b200.%s, pinned host numpy in/out" % (
                        ("lws(%d, %d, mode='music', batch_iterations=%d).run_lws(A, out=Y)" % (fs, hop, it)) if music
                        else ("lws(%d, %d).batch_lws(A, out=Y)" % (fs, hop))),
                    "plain_call_ms": 1e3 * plain_s,
                    "plain_call": "the same call with pageable numpy in and a fresh numpy result (no out=)"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks, "per_rank": per_rank, "wall_ms_per_step": wall_ms / args.steps,
            "plan": plan, "cycles_cluster0": cycles,
            "kernel_share_of_step": sum(stages.values()) * args.steps / dev_ms if dev_ms else None,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    ranks.close()
    return 0


def run_inprocess_arm(args):
    """--sharding inprocess: ONE process, the product's own multi-GPU path -- lws_b200.lws(..., device=[0 .. N-1]) splits
    the batch over the GPUs (longest-processing-time-first over frame counts, one host thread and context per device, no
    collective).  Host numpy in, host numpy out: the number is end to end by construction.  Run as plain
    `python bench.py --gpus N --sharding inprocess` (not under torchrun)."""
    import torch
    import lws_b200
    name = args.workload
    idx, B, n, fs, hop, it = WORKLOADS[name]
    N = args.gpus
    if torch.cuda.device_count() < N:
        raise RuntimeError("--gpus %d but %d visible" % (N, torch.cuda.device_count()))
    thr = thresholds_for(name, args.thresholds)
    p = make_plugin(lws_b200, name, device=list(range(N)))
    p0 = make_plugin(lws_b200, name, device=0)
    x = np.concatenate([signals(name, r) for r in range(N)])     # N * B utterances: the same per-GPU load as the torchrun arm
    A = np.abs(p0.stft(x[:B]))
    A = np.concatenate([A] + [np.abs(p0.stft(x[r * B:(r + 1) * B])) for r in range(1, N)])
    Bn, T, Nreal = A.shape
    A_pin = torch.from_numpy(A).pin_memory().numpy()
    Y_pin = torch.empty(A.shape, dtype=torch.complex128).pin_memory().numpy()
    for _ in range(args.warmup):
        hot_path(p, name, A_pin, thr, out=Y_pin)
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hot_path(p, name, A_pin, thr, out=Y_pin)
        chk = float(np.abs(Y_pin[0, 0, 0]))
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    Y1 = hot_path(p0, name, A_pin[:2], thr)                      # sharded result == single-GPU result
    assert np.array_equal(np.asarray(Y1), Y_pin[:2]) and np.array_equal(hot_path(p0, name, A_pin[-1:], thr)[0], Y_pin[-1])
    value = Bn * T * Nreal * args.steps / dt
    launches = sum(lws_b200.api._context(d).launch_count() for d in range(N))
    line = {"metric": metric_name(name), "value": value, "unit": "bins/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload_desc(name, args.thresholds), sharding="in-process: lws_b200.lws(..., device=[0..%d]), "
                                                "%d utterances split over the GPUs by frames (LPT), one host thread per GPU" % (N - 1, Bn)),
            "e2e": {"value": value, "unit": "bins/s", "h2d_bytes_per_step": int(A.nbytes), "d2h_bytes_per_step": int(Y_pin.nbytes),
                    "ms_per_step": 1e3 * dt / args.steps, "api": "lws_b200.lws(..., device=list(range(%d))) through the reference API, pinned host numpy in/out" % N},
            "gpu_launches": int(launches), "clocks": clocks,
            "note": "wall-clock around the public call (the path has no device-resident variant); compare with the torchrun arm's e2e"}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--thresholds", default="default", choices=["default", "zero"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--sharding", default="ranks", choices=["ranks", "inprocess"],
                    help="ranks: one process per GPU (torchrun, the driver's contract); inprocess: one process, device=[0..N-1]")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # ONE JSON line on stdout, whatever the libraries underneath print (NCCL writes its version banner to fd 1): file
    # descriptor 1 points at stderr while the arm runs, and the arm's own print() calls are collected and replayed
    # on the real stdout afterwards.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import io
    buf = io.StringIO()
    sys.stdout = buf
    try:
        rc = run_reference_arm(args) if args.impl == "reference" else (
            run_inprocess_arm(args) if args.sharding == "inprocess" else run_b200_arm(args))
    finally:
        sys.stdout = sys.__stdout__
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
        lines = [l for l in buf.getvalue().splitlines() if l.strip()]
        for l in lines:
            print(l, flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
