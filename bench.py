#!/usr/bin/env python
"""bench.py -- spectrogram bins/s of batch_lws (100 iterations) on B200, with the reference's CPU
path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg1|cfg2|cfg5] [--thresholds default|zero] [--cpu-seconds S]

A *step* is one full ``batch_lws`` call on one batch of synthetic magnitude spectrograms
(BASELINE.json configs[1] by default: 64 utterances x 10 s at 16 kHz, 1024-pt STFT, hop 256,
100 iterations, default thresholds).  One JSON line on stdout (rank 0):

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
    # name: (config index in BASELINE.json, utterances, samples, fsize, hop, iterations)
    "cfg1": (0, 1, 32000, 512, 128, 100),
    "cfg2": (1, 64, 160000, 1024, 256, 100),
    "cfg5": (4, 4, 1440000, 2048, 256, 200),
}
ALGO_BYTES_PER_BIN_ITER = 40.0  # 16 B read + 16 B write of the complex128 state + 8 B amplitude


def workload_desc(name, thresholds):
    idx, B, n, fs, hop, it = WORKLOADS[name]
    return {"workload": "BASELINE.json configs[%d]: %d utterances x %d samples, %d-pt STFT hop %d, batch_lws %d iters, "
                        "%s thresholds" % (idx, B, n, fs, hop, it, thresholds),
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
    t0 = time.perf_counter()
    mod.lws(fs, hop).batch_lws(A, thresholds=thresholds_for(name, thr_mode))
    return time.perf_counter() - t0, A.size


def cpu_baseline_one_core(name, thr_mode, budget_s):
    """bins/s of the reference on ONE core over a bounded sample of the workload's utterances."""
    mod, kind = load_reference()
    idx, B, n, fs, hop, it = WORKLOADS[name]
    p = mod.lws(fs, hop)
    thr = thresholds_for(name, thr_mode)
    bins, secs, used = 0, 0.0, 0
    for b in range(B):
        x = np.random.default_rng(1000 * (idx + 1) + b).standard_normal(n)
        A = np.abs(p.stft(x))
        if name == "cfg5":  # 158 s per utterance: time a 1/16 slice of the frames instead
            A = A[: A.shape[0] // 16]
        t0 = time.perf_counter()
        p.batch_lws(A, thresholds=thr)
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
    p = mod.lws(fs, hop)
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
    line = {"impl": "reference", "metric": "spectrogram bins/sec (batch_lws, %d iters)" % it, "value": value,
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
            os.environ.pop("NCCL_DEBUG", None)  # no NCCL banner (stdout is guarded as well, see main())
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
    thr = thresholds_for(name, args.thresholds)
    p = lws_b200.lws(fs, hop, device=local)

    # synthetic magnitudes (untimed set-up): |STFT| of white noise, computed by the library's own stft
    x = signals(name, rank)
    A_host = np.abs(p.stft(x))                       # (B, T, Nreal) float64
    Bn, T, Nreal = A_host.shape
    bins_rank = Bn * T * Nreal

    # ---- value: data resident in HBM, CUDA events on the stream the kernels run on
    stream = torch.cuda.current_stream()
    ctx = _native.Context(local, stream.cuda_stream)
    ctx.set_weights(_native.W, p.W)
    A_dev = torch.from_numpy(A_host).cuda()
    Y_dev = torch.empty((Bn, T, Nreal), dtype=torch.complex128, device="cuda")
    in_ptrs = [A_dev[b].data_ptr() for b in range(Bn)]
    out_ptrs = [Y_dev[b].data_ptr() for b in range(Bn)]
    Ts = [T] * Bn
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if name == "cfg1" else None

    def step_device():
        ctx.load_device(in_ptrs, Ts, Nreal, _native.F64)
        ctx.batch(thr)
        ctx.store_device(out_ptrs)

    barrier = ranks.barrier

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = ctx.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kernel_ms = []
    t_wall0 = time.perf_counter()
    ev[0].record(stream)
    dev_ms = 0.0
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        step_device()
        kernel_ms.append(ctx.last_compute_ms())
    ev[1].record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    dev_ms = ev[0].elapsed_time(ev[1])
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms_max = ranks.max(dev_ms)
    bins_all = ranks.sum(bins_rank)  # every rank holds its own utterances (weak scaling)
    value = throughput(bins_all / world, args.steps, world, dev_ms_max * 1e-3)

    # ---- e2e: public API, pinned host in/out, H2D + D2H inside the timed region
    A_pin = torch.from_numpy(A_host).pin_memory()
    Y_pin = torch.empty((Bn, T, Nreal), dtype=torch.complex128).pin_memory()
    A_np, Y_np = A_pin.numpy(), Y_pin.numpy()
    e2e_warm = max(1, min(args.warmup, 2))
    for _ in range(e2e_warm):
        p.batch_lws(A_np, thresholds=thr, out=Y_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p.batch_lws(A_np, thresholds=thr, out=Y_np)
        chk = float(np.abs(Y_np[0, 0, 0]))  # the step's result is read on the host
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_value = throughput(bins_all / world, args.steps, world, ranks.max(e2e_s))

    # sanity: device-resident and host paths agree bit for bit, magnitudes preserved
    Yd = Y_dev.cpu().numpy()
    assert np.array_equal(Yd, Y_np), "device-resident and host-API results differ"
    assert np.allclose(np.abs(Y_np[0]), A_host[0], rtol=1e-10, atol=1e-12 * A_host.max())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k_ms = statistics.mean(kernel_ms)
        achieved = ALGO_BYTES_PER_BIN_ITER * bins_rank * it / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % name)
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        cpu = cpu_baseline_one_core(name, args.thresholds, args.cpu_seconds)
        line = {
            "metric": "spectrogram bins/sec (batch_lws, %d iters)" % it, "value": value, "unit": "bins/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_desc(name, args.thresholds),
            "bin_iters_per_s": value * it,
            "e2e": {"value": e2e_value, "unit": "bins/s", "h2d_bytes_per_step": int(A_host.nbytes),
                    "d2h_bytes_per_step": int(Y_np.nbytes), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "lws_b200.lws(%d, %d).batch_lws(A, out=Y), pinned host numpy in/out" % (fs, hop)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "batch sweep kernel", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_BIN_ITER * bins_rank * it,
                         "note": "bit-exact fp64 Gauss-Seidel stencil: bounded by shared-memory bandwidth and in-order fp64 issue of the few warps the ring leaves room for, not by HBM (DESIGN.md section 5)"},
            "cpu_baseline": cpu,
            "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
            "plan": ctx.last_batch_plan(), "cycles_cluster0": ctx.last_batch_cycles(),
            "kernel_share_of_step": k_ms * args.steps / dev_ms if dev_ms else None,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    ranks.close()
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
        rc = run_reference_arm(args) if args.impl == "reference" else run_b200_arm(args)
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
