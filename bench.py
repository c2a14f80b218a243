#!/usr/bin/env python
"""Throughput of the IQ sample chain FreqShifter -> Filter -> Downsampler on B200.

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (the reference's algorithm on host cores)

Workload (BASELINE.json configs[2], the one the metric's "1/2/4/8 B200" sharding
is defined on): a channelizer batch of independent 2.4 MS/s complex-f32
streams, each FreqShifter(per-stream shift) -> Filter(3 kHz low-pass, n = 4096)
-> Downsampler(48 kS/s, bandwidth 6 kHz).  One step = one push of
CHUNKS_PER_STEP chunks for every stream.  Streams are sharded over ranks with no
data-path collective (weak scaling: STREAMS_PER_GPU per rank).

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 2_400_000.0
CHUNK_LEN = 4096
OUT_RATE = 48_000.0
BANDWIDTH = 6_000.0
CUTOFF = 3_000.0
OUT_CHUNK = 2048
BYTES_PER_SAMPLE = 8.0 * (1.0 + OUT_RATE / SAMPLE_RATE)  # read once + written once
METRIC = "complex MS/s via FreqShift->Filter->Downsampler"
UNIT = "MS/s"


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the `ncu --set full`
# capture of this very command (profiles/r01_ncu_k_front_k_poly2.txt), keyed by (kernel, streams, chunks)
NCU_TRAFFIC = {("k_front", 4096, 50): 6.715032e9 + 1.591369e9}


def stream_shift(stream_id: int) -> float:
    """SURVEY.md 8(d), config C3."""
    return float((stream_id * 577) % 2_400_000 - 1_200_000)


def lowpass(cutoff):
    def f(_bin, freq):
        return 1.0 + 0.0j if abs(freq) <= cutoff else 0.0j

    return f


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self, name):
        """wall-clock bracket of the timed region: only samples inside it are reported"""
        setattr(self, name, time.time())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        inside = [ln for ts, ln in self.lines if t0 is not None and t1 is not None and t0 <= ts <= t1 + 0.02]
        scope = "timed region" if inside else "around the timed region"
        for ln in (inside or [ln for _, ln in self.lines]):
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "scope": scope,
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle restatement) on the host cores
# ---------------------------------------------------------------------------
_CPU_INPUT = {}


def cpu_chain_run(cores: int, streams_per_core: int, n_chunks: int):
    """Times the oracle's C restatement of the reference loops (oracle/radiorust_oracle.c) on `cores`
    threads over independent streams; returns (MS/s, seconds, sample description)."""
    from oracle import oracle_c
    from oracle import radiorust_oracle as orc

    S = cores * streams_per_core
    key = (S, n_chunks)
    if key not in _CPU_INPUT:  # synthetic input of the sample: generated once, reused by every step
        base = orc.synth_noise(20260000 + 3 * 100000, n_chunks * CHUNK_LEN, "f32")
        _CPU_INPUT.clear()
        _CPU_INPUT[key] = np.stack([np.roll(base, 977 * s) for s in range(S)])
        kw0 = dict(shifts=[stream_shift(s) for s in range(cores)], freq_resp=orc.lowpass(CUTOFF), down=(OUT_RATE, BANDWIDTH, 3.0),
                   n_threads=cores)
        oracle_c.chain(_CPU_INPUT[key][:cores, : 2 * CHUNK_LEN], "f32", SAMPLE_RATE, CHUNK_LEN, **kw0)  # warm up
    x = _CPU_INPUT[key]
    shifts = [stream_shift(s) for s in range(S)]
    kw = dict(shifts=shifts, freq_resp=orc.lowpass(CUTOFF), down=(OUT_RATE, BANDWIDTH, 3.0), n_threads=cores)
    t = {}
    oracle_c.chain(x, "f32", SAMPLE_RATE, CHUNK_LEN, timing=t, **kw)
    dt = t["seconds"]
    samples = S * n_chunks * CHUNK_LEN
    desc = (f"{S} streams x {n_chunks} chunks x {CHUNK_LEN} samples of the same chain; C restatement of the reference loops "
            f"(oracle/radiorust_oracle.c, gcc -O3, radix-2 FFT standing in for rustfft), {cores} pthreads over streams")
    return samples / dt / 1e6, dt, desc


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, desc = cpu_chain_run(cores, 8, 1000)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": "configs[2] channelizer batch: independent 2.4 MS/s complex-f32 streams, FreqShifter(per-stream shift) -> "
                    "Filter(3 kHz low-pass, n=4096, N=8192) -> Downsampler(48 kS/s, bw 6 kHz, L=343)",
        "streams_per_gpu": args.streams, "streams_total": args.streams * world, "chunk_len": CHUNK_LEN,
        "chunks_per_step": args.chunks, "sample_rate": SAMPLE_RATE, "output_rate": OUT_RATE,
        "l2_policy": "inputs larger than L2 (%.2f GB per step per GPU), no flush" % (args.streams * args.chunks * CHUNK_LEN * 8 / 1e9),
        "sharding": "streams block-partitioned over ranks, no collective on the data path",
    }


def bind_to_gpu_cpus(index):
    """Keep this rank's threads (and, by first touch, its pinned host buffers) on the CPUs next to its GPU, as a
    multi-GPU host application would: eight ranks copying from one NUMA node share that node's memory bandwidth.
    Returns the number of CPUs bound to, or None when NVML / the affinity call is not available."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import radiorust_b200 as rr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_cpus(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")  # the gather runs under the next push's kernels
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    S, C = args.streams, args.chunks
    length = C * CHUNK_LEN
    ctx = rr.Context(local)
    stages = [rr.FreqShifter(0.0), rr.Filter.new(lowpass(CUTOFF)), rr.Downsampler(OUT_CHUNK, OUT_RATE, BANDWIDTH)]
    chain = rr.Chain(ctx, stages, "f32", n_streams=S)
    chain.set_shifts(0, [stream_shift(rank * S + s) for s in range(S)])

    gen = torch.Generator(device="cuda")
    gen.manual_seed(20260000 + 3 * 100000 + rank)
    x = torch.randn((S, length, 2), device="cuda", dtype=torch.float32, generator=gen)
    cap = chain.max_output(SAMPLE_RATE, CHUNK_LEN, C + 1) + OUT_CHUNK
    y = torch.zeros((S, cap, 2), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()

    ext = torch.cuda.ExternalStream(chain.cuda_stream)

    def step():
        return chain.push_device(SAMPLE_RATE, CHUNK_LEN, C, x.data_ptr(), length, y.data_ptr(), cap, cap)

    for _ in range(args.warmup):
        step()
    chain.sync()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    launches0 = rr.kernel_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    chain.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out_total = 0
    if rank == 0:
        time.sleep(0.05)  # let the sampler come up
        sampler.mark("t0")
    with torch.cuda.stream(ext):
        e0.record(ext)
        for _ in range(args.steps):
            cnt, _rate = step()
            out_total += cnt
        e1.record(ext)
    chain.sync()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.mark("t1")
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = rr.kernel_launch_count() - launches0
    elapsed_ms = e0.elapsed_time(e1)
    k_ms, k_n, k_name = chain.kernel_time()
    k_all = chain.kernel_breakdown()
    chain.set_timing(False)
    plan = chain.plan
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    samples_step = S * length
    value = samples_step * world * args.steps / (elapsed_ms * 1e-3) / 1e6

    # ---- the one collective of the path: gathering the channelizer outputs --------------
    # Outside the timed region and reported separately: every rank contributes its streams' decimated
    # samples of one step; NCCL all-gather over NVLink, timed on the device, max over ranks.
    gather = None
    if dist is not None:
        per_stream = out_total // max(args.steps, 1)
        mine = y[:, :per_stream].contiguous()
        everyone = torch.empty((world * S, per_stream, 2), device="cuda", dtype=torch.float32)
        dist.all_gather_into_tensor(everyone, mine)  # warm-up (communicator set-up)
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        g0.record()
        for _ in range(reps):
            dist.all_gather_into_tensor(everyone, mine)
        g1.record()
        torch.cuda.synchronize()
        t = torch.tensor([g0.elapsed_time(g1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gms = float(t.item())
        same = bool(torch.equal(everyone[rank * S:(rank + 1) * S], mine))
        gather = {"collective": "nccl all_gather of one step's outputs (outside the timed region)", "ms": gms,
                  "bytes_per_rank": mine.numel() * 4, "bytes_total": everyone.numel() * 4,
                  "share_of_step": gms / (elapsed_ms / args.steps), "own_slice_intact": same}
        # the same steps again with every step's gather running on a second stream while the next step computes
        # (two output buffers; a push waits for the gather that last read its buffer)
        ys = [y, torch.zeros_like(y)]
        side = torch.cuda.Stream(priority=-1)  # its kernels take freed SM slots ahead of the next push's CTAs
        staged = torch.empty_like(mine)
        busy = [None, None]

        def overlapped(k_steps):
            for i in range(k_steps):
                yi = ys[i & 1]
                with torch.cuda.stream(ext):
                    if busy[i & 1] is not None:
                        ext.wait_event(busy[i & 1])
                    chain.push_device(SAMPLE_RATE, CHUNK_LEN, C, x.data_ptr(), length, yi.data_ptr(), cap, cap)
                    done = torch.cuda.Event()
                    done.record(ext)
                with torch.cuda.stream(side):
                    side.wait_event(done)
                    staged.copy_(yi[:, :per_stream])
                    dist.all_gather_into_tensor(everyone, staged)
                    busy[i & 1] = torch.cuda.Event()
                    busy[i & 1].record(side)

        overlapped(2)
        torch.cuda.synchronize()
        dist.barrier()
        o0, o1, o2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        o0.record(ext)
        overlapped(args.steps)
        o1.record(ext)
        o2.record(side)
        chain.sync()
        torch.cuda.synchronize()
        t = torch.tensor([max(o0.elapsed_time(o1), o0.elapsed_time(o2))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        oms = float(t.item())
        gather["overlapped"] = {"ms_per_step": oms / args.steps, "value": S * length * world * args.steps / (oms * 1e-3) / 1e6, "unit": UNIT,
                                "what": "steps with each step's all_gather on a second stream under the next step's kernels"}
        del everyone, mine, staged, ys

    # ---- end to end: pinned host chunks in, host result out, through rr_chain_push ----
    e2e = None
    if not args.no_e2e:
        e_steps = max(1, min(args.steps, args.e2e_steps))
        hx = torch.empty((S, length, 2), dtype=torch.float32, pin_memory=True)
        blk = torch.randn((min(S, 64), length, 2), dtype=torch.float32)
        for s0 in range(0, S, blk.shape[0]):
            hx[s0 : s0 + blk.shape[0]] = blk[: min(blk.shape[0], S - s0)]
        hy = torch.empty((S, cap, 2), dtype=torch.float32, pin_memory=True)

        def estep():
            cnt, _ = chain.push_host_async(SAMPLE_RATE, CHUNK_LEN, C, hx.data_ptr(), length, hy.data_ptr(), cap, cap)
            chain.sync()
            return cnt

        estep()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        d2h = 0
        with torch.cuda.stream(ext):
            e0.record(ext)
            for _ in range(e_steps):
                d2h += estep() * S * 8
            e1.record(ext)
        chain.sync()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ems], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {
            "value": samples_step * world * e_steps / (ems * 1e-3) / 1e6, "unit": UNIT,
            "h2d_bytes_per_step": samples_step * 8, "d2h_bytes_per_step": d2h // e_steps, "steps": e_steps,
            "api": "rr_chain_push (pinned host chunks -> H2D -> kernels -> D2H), per rank",
            "cpus_bound_per_rank": numa,
        }
        del hx, hy

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peaks()
    k_avg_ms = k_ms / max(k_n, 1)
    achieved = samples_step * BYTES_PER_SAMPLE / (k_avg_ms * 1e-3) / 1e9 if k_n else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (torch.randn on device, seeded)",
        "config": workload_config(args, world),
        "plan": plan,
        "clocks": clocks,
        "e2e": e2e,
        "gather": gather,
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "traffic": args.traffic if args.traffic is not None else NCU_TRAFFIC.get((k_name, S, C)), "peak_kind": peak_kind, "kernel": k_name, "kernel_ms": k_avg_ms, "kernel_launches": k_n,
            "algorithmic_bytes_per_launch": samples_step * BYTES_PER_SAMPLE,
            "kernel_share_of_step": (k_ms / elapsed_ms) if elapsed_ms else None,
            "kernels": {nm: {"ms_per_launch": ms / max(n, 1), "launches": n, "share_of_step": ms / elapsed_ms}
                        for nm, (ms, n) in k_all.items()},
            # the same algorithmic bytes over the whole step (all kernels of a push), for comparison
            "step_achieved": value * 1e6 * BYTES_PER_SAMPLE / world / 1e9, "step_frac": value * 1e6 * BYTES_PER_SAMPLE / world / 1e9 / peak,
        },
        "output_samples_per_step": out_total // max(args.steps, 1),
    }
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        v, dt, desc = cpu_chain_run(cores, 8, 1000)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc, "seconds": dt}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU")
    ap.add_argument("--chunks", type=int, default=50, help="chunks per stream per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch from an ncu --set full capture")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
