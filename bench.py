#!/usr/bin/env python3
"""bench.py -- IMDCT + window + overlap-add throughput on synthetic 2048-sample long blocks.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W          our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                   the CPU baseline arm (oracle port, host cores)

One step = one pass of the hot path over one batch: S independent stereo streams x P consecutive
long/long packets (spectrum [S][P][2][1024] f32, device-resident, > L2), through
lwb_decode_chains (fused kernel k_long).  `value` = channel-samples per second over all ranks,
timed with CUDA events on the library's stream, max over ranks.  `e2e` = the same call with HOST
(pinned) buffers: H2D of the spectrum and D2H of the PCM inside the timed region.
"""
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

N2 = 1024
ALG_BYTES_PER_SAMPLE = 8          # 4 B spectrum read + 4 B f32 PCM write


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(streams, packets):
    """DRAM bytes (read + write) per k_long launch from the committed `ncu --set full` capture of this
    same workload (profiles/*_k_long_ncu_summary.txt, newest round); None for any other workload."""
    if (streams, packets) != (4096, 16):
        return None, None
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_long_ncu_summary.txt")))
    if not files:
        return None, None
    txt = open(files[-1]).read()
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(re.escape(key) + r"\s+([0-9.]+)\s+(\w+)", txt)
        if not m or m.group(2) not in unit:
            return None, None
        tot += float(m.group(1)) * unit[m.group(2)]
    return tot, os.path.relpath(files[-1], ROOT)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads():
    """Threads the CPU arm can really use: the smaller of the logical CPUs, this process's affinity
    mask and the container's CPU quota (cgroup v2 cpu.max)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return n


def cpu_reference(streams, packets, threads, target_sec=2.0, seed=1234):
    """The oracle port (lewton-equivalent C restatement) on the host cores; returns
    (samples/s, seconds, reps).  The same synthetic input is swept `reps` times so that the timed
    region lasts about target_sec of wall clock on all `threads` cores."""
    from oracle import oracle
    oracle.build()
    rng = np.random.default_rng(seed)
    chains = streams * 2
    spec = (rng.standard_normal((chains, packets, N2)) * 1e-2).astype(np.float32)
    sec1, _ = oracle.bench_chains(11, spec, threads, 1)            # calibration pass (also warms caches)
    reps = max(1, int(target_sec / max(sec1, 1e-4)))
    sec, _ = oracle.bench_chains(11, spec, threads, reps)
    samples = chains * (packets - 1) * N2 * reps     # the first packet of a fresh chain emits nothing
    return samples / sec, sec, reps


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the crate is
    Rust and cannot be built here), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    streams, packets = 8 * threads, 17            # 16 chains per thread, swept ~0.5 s per step
    vals = []
    for i in range(args.warmup + args.steps):
        v, sec, reps = cpu_reference(streams, packets, threads, target_sec=0.5, seed=1234 + i)
        if i >= args.warmup:
            vals.append((v, sec))
    v = float(np.mean([a for a, _ in vals]))
    ms = float(np.mean([b for _, b in vals])) * 1e3
    line = {"impl": "reference", "metric": "Msamples/s IMDCT+window+OLA, 2048-pt long blocks", "value": v / 1e6,
            "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "stereo long-block (n=2048) packets, IMDCT+window+OLA, CPU oracle port",
                       "streams": streams, "packets_per_stream": packets, "channels": 2},
            "cpu_baseline": {"value": v / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
                             "sample": f"{streams} stereo streams x {packets} long packets swept ~0.5 s per step, "
                                       "lewton-equivalent C restatement (oracle/), the crate itself is Rust"},
            "e2e": {"value": v / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="stereo streams per GPU per step")
    ap.add_argument("--packets", type=int, default=16, help="consecutive long packets per stream per step")
    ap.add_argument("--e2e-streams", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import lewton_b200 as L
    from lewton_b200 import _cabi as cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION on some boxes) off it
        os.environ["NCCL_DEBUG"] = os.environ.get("LWB_NCCL_DEBUG", "WARN")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    ctx = L.Context(local)
    # weak scaling: every rank owns `--streams` streams (global stream ids [lo, hi), contiguous ranges,
    # lewton_b200/sharding.py); no data-path collective
    from lewton_b200.sharding import stream_range
    lo, hi = stream_range(args.streams * world, world, rank)
    S, P, C = hi - lo, args.packets, 2
    su = L.Setup(ctx, C, 8, 11, [L.FloorTypeOne(1, [0, 128])], [L.Mapping(C)], [L.ModeInfo(False), L.ModeInfo(True)])
    stream = torch.cuda.ExternalStream(ctx.cuda_stream, device=torch.device("cuda", local))

    # synthetic spectrum, device resident before timing: N(0,1)*1e-2, seed 1234 (+rank)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    spec = torch.randn((S, P, C, N2), generator=gen, device="cuda", dtype=torch.float32) * 1e-2
    stride = P * N2
    pcm = torch.empty((S, C, stride), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    pwrs = [L.PreviousWindowRight(su) for _ in range(S)]
    modes = np.ones(P, np.uint8)
    chains = [L.ChainSpec(pwrs[s], modes, coeff_offset=s * P * C * N2, out_offset=s * C * stride, out_stride=stride)
              for s in range(S)]

    batch = L.Batch(ctx, chains, cabi.ENTRY_SPECTRUM, cabi.MEM_DEVICE, spec.data_ptr(), pcm.data_ptr(),
                    cabi.OUT_F32_PLANAR)
    step = batch.run

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    h0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    host_us = (time.perf_counter() - h0) * 1e6 / args.steps      # host cost of one submission (enqueue only)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - l0
    # The timed region lasts a few milliseconds, far below nvidia-smi's sampling period: keep the same
    # step running (untimed) for ~0.5 s so that the clock / throttle samples are taken under this load.
    try:
        t_probe = time.perf_counter()
        while time.perf_counter() - t_probe < 0.5:
            for _ in range(64):
                step()
            ctx.synchronize()
    except Exception:          # the probe must never cost the measurement
        pass
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled"] = "during the timed steps and ~0.5 s of the same step repeated right after them"
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    # after warm-up every stream has history: every packet emits 1024 samples per channel
    samples_step = S * P * C * N2 * world
    value = samples_step / (ms_step * 1e-3)

    # ---- e2e: host (pinned) buffers through the same call --------------------------------------
    Se = min(args.e2e_streams, S)
    h_spec = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_float * (Se * P * C * N2)).from_address(
        cabi.lib().lwb_host_alloc(Se * P * C * N2 * 4)))
    h_pcm = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_float * (Se * C * stride)).from_address(
        cabi.lib().lwb_host_alloc(Se * C * stride * 4)))
    h_spec[:] = (np.random.default_rng(99 + rank).standard_normal(h_spec.size) * 1e-2).astype(np.float32)
    e_pwrs = [L.PreviousWindowRight(su) for _ in range(Se)]
    e_chains = [L.ChainSpec(e_pwrs[s], modes, coeff_offset=s * P * C * N2, out_offset=s * C * stride,
                            out_stride=stride) for s in range(Se)]

    e_batch = L.Batch(ctx, e_chains, cabi.ENTRY_SPECTRUM, cabi.MEM_HOST, h_spec, h_pcm, cabi.OUT_F32_PLANAR)
    e2e_step = e_batch.run

    for _ in range(3):
        e2e_step()
    barrier()
    e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    barrier()
    e_sec = time.perf_counter() - t0
    te = torch.tensor([e_sec], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_sec = float(te.item())
    e2e_value = Se * P * C * N2 * world * e_steps / e_sec

    if rank == 0:
        peak, peak_src = peaks()
        per_gpu = value / world
        achieved = per_gpu * ALG_BYTES_PER_SAMPLE / 1e9
        traffic, traffic_src = ncu_traffic(S, P)
        line = {"metric": "Msamples/s IMDCT+window+OLA, 2048-pt long blocks", "value": value / 1e6,
                "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "stereo 44.1 kHz long-block (n=2048) packets, batched IMDCT+OLA "
                                       "(BASELINE.json configs[1]); per GPU: streams x packets x 2 channels",
                           "streams_per_gpu": S, "packets_per_stream": P, "channels": C,
                           "bytes_in_per_step_per_gpu": S * P * C * N2 * 4,
                           "l2_policy": "inputs+outputs per step (1 GiB at defaults) exceed the 126 MB L2",
                           "parallelism": f"streams sharded over {world} rank(s), no data-path collective"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                             "traffic_source": traffic_src, "algorithmic_bytes_per_launch": S * P * C * N2 * ALG_BYTES_PER_SAMPLE,
                             "peak_source": peak_src,
                             "kernel": "k_long", "algorithmic_bytes_per_sample": ALG_BYTES_PER_SAMPLE},
                "e2e": {"value": e2e_value / 1e6, "unit": "Msamples/s",
                        "h2d_bytes_per_step": Se * P * C * N2 * 4, "d2h_bytes_per_step": Se * C * stride * 4,
                        "streams": Se, "steps": e_steps, "timer": "host wall clock around synchronous calls"},
                "gpu_launches": int(launches), "host_enqueue_us_per_step": host_us, "clocks": clocks}
        if not args.no_cpu_baseline:
            threads = host_threads()
            v1, _, _ = cpu_reference(8, 17, 1, target_sec=0.5)
            v, sec, reps = cpu_reference(8 * threads, 17, threads, target_sec=2.0)
            line["cpu_baseline"] = {"value": v / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
                                    "single_thread_value": v1 / 1e6,
                                    "sample": f"{8 * threads} stereo streams x 17 long packets swept {reps}x = "
                                              f"{sec:.2f} s wall on {threads} threads ({sec * threads:.0f} core-s); "
                                              "lewton-equivalent C restatement (oracle/), the crate itself is Rust"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
