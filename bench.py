#!/usr/bin/env python3
"""bench.py — BASELINE.json's headline metric: contains_re(r'\\b\\w{4,}\\b') over the C2 corpus
(10 M strings / 1 GiB chars per GPU), strings/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rows R --bytes B]

* a "step" is one contains_re call over one device-resident column shard (device bool results).
* N > 1: launched by torchrun, one rank per GPU; every rank owns an independent 10 M-row shard (row sharding,
  no data-path collective => "weak" scaling); the timed region is bracketed by barrier + synchronize and the
  reported time is the MAX over ranks (all_reduce MAX over NCCL).
* `value`  : whole-job strings/s with inputs resident in HBM (CUDA events).
* `e2e`    : same metric through the public API from HOST buffers: nvstrings.from_offsets(pinned host) ->
             contains(devptr) -> device->host copy of the bool results, all inside the timed region.
* `roofline`: the dominant kernel's algorithmic bytes (chars + offsets + validity + results) over its
             average device time measured with CUDA events on the launch stream (custr_set_profiling), against the
             measured HBM copy peak in MEASURED_PEAKS.json.
* `cpu_baseline`: the reference's own CPU build (oracle/_ref) on a bounded sample, 1 thread, rank 0, N=1 only.
* `--impl reference`: the reference's CPU implementation on all host threads over bounded samples per step.
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

PATTERN = r"\b\w{4,}\b"
METRIC = "contains_re strings/sec (10M strings, 1 GiB chars per GPU)"


def algorithmic_bytes(n_rows, n_chars):
    # read chars + offsets + validity, write 1 result byte per row
    return n_chars + 4 * (n_rows + 1) + (n_rows + 7) // 8 + n_rows


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        try:  # NVML in-process: a sample every few ms (an nvidia-smi spawn takes longer than the whole timed region)
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.samples.append([str(sm), str(mx)] + ["Active" if r & bits[k] else "Not Active"
                                                          for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                time.sleep(0.004)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def reference_rate(chars, offsets, validity, rows, procs):
    """strings/s of the reference CPU build (oracle/_ref) over the first `rows` rows with `procs` processes."""
    from custrings_b200.workloads import slice_rows
    import multiprocessing as mp
    bounds = [rows * i // procs for i in range(procs + 1)]
    shards = [slice_rows(chars, offsets, validity, bounds[i], bounds[i + 1]) for i in range(procs)]
    if procs == 1:
        t0 = time.perf_counter()
        _ref_worker(shards[0])
        return rows / (time.perf_counter() - t0)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_ref_touch, range(procs))  # load the .so in every worker before timing
        t0 = time.perf_counter()
        pool.map(_ref_worker, shards)
        dt = time.perf_counter() - t0
    return rows / dt


def _ref_touch(_):
    from oracle import ref
    ref.lib()
    return 0


def _ref_worker(shard):
    from oracle import ref
    c, o, v, nulls = shard
    col = ref.RefStrings.from_arrays(c, o, v, nulls)
    _, cnt = col.contains_re(PATTERN)
    return cnt


def run_reference(args, rank, world):
    if rank != 0:
        return
    from custrings_b200.workloads import c2_corpus
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so not built"}))
        return
    cores = os.cpu_count() or 1
    sample_rows = min(args.rows, max(200_000, 150_000 * cores))
    chars, offsets, validity, nulls = c2_corpus(sample_rows, int(sample_rows * (args.bytes / args.rows)), seed=20240917)
    for _ in range(args.warmup):
        reference_rate(chars, offsets, validity, min(sample_rows, 20_000 * cores), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_rate(chars, offsets, validity, sample_rows, cores)
    dt = time.perf_counter() - t0
    value = sample_rows * args.steps / dt
    sample = "first %d rows (%d chars) of the C2 generator per step, %d processes" % (sample_rows, offsets[-1], cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "strings/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2 contains_re(\\b\\w{4,}\\b), reference CPU build, bounded sample", "pattern": PATTERN,
                   "rows_per_step": sample_rows},
        "cpu_baseline": {"value": value, "unit": "strings/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "strings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="custr")
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-rows", type=int, default=2_000_000)
    ap.add_argument("--tier", type=int, default=0, help="0 auto, 1 force the exact Pike VM")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from custrings_b200 import nvstrings
    from custrings_b200._lib import lib
    from custrings_b200.workloads import c2_corpus

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (custrings_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    L = lib()
    L.custr_set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the communicator comes up: stdout carries exactly ONE JSON line,
        # so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    L.custr_set_stream(torch.cuda.current_stream().cuda_stream)
    L.custr_set_regex_tier(args.tier)

    # ---- synthetic shard of this rank (different seed per rank: independent rows, same distribution)
    t_gen = time.perf_counter()
    chars, offsets, validity, nulls = c2_corpus(args.rows, args.bytes, seed=20240917 + rank)
    t_gen = time.perf_counter() - t_gen
    n = args.rows
    h_chars = torch.from_numpy(chars).pin_memory()
    h_off = torch.from_numpy(offsets).pin_memory()
    h_val = torch.from_numpy(validity).pin_memory()
    h_res = torch.empty(n, dtype=torch.uint8).pin_memory()

    col = nvstrings.from_offsets(h_chars, h_off, n, h_val, nulls, bdevmem=False)
    d_res = torch.empty(n, dtype=torch.uint8, device="cuda")

    def step():
        return L.custr_contains_re(col.m_cptr, PATTERN.encode(), d_res.data_ptr(), 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up on, through the device-resident timed region and the
    # end-to-end region (a single nvidia-smi query takes longer than the 20-step timed region itself)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        matches = step()
    tier = L.custr_last_regex_tier().decode()

    # ---- device-resident timed region
    launches0 = L.custr_launch_count()
    L.custr_set_profiling(1)
    kernel_ms = []
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
        kernel_ms.append(float(L.custr_last_kernel_ms()))
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    L.custr_set_profiling(0)
    launches = L.custr_launch_count() - launches0

    # ---- end-to-end through the public API from host buffers
    def e2e_step():
        c = nvstrings.from_offsets(h_chars, h_off, n, h_val, nulls, bdevmem=False)
        c.contains(PATTERN, devptr=d_res.data_ptr())
        h_res.copy_(d_res, non_blocking=True)
        torch.cuda.synchronize()
        del c

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_matches = int(h_res.sum().item())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    times = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = times.tolist()

    if rank == 0:
        value = world * n * args.steps / (elapsed_ms / 1e3)
        e2e_value = world * n * args.e2e_steps / (e2e_ms / 1e3)
        peak, peak_src = hbm_peak()
        k_ms = float(np.mean(kernel_ms))
        alg = algorithmic_bytes(n, int(offsets[-1]))
        achieved = alg / (k_ms / 1e3) / 1e9
        traffic = None
        try:  # dram read+write of the dominant kernel from the committed ncu --set full capture (same workload)
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tj = json.load(f)
            if tier == "bitstream" and args.rows == 10_000_000 and args.bytes == 1 << 30:
                traffic = tj["traffic"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "strings/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "C2 contains_re(\\b\\w{4,}\\b) over %d strings / %d chars per GPU" % (n, int(offsets[-1])),
                       "pattern": PATTERN, "rows_per_gpu": n, "chars_per_gpu": int(offsets[-1]), "sharding": "contiguous row ranges, no collective",
                       "l2": "inputs (1 GiB) larger than the 126 MB L2, no flush needed", "regex_tier": tier,
                       "matches_rank0": int(matches), "datagen_s": round(t_gen, 1)},
            "e2e": {"value": e2e_value, "unit": "strings/s", "h2d_bytes_per_step": int(chars.nbytes + offsets.nbytes + validity.nbytes),
                    "d2h_bytes_per_step": int(n), "steps": args.e2e_steps, "matches_rank0": e2e_matches},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel_ms": k_ms, "algorithmic_bytes": alg, "peak_source": peak_src,
                         "kernel": "bits::k_chain64<4,1,3> (+ memset, + k_vm_bool_rows for rows holding NUL)" if tier == "bitstream" else "k_vm_bool<32>"},
            "clocks": sampler.summary(),
        }
        if world == 1:
            try:
                from oracle import ref
                if ref.available():
                    rows = min(args.cpu_sample_rows, n)
                    rate = reference_rate(chars, offsets, validity, rows, 1)
                    line["cpu_baseline"] = {"value": rate, "unit": "strings/s", "cores": 1, "kind": "reference",
                                            "sample": "first %d rows of the same column, reference CPU build (oracle/_ref), 1 thread" % rows}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "strings/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
            except Exception as e:  # the baseline must never break the bench line
                line["cpu_baseline"] = {"value": None, "unit": "strings/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
