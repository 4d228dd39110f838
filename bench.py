#!/usr/bin/env python
"""bench.py -- contract benchmark of the brute-force 1-NN path (BASELINE.json metric: query.ref
pairs/s and ms/call).

    python bench.py --gpus N --steps K --warmup W              # our arm (CUDA, libnn_b200.so)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU v0 on the host cores

Workload (config.workload): BASELINE.json configs[1] = k=16, m=4096 queries, n=1,048,576
references per GPU.  One "step" = one complete search: keys_init -> fused distance+argmin ->
[all-reduce(min) of packed keys over NCCL when N > 1] -> keys_unpack.  With N GPUs every rank owns
one 2^20-reference shard of an N*2^20 reference set (weak scaling; the queries are replicated),
which is how the path shards (v8, /root/reference/sources/src/core.cu:875-883).

`value`  : pairs/s with inputs resident in HBM, timed with CUDA events per step on the launching
           stream (max over ranks), L2 flushed between steps, one untimed priming step after the sync.
`e2e`    : pairs/s through the reference-facing C-ABI call (nn_b200_search_host = the body of
           cudaCallback) with pinned HOST buffers: H2D of queries+references, search, merge, D2H of
           the indices, all inside the timed region (wall clock; the call is synchronous).
           `e2e.resident_index`: the same call against nn_b200_index_search (references already in HBM).
`roofline`: the dominant kernel against the SLOWER of the two bounds north_star names -- 3*k FP32
           lane-ops per pair at SMs*128*max clock (also measured live with non-fused FADD/FMUL) and
           n*k*4 reference bytes at the measured HBM copy bandwidth; `traffic` = DRAM bytes of that
           kernel from the committed ncu capture (profiles/traffic.json).
`cpu_baseline`: the reference's own v0 (oracle/_ref, built with its -Ofast flags) -- or the oracle
           port where /root/reference was never available -- on a bounded query sample.
--workload cfg1..cfg5 selects another BASELINE config; --scaling strong shards the workload's n
over the ranks instead of giving every rank n references.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (k, m, n per GPU)
    "cfg2": (16, 4096, 1 << 20),
    "cfg1": (3, 1024, 65536),
    "cfg3": (8, 8, 1 << 26),
    "cfg4": (16, 65536, 1 << 24),
    "cfg5": (3, 1 << 20, 1 << 20),
}
DESCR = {
    "cfg2": "BASELINE configs[1]: k=16, m=4096 queries, n=1,048,576 refs per GPU",
    "cfg1": "BASELINE configs[0]: k=3, m=1024, n=65536",
    "cfg3": "BASELINE configs[2]: k=8, m=8, n=67,108,864 per GPU",
    "cfg4": "BASELINE configs[3]: k=16, m=65536, n=16,777,216 per GPU",
    "cfg5": "BASELINE configs[4]: k=3, m=1,048,576, n=1,048,576 per GPU",
}
METRIC = "query*ref pairs/s (brute-force 1-NN, bit-exact vs v0)"
UNIT = "pairs/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs"), "sm_max_mhz": d.get("sm_max_mhz"), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons = index, [], set()
        self._stop = threading.Event()
        self._t = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return self
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for nme, bit in names.items():
                    if r & bit:
                        self.reasons.add(nme)
            except Exception:
                pass
            self._stop.wait(0.02)

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_arm(k, m, n, S, R, budget_s=12.0):
    """Times the reference's CPU implementation (v0) on a bounded sample of the workload's queries
    against the FULL reference set, all host threads.  Returns the cpu_baseline dict."""
    from oracle import oracle
    import numpy as np
    cores = host_cores()
    use_ref = oracle.ref_available(fast=True)
    kind = "reference" if use_ref else "port"

    def run(q):
        t0 = time.perf_counter()
        if use_ref:
            _, used = oracle.ref_v0(S[:q], R, k, threads=0, fast=True)
        else:
            oracle.v0(S[:q], R, k, threads=0)
            used = min(cores, q)
        return time.perf_counter() - t0, used

    probe_q = max(1, min(m, cores))
    t_probe, used = run(probe_q)  # also warms the pages
    rate = probe_q * n / max(t_probe, 1e-9)
    q = int(max(probe_q, min(m, (budget_s * rate / n) // max(1, cores) * max(1, cores))))
    q = max(1, min(q, m))
    t, used = run(q)
    return {"value": q * n / t, "unit": UNIT, "cores": used, "kind": kind,
            "sample": f"first {q} of {m} queries x all {n} references, k={k}, {t:.2f} s; "
                      f"{'oracle/_ref = reference v0 (core.cu:25-63) built -Ofast' if use_ref else 'oracle/nn_oracle.c port, -O2 -ffp-contract=off'}"
                      f", queries split over {used} host threads",
            "seconds": t}


def make_inputs(k, m, n, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    S = torch.rand((m, k), generator=g, device=device, dtype=torch.float32)
    g.manual_seed(seed + 7919)
    R = torch.rand((n, k), generator=g, device=device, dtype=torch.float32)
    return S, R


def run_reference(args):
    """--impl reference: the reference's own CPU path (v0) on this box's host cores."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k, m, n1 = WORKLOADS[args.workload]
    n = n1 * args.gpus
    rng = np.random.default_rng(1000)
    S = rng.random((m, k), dtype=np.float32)
    R = rng.random((n, k), dtype=np.float32)
    per_step_budget = max(1.0, min(8.0, 150.0 / max(1, args.steps + args.warmup)))
    res = None
    vals = []
    for i in range(args.warmup + args.steps):
        res = cpu_arm(k, m, n, S, R, budget_s=per_step_budget)
        if i >= args.warmup:
            vals.append(res)
    tot_pairs = sum(v["value"] * v["seconds"] for v in vals)
    tot_s = sum(v["seconds"] for v in vals)
    value = tot_pairs / tot_s
    cb = dict(vals[-1])
    cb["value"] = value
    cb.pop("seconds", None)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": DESCR[args.workload] + f" x {args.gpus} GPU shard(s)", "k": k, "m": m, "n": n,
                   "note": "CPU arm: every step is a bounded sample of the workload's queries against the full "
                           "reference set; pairs/s = sample pairs / wall time"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries print to stdout on their own (NCCL's version banner under torchrun).  The contract is ONE
    JSON line on stdout: everything else written to file descriptor 1 during the run is sent to
    stderr, and emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (contract default): every rank owns the workload's n references; strong: the "
                         "workload's n references are sharded over the ranks (BASELINE configs[3] at 1/2/4/8 GPUs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    _quiet_stdout()

    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import multicore_hw2_b200 as nn
    from multicore_hw2_b200 import device, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # host-side rendezvous for the phases in which only rank 0 works: an NCCL barrier would leave
        # a spinning kernel on every other GPU, and rank 0's e2e call drives those GPUs itself
        host_group = dist.new_group(backend="gloo")
    nn.lib()

    k, m, n_local = WORKLOADS[args.workload]
    n_total = n_local * world
    if args.scaling == "strong":
        n_total = n_local
        n_local = sharded.ShardedSearch(n_total, rank, world).count
    peaks = measured_peaks()

    # ---- inputs: queries replicated, this rank's reference shard; resident in HBM ----------------
    S, _ = make_inputs(k, m, 4, 1000, dev)
    _, R = make_inputs(k, 4, n_local, 2000 + rank, dev)
    shard = sharded.ShardedSearch(n_total, rank, world)
    # weak scaling keeps the per-rank shard at exactly the workload's n references
    begin = shard.begin
    assert shard.count == n_local and (args.scaling == "strong" or begin == rank * n_local)
    keys = device.new_keys(m, dev)
    out = torch.empty(m, dtype=torch.int32, device=dev)
    # L2 flush between steps: 256 MiB written, then another 256 MiB read, so that the 126 MB L2 holds
    # neither the previous step's references nor dirty lines whose write-back would ride on the
    # timed kernel's HBM stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)

    def l2_flush():
        flush.zero_()
        flush_rd.sum()

    def step():
        device.keys_init(keys)
        device.nearest_keys(S, R, keys, begin)
        sharded.merge_keys(keys)
        device.keys_unpack(keys, out)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    mev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # clocks are sampled on rank 0 only: NVML queries take driver locks, and with one poller per rank
    # they delayed launches enough to show up as all-reduce skew (0.20 ms -> 0.05 ms per step at N = 8)
    sampler = ClockSampler(local if rank == 0 else -1).start()
    import gc
    gc.disable()
    # one more untimed step AFTER the synchronize, so that the host is enqueueing ahead of the device
    # when the first timed step starts (right after a host sync every launch latency of the first
    # step would be exposed: it measured 2-3x the others on the sub-millisecond workloads)
    l2_flush()
    step()
    launches0 = nn.launch_count()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        l2_flush()  # outside the per-step event window
        ev[i][0].record()
        device.keys_init(keys)
        kev[i][0].record()
        device.nearest_keys(S, R, keys, begin)
        kev[i][1].record()
        sharded.merge_keys(keys)
        mev[i].record()
        device.keys_unpack(keys, out)
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    gc.enable()
    launches = nn.launch_count() - launches0
    clocks = sampler.stop()

    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    merge_ms = [kev[i][1].elapsed_time(mev[i]) for i in range(args.steps)]  # incl. waiting for the slowest rank
    kern_all = torch.tensor([sum(kern_ms) / len(kern_ms)], dtype=torch.float64, device=dev)
    kern_lo, kern_hi = kern_all.clone(), kern_all.clone()
    if world > 1:
        dist.all_reduce(kern_lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(kern_hi, op=dist.ReduceOp.MAX)
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    pairs_per_step = float(m) * float(n_total)
    value = pairs_per_step / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (this rank's fused distance+argmin launch) --------------
    # north_star: the slower of 3k non-fused FP32 lane-ops per pair at the FP32 issue peak and
    # n*k*4 reference bytes at HBM bandwidth.
    kern_ms_avg = sum(kern_ms) / len(kern_ms)
    kern_s = kern_ms_avg * 1e-3
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    ops = 3.0 * k * m * n_local
    ref_bytes = float(n_local) * k * 4
    fp32_peak = sms * 128 * peaks["sm_max_mhz"] * 1e6
    hbm_peak = peaks["hbm_gbs"] * 1e9
    t_fp32, t_hbm = ops / fp32_peak, ref_bytes / hbm_peak
    plan = nn.describe_plan(k, m, n_local)
    kernel_name = {"qreg": "nn_qreg_kernel", "rreg": "nn_rreg_kernel", "rtma": "nn_rtma_kernel"}.get(plan.split()[0], plan.split()[0])
    fp32_part = {"achieved_tlops": ops / kern_s / 1e12, "peak_tlops": fp32_peak / 1e12, "frac": t_fp32 / kern_s,
                 "bound_ms": t_fp32 * 1e3,
                 "peak_source": f"{sms} SMs x 128 lanes x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} sm_max_mhz)"}
    hbm_part = {"achieved_gbs": ref_bytes / kern_s / 1e9, "peak_gbs": peaks["hbm_gbs"], "frac": t_hbm / kern_s,
                "bound_ms": t_hbm * 1e3, "peak_source": f"{peaks['source']} hbm_gbs (measured copy bandwidth)"}
    if t_fp32 >= t_hbm:
        line_roof = {"bound": "fp32", "kernel": kernel_name, "achieved": fp32_part["achieved_tlops"],
                     "peak": fp32_part["peak_tlops"], "unit": "TFLOP/s (non-fused FP32 lane-ops: 3k per pair)",
                     "frac": fp32_part["frac"]}
    else:
        line_roof = {"bound": "hbm", "kernel": kernel_name, "achieved": hbm_part["achieved_gbs"],
                     "peak": hbm_part["peak_gbs"], "unit": "GB/s", "frac": hbm_part["frac"]}
    line_roof.update({"kernel_ms": kern_ms_avg, "fp32": fp32_part, "hbm": hbm_part,
                      "algorithmic": {"lane_ops_per_launch": ops, "bytes_per_launch": ref_bytes},
                      "traffic": None})
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(args.workload)
        if t and t.get("kernel") == kernel_name:
            line_roof["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
            line_roof["traffic_source"] = t["source"]
    if rank == 0:
        try:
            meas = nn.probe_fp32(0)
            meas2 = nn.probe_fp32(1)
            line_roof["peak_measured"] = max(meas, meas2) / 1e12
            line_roof["peak_measured_scalar"] = meas / 1e12
            line_roof["peak_measured_f32x2"] = meas2 / 1e12
            line_roof["frac_of_measured"] = (ops / kern_s) / max(meas, meas2)
        except Exception as e:  # measurement aid only
            line_roof["peak_measured_error"] = str(e)

    # ---- e2e: the C-ABI host entry with pinned host buffers, rank 0 drives all N GPUs -------------
    e2e = None
    if not args.no_e2e:
        Rh_parts = [R.cpu()]
        if world > 1:
            gathered = [torch.empty_like(R) for _ in range(world)] if rank == 0 else None
            dist.gather(R, gathered, dst=0)
            if rank == 0:
                Rh_parts = [g.cpu() for g in gathered]
                del gathered
            torch.cuda.synchronize()
            dist.barrier(group=host_group)  # the other ranks now wait on the CPU, their GPUs are idle
        if rank == 0:
            Sh = S.cpu().pin_memory()
            Rh = torch.cat(Rh_parts).pin_memory()
            res = np.empty(m, dtype=np.int32)
            for _ in range(2):
                nn.search_host(Sh, Rh, k, num_gpus=world, out=res)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                nn.search_host(Sh, Rh, k, num_gpus=world, out=res)
            t_e2e = (time.perf_counter() - t0) / args.steps
            torch.cuda.synchronize()
            same = bool(np.array_equal(res, out.cpu().numpy()))
            e2e = {"value": pairs_per_step / t_e2e, "unit": UNIT, "ms_per_call": t_e2e * 1e3,
                   "h2d_bytes_per_step": int((m * k + n_total * k) * 4), "d2h_bytes_per_step": int(m * 4),
                   "api": "nn_b200_search_host (body of cudaCallback), pinned host buffers, "
                          f"{world} GPU(s) driven from one process",
                   "matches_device_resident_result": same}
            # the same call against a RESIDENT reference index (build once, query many): only the
            # queries and the indices cross PCIe
            with nn.Index(Rh, k, num_gpus=world) as ix:
                res2 = np.empty(m, dtype=np.int32)
                for _ in range(2):
                    ix.search(Sh, out=res2)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    ix.search(Sh, out=res2)
                t_ix = (time.perf_counter() - t0) / args.steps
            e2e["resident_index"] = {"value": pairs_per_step / t_ix, "unit": UNIT, "ms_per_call": t_ix * 1e3,
                                     "h2d_bytes_per_step": int(m * k * 4), "d2h_bytes_per_step": int(m * 4),
                                     "api": "nn_b200_index_search (references resident in HBM)",
                                     "matches": bool(np.array_equal(res2, res))}
            del Rh, Sh
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=host_group)

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Sh, Rh = S.cpu().numpy(), R.cpu().numpy()
        cpu = cpu_arm(k, m, n_local, Sh, Rh, budget_s=12.0)
        cpu.pop("seconds", None)
        # the same sample doubles as a parity spot-check of the measured result
        from oracle import oracle
        rows = np.arange(0, m, max(1, m // 16))[:16]
        want = oracle.v0(Sh[rows], Rh, k, threads=0)
        cpu["parity_spot_check"] = bool(np.array_equal(out.cpu().numpy()[rows], want))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": DESCR[args.workload] + (f"; {world} shards, n_total={n_total}" if world > 1 else ""),
                       "k": k, "m": m, "n_per_gpu": n_local, "n_total": n_total,
                       "parallelism": f"reference shards x{world}, NCCL all-reduce(min) of uint64 keys" if world > 1 else "1 GPU",
                       "l2": "flushed between steps (256 MiB written then 256 MiB read, outside the event window)",
                       "timing": "CUDA events per step on the launching stream, summed over the K steps, max over ranks; "
                                 "one untimed priming step between the synchronize and the first timed step",
                       "plan": plan, "uniform [0,1) float32": True},
            "roofline": line_roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "step_ms_min_max": [min(step_ms), max(step_ms)], "step_ms": [round(x, 4) for x in step_ms[:32]],
            "merge_ms": sum(merge_ms) / len(merge_ms),
            "kernel_ms_min_max_over_ranks": [float(kern_lo.item()), float(kern_hi.item())],
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
