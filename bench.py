#!/usr/bin/env python
"""bench.py -- contract benchmark of the brute-force 1-NN path (BASELINE.json metric: query.ref
pairs/s and ms/call at 1/2/4/8 B200 vs roofline, v0 CPU baseline beside it).

    python bench.py --gpus N --steps K --warmup W              # our arm (CUDA, libnn_b200.so)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU v0 on the host cores

Workload (config.workload): BASELINE configs[3] = k=16, m=65,536 queries, n=16,777,216 references --
the configuration `north_star` quotes the 1/2/4/8-GPU metric on; it fits one GPU (1 GiB of
references).  With N GPUs the ONE reference set is sharded over the ranks (strong scaling, v8's job:
/root/reference/sources/src/core.cu:875-883) and the per-query packed keys are merged with one
u64-min exchange (NVLink peer atomics or NCCL all-reduce, see --merge).  --workload cfg1..cfg5 /
--scaling weak select other runs.

One "step" = one complete search with inputs resident in HBM:
  N = 1: ONE kernel launch (nn_b200_search_device: distance + argmin + merge + index store);
  N > 1: one launch per rank with the exchange fused into the search kernels over NVLink peer memory
         (nn_b200_peer_search) where the merge is latency-bound, else search (packed keys out) ->
         all-reduce(min) -> keys_unpack (--merge auto|peer|nccl; the other one is timed beside it).

`value`   : pairs/s (m * n_total / step time), CUDA events per step on the launching stream, summed over
            the K steps, max over ranks; L2 flushed between steps.
`e2e`     : the same metric through the reference-facing entry point `cudaCallback` with malloc'ed
            (pageable) HOST arrays, exactly what the reference's harness passes and times
            (main.cu:69-73, generator.h:37/44): H2D + search + merge + D2H + malloc inside the timed
            region, rank 0 driving all N GPUs in one process as v8 does; wall clock per call, the MEDIAN
            of the timed calls (`e2e.ms_spread` holds min / median / mean / max).  `e2e.pinned` is the same call
            (error-code variant) with pinned buffers, `e2e.resident_index` the build-once/query-many API.
`roofline`: the fused search kernel against the SLOWER of the two bounds north_star names -- 3k FP32
            lane-ops per pair at SMs*128*max clock and n*k*4 reference bytes at the measured HBM copy
            bandwidth; kernel time from CUDA events around the launch; `traffic` from the committed ncu
            capture (profiles/traffic.json).
`parity_spot_check`: 16 queries of the measured result re-computed by the CPU oracle (v0 restatement)
            against the full reference set -- at every N, for the device-resident result (NCCL merge)
            and for cudaCallback's (in-kernel NVLink merge), plus the in-process NCCL merge at N > 1.
`all_configs` (N = 1): a short device-resident + e2e + oracle-checked leg for each of the other BASELINE
            configs, so that every config has a driver-visible roofline fraction.
`cpu_baseline`: the reference's own v0 (oracle/_ref, built with its -Ofast flags; the oracle port where
            /root/reference was never available) on a bounded query sample, all host threads
            (CPU affinity count, not OMP_NUM_THREADS) -- and the 1-thread figure, which is what v0 is.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (k, m, n)
    "cfg1": (3, 1024, 65536),
    "cfg2": (16, 4096, 1 << 20),
    "cfg3": (8, 8, 1 << 26),
    "cfg4": (16, 65536, 1 << 24),
    "cfg5": (3, 1 << 20, 1 << 20),
}
DESCR = {
    "cfg1": "BASELINE configs[0]: k=3, m=1,024 queries, n=65,536 refs (TA sample 6 shape)",
    "cfg2": "BASELINE configs[1]: k=16, m=4,096 queries, n=1,048,576 refs",
    "cfg3": "BASELINE configs[2]: k=8, m=8 queries, n=67,108,864 refs",
    "cfg4": "BASELINE configs[3]: k=16, m=65,536 queries, n=16,777,216 refs, reference shards over the GPUs",
    "cfg5": "BASELINE configs[4]: k=3, m=1,048,576 queries, n=1,048,576 refs",
}
METRIC = "query*ref pairs/s (brute-force 1-NN, bit-exact vs v0)"
UNIT = "pairs/s"


def workload_string(name: str, scaling: str, world: int) -> str:
    """Identical for both arms and -- under strong scaling -- for every N."""
    if scaling == "weak" and world > 1:
        return DESCR[name] + f" PER GPU (weak scaling, x{world})"
    return DESCR[name]


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

    def reset(self):
        """Forget what was sampled so far (the warm-up): only the timed region counts."""
        self.samples, self.reasons = [], set()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_cores() -> int:
    """Host threads this process may use: its CPU affinity, NOT OMP_NUM_THREADS (torchrun sets that to 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------
# CPU arm (the only place, besides the parity spot checks, where oracle/ is executed)
# ---------------------------------------------------------------------------------------------------
def cpu_arm(k, m, n, S, R, budget_s=12.0, one_thread=True):
    """Times the reference's CPU implementation (v0) on a bounded sample of the workload's queries
    against the FULL reference set on all host threads (and once on one thread).  Returns the
    cpu_baseline dict."""
    from oracle import oracle
    cores = host_cores()
    use_ref = oracle.ref_available(fast=True)
    kind = "reference" if use_ref else "port"

    def run(q, threads):
        t0 = time.perf_counter()
        if use_ref:
            _, used = oracle.ref_v0(S[:q], R, k, threads=threads, fast=True)
        else:
            oracle.v0(S[:q], R, k, threads=threads)
            used = min(threads, q)
        return time.perf_counter() - t0, used

    probe_q = max(1, min(m, cores))
    t_probe, used = run(probe_q, cores)  # also warms the pages
    rate = probe_q * n / max(t_probe, 1e-9)
    q = int(max(probe_q, min(m, (budget_s * rate / n) // max(1, cores) * max(1, cores))))
    q = max(1, min(q, m))
    t, used = run(q, cores)
    out = {"value": q * n / t, "unit": UNIT, "cores": used, "kind": kind,
           "sample": f"first {q} of {m} queries x all {n} references, k={k}, {t:.2f} s; "
                     f"{'oracle/_ref = reference v0 (core.cu:25-63) built -Ofast' if use_ref else 'oracle/nn_oracle.c port, -O2 -ffp-contract=off'}"
                     f", queries split over {used} host threads (CPU affinity count; OMP_NUM_THREADS ignored)",
           "seconds": t}
    if one_thread:
        q1 = max(1, min(m, int(q / max(1, used) / 4) or 1))
        t1, _ = run(q1, 1)
        out["value_1_thread"] = q1 * n / t1
        out["sample_1_thread"] = f"first {q1} queries, 1 thread (v0 as written: serial, core.cu:27-62), {t1:.2f} s"
    return out


def reference_gpu_leg():
    """The reference's OWN CUDA path (core.cu recompiled for sm_100a, oracle/_ref/libref_gpu.so) timed on
    this B200 on TA samples 6 and 7 -- the shapes on which it is valid (v8 -> v7, core.cu:871-872);
    sample 6 is BASELINE config 1.  Runs in a subprocess (static initialisers, thrust aborts)."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "ref_gpu_time.py")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so")):
        return {"unavailable": "oracle/_ref/libref_gpu.so not built (needs /root/reference at build time)"}
    try:
        env = dict(os.environ)
        env.pop("NN_B200_GPUS", None)
        r = subprocess.run([sys.executable, exe, "6", "7"], capture_output=True, text=True, timeout=180, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": f"rc={r.returncode}: {(r.stderr or r.stdout)[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:  # a baseline beside the number, never fatal
        return {"unavailable": str(e)[:300]}


def make_inputs(k, m, n, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    S = torch.rand((m, k), generator=g, device=device, dtype=torch.float32)
    g.manual_seed(seed + 7919)
    R = torch.rand((n, k), generator=g, device=device, dtype=torch.float32)
    return S, R


def run_reference(args):
    """--impl reference: the reference's own CPU path (v0) on this box's host cores, same config."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    k, m, n = WORKLOADS[args.workload]
    n_total = n * world if args.scaling == "weak" else n
    rng = np.random.default_rng(1000)
    S = rng.random((m, k), dtype=np.float32)
    R = rng.random((n_total, k), dtype=np.float32)
    per_step_budget = max(1.0, min(8.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        res = cpu_arm(k, m, n_total, S, R, budget_s=per_step_budget, one_thread=(i == args.warmup + args.steps - 1))
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
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, args.scaling, world), "k": k, "m": m, "n_total": n_total,
                   "note": "CPU arm: every step is a bounded sample of the workload's queries against the full "
                           "reference set; pairs/s = sample pairs / wall time; same thread count at every N"},
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


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class L2Flush:
    """256 MiB written, then another 256 MiB read, so that the 126 MB L2 holds neither the previous
    step's references nor dirty lines whose write-back would ride on the timed kernel's HBM stream."""

    def __init__(self, dev):
        import torch
        self.w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.r = torch.zeros(64 << 20, dtype=torch.float32, device=dev)

    def __call__(self):
        self.w.zero_()
        self.r.sum()


def roofline_of(k, m, n_local, kernel_ms, peaks, sms, plan, workload):
    """north_star: the slower of 3k non-fused FP32 lane-ops per pair at the FP32 issue peak and
    n*k*4 reference bytes at HBM bandwidth."""
    kern_s = kernel_ms * 1e-3
    ops = 3.0 * k * m * n_local
    ref_bytes = float(n_local) * k * 4
    fp32_peak = sms * 128 * peaks["sm_max_mhz"] * 1e6
    hbm_peak = peaks["hbm_gbs"] * 1e9
    t_fp32, t_hbm = ops / fp32_peak, ref_bytes / hbm_peak
    kernel_name = {"qreg": "nn_qreg_kernel", "rreg": "nn_rreg_kernel", "rtma": "nn_rtma_kernel",
                   "qflex": "nn_qflex_kernel"}.get(plan.split()[0], plan.split()[0])
    fp32_part = {"achieved_tlops": ops / kern_s / 1e12, "peak_tlops": fp32_peak / 1e12, "frac": t_fp32 / kern_s,
                 "bound_ms": t_fp32 * 1e3,
                 "peak_source": f"{sms} SMs x 128 lanes x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} sm_max_mhz)"}
    hbm_part = {"achieved_gbs": ref_bytes / kern_s / 1e9, "peak_gbs": peaks["hbm_gbs"], "frac": t_hbm / kern_s,
                "bound_ms": t_hbm * 1e3, "peak_source": f"{peaks['source']} hbm_gbs (measured copy bandwidth)"}
    if t_fp32 >= t_hbm:
        roof = {"bound": "fp32", "kernel": kernel_name, "achieved": fp32_part["achieved_tlops"],
                "peak": fp32_part["peak_tlops"], "unit": "TFLOP/s (non-fused FP32 lane-ops: 3k per pair)",
                "frac": fp32_part["frac"]}
    else:
        roof = {"bound": "hbm", "kernel": kernel_name, "achieved": hbm_part["achieved_gbs"],
                "peak": hbm_part["peak_gbs"], "unit": "GB/s", "frac": hbm_part["frac"]}
    roof.update({"kernel_ms": kernel_ms, "fp32": fp32_part, "hbm": hbm_part,
                 "algorithmic": {"lane_ops_per_launch": ops, "bytes_per_launch": ref_bytes},
                 "traffic": None})
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(workload)
        if t and t.get("kernel") == kernel_name:
            roof["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
            roof["traffic_source"] = t["source"]
    return roof


def device_leg(nn, k, m, n_total, steps, warmup, dev, world=1, rank=0, scaling="strong", seed=1000, sampler=None,
               merge="peer", peer=None):
    """K timed steps of the device-resident search on this rank's shard.  Returns a dict of raw timings
    plus the tensors (queries, shard, result) for the later legs.
    merge (N > 1): "peer" = the exchange fused into the search kernels over NVLink peer memory (one launch
    per rank; result on rank 0, as v8 leaves it on its host), "nccl" = search -> all-reduce(min) -> unpack."""
    import torch
    import torch.distributed as dist
    from multicore_hw2_b200 import device, sharded
    shard = sharded.ShardedSearch(n_total, rank, world)
    begin, n_local = shard.begin, shard.count
    S, _ = make_inputs(k, m, 4, seed, dev)
    _, R = make_inputs(k, 4, max(n_local, 1), seed + 1000 + rank, dev)
    R = R[:n_local]
    ws = device.Workspace(m, dev)
    out = torch.empty(m, dtype=torch.int32, device=dev)
    keys = torch.empty(m, dtype=torch.int64, device=dev) if world > 1 else None
    flush = L2Flush(dev)

    def step(e0=None, e1=None, e2=None, e3=None):
        if e0 is not None:
            e0.record()
        if world == 1:
            device.search(S, R, ws, out=out, index_base=begin)          # ONE launch
        elif merge == "peer":
            peer.search(S, R, begin, out)                               # ONE launch per rank; rank 0 gets `out`
            if e1 is not None:
                e1.record()
                e2.record()
        else:
            device.search(S, R, ws, keys_out=keys, index_base=begin)    # search -> final keys of this shard
            if e1 is not None:
                e1.record()
            sharded.merge_keys(keys)                                     # all-reduce(min) over NCCL
            if e2 is not None:
                e2.record()
            device.keys_unpack(keys, out)
        if e3 is not None:
            e3.record()

    if sampler is not None:
        sampler.start()  # (before the warm-up: starting NVML takes ~0.1 s of host time on this rank)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.reset()
    gc.disable()
    # one more untimed step AFTER the synchronize, so that the host is enqueueing ahead of the device
    # when the first timed step starts (right after a host sync every launch latency of the first
    # step would be exposed)
    flush()
    step()
    launches0 = nn.launch_count()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        flush()  # outside the per-step event window
        step(*ev[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    gc.enable()
    launches = nn.launch_count() - launches0
    step_ms = [e[0].elapsed_time(e[3]) for e in ev]
    if world == 1 or merge == "peer":
        kern_ms, merge_ms = list(step_ms), [0.0] * steps
    else:
        kern_ms = [e[0].elapsed_time(e[1]) for e in ev]
        merge_ms = [e[1].elapsed_time(e[2]) for e in ev]  # incl. waiting for the slowest rank
    return {"S": S, "R": R, "out": out, "begin": begin, "n_local": n_local, "step_ms": step_ms, "kern_ms": kern_ms,
            "merge_ms": merge_ms, "launches": launches, "wall_s": t_wall, "merge": merge if world > 1 else None,
            "plan": nn.describe_plan(k, m, max(n_local, 1))}


def e2e_leg(nn, k, m, n_total, Sh, Rh, calls, world, full=True):
    """End to end through the host entry points, rank 0 driving `world` GPUs in one process.
    Sh, Rh: pageable numpy arrays (malloc'ed memory, as the reference's harness passes)."""
    import numpy as np
    import torch
    pairs = float(m) * float(n_total)
    os.environ["NN_B200_GPUS"] = str(world)   # cudaCallback has no GPU-count argument (core.h:71)

    spread = {}

    def timed(fn, reps, name=None):
        """Median wall clock of `reps` calls after two untimed ones (a call of 0.1 ms is at the mercy of a
        single host hiccup if averaged; for the long calls median and mean coincide).  The spread of the
        headline leg is kept beside it."""
        fn()
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        if name:
            spread[name] = {"min": ts[0] * 1e3, "median": ts[len(ts) // 2] * 1e3, "mean": sum(ts) / len(ts) * 1e3,
                            "max": ts[-1] * 1e3}
        return ts[len(ts) // 2], r

    t_cb, res = timed(lambda: nn.cudaCallback(k, m, n_total, Sh, Rh), calls, "cudaCallback")
    bytes_in, bytes_out = int((m * k + n_total * k) * 4), int(m * 4)
    e2e = {"value": pairs / t_cb, "unit": UNIT, "ms_per_call": t_cb * 1e3, "calls_timed": calls,
           "statistic": "median of the timed calls", "ms_spread": spread["cudaCallback"],
           "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
           "api": "cudaCallback(k, m, n, searchPoints, referencePoints, &results) -- the reference's entry point "
                  "(core.h:71), malloc'ed pageable host arrays as main.cu:69-73 times it; wall clock around the "
                  f"call incl. the malloc of the result; {world} GPU(s) driven from one process"}
    results = {"cudaCallback": res}
    if full:
        Sp, Rp = torch.from_numpy(Sh).pin_memory(), torch.from_numpy(Rh).pin_memory()
        buf = np.empty(m, dtype=np.int32)
        t_pin, _ = timed(lambda: nn.search_host(Sp, Rp, k, num_gpus=world, out=buf), calls)
        results["search_host_pinned"] = buf.copy()
        e2e["pinned"] = {"value": pairs / t_pin, "unit": UNIT, "ms_per_call": t_pin * 1e3,
                         "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
                         "api": "nn_b200_search_host (same body, error code instead of exit), PINNED host buffers"}
        if world > 1:
            nn.set_option("p2p_merge", 0)
            try:
                buf2 = np.empty(m, dtype=np.int32)
                t_nccl, _ = timed(lambda: nn.search_host(Sp, Rp, k, num_gpus=world, out=buf2), min(calls, 3))
                results["search_host_nccl"] = buf2
                e2e["pinned_nccl_merge"] = {"ms_per_call": t_nccl * 1e3,
                                            "api": "same with option p2p_merge=0: in-process ncclAllReduce(min, u64)"}
            finally:
                nn.set_option("p2p_merge", 1)
        with nn.Index(Rp, k, num_gpus=world) as ix:
            buf3 = np.empty(m, dtype=np.int32)
            t_ix, _ = timed(lambda: ix.search(Sp, out=buf3), calls)
        results["index"] = buf3
        e2e["resident_index"] = {"value": pairs / t_ix, "unit": UNIT, "ms_per_call": t_ix * 1e3,
                                 "h2d_bytes_per_step": int(m * k * 4), "d2h_bytes_per_step": bytes_out,
                                 "api": "nn_b200_index_search (references resident in HBM, build once / query many)"}
        del Sp, Rp
    return e2e, results


def spot_check(k, m, Sh, Rh, results: dict, nq=16):
    """`nq` evenly spaced queries re-computed by the CPU oracle against the FULL reference set and
    compared with every result array given.  Returns (dict name -> bool, rows)."""
    import numpy as np
    from oracle import oracle
    rows = np.unique(np.linspace(0, m - 1, num=min(nq, m)).astype(np.int64))
    want = oracle.v0(np.ascontiguousarray(Sh[rows]), Rh, k, threads=host_cores())
    return {name: bool(np.array_equal(np.asarray(r)[rows], want)) for name, r in results.items()}, len(rows)


def gather_references(R, world, rank, host_group):
    """All shards on rank 0's host as ONE pageable numpy array (None elsewhere)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if world == 1:
        return R.cpu().numpy()
    counts = [torch.zeros(1, dtype=torch.int64, device=R.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([R.shape[0]], dtype=torch.int64, device=R.device))
    counts = [int(c.item()) for c in counts]
    cap = max(counts)
    pad = torch.zeros((cap, R.shape[1]), dtype=R.dtype, device=R.device)
    pad[:R.shape[0]] = R
    gathered = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, gathered, dst=0)
    torch.cuda.synchronize()
    out = None
    if rank == 0:
        out = np.concatenate([g[:c].cpu().numpy() for g, c in zip(gathered, counts)])
        del gathered
    del pad
    torch.cuda.empty_cache()
    dist.barrier(group=host_group)  # the other ranks now wait on the CPU, their GPUs are idle
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-all-configs", action="store_true")
    ap.add_argument("--merge", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: peer = the exchange fused into the search kernels over NVLink peer memory (CUDA IPC, "
                         "system-scope atomicMin into rank 0's keys); nccl = search -> all-reduce(min) -> unpack; "
                         "auto = peer for latency-bound merges (m <= 16384 and >= 40 us of search per rank), else nccl")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (contract default): the workload's ONE reference set is sharded over the ranks "
                         "(v8, core.cu:875-883); weak: every rank owns the workload's n references")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    _quiet_stdout()

    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import multicore_hw2_b200 as nn

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
    peaks = measured_peaks()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count

    k, m, n_wl = WORKLOADS[args.workload]
    n_total = n_wl * world if args.scaling == "weak" else n_wl

    # ---- main leg: device-resident, K timed steps ------------------------------------------------
    # clocks are sampled on rank 0 only: NVML queries take driver locks, and one poller per rank
    # delayed launches enough to show up as all-reduce skew
    sampler = ClockSampler(local if rank == 0 else -1)
    merge, peer, peer_note = args.merge, None, None
    if merge == "auto":
        # Measured on 8 B200s (profiles/README.md): the peer merge pushes one remote atomic per query and rank and
        # rank 0's last CTA unpacks all m keys, so from ~16K queries on the bandwidth-optimal all-reduce wins
        # (config 5, m = 2^20: 35.7 vs 33.6 ms; config 4, m = 65536: 185.17 vs 184.94 ms); and when a rank's search
        # is only a few microseconds (config 1 on 8 GPUs: 12 us) the ranks that are not rank 0 run ahead and spend
        # the difference waiting inside their kernels (81 vs 55 us per step by the max-over-ranks clock).
        est_us = 3.0 * k * m * (n_total / world) / (0.85 * sms * 128 * peaks["sm_max_mhz"] * 1e6) * 1e6
        merge = "peer" if (m <= 16384 and est_us >= 40.0) else "nccl"
    if world > 1 and merge == "peer":
        try:
            from multicore_hw2_b200 import sharded
            peer = sharded.PeerMerge(m, group=host_group)
        except Exception as e:  # no CUDA IPC between the ranks (not the case on an NVSwitch box): say so, use NCCL
            peer_note = f"peer merge unavailable ({str(e)[:200]}): fell back to the NCCL all-reduce"
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            merge, peer = "nccl", None
    leg = device_leg(nn, k, m, n_total, args.steps, args.warmup, dev, world, rank, args.scaling, sampler=sampler,
                     merge=merge, peer=peer)
    clocks = sampler.stop()
    nccl_leg = None
    peer_leg = None
    if world > 1 and merge == "nccl" and args.merge == "auto":
        # the other merge beside it (short)
        try:
            from multicore_hw2_b200 import sharded
            pm2 = sharded.PeerMerge(m, group=host_group)
            lg2 = device_leg(nn, k, m, n_total, max(3, min(args.steps, 10)), 3, dev, world, rank, args.scaling, merge="peer",
                             peer=pm2)
            t2 = torch.tensor([sum(lg2["step_ms"]) / len(lg2["step_ms"])], dtype=torch.float64, device=dev)
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            same = torch.tensor([1 if (rank != 0 or torch.equal(lg2["out"], leg["out"])) else 0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            peer_leg = {"ms_per_step": float(t2.item()), "steps": len(lg2["step_ms"]),
                        "same_result_as_nccl_merge": bool(int(same.item())), "timed_out": pm2.error(),
                        "step": "one launch per rank: nn_b200_peer_search (merge inside the kernels over NVLink peer memory)"}
            del lg2
            pm2.close()
        except Exception as e:
            peer_leg = {"unavailable": str(e)[:200]}
    if world > 1 and merge == "peer":
        # the same steps with the exchange as a collective, for comparison (fewer steps)
        lg2 = device_leg(nn, k, m, n_total, max(3, min(args.steps, 10)), 3, dev, world, rank, args.scaling, merge="nccl")
        t2 = torch.tensor([sum(lg2["step_ms"]) / len(lg2["step_ms"])], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        same = torch.tensor([1 if (rank != 0 or torch.equal(lg2["out"], leg["out"])) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        nccl_leg = {"ms_per_step": float(t2.item()), "steps": len(lg2["step_ms"]),
                    "kernel_ms": sum(lg2["kern_ms"]) / len(lg2["kern_ms"]), "merge_ms": sum(lg2["merge_ms"]) / len(lg2["merge_ms"]),
                    "same_result_as_peer_merge": bool(int(same.item())),
                    "step": "search (final keys of the shard) -> ncclAllReduce(min, u64) -> keys_unpack"}
        del lg2
        if peer.error():
            peer_note = "a device-side wait of the peer merge timed out"
    step_ms, kern_ms, merge_ms = leg["step_ms"], leg["kern_ms"], leg["merge_ms"]
    t = torch.tensor([sum(step_ms), sum(kern_ms) / len(kern_ms)], dtype=torch.float64, device=dev)
    t_max, t_min = t.clone(), t.clone()
    per_rank_steps = None
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
        mine = torch.tensor(step_ms[:16], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank_steps = [[round(float(x), 4) for x in a.tolist()] for a in allr]
    total_ms = float(t_max[0].item())
    ms_per_step = total_ms / args.steps
    pairs_per_step = float(m) * float(n_total)
    value = pairs_per_step / (ms_per_step * 1e-3)
    kern_ms_avg = sum(kern_ms) / len(kern_ms)
    roof = roofline_of(k, m, leg["n_local"], kern_ms_avg, peaks, sms, leg["plan"], args.workload)
    if rank == 0:
        try:
            meas, meas2 = nn.probe_fp32(0), nn.probe_fp32(1)
            roof["peak_measured"] = max(meas, meas2) / 1e12
            roof["peak_measured_scalar"] = meas / 1e12
            roof["peak_measured_f32x2"] = meas2 / 1e12
            roof["frac_of_measured"] = (3.0 * k * m * leg["n_local"] / (kern_ms_avg * 1e-3)) / max(meas, meas2)
        except Exception as e:  # measurement aid only
            roof["peak_measured_error"] = str(e)

    # ---- e2e + oracle spot check (rank 0 holds the whole reference set on the host) -----------------
    e2e, parity, parity_n = None, None, 0
    need_host = not args.no_e2e
    if need_host:
        Rh = gather_references(leg["R"], world, rank, host_group)
        if rank == 0:
            Sh = leg["S"].cpu().numpy()
            calls = max(1, min(args.steps, 5 if pairs_per_step > 2e11 else 20))
            e2e, results = e2e_leg(nn, k, m, n_total, Sh, Rh, calls, world)
            results["device_resident" + (f"_{merge}_merge" if world > 1 else "")] = leg["out"].cpu().numpy()
            parity, parity_n = spot_check(k, m, Sh, Rh, results)
            e2e["matches_device_resident_result"] = bool(np.array_equal(results["cudaCallback"], leg["out"].cpu().numpy()))
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=host_group)

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Sh2 = leg["S"].cpu().numpy()
        Rh2 = Rh if need_host else leg["R"].cpu().numpy()
        cpu = cpu_arm(k, m, n_total, Sh2, Rh2, budget_s=12.0)
        cpu.pop("seconds", None)
    main_S, main_out = leg["S"], leg["out"]
    launches, plan, wall_s = leg["launches"], leg["plan"], leg["wall_s"]
    del leg
    if need_host and rank == 0:
        del Rh
    torch.cuda.empty_cache()

    # ---- every other BASELINE config, short (N = 1) -----------------------------------------------
    all_cfg = None
    if rank == 0 and world == 1 and not args.no_all_configs:
        all_cfg = {}
        for name in sorted(WORKLOADS):
            if name == args.workload:
                continue
            kk, mm, nn_ = WORKLOADS[name]
            st = max(3, min(args.steps, 10 if mm * nn_ < 5e11 else 5))
            lg = device_leg(nn, kk, mm, nn_, st, 3, dev, seed=3000)
            km = sum(lg["kern_ms"]) / len(lg["kern_ms"])
            rf = roofline_of(kk, mm, nn_, km, peaks, sms, lg["plan"], name)
            entry = {"workload": DESCR[name], "k": kk, "m": mm, "n": nn_, "steps": st, "ms_per_step": km,
                     "value": float(mm) * nn_ / (km * 1e-3), "unit": UNIT, "plan": lg["plan"],
                     "gpu_launches_per_step": lg["launches"] / st,
                     "roofline": {kx: rf[kx] for kx in ("bound", "kernel", "achieved", "peak", "unit", "frac", "kernel_ms",
                                                        "traffic")},
                     "step_ms_min_max": [min(lg["step_ms"]), max(lg["step_ms"])]}
            if not args.no_e2e:
                Sh, Rh = lg["S"].cpu().numpy(), lg["R"].cpu().numpy()
                ee, res = e2e_leg(nn, kk, mm, nn_, Sh, Rh, 3 if float(mm) * nn_ > 1e10 else 15, 1, full=False)
                res["device_resident"] = lg["out"].cpu().numpy()
                ok, nq = spot_check(kk, mm, Sh, Rh, res)
                entry["e2e"] = {kx: ee[kx] for kx in ("value", "unit", "ms_per_call", "calls_timed", "statistic",
                                                      "ms_spread", "h2d_bytes_per_step", "d2h_bytes_per_step")}
                entry["e2e"]["api"] = "cudaCallback, malloc'ed pageable host arrays"
                entry["parity_spot_check"] = all(ok.values())
                entry["parity_detail"] = {"oracle_queries": nq, **ok}
                del Sh, Rh
            all_cfg[name] = entry
            del lg
            torch.cuda.empty_cache()
        # the reference's own GPU kernels, recompiled for this GPU, on the shapes where they are valid
        torch.cuda.synchronize()
        all_cfg["reference_gpu_on_this_b200"] = reference_gpu_leg()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, args.scaling, world),
                       "k": k, "m": m, "n_total": n_total,
                       "parallelism": (f"{world} reference shards of one set ({args.scaling} scaling), one process per GPU; merge: "
                                       + ("inside the search kernels, system-scope u64 atomicMin into rank 0's keys over NVLink "
                                          "peer memory (CUDA IPC), result on rank 0" if merge == "peer"
                                          else "NCCL all-reduce(min) of uint64 keys") if world > 1 else "1 GPU"),
                       "merge": merge if world > 1 else None, "merge_note": peer_note,
                       "l2": "flushed between steps (256 MiB written then 256 MiB read, outside the event window)",
                       "timing": "CUDA events per step on the launching stream, summed over the K steps, max over ranks; "
                                 "one untimed priming step between the synchronize and the first timed step",
                       "step": "one launch: nn_b200_search_device (search + merge + index store)" if world == 1 else
                               ("one launch per rank: nn_b200_peer_search (search + cross-GPU merge + index store on rank 0)"
                                if merge == "peer" else
                                "search (one launch, final keys of the shard) -> all-reduce(min) -> keys_unpack"),
                       "plan": plan, "data": "uniform [0,1) float32, seeded"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "parity_spot_check": (all(parity.values()) if parity else None),
            "parity_detail": ({"oracle_queries": parity_n, "oracle": "oracle/nn_oracle.c v0 restatement, full reference set",
                               **parity} if parity else None),
            "all_configs": all_cfg, "nccl_merge": nccl_leg, "peer_merge": peer_leg,
            "clocks": clocks, "wall_ms_per_step_incl_flush": 1e3 * wall_s / args.steps,
            "step_ms_min_max": [min(step_ms), max(step_ms)], "step_ms": [round(x, 4) for x in step_ms[:32]],
            "merge_ms": sum(merge_ms) / len(merge_ms),
            "kernel_ms_min_max_over_ranks": [float(t_min[1].item()), float(t_max[1].item())],
            "step_ms_per_rank": per_rank_steps,
        }
        emit(line)
    if world > 1:
        dist.barrier(group=host_group)
        if peer is not None:
            peer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
