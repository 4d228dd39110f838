#!/usr/bin/env python
"""Turns the ncu captures of scripts/gpu_profile.sh (gpurun_out/<tag>_*.ncu-rep) into the committed
evidence under profiles/: <tag>_ncu_summary.txt (side-by-side metrics + stall reasons),
<tag>_launches_bench_cfg2.csv (copied) and traffic.json (dram bytes per launch, read by bench.py).

    python scripts/make_profiles.py r01
"""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
caps = ["cfg1_qreg", "cfg2_qreg", "cfg3_rtma", "cfg4_qreg", "cfg4q_qreg_super", "cfg5_qreg", "cfg4s_qreg", "cfg5s_qreg", "m100k8_qflex", "m100k3_qflex",
        "cfg3shard_rtma", "m1_rreg", "repack_k16"]
paths = [os.path.join(ROOT, "gpurun_out", f"{tag}_{c}.raw.csv") for c in caps]
paths = [p for p in paths if os.path.exists(p)]
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py")] + paths, capture_output=True, text=True).stdout
hdr = (f"# ncu --set full --clock-control none captures ({tag}), one launch each, nn_bench command lines in scripts/gpu_profile.sh\n"
       "# all five BASELINE configs at FULL size, through the one-launch search (nn_b200_search_device); m100* = 100 queries x 2^22 references (phased kernel)\n"
       "# cfg4q_qreg_super = config 4 at a quarter of its references (k=16, m=65536, n=2^22) with the super-chunk loop form that config 4 now runs;\n"
       "#   cfg4_qreg is the earlier full-size capture of the per-chunk form (same tiles, grid order and DRAM traffic)\n")
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.txt"), "w").write(hdr + out)
print(out)

def raw(path):
    o = open(path).read()
    rows = list(csv.reader(o.splitlines()))
    return {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}

def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

traffic = {}
for wl, cap, kern, alg in [("cfg1", "cfg1_qreg", "nn_qreg_kernel", 65536 * 3 * 4), ("cfg2", "cfg2_qreg", "nn_qreg_kernel", (1 << 20) * 16 * 4),
                           ("cfg3", "cfg3_rtma", "nn_rtma_kernel", (1 << 26) * 8 * 4), ("cfg4", "cfg4_qreg", "nn_qreg_kernel", (1 << 24) * 16 * 4),
                           ]:
    p = os.path.join(ROOT, "gpurun_out", f"{tag}_{cap}.raw.csv")
    if not os.path.exists(p):
        continue
    d = raw(p)
    traffic[wl] = {"kernel": kern, "dram_bytes_read": to_bytes(*d["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(*d["dram__bytes_write.sum"]),
                   "algorithmic_bytes": alg, "gpu_time_us": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[d["gpu__time_duration.sum"][1]],
                   "source": f"profiles/{tag}_ncu_summary.txt ({tag}_{cap}.ncu-rep, ncu --set full, one launch)"}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
for name in (f"{tag}_launches_bench_cfg4.csv", f"{tag}_launches_bench_cfg2.csv"):
    src = os.path.join(ROOT, "gpurun_out", name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(ROOT, "profiles", name))
print(json.dumps(traffic, indent=1))
