#!/bin/bash
# the driver's scaling command at N GPUs (both arms), final code
N=${1:-4}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_final_n$N.json 2> gpurun_out/r02_final_n$N.err ) 2>&1 | grep real
python - gpurun_out/r02_final_n$N.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d["e2e"]
print("N", d["n_gpus"], "merge", d["config"]["merge"], "ms/step %.3f value %.4e frac %.3f" % (d["ms_per_step"], d["value"], d["roofline"]["frac"]),
      "| e2e cudaCallback %.1f ms pinned %.1f nccl-merge %.1f index %.1f" % (e["ms_per_call"], e["pinned"]["ms_per_call"], e["pinned_nccl_merge"]["ms_per_call"], e["resident_index"]["ms_per_call"]),
      "| peer leg", d.get("peer_merge"), "| parity", d["parity_spot_check"], d["parity_detail"], "| launches", d["gpu_launches"], "clocks", d["clocks"])
PY
tail -c 200 gpurun_out/r02_final_n$N.err | grep -v "^$" | tail -2
