#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for w in cfg4 cfg3; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --workload $w > gpurun_out/r02_auto_n2_$w.json 2> gpurun_out/r02_auto_n2_$w.err ) 2>&1 | grep real
  python - gpurun_out/r02_auto_n2_$w.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"][:40], "| merge", d["config"]["merge"], "ms/step %.4f" % d["ms_per_step"], "launches", d["gpu_launches"], "nccl leg", d.get("nccl_merge") and d["nccl_merge"]["ms_per_step"], "peer leg", d.get("peer_merge"), "parity", d["parity_spot_check"], d["parity_detail"])
PY
  tail -c 300 gpurun_out/r02_auto_n2_$w.err | grep -v "^$" | tail -2
done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) 2>&1 | cut -c1-300 | tail -4
