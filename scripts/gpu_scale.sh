#!/bin/bash
# What the driver's scaling run does, for the GPU counts available on this box: bench.py at N = 1, 2, 4[, 8].
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
G=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -le $G ] || continue
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  python - <<PY
import json
for l in open("gpurun_out/scale_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); e=d["e2e"]
        print("N=%d value %.4e pairs/s  step %.4f ms  kernel %.4f ms  merge %.4f ms  e2e %.3f ms same=%s launches %d" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["merge_ms"], e["ms_per_call"], e["matches_device_resident_result"], d["gpu_launches"]))
PY
done
