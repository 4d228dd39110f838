#!/bin/bash
# N-GPU confirmation: multi-GPU host entry + resident index tests, then torchrun bench at N ranks
# (weak cfg2 = what the driver's scaling run does; strong cfg4 = BASELINE configs[3]).
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/gpus.txt; wc -l gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu or resident_index" > gpurun_out/pytest_multi.log 2>&1; tail -2 gpurun_out/pytest_multi.log
run() { # tag, args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $2 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/bench_$1.err | tail -2 | cut -c1-300
  python - <<PY
import json
for l in open("gpurun_out/bench_$1.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; e=d["e2e"]
        print("$1 N=%d %s value %.4e pairs/s  step %.4f ms  kernel %.4f ms frac %.3f merge %.4f ms  e2e %s" % (d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["merge_ms"], e and ("%.4e (%.3f ms) same=%s resident %.3f ms" % (e["value"], e["ms_per_call"], e["matches_device_resident_result"], e["resident_index"]["ms_per_call"]))))
        print("   steps", d["step_ms"])
PY
}
run n${N}_cfg2_weak "--steps 10"
run n${N}_cfg4_strong "--workload cfg4 --scaling strong --steps 5"
