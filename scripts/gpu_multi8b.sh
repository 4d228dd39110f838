#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { # tag, env, args
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --no-e2e $3 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_$1.err | tail -3 | cut -c1-300
  python - <<PY
import json
for l in open("gpurun_out/bench_$1.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("$1 N=%d %s value %.4e  step %.4f ms  kernel %.4f ms (ranks %s) merge %.4f ms steps %s clocks %s" % (d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], r["kernel_ms"], d["kernel_ms_min_max_over_ranks"], d["merge_ms"], d["step_ms_min_max"], d["clocks"]))
PY
}
run n${N}_cfg2_a "A=1" ""
run n${N}_cfg2_nosampler "NN_BENCH_NO_SAMPLER=1" ""
run n${N}_cfg2_ll "NCCL_PROTO=LL NCCL_ALGO=Ring NN_BENCH_NO_SAMPLER=1" ""
