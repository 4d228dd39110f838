#!/bin/bash
# round 2, N GPUs of one box: multi-GPU parity tests, strong-scaling bench lines, GPU-count sweep
N=${1:-4}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi topo -m > gpurun_out/r2_topo_$N.txt 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu or resident_index or harness" 2>&1 | tail -5 ) > gpurun_out/r2_pytest_multi_$N.log
tail -3 gpurun_out/r2_pytest_multi_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # tag, port, args...
  local tag=$1 port=$2; shift 2
  ( time timeout 600 $TR --master-port $port bench.py --gpus $N "$@" > gpurun_out/r2_bench_n${N}_$tag.json 2> gpurun_out/r2_bench_n${N}_$tag.err ) 2>&1 | grep real
  python - "$tag" "gpurun_out/r2_bench_n${N}_$tag.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], "NO LINE", e); sys.exit(0)
e = d.get("e2e") or {}
print(sys.argv[1], "value %.3e ms/step %.4f kernel %.4f frac %.3f merge %.4f minmax %s e2e %s parity %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["merge_ms"], d["step_ms_min_max"],
    e.get("ms_per_call"), d.get("parity_detail")))
if d.get("step_ms_per_rank"):
    for r, s in enumerate(d["step_ms_per_rank"]):
        print("   rank", r, s[:12])
PY
}
run cfg4 29511 --steps 5 --warmup 3
run cfg3 29512 --workload cfg3 --steps 20 --warmup 5
run cfg1 29513 --workload cfg1 --steps 20 --warmup 5 --no-e2e
run cfg2 29514 --workload cfg2 --steps 20 --warmup 5 --no-e2e
run cfg5 29515 --workload cfg5 --steps 5 --warmup 3 --no-e2e
tail -c 400 gpurun_out/r2_bench_n${N}_cfg4.err
python scripts/gpu_count_sweep.py 2>&1 | tail -40
