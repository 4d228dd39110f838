#!/bin/bash
# round 2, 8 GPUs of one box: multi-GPU parity, the strong-scaling bench lines (peer merge + NCCL leg), GPU-count sweep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi topo -m > gpurun_out/r02_topo_8.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q -k "peer_merge or multi_gpu or every_gpu" 2>&1 | tail -8 ) > gpurun_out/r02_pytest_multi_8.log
tail -4 gpurun_out/r02_pytest_multi_8.log
run() { # N, tag, port, args...
  local N=$1 tag=$2 port=$3; shift 3
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/r02_bench_n${N}_$tag.json 2> gpurun_out/r02_bench_n${N}_$tag.err ) 2>&1 | grep real
  python - "$N $tag" "gpurun_out/r02_bench_n${N}_$tag.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], "NO LINE", e); sys.exit(0)
e = d.get("e2e") or {}
n = d.get("nccl_merge") or {}
print("N", sys.argv[1], "merge=%s ms/step %.4f (kernel on rank0 %.4f, frac %.3f, kernel min/max over ranks %s) nccl-leg %.4f (kernel %.4f merge %.4f same %s) e2e %s parity %s note %s" % (
    d["config"]["merge"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], [round(x, 4) for x in d["kernel_ms_min_max_over_ranks"]],
    n.get("ms_per_step", -1), n.get("kernel_ms", -1), n.get("merge_ms", -1), n.get("same_result_as_peer_merge"), e.get("ms_per_call"), d.get("parity_detail"), d["config"].get("merge_note")))
PY
}
run 8 cfg4 29511 --steps 10 --warmup 5
run 8 cfg3 29512 --workload cfg3 --steps 20 --warmup 5
run 8 cfg1 29513 --workload cfg1 --steps 20 --warmup 5 --no-e2e
run 8 cfg2 29514 --workload cfg2 --steps 20 --warmup 5 --no-e2e
run 8 cfg5 29515 --workload cfg5 --steps 5 --warmup 3 --no-e2e
run 4 cfg4 29516 --steps 3 --warmup 3 --no-e2e
run 2 cfg4 29517 --steps 3 --warmup 3 --no-e2e
timeout 400 python scripts/gpu_count_sweep.py 2>&1 | tail -20
rm -f gpurun_out/*.err.tmp; du -sh gpurun_out
