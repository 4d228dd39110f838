#!/bin/bash
# round 2, N GPUs: the multi-process peer merge (parity under torchrun, then the bench lines with it)
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q -k "peer_merge or multi_gpu or every_gpu or resident_index" 2>&1 | tail -25 ) > gpurun_out/r2_pytest_peer_$N.log
tail -12 gpurun_out/r2_pytest_peer_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # tag, port, args...
  local tag=$1 port=$2; shift 2
  ( time timeout 600 $TR --master-port $port bench.py --gpus $N "$@" > gpurun_out/r2_peer_n${N}_$tag.json 2> gpurun_out/r2_peer_n${N}_$tag.err ) 2>&1 | grep real
  python - "$tag" "gpurun_out/r2_peer_n${N}_$tag.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], "NO LINE", e); sys.exit(0)
e = d.get("e2e") or {}
print(sys.argv[1], "merge=%s value %.3e ms/step %.4f kernel %.4f frac %.3f minmax %s e2e %s parity %s note %s" % (
    d["config"]["merge"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["step_ms_min_max"],
    e.get("ms_per_call"), d.get("parity_detail"), d["config"].get("merge_note")))
print("   nccl leg:", d.get("nccl_merge"))
if d.get("step_ms_per_rank"):
    for r, s in enumerate(d["step_ms_per_rank"]):
        print("   rank", r, s[:10])
PY
  tail -c 300 gpurun_out/r2_peer_n${N}_$tag.err | grep -v "^$" | tail -3
}
run cfg3 29512 --workload cfg3 --steps 20 --warmup 5
run cfg1 29513 --workload cfg1 --steps 20 --warmup 5 --no-e2e
run cfg2 29514 --workload cfg2 --steps 20 --warmup 5 --no-e2e
run cfg4 29511 --steps 3 --warmup 3
