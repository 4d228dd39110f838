#!/bin/bash
# round 2: GPU suite, the new contract bench line at N=1 (cfg4 strong + all_configs), family crossover sweep, one ncu look at the phased kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_pytest.log
tail -3 gpurun_out/r2_pytest.log
( time python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f frac %.3f e2e %.1f ms pinned %.1f ms index %.1f ms parity %s launches %d" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_call"], d["e2e"]["pinned"]["ms_per_call"],
    d["e2e"]["resident_index"]["ms_per_call"], d["parity_detail"], d["gpu_launches"]))
print("cpu", d["cpu_baseline"])
for k, v in (d["all_configs"] or {}).items():
    print(k, "ms %.4f frac %.3f e2e %.3f ms parity %s launches/step %.1f | %s" % (v["ms_per_step"], v["roofline"]["frac"],
          v["e2e"]["ms_per_call"], v["parity_detail"], v["gpu_launches_per_step"], v["plan"][:70]))
PY
B=./multicore-hw2_b200/nn_bench
$B --sweep cross --iters 7 --warmup 2 2> gpurun_out/r2_cross.err | grep -v '"device"' > gpurun_out/r2_cross.jsonl
python scripts/make_crossover.py gpurun_out/r2_cross.jsonl gpurun_out/r02_fewquery_crossover.json
C="$B --k 8 --m 32 --n 4194304 --variant 5 --iters 2 --warmup 1"
$C > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_qflex -s 1 -c 1 -f -o gpurun_out/r02_flex_k8m32 $C > gpurun_out/ncu_flex.log 2>&1
tail -2 gpurun_out/ncu_flex.log | cut -c1-200
