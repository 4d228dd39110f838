#!/bin/bash
# after the 2-D grouped CTA order: full GPU suite, N=1 bench line, full-size ncu capture of config 4 (bench.py's plan)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tail -9
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("cfg4 value %.4e ms/step %.3f frac %.3f e2e %.1f pinned %.1f index %.1f parity %s launches %d" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_call"], d["e2e"]["pinned"]["ms_per_call"],
    d["e2e"]["resident_index"]["ms_per_call"], d["parity_spot_check"], d["gpu_launches"]))
for k, v in (d["all_configs"] or {}).items():
    if "ms_per_step" in v:
        print(k, "ms %.4f frac %.3f e2e %.3f ms parity %s" % (v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["ms_per_call"], v["parity_spot_check"]))
PY
B=./multicore-hw2_b200/nn_bench
TAG=r02
cap() {
  local C="$B $3 --iters 2 --warmup 1"
  $C > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/${TAG}_$1 $C > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-150
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1.raw.csv 2>/dev/null
  if [ "$4" != "keep" ]; then rm -f gpurun_out/${TAG}_$1.ncu-rep; fi
}
cap cfg2_qreg nn_qreg "--k 16 --m 4096 --n 1048576 --fused 1"
cap cfg5s_qreg nn_qreg "--k 3 --m 65536 --n 1048576 --fused 1"
cap cfg4_qreg nn_qreg "--k 16 --m 65536 --n 16777216 --fused 1" keep
du -sh gpurun_out
