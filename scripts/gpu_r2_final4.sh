#!/bin/bash
# Round-end pass after the super-chunk loop form: smoke, the GPU suite, the driver's bench command, the ncu
# launch list of the bench command and one full capture of the super-chunk kernel (quarter-size config 4).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) 2>&1 | tail -5
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tail -9
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("cfg4 ms/step %.3f frac %.3f e2e %.1f pinned %.1f parity %s launches %s" % (d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_call"], d["e2e"]["pinned"]["ms_per_call"], d["parity_spot_check"], d["gpu_launches"]))
for k, v in (d["all_configs"] or {}).items():
    if "ms_per_step" in v:
        print(k, "ms %.4f frac %.3f e2e %.3f ms parity %s" % (v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["ms_per_call"], v["parity_spot_check"]))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-all-configs --no-e2e"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_cfg4.csv $CMD > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
B=./multicore-hw2_b200/nn_bench
C="$B --k 16 --m 65536 --n 4194304 --fused 1 --iters 2 --warmup 1"
$C > gpurun_out/plain_super.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_qreg -s 1 -c 1 -f -o gpurun_out/r02_cfg4q_qreg_super $C > gpurun_out/ncu_super.log 2>&1
tail -1 gpurun_out/ncu_super.log | cut -c1-150
ncu -i gpurun_out/r02_cfg4q_qreg_super.ncu-rep --page raw --csv > gpurun_out/r02_cfg4q_qreg_super.raw.csv 2>/dev/null
rm -f gpurun_out/plain_*.log; du -sh gpurun_out
