#!/bin/bash
# round 2: parity of the one-launch search + its effect on config 1; fixed overheads of each kernel family
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
B=./multicore-hw2_b200/nn_bench
O=gpurun_out/r2_fused.jsonl
: > $O
run() { $B "$@" | grep -v device >> $O; }
for f in 0 1; do
  run --k 3 --m 1024 --n 65536 --fused $f --iters 31 --warmup 5 --check 1 --tag "cfg1 fused=$f"
  for s in 37 74 148; do run --k 3 --m 1024 --n 65536 --fused $f --q 2 --splits $s --iters 31 --warmup 5 --tag "cfg1 q2 s$s fused=$f"; done
  for s in 74 148 222; do run --k 3 --m 1024 --n 65536 --fused $f --q 4 --splits $s --iters 31 --warmup 5 --tag "cfg1 q4 s$s fused=$f"; done
  run --k 3 --m 1024 --n 64 --fused $f --iters 31 --warmup 5 --tag "fixed qreg fused=$f"
  run --k 3 --m 1024 --n 4096 --fused $f --iters 31 --warmup 5 --tag "fixed qreg fused=$f"
  run --k 3 --m 1024 --n 16384 --fused $f --iters 31 --warmup 5 --tag "fixed qreg fused=$f"
  run --k 3 --m 1024 --n 32768 --fused $f --iters 31 --warmup 5 --tag "fixed qreg fused=$f"
  run --k 3 --m 1024 --n 131072 --fused $f --iters 31 --warmup 5 --tag "fixed qreg fused=$f"
  run --k 8 --m 8 --n 1536 --fused $f --iters 31 --warmup 5 --tag "fixed rtma fused=$f"
  run --k 8 --m 8 --n 227328 --fused $f --iters 31 --warmup 5 --tag "fixed rtma fused=$f"
  run --k 8 --m 8 --n 2273280 --fused $f --iters 31 --warmup 5 --tag "fixed rtma fused=$f"
  run --k 8 --m 8 --n 8388608 --fused $f --iters 31 --warmup 5 --check 1 --tag "cfg3 shard fused=$f"
  run --k 8 --m 1 --n 64 --fused $f --iters 31 --warmup 5 --tag "fixed rreg fused=$f"
  run --k 16 --m 1024 --n 65536 --fused $f --iters 31 --warmup 5 --check 1 --tag "ta7 fused=$f"
  run --k 16 --m 4096 --n 1048576 --fused $f --iters 7 --check 1 --tag "cfg2 fused=$f"
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_fused.jsonl"):
    d = json.loads(l)
    print(f"{d['tag']:26s} k={d['k']:2d} m={d['m']:5d} n={d['n']:9d} med {d['ms_med']*1e3:9.1f} us best {d['ms_best']*1e3:9.1f} us fp32 {d['fp32_frac_maxclk']:.3f} mism {d['mismatch_vs_plain']} | {d['plan'][:80]}")
PY
