#!/bin/bash
# round 2: after the cheaper finish protocol, rtma early-return restore and the flex winners-only epilogue
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
B=./multicore-hw2_b200/nn_bench
O=gpurun_out/r2_flex2.jsonl
: > $O
run() { $B "$@" 2>>gpurun_out/r2_flex2.err | grep -v '"device"' >> $O; }
for f in 0 1; do
  run --k 3 --m 1024 --n 65536 --fused $f --iters 31 --warmup 5 --check 1 --tag "cfg1 fused=$f"
  run --k 3 --m 1024 --n 64 --fused $f --iters 31 --warmup 5 --tag "fixed qreg fused=$f"
  run --k 8 --m 8 --n 227328 --fused $f --iters 31 --warmup 5 --tag "fixed rtma fused=$f"
  run --k 8 --m 8 --n 8388608 --fused $f --iters 31 --warmup 5 --check 1 --tag "cfg3 shard fused=$f"
  run --k 8 --m 8 --n 67108864 --fused $f --iters 11 --warmup 3 --tag "cfg3 fused=$f"
  run --k 8 --m 1 --n 64 --fused $f --iters 31 --warmup 5 --tag "fixed rreg fused=$f"
  run --k 16 --m 1024 --n 65536 --fused $f --iters 31 --warmup 5 --check 1 --tag "ta7 fused=$f"
  run --k 16 --m 4096 --n 1048576 --fused $f --iters 7 --tag "cfg2 fused=$f"
done
for k in 3 8 16; do for m in 9 16 32 64 100 112 128 200 300; do
  run --k $k --m $m --n 4194304 --variant 5 --iters 7 --check 1 --tag "flex"
done; done
for k in 3 8 16; do for q in 2 4 8; do for m in 16 32 64; do
  run --k $k --m $m --n 4194304 --variant 5 --q $q --iters 7 --tag "flex q$q"
done; done; done
for k in 3 8 16; do for m in 9 16 32; do
  run --k $k --m $m --n 4194304 --variant 4 --iters 7 --tag "rtma"
done; done
python - <<'PY'
import json
for l in open("gpurun_out/r2_flex2.jsonl"):
    d = json.loads(l)
    print(f"{d['tag']:24s} k={d['k']:2d} m={d['m']:5d} n={d['n']:9d} med {d['ms_med']*1e3:9.1f} us best {d['ms_best']*1e3:9.1f} fp32 {d['fp32_frac_maxclk']:.3f} {d['GBps']:6.0f} GB/s mism {d['mismatch_vs_plain']} | {d['plan'][:100]}")
PY
tail -5 gpurun_out/r2_flex2.err
