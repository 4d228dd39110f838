#!/bin/bash
# round 2: parity (full GPU suite), one-launch search timings, phased query-register kernel over the 9..500 query band
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
B=./multicore-hw2_b200/nn_bench
O=gpurun_out/r2_flex.jsonl
: > $O
run() { $B "$@" 2>>gpurun_out/r2_flex.err | grep -v '"device"' >> $O; }
for f in 0 1; do
  run --k 3 --m 1024 --n 65536 --fused $f --iters 31 --warmup 5 --check 1 --tag "cfg1 fused=$f"
  for s in 37 74 111 148 185; do run --k 3 --m 1024 --n 65536 --fused $f --q 2 --splits $s --iters 31 --warmup 5 --tag "cfg1 q2 s$s fused=$f"; done
  for s in 74 148; do run --k 3 --m 1024 --n 65536 --fused $f --q 4 --splits $s --iters 31 --warmup 5 --tag "cfg1 q4 s$s fused=$f"; done
  run --k 3 --m 1024 --n 64 --fused $f --iters 31 --warmup 5 --tag "fixed qreg fused=$f"
  run --k 8 --m 8 --n 227328 --fused $f --iters 31 --warmup 5 --tag "fixed rtma fused=$f"
  run --k 8 --m 8 --n 2273280 --fused $f --iters 31 --warmup 5 --tag "fixed rtma fused=$f"
  run --k 8 --m 8 --n 8388608 --fused $f --iters 31 --warmup 5 --check 1 --tag "cfg3 shard fused=$f"
  run --k 8 --m 8 --n 67108864 --fused $f --iters 11 --warmup 3 --tag "cfg3 fused=$f"
  run --k 8 --m 1 --n 64 --fused $f --iters 31 --warmup 5 --tag "fixed rreg fused=$f"
  run --k 8 --m 1 --n 67108864 --fused $f --iters 11 --warmup 3 --tag "m1 fused=$f"
  run --k 16 --m 1024 --n 65536 --fused $f --iters 31 --warmup 5 --check 1 --tag "ta7 fused=$f"
done
for k in 3 8 16; do for m in 5 8 9 16 32 64 100 112 128 200 256 500; do
  run --k $k --m $m --n 4194304 --variant 5 --iters 7 --check 1 --tag "flex"
done; done
for k in 3 8 16; do for m in 128 200 256 500; do
  run --k $k --m $m --n 4194304 --variant 1 --iters 7 --tag "qreg"
done; done
for k in 3 8 16; do for q in 2 4 8; do
  run --k $k --m 100 --n 4194304 --variant 5 --q $q --iters 7 --tag "flex q$q"
done; done
for k in 3 8 16; do for n in 65536 1048576; do for m in 32 100; do
  run --k $k --m $m --n $n --variant 5 --iters 11 --tag "flex small-n"
done; done; done
run --k 8 --m 8 --n 67108864 --variant 5 --iters 7 --check 1 --tag "cfg3 flex"
python - <<'PY'
import json
for l in open("gpurun_out/r2_flex.jsonl"):
    d = json.loads(l)
    print(f"{d['tag']:24s} k={d['k']:2d} m={d['m']:5d} n={d['n']:9d} med {d['ms_med']*1e3:9.1f} us best {d['ms_best']*1e3:9.1f} fp32 {d['fp32_frac_maxclk']:.3f} {d['GBps']:6.0f} GB/s mism {d['mismatch_vs_plain']} | {d['plan'][:96]}")
PY
tail -5 gpurun_out/r2_flex.err
