#!/bin/bash
# round 2, first look: cfg1 tile/split landscape, the 9..112-query band, cfg3's 8-GPU shard on one GPU
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
O=gpurun_out/r2_explore.jsonl
: > $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv >> gpurun_out/r2_box.txt
for q in 2 4 8; do for s in 0 37 74 111 148 185 222 296 370; do
  $B --k 3 --m 1024 --n 65536 --variant 1 --q $q --splits $s --iters 21 --warmup 5 | grep -v device | sed "s/\"single\"/\"cfg1 q$q s$s\"/" >> $O
done; done
for k in 3 8 16; do for m in 9 16 32 64 100 112; do for v in 1 4; do
  $B --k $k --m $m --n 4194304 --variant $v --iters 7 | grep -v device | sed "s/\"single\"/\"few v$v\"/" >> $O
done; done; done
for n in 8388608 16777216 33554432; do
  $B --k 8 --m 8 --n $n --iters 15 | grep -v device | sed "s/\"single\"/\"cfg3 shard\"/" >> $O
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_explore.jsonl"):
    d = json.loads(l)
    print(f"{d['tag']:16s} k={d['k']:2d} m={d['m']:5d} n={d['n']:9d} med {d['ms_med']*1e3:9.1f} us best {d['ms_best']*1e3:9.1f} us fp32 {d['fp32_frac_maxclk']:.3f} {d['GBps']:7.0f} GB/s | {d['plan'][:90]}")
PY
