#!/bin/bash
# A/B of build variants by kernel duration under ncu (metric-only pass): ARGS="--k 3 --m 1024 --n 65536 --q 2" KERN=nn_qreg
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for d in multicore-hw2_b200 build/alt_*; do
  [ -x $d/nn_bench ] || continue
  C="$d/nn_bench $ARGS --iters 6 --warmup 2"
  $C > gpurun_out/plain_ab.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:${KERN:-nn_qreg} --csv $C 2>/dev/null | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>6 and r[0].isdigit()]
d=[float(r[-1].replace(',','')) for r in rows if 'gpu__time' in r[-3]]
p=[float(r[-1]) for r in rows if 'pipe_fma' in r[-3]]
d.sort()
print('$(basename $d)', '$ARGS', 'kernel us: median %.2f min %.2f  fma pipe %.1f%%' % (d[len(d)//2]/1e3 if d[0]>1000 else d[len(d)//2], d[0]/1e3 if d[0]>1000 else d[0], sum(p)/len(p)))"
done
