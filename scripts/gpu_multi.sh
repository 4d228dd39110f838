#!/bin/bash
# Multi-GPU check (gpurun --gpus N): in-process sharded host entry + torchrun bench at N ranks.
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu or config1" > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  tail -3 gpurun_out/bench_n$n.err | cut -c1-300; cat gpurun_out/bench_n$n.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N',d['n_gpus'],'value %.4e'%d['value'],'ms/step %.3f'%d['ms_per_step'],'roofline %.3f'%d['roofline']['frac'],'e2e',d['e2e'] and ('%.4e pairs/s %.3f ms'%(d['e2e']['value'],d['e2e']['ms_per_call'])), 'same', d['e2e'] and d['e2e']['matches_device_resident_result'])
"
done
