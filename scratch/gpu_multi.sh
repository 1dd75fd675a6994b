#!/bin/bash
# multi-GPU bench: bash scratch/gpu_multi.sh <tag> <N> [<N> ...]
T=$1; shift
mkdir -p gpurun_out
for N in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
     bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_n${N}_bench.log 2> gpurun_out/${T}_n${N}_bench.err
  echo "N=$N rc=$?"
  tail -1 gpurun_out/${T}_n${N}_bench.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('N=%d value %.1f e2e %.1f ms %.2f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']), d['run'])
except Exception as e: print('parse failed', e)
"
  tail -3 gpurun_out/${T}_n${N}_bench.err
done
