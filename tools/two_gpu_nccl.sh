#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/nccl_$tag.json 2> $O/nccl_$tag.err
  python - <<PY
import json
txt=open('$O/nccl_$tag.json').read().splitlines()
js=[l for l in txt if l.startswith('{')]
print('$tag', 'stdout lines', len(txt), end=' ')
if js:
    d=json.loads(js[-1]); print(d['ms_per_step'], d.get('transpose_ms_per_step'))
else:
    print('no json')
PY
}
run default A=1
run chan32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run chan16 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=16
run memcpy NCCL_P2P_USE_CUDA_MEMCPY=1
