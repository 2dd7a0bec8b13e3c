#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for v in "" rs223 rs224; do
  tag=${v:-default}
  PYPDE_B200_LIB=${v:+$PWD/_ab/lib$v.so} timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench7_$tag.json 2> $O/bench7_$tag.err
  python -c "
import json; d=json.load(open('$O/bench7_$tag.json')); print('$tag', d['ms_per_step'], d['kernel_ms_per_step']['pde_banded_multi'])"
done
