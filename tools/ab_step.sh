#!/bin/bash
# A/B of library variants (_ab/lib<name>.so, tools/build_variant.sh) inside the rbc2048 step (GPU box): tools/ab_dct.sh name ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { echo "== $1"; shift; env "$@" $B 2> gpurun_out/ab_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernel_ms_per_step']
print(json.dumps({'ms': d['ms_per_step'], 'dct1': k.get('pde_dct1_multi[axis1]'), 'dct0': k.get('pde_dct1_multi[axis0]'), 'gemm': k.get('pde_gemm_f64'), 'passes': {n[5:-1]: v for n, v in k.items() if n.startswith('pass[')}}))
" || tail -5 gpurun_out/ab_err.log; }
run default X=1
for v in "$@"; do run $v PYPDE_B200_LIB=$PWD/_ab/lib$v.so; done
