#!/bin/bash
# A/B of the DCT row kernels inside the rbc2048 step (GPU box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { echo "== $1"; shift; env "$@" PDE_DCT_DEBUG=1 $B 2> gpurun_out/ab_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernel_ms_per_step']
print(json.dumps({'ms': d['ms_per_step'], 'dct1': k.get('pde_dct1_multi[axis1]'), 'dct0': k.get('pde_dct1_multi[axis0]'), 'products': k.get('pde_conv_products')}))
" || tail -5 gpurun_out/ab_err.log; grep k_dct_row gpurun_out/ab_err.log; }
run carve85 PDE_DCT_CARVE=85
run carve86 PDE_DCT_CARVE=86
run carve87 PDE_DCT_CARVE=87
run carve100 PDE_DCT_CARVE=100
