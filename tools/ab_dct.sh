#!/bin/bash
# A/B of the DCT kernels inside the rbc2048 step (GPU box): default build (persistent TMA row kernel) vs the
# previous row kernel (PDE_DCT_TMA=0), two columns per CTA on axis 0, and the twiddle-product build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_primitives.py -q -x -m gpu -k "dct or fft" 2>&1 | tail -5
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { echo "== $1"; shift; env "$@" $B 2> gpurun_out/ab_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernel_ms_per_step']
print(json.dumps({'ms': d['ms_per_step'], 'dct1': k.get('pde_dct1_multi[axis1]'), 'dct0': k.get('pde_dct1_multi[axis0]'), 'products': k.get('pde_conv_products')}))
" || tail -5 gpurun_out/ab_err.log; }
run default X=1
run old_rows PDE_DCT_TMA=0
run axis0_S2 PDE_FFT_AXIS0_S=2
run twpow PYPDE_B200_LIB=$PWD/pypde_b200/_lib/libpypde_b200_twpow.so
run twpow_S2 PYPDE_B200_LIB=$PWD/pypde_b200/_lib/libpypde_b200_twpow.so PDE_FFT_AXIS0_S=2
python -m pytest tests/test_gpu_large.py tests/test_gpu_rbc.py -q -x -m gpu 2>&1 | tail -5
