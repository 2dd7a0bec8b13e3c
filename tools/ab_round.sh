#!/bin/bash
# One GPU visit: parity tests, A/B timings of the base library (_ab/libbase.so) against the current
# build, a bench line and a short ncu capture of the DCT kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
if [ "$1" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
fi
for lib in _ab/libbase.so ""; do
  tag=new; [ -n "$lib" ] && tag=base
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  PYPDE_B200_LIB=${lib:+$PWD/$lib} timeout 300 python tools/prof_dct.py > $O/prof_dct_$tag.log 2>&1
  PYPDE_B200_LIB=${lib:+$PWD/$lib} timeout 300 python tools/gpu_probe.py > $O/gpu_probe_$tag.log 2>&1
done
for sh in 0 1 2; do
  PDE_GEMM_SHAPE=$sh timeout 200 python tools/gpu_probe.py 2>&1 | grep "2046\|8192" > $O/gpu_probe_shape$sh.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_rbc2048.json 2> $O/bench_rbc2048.err
timeout 300 ncu --set full --clock-control none -k regex:k_dct_fft_t -s 6 -c 2 -o $O/ncu_dct_new -f python tools/prof_dct.py > $O/ncu_dct.log 2>&1
ncu -i $O/ncu_dct_new.ncu-rep --page raw --csv > $O/ncu_dct_new_raw.csv 2>/dev/null
tail -5 $O/pytest_gpu.log; cat $O/prof_dct_*.log | grep -i "dct\|diff\|fdma\|from_cheb"; grep -H "2046\|_8192" $O/gpu_probe_*.log; cat $O/bench_rbc2048.json | head -c 3000
