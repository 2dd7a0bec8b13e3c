#!/bin/bash
# GPU visit 2: tests, A/B of radix-12 tail / twiddle reuse / L2 prefetch distance, bench, ncu of the axis-1 DCT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
run() { # tag, lib, env...
  tag=$1; lib=$2; shift 2
  env "$@" PYPDE_B200_LIB=${lib:+$PWD/$lib} timeout 300 python tools/prof_dct.py 2>&1 | grep -i "dct\|diff\|fdma\|from_cheb" | tail -8 | sed "s/^/$tag: /" >> $O/ab2.log
}
: > $O/ab2.log
run default "" A=1
run radix4x3 "" PDE_FFT_RADIX12=0
run notw _ab/libnotw.so A=1
run pf0 _ab/libpf0.so A=1
run pf16 _ab/libpf16.so A=1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_rbc2048.json 2> $O/bench_rbc2048.err
PYPDE_B200_LIB=$PWD/_ab/libpf0.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_rbc2048_pf0.json 2> $O/bench_rbc2048_pf0.err
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"k_dct_fft_t<3072, 1" -s 4 -c 1 -o $O/ncu_dct_ax1 -f python tools/prof_dct.py > $O/ncu_dct1.log 2>&1
ncu -i $O/ncu_dct_ax1.ncu-rep --page raw --csv > $O/ncu_dct_ax1_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"k_dct_fft_t<3072, 4" -s 4 -c 1 -o $O/ncu_dct_ax0 -f python tools/prof_dct.py > $O/ncu_dct0.log 2>&1
ncu -i $O/ncu_dct_ax0.ncu-rep --page raw --csv > $O/ncu_dct_ax0_raw.csv 2>/dev/null
tail -3 $O/pytest_gpu.log; cat $O/ab2.log; head -c 600 $O/bench_rbc2048.json; echo; head -c 600 $O/bench_rbc2048_pf0.json
