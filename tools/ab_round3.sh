#!/bin/bash
# GPU visit 3: tests, strip-banded A/B, ncu --set full of one eager stage (DCT, banded, sweeps), launch list
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench3_default.json 2> $O/bench3_default.err
PDE_BANDED_STRIP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench3_nostrip.json 2> $O/bench3_nostrip.err
# one eager stage = 41 launches; skip the first full step (3 stages) and capture the next stage
timeout 900 ncu --set full --clock-control none -k regex:"k_dct_fft_t|k_banded|k_sweep|k_to_cheb|k_lincomb|k_conv" -s 123 -c 45 -o $O/ncu_stage -f python tools/prof_sweeps.py > $O/ncu_stage.log 2>&1
ncu -i $O/ncu_stage.ncu-rep --page raw --csv > $O/ncu_stage_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_rbc2048.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > $O/launches_bench.log 2>&1
tail -3 $O/pytest_gpu.log; head -c 400 $O/bench3_default.json; echo; head -c 400 $O/bench3_nostrip.json; echo; tail -3 $O/ncu_stage.log; ls -la $O
