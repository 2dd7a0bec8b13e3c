#!/bin/bash
# GPU visit 4: tests (incl. tiled sweeps, strip banded), bench A/B tile on/off, ncu of one eager stage (CSV only)
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench4_default.json 2> $O/bench4_default.err
PDE_SWEEP_TILE=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench4_notile.json 2> $O/bench4_notile.err
# one eager stage: skip the first step (3 stages), capture the next stage; keep only the CSV
timeout -s KILL 900 ncu --set full --clock-control none -k regex:"k_dct_fft_t|k_banded|k_sweep|k_to_cheb|k_lincomb|k_conv|k_gemm" -s 123 -c 42 -o /tmp/ncu_stage -f python tools/prof_sweeps.py > $O/ncu_stage.log 2>&1
ncu -i /tmp/ncu_stage.ncu-rep --page raw --csv > $O/ncu_stage_raw.csv 2>/dev/null
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_rbc2048.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > $O/launches_bench.log 2>&1
grep -E "passed|failed|rc=" $O/pytest_gpu.log | tail -3; grep -E "^FAILED|Error|assert" $O/pytest_gpu.log | head -10; head -c 330 $O/bench4_default.json; echo; head -c 330 $O/bench4_notile.json; echo; du -sh $O
