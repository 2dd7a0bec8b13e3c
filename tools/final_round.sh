#!/bin/bash
# Round-end validation on one B200: parity tests, smoke, the bench line, the ncu launch list of the same command
# and one ncu --set full capture of an eager stage (CSV only; the .ncu-rep stays on the box)
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > $O/bench_final_rbc2048.json 2> $O/bench_final_rbc2048.err
for wl in rbc512 rbc64 diff1024; do
  timeout -s KILL 600 python bench.py --workload $wl --steps 20 --warmup 3 > $O/bench_final_$wl.json 2> $O/bench_final_$wl.err
done
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_rbc2048.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > $O/launches_bench.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none -k regex:"k_dct_fft_t|k_banded|k_sweep|k_to_cheb|k_lincomb|k_conv|k_gemm" -s 117 -c 39 -o /tmp/ncu_stage -f python tools/prof_sweeps.py > $O/ncu_stage.log 2>&1
ncu -i /tmp/ncu_stage.ncu-rep --page raw --csv > $O/ncu_stage_raw.csv 2>/dev/null
timeout -s KILL 600 python tools/bench_dct.py > $O/dct_sweep.log 2>&1
grep -E "passed|failed|rc=|^FAILED" $O/pytest_gpu.log | tail -4; tail -2 $O/smoke.log; for f in $O/bench_final_*.json; do head -c 260 $f; echo; done; du -sh $O
