#!/bin/bash
# Round-2 validation on one B200: parity tests, smoke, the bench lines, the ncu launch list of the bench command and
# one ncu --set full capture of an eager stage (CSV only; the .ncu-rep stays on the box).
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2.log 2>&1; echo "smoke rc=$?" >> $O/smoke_r2.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > $O/bench_r2_rbc2048.json 2> $O/bench_r2_rbc2048.err
for wl in rbc512 rbc64 diff1024 dct; do
  timeout -s KILL 600 python bench.py --workload $wl --steps 10 --warmup 3 > $O/bench_r2_$wl.json 2> $O/bench_r2_$wl.err
done
timeout -s KILL 600 python bench.py --workload ens128 --steps 10 --warmup 3 > $O/bench_r2_ens128.json 2> $O/bench_r2_ens128.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_r2_rbc2048.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/launches_r2_bench.log 2>&1
# one eager stage (15 launches) after 3 warm-up steps = 135 launches of our kernels
timeout -s KILL 900 ncu --set full --clock-control none -k regex:"k_pass|k_dct|k_gemm|k_conv" -s 135 -c 15 -o /tmp/ncu_stage_r2 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_stage_r2.log 2>&1
ncu -i /tmp/ncu_stage_r2.ncu-rep --page raw --csv > $O/ncu_stage_r2_raw.csv 2>/dev/null
python tools/summarise_ncu.py raw $O/ncu_stage_r2_raw.csv $O/r02_ncu_stage.csv "ncu --set full --clock-control none -k regex:k_pass|k_dct|k_gemm|k_conv -s 135 -c 15 python bench.py --steps 1 --warmup 3 --no-graph (ONE eager rbc2048 RK3 stage of the round-2 stepper: 15 launches in order PX1 PY2 DCTx DCTy products DCTy DCTx PX3 PY4 PX5 GEMM PX6 GEMM PY7 PX8; per-launch, cold cache, serialised)"
python tools/summarise_ncu.py list $O/launches_r2_rbc2048.csv $O/r02_ncu_launches_rbc2048_summary.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 300 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline (round 2, PassStepper: 15 launches per stage; first 300 launches: warm-up + timed steps; cold-cache serialised: compare SHARES)"
python tools/bench_pass.py 2048 > $O/bench_pass_r2.json 2> $O/bench_pass_r2.err
grep -E "passed|failed|rc=|^FAILED" $O/pytest_gpu_r2.log | tail -4; tail -2 $O/smoke_r2.log; for f in $O/bench_r2_*.json; do head -c 200 $f; echo; done; du -sh $O
