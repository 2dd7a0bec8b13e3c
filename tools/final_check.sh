#!/bin/bash
# short round-end check: GPU parity tests, smoke, the default bench line
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout -s KILL 600 python bench.py > $O/bench_check.json 2> $O/bench_check.err; echo "bench rc=$?" >> $O/bench_check.err
grep -E "passed|failed|rc=|^FAILED" $O/pytest_gpu.log | tail -4; tail -2 $O/smoke.log; head -c 400 $O/bench_check.json; echo; tail -1 $O/bench_check.err
