#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 200 python tools/dbg_tile.py 2046 > $O/dbg_tile_on.log 2>&1
PDE_SWEEP_TILE=0 timeout -s KILL 200 python tools/dbg_tile.py 2046 > $O/dbg_tile_off.log 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench5_default.json 2> $O/bench5_default.err
cat $O/dbg_tile_on.log; echo ---; cat $O/dbg_tile_off.log; grep -E "passed|failed|rc=|^FAILED" $O/pytest_gpu.log | tail -5; python -c "
import json; d=json.load(open('$O/bench5_default.json')); print(d['ms_per_step']); print({k:v for k,v in d['kernel_ms_per_step'].items() if 'sweep' in k})"
