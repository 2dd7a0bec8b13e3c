#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout -s KILL 300 python tools/prof_dct.py 2>&1 | grep "dct axis" > $O/prof6_fused.log
PYPDE_B200_LIB=$PWD/_ab/libnofuse.so timeout -s KILL 300 python tools/prof_dct.py 2>&1 | grep "dct axis" > $O/prof6_nofuse.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench6_default.json 2> $O/bench6_default.err
grep -E "passed|failed|rc=|^FAILED" $O/pytest_gpu.log | tail -5; cat $O/prof6_fused.log; echo ---; cat $O/prof6_nofuse.log; python -c "
import json; d=json.load(open('$O/bench6_default.json')); print(d['ms_per_step']); print(d['kernel_ms_per_step'])"
