#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout -s KILL 600 python tools/bench_dct.py > $O/dct_sweep.log 2>&1
grep -E "passed|failed|rc=|^FAILED" $O/pytest_gpu.log | tail -4
python - <<'PY'
import json
d=json.load(open('gpurun_out/dct_sweep.json'))
for r in d['sweep']:
    print(r['N'], r['algo'], "axis0 %.0f axis1 %.0f GB/s  roundtrip %.1e" % (r['bwd_axis0_gbs'], r['bwd_axis1_gbs'], r['roundtrip_rel']))
PY
