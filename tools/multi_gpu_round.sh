#!/bin/bash
# tools/multi_gpu_round.sh N: slab parity check + the rbc2048 and ens128 bench lines on N GPUs of one node
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --master-port 29513 tests/dist_slab_check.py > $O/slab_check_$N.log 2>&1; echo "slab check rc=$?"; tail -2 $O/slab_check_$N.log
timeout -s KILL 240 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench rc=$?"
timeout -s KILL 240 $TR --master-port 29512 bench.py --gpus $N --workload ens128 --steps 10 --warmup 3 > $O/bench_ens128_${N}gpu.json 2> $O/bench_ens128_${N}gpu.err; echo "ens rc=$?"
python - <<PY
import json
for f in ("$O/bench_${N}gpu.json", "$O/bench_ens128_${N}gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"], 1), d.get("parity_vs_single_gpu"), d.get("transpose_ms_per_step"), d.get("member_steps_per_sec"), d["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
