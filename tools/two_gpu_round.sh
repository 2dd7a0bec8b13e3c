#!/bin/bash
# 2-GPU validation of the slab-sharded path: parity check against the oracle (tests/dist_slab_check.py), bench line
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_rbc.py -m gpu -q -k "slab" > $O/pytest_slab.log 2>&1; echo "pytest rc=$?" >> $O/pytest_slab.log
timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?" >> $O/bench_2gpu.err
tail -3 $O/pytest_slab.log; head -c 600 $O/bench_2gpu.json; echo; tail -1 $O/bench_2gpu.err
