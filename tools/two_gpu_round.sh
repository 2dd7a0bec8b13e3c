#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/smi2.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_rbc.py -m gpu -q -k "slab" > $O/pytest_slab.log 2>&1; echo "pytest rc=$?" >> $O/pytest_slab.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?" >> $O/bench_2gpu.err
tail -3 $O/pytest_slab.log; head -c 1500 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
