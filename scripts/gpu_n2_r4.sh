#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 300 python -m pytest tests/test_distributed.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/r4_bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-300
