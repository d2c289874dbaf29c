#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/r2_bench_n$N.json | cut -c1-2500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 scratch/gpu_gather_sweep.py 2>&1 | grep -v Warning | tail -14 | tee $OUT/r2z_gather_sweep_n$N.log
