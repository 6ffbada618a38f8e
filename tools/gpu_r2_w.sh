#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q -m gpu 2>&1 | tail -5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2w_bench2.log 2>&1
tail -1 gpurun_out/r2w_bench2.log | cut -c1-3000
