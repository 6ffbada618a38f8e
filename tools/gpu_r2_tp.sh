#!/bin/bash
# 2-GPU evidence (gpurun --gpus 2): tensor-parallel parity tests (Megatron split, lost peer, row-split C ABI, llama.cpp --split-mode row)
# and the N=2 bench line (value = one Llama-3-70B stream row-split over the GPUs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tp.py tests/test_gpu_reference_parity.py -q -m gpu -k "tp or lost_peer or split" 2>&1 | tail -8 | tee gpurun_out/r2_pytest_tp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 8 --warmup 1 > gpurun_out/r2_bench2_reference.json 2>/dev/null
tail -c 700 gpurun_out/r2_bench2_reference.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -1 gpurun_out/r2_bench2.json | cut -c1-1500
