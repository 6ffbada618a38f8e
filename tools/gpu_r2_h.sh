#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_llama_step.py -x -q -s -m gpu -k "decode_step_kernel" > gpurun_out/r2h_dstep.log 2>&1; echo "dstep rc=$?" >> gpurun_out/r2h_dstep.log
tail -25 gpurun_out/r2h_dstep.log | cut -c1-400
