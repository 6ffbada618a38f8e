#!/bin/bash
# round 2, lease A: new parity tests through the plugin + exact-mode FA + scratch-growth regression
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_reference_parity.py -x -q -s -m gpu > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2a_parity.log
timeout 600 python -m pytest tests/test_gpu_fattn.py tests/test_gpu_llama_step.py -q -s -m gpu -k "exact or scratch or replay" > gpurun_out/r2a_fa.log 2>&1; echo "fa rc=$?" >> gpurun_out/r2a_fa.log
grep -h "PARITY\|passed\|failed\|rc=\|fa_exact\|Error\|error" gpurun_out/r2a_parity.log gpurun_out/r2a_fa.log | tail -40
