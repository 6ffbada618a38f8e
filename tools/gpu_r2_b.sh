#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_reference_parity.py -q -s -m gpu > gpurun_out/r2b_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2b_parity.log
timeout 900 python -m pytest tests/test_gpu_fattn.py tests/test_gpu_llama_step.py tests/test_gpu_glue.py -q -s -m gpu -k "exact or scratch or replay or swiglu" > gpurun_out/r2b_fa.log 2>&1; echo "fa rc=$?" >> gpurun_out/r2b_fa.log
grep -h "PARITY\|SPLITS\|passed\|failed\|rc=\|fa_exact\|Error" gpurun_out/r2b_parity.log gpurun_out/r2b_fa.log | cut -c1-700 | tail -40
