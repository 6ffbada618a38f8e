#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mulmat.py tests/test_gpu_fattn.py -q -s -m gpu -k "exact" > gpurun_out/r2d_exact.log 2>&1; echo "exact rc=$?" >> gpurun_out/r2d_exact.log
timeout 1500 python -m pytest tests/test_gpu_reference_parity.py -q -s -m gpu > gpurun_out/r2d_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2d_parity.log
grep -h "PARITY\|SPLITS\|passed\|failed\|rc=\|fa_exact\|Error" gpurun_out/r2d_exact.log gpurun_out/r2d_parity.log | cut -c1-600 | tail -40
