#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tp.py tests/test_gpu_reference_parity.py -x -q -m gpu -k "split_mode or row_split" 2>&1 | tail -30
