#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_llama_step.py tests/test_gpu_glue.py tests/test_gpu_mulmat.py -x -q -m gpu 2>&1 | tail -25
