#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/cortex.llamacpp_b200:$PWD/oracle/_ref
python tools/make_gguf.py --model tinyllama --ftype q4_0 --out /tmp/tl.gguf 2>/dev/null
D=oracle/_ref/logits_dump
LOGITS_DUMP_SEED=9 LOGITS_DUMP_NODES=/tmp/n_cpu.bin $D /tmp/tl.gguf /tmp/o_cpu.bin 0 32 2 f16 1 16 > /dev/null 2>&1
LOGITS_DUMP_SEED=9 GGML_B200_CPU_EXACT=1 LOGITS_DUMP_NODES=/tmp/n_gpu.bin GGML_BACKEND_PATH=$PWD/cortex.llamacpp_b200/libggml-b200.so $D /tmp/tl.gguf /tmp/o_gpu.bin 99 32 2 f16 1 4 > /dev/null 2>&1
python tools/compare_nodes.py /tmp/n_cpu.bin /tmp/n_gpu.bin 1e-12 > gpurun_out/r2e_nodes.txt 2>&1
python tools/compare_logits.py /tmp/o_cpu.bin /tmp/o_gpu.bin >> gpurun_out/r2e_nodes.txt 2>&1
head -60 gpurun_out/r2e_nodes.txt; tail -2 gpurun_out/r2e_nodes.txt
