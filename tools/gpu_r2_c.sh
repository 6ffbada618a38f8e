#!/bin/bash
# node-level diagnosis: first op whose output deviates between the CPU backend and the plugin
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/cortex.llamacpp_b200:$PWD/oracle/_ref
python tools/make_gguf.py --model llama3-8b --layers 2 --ftype q4_k_m --out /tmp/l3x2.gguf 2>/dev/null
D=oracle/_ref/logits_dump
LOGITS_DUMP_SEED=10 LOGITS_DUMP_NODES=/tmp/n_cpu.bin $D /tmp/l3x2.gguf /tmp/o_cpu.bin 0 32 3 q8_0 1 16 > /dev/null 2>&1
LOGITS_DUMP_SEED=10 LOGITS_DUMP_NODES=/tmp/n_gpu.bin GGML_BACKEND_PATH=$PWD/cortex.llamacpp_b200/libggml-b200.so $D /tmp/l3x2.gguf /tmp/o_gpu.bin 99 32 3 q8_0 1 4 > /dev/null 2>&1
python tools/compare_nodes.py /tmp/n_cpu.bin /tmp/n_gpu.bin > gpurun_out/r2c_nodes.txt 2>&1
python tools/compare_logits.py /tmp/o_cpu.bin /tmp/o_gpu.bin >> gpurun_out/r2c_nodes.txt 2>&1
head -120 gpurun_out/r2c_nodes.txt
