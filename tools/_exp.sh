HERE=/root/repo
export LD_LIBRARY_PATH="$HERE/cortex.llamacpp_b200:$HERE/oracle/_ref:$LD_LIBRARY_PATH"
export GGML_BACKEND_PATH=$HERE/cortex.llamacpp_b200/libggml-b200.so GGML_B200_NO_OFFLOAD=1
D="$HERE/oracle/_ref/logits_dump"
ft=${1:-q4_k_m}; kv=${2:-q8_0}
G=/tmp/x_$ft.gguf
python $HERE/tools/make_gguf.py --model tiny-d128 --ftype $ft --out $G 2>/dev/null
LOGITS_DUMP_NODES=/tmp/n_cpu.bin $D $G /tmp/cpu.bin 0 40 2 $kv 1 8 >/dev/null 2>&1
LOGITS_DUMP_NODES=/tmp/n_gpu.bin $D $G /tmp/gpu.bin 99 40 2 $kv 1 4 >/dev/null 2>&1
python $HERE/tools/compare_logits.py /tmp/cpu.bin /tmp/gpu.bin | cut -c1-200
python $HERE/tools/compare_nodes.py /tmp/n_cpu.bin /tmp/n_gpu.bin 0 > /tmp/nodes.txt; head -${3:-70} /tmp/nodes.txt
