#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 8 --warmup 1 > gpurun_out/r2_bench${N}_reference.json 2>/dev/null
tail -c 500 gpurun_out/r2_bench${N}_reference.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench${N}.json 2> gpurun_out/r2_bench${N}.err
tail -1 gpurun_out/r2_bench${N}.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d.get(k) for k in ['metric','value','ms_per_step','scaling','n_gpus','clocks']}); print('e2e',d['e2e']['value']); print('roofline',d['roofline']['frac']); print('replicas',d.get('replicas',{}).get('value')); print('tp breakdown',d['tp'].get('step_ms_with_kernel_classes_not_launched'), d['tp'].get('logits_identical_on_all_ranks'))"
tail -3 gpurun_out/r2_bench${N}.err
