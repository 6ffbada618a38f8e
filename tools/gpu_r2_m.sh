#!/bin/bash
mkdir -p gpurun_out
for m in bs32 pp512; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2m_$m.csv python tools/batched_prof.py $m 4 1 > gpurun_out/r2m_$m.log 2>&1
  python tools/summarize_launches.py gpurun_out/r2m_$m.csv > gpurun_out/r2m_${m}_summary.txt 2>&1
  cat gpurun_out/r2m_${m}_summary.txt
done
python tools/batched_prof.py bs32 32 8; python tools/batched_prof.py pp512 32 4
