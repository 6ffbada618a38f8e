#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -s ) > gpurun_out/r2g_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2g_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2g_smoke.log
grep -h "passed\|failed\|rc=\|Error\|real\|smoke step\|fast mode\|PARITY\|BACKEND_OPS" gpurun_out/r2g_all.log gpurun_out/r2g_smoke.log | cut -c1-300 | tail -80
