#!/bin/bash
cd "$(dirname "$0")/.."
for shp in "4096 4096" "14336 4096" "4096 14336" "128256 4096"; do echo "== prof $shp"; timeout 120 python tools/gemv_prof.py $shp 2>&1 | tail -9; done
