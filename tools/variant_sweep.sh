#!/bin/bash
# Tuning experiment (GPU box): bench every tuning build under variants/ (AOG_LIB selects the library).
for lib in "" variants/*.so; do
  echo -n "${lib:-default}  "
  AOG_LIB=${lib:+$PWD/$lib} timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | \
    python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("step ms", round(d["ms_per_step"],3), d["roofline"].get("kernel_ms"))'
done
