#!/bin/bash
# Tuning experiment (GPU box): time the MFT kernels with parts switched off.  AOG_TC_DEBUG bits: 1 no MMA,
# 2 no TMA (stage 2), 4 no epilogue work, 32 no field arithmetic, 64 no phase tiles, 128 no twiddle tiles (stage 1).
# Results are wrong by construction; only the per-kernel milliseconds matter.
for v in ${@:-0 1 4 32 64 128 5 33 36 37 96 101 133 229}; do
  echo -n "AOG_TC_DEBUG=$v  "
  AOG_TC_DEBUG=$v timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("step ms", round(d["ms_per_step"],3), d["roofline"].get("kernel_ms"))'
done
