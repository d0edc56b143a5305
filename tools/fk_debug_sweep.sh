#!/bin/bash
# Tuning experiment (GPU box): time the phase kernel with parts switched off.  AOG_FK_DEBUG bits: 1 no phi stores
# (tensor path only), 2 no phase prefetch, 4 no TMEM load.  Results are wrong by construction.
for d in ${@:-0 2 4 6}; do
  echo -n "AOG_FK_DEBUG=$d  "
  AOG_FK_DEBUG=$d timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-mft-arm 2>&1 | tail -1 | \
    python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("step ms", round(d["ms_per_step"],3), d["roofline"]["ms_per_launch"])'
done
