#!/bin/bash
# BASELINE.json configs[4], batch-size axis at the reference geometry (pupil 240^2, focal 128^2): env-steps/s of the
# headline workload against the number of environments on one B200.  Writes one line per size.
for n in ${@:-1 16 128 1024 4096 16384 65536}; do
  python bench.py --envs $n --steps 60 --warmup 5 --no-cpu-baseline --no-mft-arm --no-workloads 2>&1 | tail -1 | \
    python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({"envs": d["config"]["envs_per_gpu"], "ms_per_step": round(d["ms_per_step"],4), "env_steps_per_s": round(d["value"]), "e2e_env_steps_per_s": round(d["e2e"]["value"]), "kernel_ms": round(d["roofline"]["ms_per_launch"],4)}))'
done
