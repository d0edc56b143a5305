O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu -k "direct_extrusion or dynamic_extrusion or seeded or generated or golden" 2>&1 | tail -3
B="--no-cpu-baseline --no-mft-arm --no-workloads --steps 40 --warmup 5"
for w in dynamic_v20 dynamic_v5; do
python bench.py --workload $w $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'])"
done
