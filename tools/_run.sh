O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu -k "dynamic_extrusion or direct_extrusion or seeded" 2>&1 | tail -4 > $O/r3_g_tests.log
cat $O/r3_g_tests.log
B="--no-cpu-baseline --no-mft-arm --no-workloads --steps 40 --warmup 5"
for w in dynamic_v20 dynamic_v20_sh; do
python bench.py --workload $w $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'])"
done
B="--no-cpu-baseline --no-mft-arm --no-workloads --steps 3 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r3_g_v20.csv python bench.py --workload dynamic_v20 $B > /dev/null 2>&1
python tools/summarize_ncu.py launches $O/r3_g_v20.csv $O/r3_g_v20.md; grep -E "k_ar" $O/r3_g_v20.md
