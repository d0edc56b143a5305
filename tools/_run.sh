O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu -k "synthesis or generated_screens" 2>&1 | tail -3
B="--no-cpu-baseline --no-mft-arm --no-workloads --steps 40 --warmup 5"
python bench.py --workload semi_dynamic_64act --envs 16384 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('semi fft8', d['value'], d['ms_per_step'])"
B="--no-cpu-baseline --no-mft-arm --no-workloads --steps 21 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r3_l_semi.csv python bench.py --workload semi_dynamic_64act --envs 4096 $B > /dev/null 2>&1
python tools/summarize_ncu.py launches $O/r3_l_semi.csv $O/r3_l_semi.md; grep -E "k_scr|dgemm" $O/r3_l_semi.md
