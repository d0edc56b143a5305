O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu -k "config2 or other_detector_sizes or other_grid_sizes_fused or zero_action" 2>&1 | tail -3 > $O/r3_e_tests.log
cat $O/r3_e_tests.log
B="--no-cpu-baseline --no-mft-arm --no-workloads --steps 3 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r3_e_z6.csv python bench.py --workload zernike6_smf_ssim $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r3_e_head.csv python bench.py $B > /dev/null 2>&1
for f in z6 head; do python tools/summarize_ncu.py launches $O/r3_e_$f.csv $O/r3_e_$f.md; grep -E "finalize|phase_tc|actuators" $O/r3_e_$f.md; done
B="--no-cpu-baseline --no-mft-arm --no-workloads --steps 60 --warmup 5"
python bench.py --workload zernike6_smf_ssim $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('z6', d['value'], d['ms_per_step'])"
python bench.py $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('head', d['value'], d['ms_per_step'])"
