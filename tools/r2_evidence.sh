#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): GPU tests, the driver's bench invocation, ncu launch lists and
# --set full captures of every kernel quoted in DESIGN.md.  Everything lands in gpurun_out/ (copied to profiles/).
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/r2_z_gputests.log
python bench.py --steps 20 --warmup 5 > $O/r2_z_bench.json 2> $O/r2_z_bench.err
B="--no-cpu-baseline --no-mft-arm --no-workloads"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_z_fused_launches.csv python bench.py --steps 3 --warmup 3 $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_z_sh_launches.csv python bench.py --workload dynamic_v20_sh --steps 3 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_dm_phase_tc|k_finalize_tc|k_actuators_pack" -s 6 -c 3 -o $O/r2_z_fused_full python bench.py --steps 2 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_ar_step|k_ar_gather" -s 40 -c 2 -o $O/r2_z_ar_full python bench.py --workload dynamic_v20 --steps 2 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_sh_|k_dm_phase_tc<0, 1, 3" -s 10 -c 5 -o $O/r2_z_sh_full python bench.py --workload dynamic_v20_sh --steps 2 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_finalize_tc" -s 4 -c 1 -o $O/r2_z_fin5_full python bench.py --workload zernike6_smf_ssim --steps 2 --warmup 3 $B > /dev/null 2>&1
for f in fused ar sh fin5; do ncu -i $O/r2_z_${f}_full.ncu-rep --page raw --csv > $O/r2_z_${f}_full.raw.csv 2>/dev/null; rm -f $O/r2_z_${f}_full.ncu-rep; done
python tools/single_env_latency.py > $O/r2_z_single_env.log 2>&1
python tools/rollout_throughput.py > $O/r2_z_rollout.log 2>&1
ls -la $O | tail -20
