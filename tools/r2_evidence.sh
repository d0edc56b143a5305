#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): GPU tests, the driver's bench invocation, ncu launch lists and
# --set full captures of every kernel quoted in DESIGN.md.  Everything lands in gpurun_out/ (copied to profiles/).
# TAG selects the file prefix (r2_zz = the final state of round 2).
O=gpurun_out
T=${TAG:-r2_zz}
mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/${T}_gputests.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_1gpu.json 2> $O/${T}_bench.err
B="--no-cpu-baseline --no-mft-arm --no-workloads"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_fused_launches.csv python bench.py --steps 3 --warmup 3 $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_sh_launches.csv python bench.py --workload dynamic_v20_sh --steps 3 --warmup 3 $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_semi_launches.csv python bench.py --workload semi_dynamic_64act --envs 4096 --steps 21 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_dm_phase_tc|k_finalize_tcw|k_actuators_pack" -s 6 -c 3 -o $O/${T}_fused_full python bench.py --steps 2 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_ar_step|k_ar_noise" -s 40 -c 2 -o $O/${T}_ar_full python bench.py --workload dynamic_v20 --steps 2 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_sh_|k_dm_phase_tc<0, 1, 3" -s 10 -c 5 -o $O/${T}_sh_full python bench.py --workload dynamic_v20_sh --steps 2 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_finalize_tcw" -s 4 -c 1 -o $O/${T}_fin5_full python bench.py --workload zernike6_smf_ssim --steps 2 --warmup 3 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_scr_fft" -s 2 -c 2 -o $O/${T}_fft_full python bench.py --workload semi_dynamic_64act --envs 4096 --steps 2 --warmup 1 $B > /dev/null 2>&1
AOG_NO_GRAPH=1 ncu --set full --clock-control none -k regex:"k_small_fused" -s 30 -c 1 -o $O/${T}_small_full python tools/single_env_latency.py fused > /dev/null 2>&1
for f in fused ar sh fin5 fft small; do ncu -i $O/${T}_${f}_full.ncu-rep --page raw --csv > $O/${T}_${f}_full.raw.csv 2>/dev/null; rm -f $O/${T}_${f}_full.ncu-rep; done
python tools/single_env_latency.py > $O/${T}_single_env.log 2>&1
python tools/rollout_throughput.py > $O/${T}_rollout.log 2>&1
ls -la $O | tail -30
