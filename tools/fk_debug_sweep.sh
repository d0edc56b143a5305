for d in 0 1 2 4 8 3 7 15; do
  echo -n "dbg=$d: "; AOG_FK_DEBUG=$d python bench.py --steps 10 --warmup 3 --envs 4096 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_launch'])"
done
echo -n "ssim n5: "; python bench.py --steps 10 --warmup 3 --envs 4096 --no-cpu-baseline --workload zernike6_smf_ssim 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_launch'])"
