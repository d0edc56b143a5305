#!/usr/bin/env python
"""bench.py -- env-steps/sec of the AO-v0 step path on N B200s (one JSON line on rank 0).

A "step" is one lock-stepped ``AOVecEnv.step`` over the rank's batch of environments (B env-steps),
including the ``reset`` at every episode boundary.  Workload (BASELINE.json: "64-act quasi-static"):
quasi_static, r0 = 0.20 m, 64 disk-harmonic actuators, 2x2 photodetector, strehl_ratio reward,
30 steps/episode (configs[0] physics), batched at ``--envs`` environments per GPU (weak scaling).

  value   device-resident throughput: actions already in HBM, outputs left in HBM
  e2e     same metric through the public API with HOST buffers: pinned host actions copied
          host->device every step, obs/reward/power copied device->host every step
  roofline  the dominant kernel sequence (the matrix-Fourier-transform GEMMs) timed live with
          CUDA events on the launch stream inside the library
  cpu_baseline  the CPU oracle (NumPy FP64, unabridged reference op sequence) on this box's cores

``--impl reference`` times the reference-side CPU implementation (the oracle port; hcipy itself is
not installable here) on the host cores for the same metric / config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json metric config: 64-actuator quasi-static (configs[0] physics), batched
    'quasi_static_64act': dict(atm_type='quasi_static', atm_vel=0, atm_fried=0.20, act_type='num_actuators',
                               act_dim=64, obs_dim=2, rew_type='strehl_ratio', timesteps_per_episode=30,
                               flat_mirror_start_per_episode=True),
    # BASELINE.json configs[1]
    'zernike6_smf_ssim': dict(atm_type='quasi_static', atm_vel=0, atm_fried=0.15, act_type='zernike', act_dim=6,
                              obs_dim=5, rew_type='smf_ssim', timesteps_per_episode=20,
                              flat_mirror_start_per_episode=True),
    # BASELINE.json configs[2] as the reference runs it: semi_dynamic coerces the velocity to 0 (AO_env.py:200-203) and
    # draws a fresh von-Karman screen at every reset (:76-77)
    'semi_dynamic_64act': dict(atm_type='semi_dynamic', atm_vel=0, atm_fried=0.15, act_type='num_actuators', act_dim=64,
                               obs_dim=5, rew_type='strehl_ratio', timesteps_per_episode=20,
                               flat_mirror_start_per_episode=True),
    # ... and with the phase-screen evolution that config names (dynamic, 5 m/s: 2.4 extrusions per step)
    'dynamic_v5': dict(atm_type='dynamic', atm_vel=5, atm_fried=0.15, act_type='num_actuators', act_dim=64,
                       obs_dim=5, rew_type='strehl_ratio', timesteps_per_episode=20,
                       flat_mirror_start_per_episode=True),
    # BASELINE.json configs[3] physics without the SH loop
    'dynamic_v20': dict(atm_type='dynamic', atm_vel=20, atm_fried=0.10, act_type='num_actuators', act_dim=64,
                        obs_dim=2, rew_type='strehl_ratio', timesteps_per_episode=20,
                        flat_mirror_start_per_episode=True),
    # BASELINE.json configs[3]: the Shack-Hartmann integrator drives the mirror (SH_step + step every step)
    'dynamic_v20_sh': dict(atm_type='dynamic', atm_vel=20, atm_fried=0.10, act_type='num_actuators', act_dim=64,
                           obs_dim=2, rew_type='strehl_ratio', timesteps_per_episode=20,
                           flat_mirror_start_per_episode=True, SH_operation=True),
}

MFT_FLOP_PER_ENV = lambda Np, Nf: 8.0 * (Nf * Np * Np + Nf * Np * Nf)   # 90.44 MFLOP @ 240/128


def describe(name, envs, n_gpus):
    w = WORKLOADS[name]
    return (f"{w['atm_type']} r0={w['atm_fried']} {w['act_type']} K={w['act_dim']} obs {w['obs_dim']}x{w['obs_dim']} "
            f"{w['rew_type']} {w['timesteps_per_episode']} steps/episode, {envs} envs/GPU x {n_gpus} GPU")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows = []
        self.stop_flag = threading.Event()
        self.index = index

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in self.rows)]
        try:
            mx = int(float(self.rows[0][1]))
        except ValueError:
            mx = None
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': reasons,
                'samples': len(self.rows)}


def cpu_oracle_rate(name, budget_s, threads_note=True):
    """env-steps/s of the CPU oracle (single env, the reference's own shape) for ~budget_s."""
    import numpy as np
    from oracle.ao_oracle import OracleAOEnv
    w = WORKLOADS[name]
    env = OracleAOEnv(**w, seed=0)
    rng = np.random.default_rng(1)
    K = w['act_dim']
    T = w['timesteps_per_episode']
    env.reset()
    for _ in range(3):
        env.step(rng.uniform(-1, 1, K).astype(np.float32))
    n = 0
    t0 = time.perf_counter()
    while True:
        env.reset()
        for _ in range(T):
            env.step(rng.uniform(-1, 1, K).astype(np.float32))
            n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, n


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # K "steps", each a bounded sample of the workload: one full episode of the single CPU env.
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is to use every host thread it can
    # (rank 0 alone runs), so the BLAS pool is sized before NumPy loads.
    threads = os.environ.get('AOG_REF_THREADS', str(os.cpu_count()))
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[k] = threads
    import numpy as np
    from oracle.ao_oracle import OracleAOEnv
    w = WORKLOADS[args.workload]
    env = OracleAOEnv(**w, seed=0)
    rng = np.random.default_rng(1)
    K, T = w['act_dim'], w['timesteps_per_episode']

    def episode():
        env.reset()
        for _ in range(T):
            env.step(rng.uniform(-1, 1, K).astype(np.float32))

    for _ in range(max(args.warmup, 1)):
        episode()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        episode()
    dt = time.perf_counter() - t0
    v = args.steps * T / dt
    cores = int(threads)
    line = {
        'impl': 'reference', 'metric': 'env-steps/sec', 'value': v, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': describe(args.workload, args.envs, args.gpus),
                   'note': 'reference arm = single CPU env (the reference has no batching); one bench step = one '
                           f'{T}-step episode incl. reset'},
        'cpu_baseline': {'value': v, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{args.steps} episodes x {T} steps of one env, NumPy/OpenBLAS with {cores} threads '
                                   '(hcipy==0.5.1 is not installable here: oracle port, unabridged op sequence)'},
        'e2e': {'value': v, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs', type=int, default=4096, help='environments per GPU')
    ap.add_argument('--workload', default='quasi_static_64act', choices=list(WORKLOADS))
    ap.add_argument('--precision', default=os.environ.get('AOG_PRECISION', 'auto'), choices=['auto', 'f64', 'tensor', 'fused'])
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-mft-arm', action='store_true', help='skip the secondary timing of the tensor-core MFT path')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from adaptive_optics_gym_b200 import AOVecEnv
    from adaptive_optics_gym_b200._lib import AogError

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    w = WORKLOADS[args.workload]
    B, K, n2, T = args.envs, w['act_dim'], w['obs_dim'] ** 2, w['timesteps_per_episode']
    precision = 'fused' if args.precision == 'auto' else args.precision
    env = AOVecEnv(B, **w, device=local, seed=1234, precision=precision, env_id_base=rank * B)
    Np, Nf = env.num_pupil_pixels, env.num_focal_pixels_fiber

    # action pool: device-resident for `value`, pinned host for `e2e`
    g = torch.Generator(device='cpu').manual_seed(7 + rank)
    pool = 8
    act_host = [torch.empty((B, K), dtype=torch.float32).uniform_(-1, 1, generator=g).pin_memory() for _ in range(pool)]
    act_dev = [a.to(dev) for a in act_host]

    state = {'t': 0}

    sh_loop = bool(w.get('SH_operation'))

    def step_device(i):
        if state['t'] % T == 0:
            env.reset()
        a = env.SH_step()[0] if sh_loop else act_dev[i % pool]
        _, _, done, _, _ = env.step(a)
        state['t'] += 1

    def step_e2e(i):
        if state['t'] % T == 0:
            env.reset()
        a = act_host[i % pool].to(dev, non_blocking=True)
        if sh_loop:
            a = env.SH_step()[0]
        env.step(a)
        env.fetch()                     # obs, reward, power -> pinned host memory (one copy) + stream synchronise:
        state['t'] += 1                 # the caller needs the result before acting again

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        state['t'] = 0
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = env._h.launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = env._h.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches = timed(step_device, args.steps, args.warmup)
    value = world * B * args.steps / (ms * 1e-3)

    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    clocks = sampler.summary() if sampler else None      # sampled over both timed regions

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass

    def kernel_times(e, prec):
        """median per-kernel milliseconds of one step (CUDA events on the launch stream inside the library)"""
        e._h.set_timing(True)
        mft_ms, kms = [], []
        for i in range(6):      # no reset inside this loop
            e.step(e.SH_step()[0] if sh_loop else act_dev[i % pool])
            torch.cuda.synchronize()
            mft_ms.append(e._h.last_mft_ms())
            if prec != 'f64':
                kms.append(e._h.last_kernel_ms())
        e._h.set_timing(False)
        mft_ms = sorted(mft_ms[1:])
        km = {k: sorted(d[k] for d in kms[1:])[len(kms[1:]) // 2] for k in kms[0]} if kms else None
        return mft_ms[len(mft_ms) // 2], km

    def traffic_from_profiles(name, last_chunk):
        # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture (profiles/),
        # valid for the chunk size it was taken at
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', name)))
            if tj.get('envs_per_launch') == last_chunk:
                return tj['bytes_per_launch']
        except Exception:
            pass
        return None

    def mft_roofline(e, prec):
        """the matrix-Fourier-transform GEMMs against the tensor (or FP64) pipe"""
        mft, km = kernel_times(e, prec)
        chunk = min(e._h.chunk_size(), B)
        last_chunk = B - (B - 1) // chunk * chunk
        achieved = MFT_FLOP_PER_ENV(Np, Nf) * last_chunk / (mft * 1e-3) / 1e12
        if prec == 'tensor':
            peak = peaks.get('bf16_tflops_sustained', 1400.0)
            peak_note = ('measured cuBLAS bf16 sustained (MEASURED_PEAKS.json)' if peaks else
                         'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)')
            issued = 3.0
        else:
            peak = 37.0
            peak_note = 'nominal B200 FP64 37 TFLOP/s (no measured FP64 peak in MEASURED_PEAKS.json)'
            issued = 1.0
        r = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
             'traffic': traffic_from_profiles('mft_dram_traffic.json', last_chunk) if prec == 'tensor' else None,
             'kernel': ('matrix Fourier transform: k_field_mft1 (field formation + stage-1 product) + k_mft2 (stage-2 '
                        'product + fibre projection)') if prec == 'tensor' else 'MFT stage-1 + stage-2 complex GEMMs (FP64)',
             'ms_per_launch': mft, 'envs_per_launch': last_chunk, 'algorithmic_flop_per_env': MFT_FLOP_PER_ENV(Np, Nf),
             'issued_over_algorithmic': issued, 'peak_source': peak_note}
        if km:
            r['kernel_ms'] = km
        return r

    mft_arm = None
    if precision == 'fused':
        # dominant kernel: k_dm_phase_tc<fused> -- DM-surface GEMM (tcgen05) + phase + every reduction of the step.
        # Algorithmic bytes (fused design): one FP32-sized read of the screen per env-step, 4 P bytes.
        _, km = kernel_times(env, precision)
        chunk = min(env._h.chunk_size(), B)
        last_chunk = B - (B - 1) // chunk * chunk
        alg_bytes = 4.0 * Np * Np
        ach = alg_bytes * last_chunk / (km['field'] * 1e-3) / 1e9
        peak = peaks.get('hbm_gbs', 6500.0)
        roofline = {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': traffic_from_profiles('fused_dram_traffic.json', last_chunk),
                    'kernel': 'k_dm_phase_tc<fused>: tcgen05 DM-surface GEMM, wavefront phase, obs-arm / Strehl / '
                              'back-projected fibre-mode reductions (the whole optics chain of a step)',
                    'ms_per_launch': km['field'], 'envs_per_launch': last_chunk,
                    'algorithmic_bytes_per_env': alg_bytes,
                    'peak_source': 'measured copy bandwidth (MEASURED_PEAKS.json)' if peaks else
                                   'fallback 6.5 TB/s (B200_PROFILING.md)',
                    'note': 'the kernel is issue / SFU bound (4 MUFU per lit pixel), not HBM bound; see DESIGN.md 4.3'}
    else:
        roofline = mft_roofline(env, precision)
    if precision == 'fused' and not args.no_mft_arm and args.workload != 'dynamic_v20_sh':
        # BASELINE.json's metric also asks for the MFT tensor-pipe utilisation: time the path that runs the
        # fibre-arm matrix Fourier transform as tcgen05 GEMMs (precision='tensor') on the same workload
        env_t = AOVecEnv(B, **w, device=local, seed=1234, precision='tensor', env_id_base=rank * B)
        main_env, env = env, env_t
        ms_t, _ = timed(step_device, max(10, args.steps // 10), 3)
        n_t = max(10, args.steps // 10)
        mft_arm = {'precision': 'tensor', 'value': world * B * n_t / (ms_t * 1e-3), 'unit': 'env-steps/s',
                   'ms_per_step': ms_t / n_t, 'steps': n_t, 'roofline': mft_roofline(env_t, 'tensor')}
        env = main_env
        env_t.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, n = cpu_oracle_rate(args.workload, args.cpu_seconds)
        cpu = {'value': v, 'unit': 'env-steps/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': f'{n} steps of one env ({args.cpu_seconds:.0f} s), NumPy/OpenBLAS default threads, '
                         'unabridged reference op sequence'}

    line = {
        'metric': 'env-steps/sec', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None,
        'dtype': {'f64': 'f64', 'tensor': 'f16x3(tcgen05)+f32/f64', 'fused': 'f16x3(tcgen05 DM GEMM)+f32/f64'}[precision],
        'data': 'synthetic',
        'config': {'workload': describe(args.workload, B, world), 'precision': precision, 'envs_per_gpu': B,
                   'l2_policy': (f'inputs larger than L2: {B * Np * Np * 4 / 1e6:.0f} MB of phase-screen tiles read per step, '
                                 f'plus {B * Np * Np * 4 / 1e6:.0f} MB of phase and {B * 128 * 480 * 4 / 1e6:.0f} MB of stage-1 product '
                                 'written and re-read (126 MB L2)') if precision == 'tensor' else
                                (f'inputs larger than L2: {B * Np * Np * 4 / 1e6:.0f} MB of phase-screen tiles read per step '
                                 '(126 MB L2)') if precision == 'fused' else
                                f'inputs larger than L2: {B * Np * Np * 8 / 1e6:.0f} MB of screens read per step',
                   'timing': 'CUDA events on the launch stream, max over ranks'},
        'e2e': {'value': e2e, 'unit': 'env-steps/s', 'h2d_bytes_per_step': world * B * K * 4,
                'd2h_bytes_per_step': world * B * (n2 * 2 + 16), 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'cpu_baseline': cpu,
        'clocks': clocks,
    }
    if mft_arm:
        line['mft_gemm_path'] = mft_arm
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
