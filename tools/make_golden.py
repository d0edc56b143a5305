#!/usr/bin/env python
"""Generate tests/golden/*.npz: seeded input/output trajectories of the AO-v0 step path.

hcipy==0.5.1 is not importable in this image, so these vectors come from the CPU oracle
(oracle/ao_oracle.py, the NumPy FP64 restatement of AO_env.py + the hcipy calls it makes), NOT from the reference
itself: they pin the oracle and the CUDA path against regressions and against each other, they do not pin either
against real hcipy ("parity unpinned").  On a machine WITH hcipy, tools/export_hcipy_tables.py
is the script to run instead.

    python tools/make_golden.py            # rewrites tests/golden/*.npz
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ao_oracle as O  # noqa: E402

CASES = {
    # BASELINE.json configs[0]
    'config1_quasi_static_64act_strehl': dict(
        kw=dict(atm_type='quasi_static', atm_vel=0, atm_fried=0.20, act_type='num_actuators', act_dim=64, obs_dim=2,
                rew_type='strehl_ratio', timesteps_per_episode=4), screen_seed=100, action='uniform', episodes=2),
    # BASELINE.json configs[1] (one env of the batch)
    'config2_zernike6_smf_ssim': dict(
        kw=dict(atm_type='quasi_static', act_type='zernike', act_dim=6, obs_dim=5, rew_type='smf_ssim',
                timesteps_per_episode=4, flat_mirror_start_per_episode=False), screen_seed=101, action='normal',
        episodes=2),
    # BASELINE.json configs[2] dynamic variant (v = 5 m/s: 2-3 extrusions per step), obs 5x5
    'config3_dynamic_v5_obs5': dict(
        kw=dict(atm_type='dynamic', atm_vel=5, atm_fried=0.15, act_type='num_actuators', act_dim=64, obs_dim=5,
                rew_type='strehl_ratio', timesteps_per_episode=4), screen_seed=102, action='uniform', episodes=1,
        env_seed=21),
    # BASELINE.json configs[3]: dynamic 20 m/s, r0 = 0.10, Shack-Hartmann closed loop (noise-free camera)
    'config4_dynamic_v20_shack_hartmann': dict(
        kw=dict(atm_type='dynamic', atm_vel=20, atm_fried=0.10, act_type='zernike', act_dim=10, obs_dim=2,
                rew_type='strehl_ratio', timesteps_per_episode=3, SH_operation=True), screen_seed=103, action='sh',
        episodes=1, env_seed=22),
}


def screen(seed, r0):
    g = O.hcipy_make_pupil_grid(240, 0.5)
    cn2 = O.hcipy_Cn_squared_from_fried_parameter(r0, 2.2e-6)
    return O.von_karman_screen(g, cn2, 10.0, np.random.default_rng(seed)).astype(np.float32)


def run_case(case, scr):
    """-> dict of arrays.  Shared by this script and tests/test_golden.py (oracle side)."""
    kw = case['kw']
    env = O.OracleAOEnv(**kw, initial_screen=scr.astype(np.float64), seed=case.get('env_seed', 0))
    rng = np.random.default_rng(case['screen_seed'] + 1000)
    K, T = kw['act_dim'], kw['timesteps_per_episode']
    rec = {k: [] for k in ('actions', 'noise', 'reset_obs', 'obs', 'obs_f16', 'reward', 'power', 'aux', 'done')}
    for ep in range(case['episodes']):
        env.reset()
        rec['reset_obs'].append(env.last_obs_f64.copy())
        for t in range(T):
            if case['action'] == 'uniform':
                a = rng.uniform(-1, 1, K).astype(np.float32).astype(np.float64)
            elif case['action'] == 'normal':
                a = rng.normal(0, np.sqrt(0.5), K)
            else:
                a = np.array(env.SH_step(poisson=False)[0], dtype=np.float64)
            n_ext = env.num_extrusions_for_next_step()
            nz = rng.standard_normal((n_ext, 240))
            o16, r, d, _, info = env.step(a, extrusion_noise=nz if n_ext else None)
            rec['actions'].append(a)
            rec['noise'].append(nz.ravel())
            rec['obs'].append(env.last_obs_f64.copy())
            rec['obs_f16'].append(o16.view(np.uint16).copy())
            rec['reward'].append(r)
            rec['power'].append(info['power'])
            rec['aux'].append(env.last_strehl if kw.get('rew_type', 'strehl_ratio') == 'strehl_ratio' else env.last_ssim)
            rec['done'].append(d)
    out = {k: np.array(v) for k, v in rec.items() if k != 'noise'}
    out['noise_counts'] = np.array([x.size // 240 for x in rec['noise']])
    out['noise'] = np.concatenate(rec['noise']) if rec['noise'] else np.zeros(0)
    out['final_timestep'] = np.array(env.timestep)
    out['final_episode_no'] = np.array(env.episode_no)
    return out


def main():
    outdir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    for name, case in CASES.items():
        scr = screen(case['screen_seed'], case['kw'].get('atm_fried', 0.15))
        out = run_case(case, scr)
        np.savez_compressed(os.path.join(outdir, name + '.npz'), screen=scr, case=np.array(json.dumps(case)), **out)
        print(name, {k: v.shape for k, v in out.items()}, 'reward', out['reward'])


if __name__ == '__main__':
    main()
