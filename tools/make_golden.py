#!/usr/bin/env python
"""Generate tests/golden/*.npz: seeded input/output trajectories of the AO-v0 step path.

hcipy==0.5.1 is not importable in this image, so these vectors come from the CPU oracle
(oracle/ao_oracle.py, the NumPy FP64 restatement of AO_env.py + the hcipy calls it makes), NOT from the reference
itself: they pin the oracle and the CUDA path against regressions and against each other, they do not pin either
against real hcipy ("parity unpinned").  On a machine WITH hcipy, tools/export_hcipy_tables.py
is the script to run instead.

The autoregressive extrusion matrices A, B (hcipy `InfiniteAtmosphericLayer`) come out of a Tikhonov inverse
and an SVD of near-singular covariances: their low bits differ between LAPACK builds / CPUs, and that shows up at
the 1e-5 level in the extruded screens.  The dynamic cases therefore use ONE committed copy of the tables, rounded to
float32 (tests/golden/ar_tables.npz), injected into the oracle and into the CUDA path alike.

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
        episodes=1, env_seed=21),
}


def screen(seed, r0):
    g = O.hcipy_make_pupil_grid(240, 0.5)
    cn2 = O.hcipy_Cn_squared_from_fried_parameter(r0, 2.2e-6)
    return O.von_karman_screen(g, cn2, 10.0, np.random.default_rng(seed)).astype(np.float32)


AR_TABLES = os.path.join(ROOT, 'tests', 'golden', 'ar_tables.npz')
AR_SEED = 21


def make_ar_tables():
    """A, B, stencil of the 240 x 240 pupil grid (L0 = 10 m, unit Cn^2), rounded to float32."""
    env = O.OracleAOEnv(atm_type='dynamic', atm_vel=5, act_type='zernike', act_dim=3, seed=AR_SEED,
                        initial_screen=np.zeros(57600))
    lay = env.layer
    np.savez_compressed(AR_TABLES, ar_A=lay.A_horizontal.astype(np.float32), ar_B=lay.B_horizontal.astype(np.float32),
                        ar_stencil=np.flatnonzero(lay.stencil_left).astype(np.int32))


def load_ar_tables():
    z = np.load(AR_TABLES)
    return dict(ar_stencil=z['ar_stencil'], ar_A=z['ar_A'].astype(np.float64), ar_B=z['ar_B'].astype(np.float64))


def oracle_env(case, scr):
    """The oracle env of a case; dynamic cases get the committed AR tables."""
    kw = case['kw']
    env = O.OracleAOEnv(**kw, initial_screen=np.asarray(scr, dtype=np.float64), seed=case.get('env_seed', 0))
    if kw['atm_type'] == 'dynamic':
        t = load_ar_tables()
        st = np.zeros(env.layer.stencil_left.size, dtype=bool)
        st[t['ar_stencil']] = True
        env.layer.stencil_left, env.layer.A_horizontal, env.layer.B_horizontal = st, t['ar_A'], t['ar_B']
    return env


def run_case(case, scr):
    """-> dict of arrays.  Shared by this script and tests/test_golden.py (oracle side)."""
    kw = case['kw']
    env = oracle_env(case, scr)
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
    make_ar_tables()
    for name, case in CASES.items():
        scr = screen(case['screen_seed'], case['kw'].get('atm_fried', 0.15))
        out = run_case(case, scr)
        np.savez_compressed(os.path.join(outdir, name + '.npz'), screen=scr, case=np.array(json.dumps(case)), **out)
        print(name, {k: v.shape for k, v in out.items()}, 'reward', out['reward'])


if __name__ == '__main__':
    main()
