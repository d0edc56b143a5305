#!/usr/bin/env python
"""Export set-up tables and golden trajectories from the REAL reference (hcipy==0.5.1 + gym_AO).

Run this on a machine where the reference imports (``pip install hcipy==0.5.1 gymnasium scikit-image`` and the
reference checkout on PYTHONPATH); it is NOT runnable in the build image (no hcipy, no network).  It writes

  hcipy_tables_<config>.npz   the set-up tables whose restatement is uncertain (SURVEY App. A [VERIFY]): the DM mode
                              matrix (read through the PUBLIC hcipy API: one ``dm.actuators = e_k; dm.surface`` probe
                              per mode), the AR extrusion stencil / A / B when the layer exposes them, the SH
                              reconstruction matrix and reference slopes; names as ``AOEnv(tables=...)`` takes them
  hcipy_golden_<config>.npz   per step: the achromatic screen the optics saw, the action, observation (before the
                              float16 cast), reward, info["power"], done

Dropping both files into ``tests/golden/`` (or ``$AOG_HCIPY_GOLDEN``) turns "parity unpinned" into pinned without a
code change: ``tests/test_golden.py`` replays every ``hcipy_golden_*.npz`` through the CPU oracle (1e-9) and through
the CUDA path (FP64 1e-7, tensor / fused 1e-5).  The replay feeds the recorded screen before every step, so it does
not depend on reproducing hcipy's random draws.

``collect`` only touches attributes that the oracle (an hcipy-shaped restatement) has too, which is how the CPU test
``test_hcipy_export_dry_run`` exercises this file end to end without hcipy.

    python tools/export_hcipy_tables.py --out tests/golden --config config1
"""
import argparse
import json
import os

import numpy as np

CONFIGS = {
    'config1': dict(atm_type='quasi_static', atm_vel=0, atm_fried=0.20, act_type='num_actuators', act_dim=64,
                    obs_dim=2, rew_type='strehl_ratio', timesteps_per_episode=30),
    'config2': dict(atm_type='quasi_static', act_type='zernike', act_dim=6, obs_dim=5, rew_type='smf_ssim',
                    timesteps_per_episode=20),
    'config3': dict(atm_type='dynamic', atm_vel=5, atm_fried=0.15, act_type='num_actuators', act_dim=64, obs_dim=5,
                    rew_type='strehl_ratio', timesteps_per_episode=20),
    'config4': dict(atm_type='dynamic', atm_vel=20, atm_fried=0.10, act_type='num_actuators', act_dim=64, obs_dim=2,
                    rew_type='strehl_ratio', timesteps_per_episode=20, SH_operation=True),
}


def _screen(env):
    """Achromatic screen S (phase = S / lambda) of the env's atmospheric layer: hcipy ``phase_for(1)``."""
    return np.array(env.layer.phase_for(1.0), dtype=np.float64).ravel()


def collect(env, kw, episodes=2, steps=None, seed=0):
    """(tables, golden) from a reference-shaped env (``gym.make('AO-v0', **kw).unwrapped``)."""
    t = {}
    dm = env.deformable_mirror
    K = int(kw.get('act_dim', 64))
    cols = []
    for k in range(K):                      # public API probe: column k of the influence matrix
        a = np.zeros(K)
        a[k] = 1.0
        dm.actuators = a
        cols.append(np.array(dm.surface, dtype=np.float64).ravel())
    dm.flatten()
    m = np.stack(cols)                      # [K, P]
    t['dm_modes'] = m
    t['dm_gram'] = m @ m.T / m.shape[1] - np.outer(m.mean(1), m.mean(1))
    lay = env.layer
    if kw.get('atm_type') == 'dynamic':
        for names, key in ((('stencil_left', 'new_col_stencil'), 'ar_stencil'), (('A_horizontal',), 'ar_A'),
                           (('B_horizontal',), 'ar_B')):
            for n in names:
                if hasattr(lay, n):
                    v = np.asarray(getattr(lay, n))
                    t[key] = np.flatnonzero(v).astype(np.int32) if key == 'ar_stencil' else v.astype(np.float64)
                    break
    if kw.get('SH_operation'):
        t['sh_recon'] = np.asarray(env.reconstruction_matrix, dtype=np.float64)
        t['sh_slopes_ref'] = np.asarray(env.slopes_ref, dtype=np.float64)

    rec = dict(reset_screens=[], reset_obs=[], screens=[], actions=[], obs=[], reward=[], power=[], done=[])
    rng = np.random.default_rng(seed + 1)
    T = int(steps or kw['timesteps_per_episode'])
    for ep in range(episodes):
        env.reset()
        rec['reset_screens'].append(_screen(env))
        rec['reset_obs'].append(np.array(env.wf_wfs_after_foc_subsample.power, dtype=np.float64))
        for step in range(T):
            a = env.SH_step()[0] if kw.get('SH_operation') else rng.uniform(-1, 1, K).astype(np.float32)
            a = np.array(a, dtype=np.float64)
            o, r, d, _, info = env.step(a)
            rec['screens'].append(_screen(env))            # after the layer moved: what this step's optics saw
            rec['actions'].append(a)
            rec['obs'].append(np.array(env.wf_wfs_after_foc_subsample.power, dtype=np.float64))
            rec['reward'].append(float(r))
            rec['power'].append(float(info['power']))
            rec['done'].append(bool(d))
            if d:
                break
    g = {k: np.array(v) for k, v in rec.items()}
    g['kw'] = np.array(json.dumps(kw))
    g['episodes'] = np.array(episodes)
    return t, g


def write(name, tables, golden, out):
    np.savez_compressed(os.path.join(out, f'hcipy_tables_{name}.npz'), **tables)
    np.savez_compressed(os.path.join(out, f'hcipy_golden_{name}.npz'), **golden)


def export(name, kw, out, episodes=2, seed=0):
    import gymnasium as gym
    import gym_AO  # noqa: F401  (the reference package: registers AO-v0)
    np.random.seed(seed)                    # hcipy draws from NumPy's global generator
    env = gym.make('AO-v0', **kw).unwrapped
    tables, golden = collect(env, kw, episodes=episodes, seed=seed)
    write(name, tables, golden, out)
    print(name, 'exported', {k: v.shape for k, v in golden.items()})


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='tests/golden')
    ap.add_argument('--config', default='all', choices=['all'] + list(CONFIGS))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for name, kw in CONFIGS.items():
        if args.config in ('all', name):
            export(name, kw, args.out)
