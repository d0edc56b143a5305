#!/usr/bin/env python
"""Export set-up tables and golden trajectories from the REAL reference (hcipy==0.5.1 + gym_AO).

Run this on a machine where the reference imports (``pip install hcipy==0.5.1 gymnasium scikit-image`` and the
reference checkout on PYTHONPATH); it is NOT runnable in the build image (no hcipy, no network).  It writes

  hcipy_tables_<config>.npz   every table ``AOEnv(tables=...)`` accepts, taken from the reference env's own hcipy
                              objects (names = adaptive_optics_gym_b200._lib.TABLE_IDS plus the scalars)
  hcipy_golden_<config>.npz   screens, actions, extrusion noise, observations, rewards, power for a few episodes

which turn "parity unpinned"  into pinned: commit the .npz under tests/golden/ and point
tests/test_golden.py at them.  Noise is made reproducible by seeding NumPy's global RNG (the only RNG hcipy uses,
SURVEY.md section 5) and recording the screen before every step.

    python tools/export_hcipy_tables.py --out tests/golden --config config1
"""
import argparse
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


def export(name, kw, out, episodes=2, seed=0):
    import gymnasium as gym
    import gym_AO  # noqa: F401  (the reference package: registers AO-v0)
    np.random.seed(seed)
    env = gym.make('AO-v0', **kw).unwrapped
    grid = env.wf_wfs_fiber.electric_field.grid
    Np = int(round(np.sqrt(grid.size)))
    t = {}
    t['aperture'] = np.asarray(env.wf_wfs_fiber.electric_field != 0, dtype=np.float64)
    t['dm_modes'] = np.asarray(env.deformable_mirror.influence_functions.transformation_matrix.T.todense()
                               if hasattr(env.deformable_mirror.influence_functions.transformation_matrix, 'todense')
                               else env.deformable_mirror.influence_functions.transformation_matrix.T)
    m = t['dm_modes']
    t['dm_gram'] = m @ m.T / m.shape[1] - np.outer(m.mean(1), m.mean(1))
    for key, prop in (('fib', env.propagator_fiber), ('obs', env.propagator_fiber_subsample)):
        prop(env.wf_wfs_fiber)                                    # builds and caches the MatrixFourierTransform
        ft = prop.fourier_transform if hasattr(prop, 'fourier_transform') else prop._fourier_transform
        t[f'mft_{key}_1'], t[f'mft_{key}_2'] = np.asarray(ft.M1), np.asarray(ft.M2)
    lay = env.layer
    if kw['atm_type'] == 'dynamic':
        t['ar_stencil'] = np.flatnonzero(lay.new_col_stencil if hasattr(lay, 'new_col_stencil') else lay.stencil_left)
        t['ar_A'], t['ar_B'] = np.asarray(lay.A_horizontal), np.asarray(lay.B_horizontal)
    np.savez_compressed(os.path.join(out, f'hcipy_tables_{name}.npz'), **t)

    rec = dict(screens=[], actions=[], obs=[], reward=[], power=[], done=[])
    rng = np.random.default_rng(seed + 1)
    for ep in range(episodes):
        env.reset()
        for step in range(kw['timesteps_per_episode']):
            rec['screens'].append(np.asarray(lay.phase_for(1.0), dtype=np.float64))     # achromatic screen S
            a = env.SH_step()[0] if kw.get('SH_operation') else rng.uniform(-1, 1, kw['act_dim']).astype(np.float32)
            o, r, d, _, info = env.step(a)
            rec['actions'].append(np.asarray(a, dtype=np.float64))
            rec['obs'].append(np.asarray(env.wf_wfs_after_foc_subsample.power, dtype=np.float64))
            rec['reward'].append(float(r))
            rec['power'].append(float(info['power']))
            rec['done'].append(bool(d))
    np.savez_compressed(os.path.join(out, f'hcipy_golden_{name}.npz'), kw=np.array(repr(kw)),
                        **{k: np.array(v) for k, v in rec.items()})
    print(name, 'exported', {k: np.array(v).shape for k, v in rec.items()})


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='tests/golden')
    ap.add_argument('--config', default='all', choices=['all'] + list(CONFIGS))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for name, kw in CONFIGS.items():
        if args.config in ('all', name):
            export(name, kw, args.out)
