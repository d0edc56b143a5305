"""Tiny run of every round-2 kernel for compute-sanitizer (memcheck): one fused env with the Shack-Hartmann integrator on
a dynamic atmosphere (tensor-core SH step, extrusions, fused optics, finalize), a 128^2-pupil fused env, a reseeded reset.

    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_optics_gym_b200 import AOEnv, AOVecEnv  # noqa: E402

kw = dict(atm_type='dynamic', atm_vel=20, atm_fried=0.10, act_dim=64, obs_dim=2, rew_type='strehl_ratio',
          timesteps_per_episode=3, SH_operation=True, seed=1)
env = AOVecEnv(3, **kw, precision='fused')
env.reset(seed=5)
for _ in range(2):
    a, _ = env.SH_step()
    env.step(a)
env.close()
e = AOEnv(atm_type='quasi_static', act_type='zernike', act_dim=6, obs_dim=5, rew_type='smf_ssim', timesteps_per_episode=4,
          precision='fused', seed=2, num_pupil_pixels=128, num_focal_pixels_fiber=64)
e.reset()
for _ in range(5):
    e.step(np.random.default_rng(0).normal(0, 1, 6))
e.close()
print('sanitize_small: done')
