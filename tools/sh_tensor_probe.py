"""Tensor-core Shack-Hartmann path against the FP64 kernels (B200): camera image, integrator action, closed loop.

    python tools/sh_tensor_probe.py            # prints the error magnitudes the GPU tests assert on
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from adaptive_optics_gym_b200 import AOEnv, AOVecEnv
    from oracle.ao_oracle import hcipy_make_pupil_grid, hcipy_Cn_squared_from_fried_parameter, von_karman_screen
    g = hcipy_make_pupil_grid(240, 0.5)
    cn2 = hcipy_Cn_squared_from_fried_parameter(0.10, 2.2e-6)
    kw = dict(atm_type='dynamic', atm_vel=20, atm_fried=0.10, act_dim=64, obs_dim=2, rew_type='strehl_ratio',
              timesteps_per_episode=6, SH_operation=True, seed=3)
    for s in range(2):
        scr = von_karman_screen(g, cn2, 10.0, np.random.default_rng(50 + s))
        ef = AOEnv(**kw, precision='fused', initial_screen=scr)
        e64 = AOEnv(**kw, precision='f64', initial_screen=scr)
        ts = ef._h
        ef.reset(), e64.reset()
        img64 = e64._h.get_field('sh_image')
        imgtc = ef._h.get_field('sh_image_tc')
        err = np.abs(imgtc - img64)
        print(f'screen {s}: image max {img64.max():.4e}  max abs err / max {err.max() / img64.max():.3e}  '
              f'rel err on pixels > 1e-3 max: {(err / img64)[img64 > 1e-3 * img64.max()].max():.3e}  '
              f'sum rel {abs(imgtc.sum() - img64.sum()) / img64.sum():.3e}', flush=True)
        for t in range(5):
            a, _ = ef.SH_step(noise='none')
            b, _ = e64.SH_step(noise='none')
            da = np.abs(a - b).max() / np.abs(b).max()
            o1, r1, _, _, i1 = ef.step(a)
            o2, r2, _, _, i2 = e64.step(b)
            eo = np.abs(ef.last_obs_f64 - e64.last_obs_f64) / np.abs(e64.last_obs_f64)
            print(f'  t={t} action rel-to-max err {da:.3e}  |a|max {np.abs(b).max():.3e}  obs rel {eo.max():.3e}  '
                  f'reward rel {abs(r1 - r2) / abs(r2):.3e}  power rel {abs(i1["power"] - i2["power"]) / i2["power"]:.3e}',
                  flush=True)
        # photon noise on the device: spread of the action over draws, both samplers
        a0f, a064 = ef._h.get_sh_actuators(), e64._h.get_sh_actuators()
        sp = {}
        for name, e, a0 in (('fused', ef, a0f), ('f64', e64, a064)):
            acts = []
            for k in range(8):
                e._h.set_sh_actuators(a0)
                acts.append(e.SH_step(noise='poisson')[0])
            e._h.set_sh_actuators(a0)
            ref = e.SH_step(noise='none')[0]
            acts = np.array(acts)
            sp[name] = (np.abs(acts.mean(0) - ref).max() / np.abs(ref).max(), acts.std(0).max() / np.abs(ref).max())
        print(f'  photon noise: |mean - noise-free| / max, std / max   fused {sp["fused"][0]:.3e} {sp["fused"][1]:.3e}   '
              f'f64 {sp["f64"][0]:.3e} {sp["f64"][1]:.3e}', flush=True)
        ef.close(), e64.close()
    # timing at 4096 envs
    for prec in ('fused',):
        env = AOVecEnv(4096, **dict(kw, timesteps_per_episode=20), precision=prec)
        env.reset()
        for _ in range(3):
            env.step(env.SH_step()[0])
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        tsh = tst = 0.0
        for _ in range(10):
            e0.record()
            a = env.SH_step()[0]
            e1.record()
            env.step(a)
            e2.record()
            torch.cuda.synchronize()
            tsh += e0.elapsed_time(e1)
            tst += e1.elapsed_time(e2)
        print(f'{prec}: SH_step {tsh / 10:.3f} ms, step {tst / 10:.3f} ms at 4096 envs', flush=True)
        env.close()


if __name__ == '__main__':
    main()
