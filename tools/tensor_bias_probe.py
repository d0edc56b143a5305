"""Diagnostic: signed error of the tcgen05 stage-1 product against an FP64 product of the SAME split-fp16
operands (isolates tensor-core accumulation from operand rounding)."""
import sys, numpy as np
sys.path.insert(0, '.')
from adaptive_optics_gym_b200 import AOEnv
from adaptive_optics_gym_b200.tables import AOConfig, build_tables
from tests.test_parity_gpu import _screen
for seed, flat in ((0, False), (1, False), (2, True)):
    scr = np.zeros(57600) if flat else _screen(seed, 0.2)
    env = AOEnv(precision='tensor', atm_fried=0.2, act_dim=64, obs_dim=2, initial_screen=scr)
    env.reset()
    E = env._h.get_field('tc_pupil').reshape(240, 240)
    T = env._h.get_field('tc_stage1').reshape(128, 240)
    M1 = env.tables['mft_fib_1'] / env.tables['pupil_weight']
    # the kernel's twiddles are fp16 hi+lo of the FP64 table
    def split(x):
        hi = x.astype(np.float16).astype(np.float64); lo = (x - hi).astype(np.float16).astype(np.float64); return hi + lo
    M1s = split(M1.real) + 1j * split(M1.imag)
    Tref = M1s @ E
    mag = np.abs(Tref)
    big = mag > 0.3 * mag.max()
    rel = (np.abs(T) - mag)[big] / mag[big]
    print(f'seed {seed} flat={flat}: |T| signed rel err mean {rel.mean():+.3e} std {rel.std():.2e}; '
          f'complex rel err rms {np.sqrt(np.mean(np.abs(T-Tref)[big]**2))/np.sqrt(np.mean(mag[big]**2)):.2e}')
    env.close()
