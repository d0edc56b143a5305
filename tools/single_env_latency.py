import time, numpy as np, sys
sys.path.insert(0, '.')
from adaptive_optics_gym_b200 import AOEnv
kw = dict(atm_type='quasi_static', atm_fried=0.20, act_type='num_actuators', act_dim=64, obs_dim=2, rew_type='strehl_ratio', timesteps_per_episode=30)
for prec in (sys.argv[1:] or ('f64', 'tensor', 'fused')):
    env = AOEnv(**kw, seed=0, precision=prec)
    env.reset()
    a = np.random.default_rng(0).uniform(-1, 1, 64).astype(np.float32)
    for _ in range(20): env.step(a)
    t0 = time.perf_counter(); n = 0
    for ep in range(10):
        env.reset()
        for _ in range(30): env.step(a); n += 1
    dt = time.perf_counter() - t0
    print(prec, 'single-env steps/s', round(n / dt, 1), 'ms/step', round(1e3 * dt / n, 3))
    env.close()
