#!/usr/bin/env python
"""Rollout throughput with a policy in the loop: ``VecRolloutCollector`` on ``AOVecEnv`` with an actor
of the reference's shape (network.py:16-60: three hidden ReLU layers, diagonal Gaussian with covariance 0.5 I,
algorithm.py:107) evaluated on the device every step -- nothing crosses PCIe.  Prints one JSON line.

    python tools/rollout_throughput.py [--envs 4096] [--hidden 256] [--episodes 4]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=4096)
    ap.add_argument('--hidden', type=int, default=256)
    ap.add_argument('--episodes', type=int, default=4)
    args = ap.parse_args()
    import torch
    from adaptive_optics_gym_b200 import AOVecEnv
    from adaptive_optics_gym_b200.rollout import VecRolloutCollector
    kw = dict(atm_type='quasi_static', atm_fried=0.20, act_type='num_actuators', act_dim=64, obs_dim=2,
              rew_type='strehl_ratio', timesteps_per_episode=30)
    env = AOVecEnv(args.envs, **kw, precision='fused', seed=0)
    dev = env.device
    H = args.hidden
    net = torch.nn.Sequential(torch.nn.Linear(4, H), torch.nn.ReLU(), torch.nn.Linear(H, H), torch.nn.ReLU(),
                              torch.nn.Linear(H, H), torch.nn.ReLU(), torch.nn.Linear(H, 64)).to(dev)
    std = 0.5 ** 0.5

    def policy(obs):
        mean = net(obs)
        act = mean + std * torch.randn_like(mean)
        logp = -0.5 * (((act - mean) / std) ** 2).sum(dim=1) - 64 * (0.5 * torch.log(torch.tensor(2 * torch.pi * std ** 2)))
        return act, logp

    col = VecRolloutCollector(env, policy)
    col.rollout(1)                                        # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = col.rollout(args.episodes)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = out[0].shape[0]
    print(json.dumps({'envs': args.envs, 'hidden': H, 'episodes_per_env': args.episodes, 'transitions': n,
                      'seconds': round(dt, 4), 'env_steps_per_s': round(n / dt),
                      'mean_episode_return': round(float(col.episode_returns().mean()), 3)}))


if __name__ == '__main__':
    main()
