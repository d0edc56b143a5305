"""CPU: host-side logic of the boundary that needs no GPU -- registration, the gymnasium
stand-in, velocity coercion, env sharding (incl. a world_size-2 gloo run)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registration_and_spaces():
    import gym_AO  # noqa: F401  (registers AO-v0 like the reference's gym_AO/__init__.py)
    from adaptive_optics_gym_b200 import _gym_compat as G
    if not G.HAVE_GYMNASIUM:
        assert 'AO-v0' in G._REGISTRY and G._REGISTRY['AO-v0'][0] == 'gym_AO.envs:AOEnv'
    box = G.spaces.Box(low=-1, high=1, shape=(4,), dtype=np.float16)
    assert box.shape == (4,) and box.dtype == np.float16
    assert type(box).__name__ == 'Box'


def test_velocity_coercion_prints_like_reference(capsys):
    from adaptive_optics_gym_b200.env import _coerce_velocity
    assert _coerce_velocity('quasi_static', 5) == 0
    out = capsys.readouterr().out
    assert 'In quasi_static atmospheric condition, the velocity value should be zero.' in out
    assert _coerce_velocity('dynamic', 0) == 1
    assert 'therefore velocity value is changed to 1 m/s' in capsys.readouterr().out
    assert _coerce_velocity('dynamic', 7) == 7 and _coerce_velocity('semi_dynamic', 0) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from adaptive_optics_gym_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError, match='no CPU fallback'):
        _lib.load()


def test_shard_range_partitions_exactly():
    from adaptive_optics_gym_b200.sharding import shard_range
    for n, g in ((16384, 8), (4096, 1), (10, 4), (7, 8)):
        blocks = [shard_range(n, r, g) for r in range(g)]
        ids = [i for f, c in blocks for i in range(f, f + c)]
        assert ids == list(range(n))
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from adaptive_optics_gym_b200.sharding import shard_range, gather_episode_stats
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
first, count = shard_range(10, rank, world)
returns = torch.arange(first, first + count, dtype=torch.float64) * -1.5      # per-env episode returns
st = gather_episode_stats(returns)
full = torch.arange(10, dtype=torch.float64) * -1.5
assert st['count'] == 10 and abs(st['mean'] - full.mean().item()) < 1e-12, st
assert abs(st['std'] - full.std(unbiased=False).item()) < 1e-12 and st['min'] == -13.5 and st['max'] == 0.0, st
dist.destroy_process_group()
print('ok', rank)
'''


def test_gather_episode_stats_world_size_2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29533', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f'ok {r}' in o, o


def test_gather_episode_stats_single_process():
    import torch
    from adaptive_optics_gym_b200.sharding import gather_episode_stats
    st = gather_episode_stats(torch.tensor([1.0, 2.0, 3.0]))
    assert st == dict(count=3, mean=2.0, std=pytest.approx((2 / 3) ** 0.5), min=1.0, max=3.0)


class _FakeVecEnv:
    """CPU stand-in with AOVecEnv's interface: obs = [env id, step counter], reward = -(step + env / 100)."""

    def __init__(self, B=3, T=4, K=2):
        import torch
        from adaptive_optics_gym_b200._gym_compat import spaces
        self.num_envs, self.max_steps, self.device, self.SH_operation = B, T, torch.device('cpu'), False
        self.single_observation_space = spaces.Box(low=-1, high=1, shape=(2,), dtype=np.float16)
        self.single_action_space = spaces.Box(low=-1, high=1, shape=(K,), dtype=np.float16)
        self.t = 0

    def _obs(self):
        import torch
        return torch.stack([torch.arange(self.num_envs, dtype=torch.float16),
                            torch.full((self.num_envs,), float(self.t), dtype=torch.float16)], dim=1)

    def reset(self):
        self.t = 0
        return self._obs(), {}

    def step(self, actions):
        import torch
        assert actions.shape == (self.num_envs, self.single_action_space.shape[0])
        rew = -(self.t + torch.arange(self.num_envs, dtype=torch.float64) / 100)
        self.t += 1
        done = torch.full((self.num_envs,), self.t == self.max_steps)
        return self._obs(), rew, done, torch.zeros(self.num_envs, dtype=torch.bool), {'power': rew}


def test_vec_rollout_collector_layout_matches_reference_rollout():
    """the seven fields of algorithm.py:216-296, each env's episode a contiguous run of T rows"""
    import torch
    from adaptive_optics_gym_b200.rollout import VecRolloutCollector
    env = _FakeVecEnv(B=3, T=4, K=2)
    policy = lambda o: (o[:, :1].repeat(1, 2) * 0.5, -o[:, 1])        # action from the env id, log-prob from the step
    col = VecRolloutCollector(env, policy)
    obs, act, logp, rew, nxt, done, lens = col.rollout(episodes_per_iteration=2)
    N = 2 * 3 * 4
    assert obs.shape == (N, 2) and act.shape == (N, 2) and logp.shape == (N,) and rew.shape == (N,)
    assert nxt.shape == (N, 2) and done.shape == (N,) and lens.shape == (6,) and torch.all(lens == 4)
    rows = obs.reshape(2, 3, 4, 2)
    for e in range(2):
        for b in range(3):
            assert torch.all(rows[e, b, :, 0] == b) and rows[e, b, :, 1].tolist() == [0, 1, 2, 3]
    assert torch.all(nxt.reshape(2, 3, 4, 2)[..., 1] == torch.tensor([1., 2., 3., 4.]))
    assert torch.all(done.reshape(2, 3, 4)[..., :3] == 0) and torch.all(done.reshape(2, 3, 4)[..., 3] == 1)
    assert torch.allclose(act.reshape(2, 3, 4, 2)[0, 2], torch.full((4, 2), 1.0))
    assert torch.allclose(col.episode_returns(), torch.tensor([-6.0, -6.04, -6.08] * 2))
    assert col.num_episodes == 6
    with pytest.raises(ValueError):
        VecRolloutCollector(env, None)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm): one JSON line with the base contract's keys, timed on
    the CPU oracle with one single-thread env per worker process (2 here), no GPU and no library needed."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AOG_REF_THREADS='2')
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1'], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'env-steps/sec' and line['unit'] == 'env-steps/s'
    assert line['higher_is_better'] is True and line['value'] > 0 and line['cpu_baseline']['kind'] == 'port'
    assert line['cpu_baseline']['cores'] == 2 and line['e2e']['h2d_bytes_per_step'] == 0
    assert 'quasi_static' in line['config']['workload']


def test_fft240_index_maps_on_the_cpu(tmp_path):
    """csrc/fft240.cuh (the 16 x 15 Cooley-Tukey transform of the FFT screen synthesis) compiled for the host against a
    direct O(N^2) DFT: the radix-2 DFT-16, the Good-Thomas DFT-15 and the twiddle / slot maps agree to 1e-13."""
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not found')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'fft240_host')
    subprocess.run([nvcc, '-O2', '-std=c++17', '-Wno-deprecated-gpu-targets', '-o', exe,
                    os.path.join(root, 'tools', 'micro', 'fft240_host.cu')], check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert float(out.strip()) < 1e-13
