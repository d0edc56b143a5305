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
