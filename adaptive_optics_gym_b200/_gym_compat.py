"""gymnasium when it is installed, else the few names the ``AO-v0`` boundary needs.

The reference registers ``AO-v0`` with gymnasium (``gym_AO/__init__.py:6-11``) and its callers
only use ``gym.make``, ``spaces.Box`` and the ``Env`` base class (``main.py:280``,
``algorithm.py:32-35``).  gymnasium is absent from the build image, so this module supplies a
minimal stand-in with the same call signatures; with gymnasium present it is a pass-through.
"""
from __future__ import annotations

import importlib

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as _gym
    from gymnasium import spaces
    from gymnasium.envs.registration import register
    Env = _gym.Env
    make = _gym.make
    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f'Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})'

    class _Spaces:
        pass

    spaces = _Spaces()
    spaces.Box = Box

    class Env:
        metadata = {}
        observation_space = None
        action_space = None

        def reset(self, seed=None, options=None):
            raise NotImplementedError

        def step(self, action):
            raise NotImplementedError

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    _REGISTRY = {}

    def register(id, entry_point, **kwargs):
        _REGISTRY[id] = (entry_point, kwargs)

    def make(id, **kwargs):
        entry_point, defaults = _REGISTRY[id]
        if isinstance(entry_point, str):
            mod, attr = entry_point.split(':')
            entry_point = getattr(importlib.import_module(mod), attr)
        return entry_point(**{**defaults, **kwargs})
