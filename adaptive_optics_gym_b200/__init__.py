"""aogym-b200: the ``AO-v0`` adaptive-optics environment step path on B200 (sm_100a).

Public API: :class:`AOEnv` (reference-compatible single env), :class:`AOVecEnv` (N lock-stepped
envs on one GPU, torch tensors), :mod:`adaptive_optics_gym_b200.sharding` (multi-GPU env sharding).
"""
__version__ = '0.1.0'


def __getattr__(name):
    if name in ('AOEnv', 'AOVecEnv'):
        from . import env
        return getattr(env, name)
    raise AttributeError(name)
