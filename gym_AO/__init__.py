"""Drop-in for the reference's ``gym_AO`` package: ``import gym_AO`` registers ``AO-v0``
(reference ``gym_AO/__init__.py:6-11``) -- backed by the B200 CUDA step path."""
from adaptive_optics_gym_b200._gym_compat import register

register(
    id='AO-v0',
    entry_point='gym_AO.envs:AOEnv',
)
