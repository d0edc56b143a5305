"""Drop-in for the reference's ``gym_AO`` package: importing it makes ``gym.make('AO-v0', ...)`` resolve to the
B200 CUDA step path (the reference registers the same id and entry point in ``gym_AO/__init__.py:6-11``)."""
from adaptive_optics_gym_b200 import _gym_compat

ENV_ID = 'AO-v0'
ENTRY_POINT = 'gym_AO.envs:AOEnv'

_gym_compat.register(id=ENV_ID, entry_point=ENTRY_POINT)
