"""Reference ``gym_AO/envs/__init__.py:6`` exports ``AOEnv``; the vectorised variant rides along."""
from adaptive_optics_gym_b200.env import AOEnv, AOVecEnv  # noqa: F401
