from .create_envs import *  # noqa: F401,F403  (mirrors lifelike/sim_envs/pybullet_envs/__init__.py:1)
