"""Environment factories with the reference's names and keyword contract.

Mirrors ``lifelike/sim_envs/pybullet_envs/create_pybullet_envs.py`` (create_tracking_game :21-64,
SingleAgentWrapper :6-18, create_tracking_env :143-147).  ``--outer_env`` of the reference launcher
(bin/run_pg_actor.py:81-83) takes a dotted path to one of these callables, so pointing it at
``lifelike_agility_and_play_b200.sim_envs.create_envs.create_tracking_game`` drops the actor onto the CUDA engine.
"""
from . import primitive_level_env as _ple
from .. import spaces


class SingleAgentWrapper:
    """CPE:6-18 -- tuple-ises spaces, observations and rewards for TLeague's multi-agent protocol."""

    def __init__(self, env):
        self.env = env
        self.observation_space = spaces.Tuple([env.observation_space])
        self.action_space = spaces.Tuple([env.action_space])

    def __getattr__(self, name):          # gym.Wrapper forwards unknown attributes to the wrapped env
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kwargs):            # accepts and ignores inter_kwargs= (distill_actor.py:205)
        obs = self.env.reset()
        return (obs,)

    def step(self, action):
        obs, rwd, done, info = self.env.step(action[0])
        return (obs,), (rwd,), done, info

    def close(self):
        return self.env.close()


def create_tracking_game(**env_config):
    arena_id = env_config["arena_id"]
    assert arena_id in [
        "LeggedRobotTracking",
    ]
    env0 = _ple.PrimitiveLevelEnv(
        enable_render=env_config.get("render", False),
        control_freq=env_config.get("control_freq", 25.0),
        kp=env_config.get("kp", 50.0),
        kd=env_config.get("kd", 1.0),
        foot_lateral_friction=env_config.get('foot_lateral_friction', 0.5),
        max_tau=env_config.get("max_tau", 18.0),
        sim_freq=env_config.get("sim_freq", 500.0),
        video_path=env_config.get('video_path', None),
        enable_gui=env_config.get('enable_gui', True),
        data_path=env_config.get("data_path", ""),
        prop_type=env_config.get("prop_type", ""),
        prioritized_sample_factor=env_config.get("prioritized_sample_factor", 0.0),
        set_obstacle=env_config.get("set_obstacle", False),
        obstacle_height=env_config.get("obstacle_height", 0.0),
        reward_weights=env_config.get("reward_weights", None),
        # engine-only extras (ignored by the reference)
        seed=env_config.get("seed", None), device=env_config.get("device", 0), mocap=env_config.get("mocap", None),
    )
    return SingleAgentWrapper(env0)


def create_tracking_env(**env_config):
    env = create_tracking_game(**env_config)
    env.observation_space = env.observation_space.spaces[0]
    env.action_space = env.action_space.spaces[0]
    return env


def create_playground_game(**env_config):
    """CPE:67-101 -- element_id 0 (flat joystick arena, the shipped script default) runs on the CUDA engine."""
    arena_id = env_config["arena_id"]
    assert arena_id in [
        "Playground",
    ]
    from .playground_env import PlayGroundEnv
    env0 = PlayGroundEnv(
        enable_render=env_config["render"] if "render" in env_config else False,
        control_freq=env_config["control_freq"] if "control_freq" in env_config else 50.0,
        kp=env_config["kp"] if "kp" in env_config else 50.0,
        kd=env_config["kd"] if "kd" in env_config else 1.0,
        max_tau=env_config["max_tau"] if "max_tau" in env_config else 16.0,
        prop_type=env_config["prop_type"] if "prop_type" in env_config else None,
        max_steps=env_config["max_steps"] if "max_steps" in env_config else 1000,
        obs_randomization=env_config["obs_randomization"] if "obs_randomization" in env_config else None,
        env_randomize_config=env_config["env_randomize_config"] if "env_randomize_config" in env_config else None,
        seed=env_config.get("seed", None), device=env_config.get("device", 0),
    )
    return SingleAgentWrapper(env0)


def create_playground_env(**env_config):
    env = create_playground_game(**env_config)
    env.observation_space = env.observation_space.spaces[0]
    env.action_space = env.action_space.spaces[0]
    return env


def create_chase_tag_game(**env_config):
    """CPE:104-140 -- the shipped empty arena runs on the CUDA engine (two robots = one pair)."""
    arena_id = env_config["arena_id"]
    assert arena_id in [
        "CTG",
    ]
    from .chase_tag_game_env import ChaseTagGameEnv
    env0 = ChaseTagGameEnv(
        enable_render=env_config["render"] if "render" in env_config else False,
        control_freq=env_config["control_freq"] if "control_freq" in env_config else 25.0,
        kp=env_config["kp"] if "kp" in env_config else 50.0,
        kd=env_config["kd"] if "kd" in env_config else 1.0,
        max_tau=env_config["max_tau"] if "max_tau" in env_config else 18.0,
        prop_type=env_config["prop_type"] if "prop_type" in env_config else None,
        max_steps=env_config["max_steps"] if "max_steps" in env_config else 1000,
        obs_randomization=env_config["obs_randomization"] if "obs_randomization" in env_config else None,
        element_config=env_config.get('element_config', {}),
        env_randomize_config=env_config.get('env_randomize_config', {}),
        seed=env_config.get("seed", None), device=env_config.get("device", 0),
    )
    return env0


def create_chase_tag_env(**env_config):
    env = create_chase_tag_game(**env_config)
    env.observation_space = env.observation_space.spaces[0]
    env.action_space = env.action_space.spaces[0]
    return env
