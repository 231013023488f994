"""Generate tests/golden/epmc_reference_golden.npz: the UNMODIFIED reference PlayGroundEnv (max_game_elements/
playground_env.py) with the shipped EPMC config (train_scripts/example_epmc_train.sh:88-117, element_id 0), executed in
this container on tests/golden/pybullet_shim.py (oracle physics, independent numpy ray caster).

The reference draws its random numbers from numpy's global generator; here np.random.{uniform,rand,randint} are replaced
by a scripted source that hands out the engine's Philox streams (include/llq.h: stream 1 = reset, 2 = push randomiser,
3 = joystick command) in the order the reference consumes them.  The frozen file therefore pins, through the oracle's own
sampling path: friction / yaw / command-frequency draws, the push schedule (PR:56-87, K7), the joystick command logic,
the three perception arrays, the reward and the termination rules of the reference.

    python tests/golden/gen_golden_epmc_from_reference.py     (needs /root/reference; run from the repo root)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
REF_SRC = "/root/reference/src"

import pybullet_shim  # noqa: E402
from lifelike_agility_and_play_b200.model.compile_model import load_model_blob  # noqa: E402
from oracle import oracle  # noqa: E402

SEED = 20240917
M32 = 0xFFFFFFFF


def philox4x32_10(c, k0, k1):
    c = list(c)
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0, k1 = (k0 + 0x9E3779B9) & M32, (k1 + 0xBB67AE85) & M32
    return c


def stream_uniforms(seed, gid, episode, stream, index):
    c = [gid & M32, ((gid >> 32) & 0x00FFFFFF) | (stream << 24), episode & M32, index & M32]
    return [(x + 0.5) / 4294967296.0 for x in philox4x32_10(c, seed & M32, (seed >> 32) & M32)]


class ScriptedRandom:
    """Hands the reference the same uniforms the engine would draw (gid 0)."""

    def __init__(self):
        self.episode, self.push_draws, self.cmd_draws = -1, 0, 0
        self.push_slot, self.cmd_slot = 0, 0

    def caller(self):
        f = sys._getframe(2)
        return f.f_code.co_name

    def uniform(self, lo=0.0, hi=1.0):
        who = self.caller()
        if who == "__init__":            # PlayGroundEnv.__init__ draws a throw-away friction before the first reset (PGE:94)
            return lo + 0.5 * (hi - lo)
        if who == "reset":               # friction (PGE:209): first draw of a new episode
            self.episode += 1
            self.push_draws = self.cmd_draws = 0
            u = stream_uniforms(SEED, 0, self.episode, 1, 0)[0]
        elif who == "randomize_force":   # theta, h, v (PR:89-93)
            u = stream_uniforms(SEED, 0, self.episode, 2, self.push_draws)[self.push_slot]
            self.push_slot += 1
            if self.push_slot == 3:
                self.push_slot, self.push_draws = 0, self.push_draws + 1
        elif who == "step":              # target_angle then target_spd (PGE:308,317)
            u = stream_uniforms(SEED, 0, self.episode, 3, self.cmd_draws)[self.cmd_slot]
            self.cmd_slot += 1
            if self.cmd_slot == 2:
                self.cmd_slot, self.cmd_draws = 0, self.cmd_draws + 1
        else:
            raise RuntimeError("unexpected np.random.uniform caller " + who)
        return lo + u * (hi - lo)

    def rand(self):
        assert self.caller() == "randomize_init_states"                                  # PGE:183
        return stream_uniforms(SEED, 0, self.episode, 1, 0)[1]

    def randint(self, lo, hi):
        assert self.caller() == "reset"                                                  # PGE:223
        return lo + int(np.floor(stream_uniforms(SEED, 0, self.episode, 1, 0)[2] * (hi - lo)))


def main():
    assert os.path.isdir(REF_SRC), "reference tree not mounted"
    eng = oracle.make_engine(1, load_model_blob(), None, env_kind=1, kp=50.0, kd=0.5, max_tau=16.0, ground_friction=1.0)
    pybullet_shim.FakeBulletClient.oracle_engine = eng
    pybullet_shim.install()
    sys.path.insert(0, REF_SRC)
    import builtins
    real_print = builtins.print
    builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and (a[0].startswith("Current episodic") or a[0].startswith("Terminates"))) else real_print(*a, **k)
    from lifelike.sim_envs.pybullet_envs.create_pybullet_envs import create_playground_game
    from lifelike.utils.constants import STATES_INFO_12_RUN_0
    import copy
    init0 = copy.deepcopy(STATES_INFO_12_RUN_0)      # the env mutates this module-level dict at every reset (PGE:181-189)
    sr = ScriptedRandom()
    np.random.uniform, np.random.rand, np.random.randint = sr.uniform, sr.rand, sr.randint
    max_steps = 60
    env_config = {
        'arena_id': 'Playground', 'render': False, 'control_freq': 50.0,
        'prop_type': ['joint_pos', 'joint_vel', 'root_ang_vel_loc', 'root_lin_vel_loc', 'e_g'],
        'kp': 50.0, 'kd': 0.5, 'max_tau': 16, 'max_steps': max_steps, 'obs_randomization': {},
        'env_randomize_config': {
            'element_id': 0, 'height_range': [0.0, 0.0], 'friction_range': [0.4, 3.0],
            'disturb_force_config': {'start_time': 0.5, 'interval_time': 1.0, 'duration_time': 0.2,
                                     'horizontal_force': [0, 50], 'vertical_force': [0, 10]},
            'cmd_vary_freq_range': [25, 40], 'target_spd_range': [0.5, 3.0], 'auxiliary_radius': 0.02,
            'hole_config': {'min_gap_height': 0.25, 'max_gap_height': 0.25}},
    }
    env = create_playground_game(**env_config)
    inner = env.env
    keys = ["prop", "prop_a", "percep_2d", "percep_1d", "percep_front", "target"]
    rec = {k: [] for k in ["episode", "action", "reward", "done", "state", "aux", "reset_obs", "reset_state", "reset_aux"] + keys}
    arng = np.random.default_rng(11)

    def flat_obs(o):
        return np.concatenate([np.asarray(o[k], dtype=np.float64).reshape(-1) for k in keys])

    def flat_state(si):
        return np.concatenate([si["base_pos"], si["base_orn"], si["base_lin_vel"], si["base_ang_vel"], si["joint_pos"], si["joint_vel"]])

    def aux():
        fr = inner.force_randomizer
        return np.array([inner.counter, inner.cmd_vary_freq, inner._target_pos[0], inner._target_pos[1], inner.target_spd,
                         inner.target_angle, inner.last_pos_diff_len, inner.total_spd, inner.max_spd, fr._count,
                         fr._randomized_force[0], fr._randomized_force[1], fr._randomized_force[2],
                         inner.bullet_client.bodies[inner.legged_robot.robot_id].foot_mu, sr.push_draws, sr.cmd_draws, 0.0, 0.0], dtype=np.float64)

    for ep in range(4):
        obs = env.reset(inter_kwargs={})[0]
        rec["reset_obs"].append(flat_obs(obs)); rec["reset_state"].append(flat_state(inner.legged_robot.get_states_info()))
        rec["reset_aux"].append(aux())
        for t in range(max_steps + 5):
            scale = 0.15 if ep != 2 else 0.6          # episode 2 flails until the robot falls
            a = (scale * arng.standard_normal(12)).astype(np.float32)
            o, r, d, info = env.step([{'A_LLC': a.astype(np.float64)}])
            o = o[0]
            rec["episode"].append(ep); rec["action"].append(a); rec["reward"].append(r[0]); rec["done"].append(bool(d))
            rec["state"].append(flat_state(inner.legged_robot.get_states_info())); rec["aux"].append(aux())
            for k in keys:
                rec[k].append(np.asarray(o[k], dtype=np.float64).reshape(-1))
            if d:
                break
    out = {k: np.asarray(v) for k, v in rec.items()}
    init = init0
    out["init_state"] = np.concatenate([init["base_pos"], init["base_orn"], init["base_lin_vel"], init["base_ang_vel"], init["joint_pos"], init["joint_vel"]])
    out["seed"] = SEED
    out["max_steps"] = max_steps
    path = os.path.join(ROOT, "tests", "golden", "epmc_reference_golden.npz")
    np.savez_compressed(path, **out)
    real_print("wrote", path, "steps", len(rec["reward"]), "dones", int(np.sum(rec["done"])), "per-episode lengths",
               np.bincount(np.asarray(rec["episode"])), "size", os.path.getsize(path))


if __name__ == "__main__":
    main()
