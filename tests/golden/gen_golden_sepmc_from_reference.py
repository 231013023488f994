"""Generate tests/golden/sepmc_reference_golden.npz: the UNMODIFIED reference ChaseTagGameEnv (max_game/chase_tag_game_env.py)
with the shipped SEPMC config (train_scripts/example_sepmc_train.sh:94-117: empty arena, no cubes / hurdles / holes), executed
in this container on tests/golden/pybullet_shim.py (oracle physics per robot, independent numpy ray caster and numpy
sphere / box contact tests over the model's detection proxies).

np.random.{uniform,rand,randint} are replaced by a scripted source handing out the engine's Philox streams (include/llq.h:
stream 1 = reset draws in the order the reference consumes them, 2 = push randomiser, 4 = flag re-placement after a switch).
The frozen file pins, through the oracle's own sampling path: the reset draws (speed command, flag holder, friction, the two
poses with the shared accumulating yaw, flag position), the two-robot push schedule (PR:79-87: one fresh force per robot per
windowed sub-step), all twelve observation entries of both agents, visibility, flag switching, tagging, rewards, termination.

Scenario episodes teleport a robot through the reference's own LeggedRobot.set_states_info() (fp32-representable states; the
replay applies the same teleports): next to the flag (switch), next to the other robot (tag), flag right in front of a robot's
head handle (occlusion), into a wall (wall contact).

    python tests/golden/gen_golden_sepmc_from_reference.py     (needs /root/reference; run from the repo root)"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
REF_SRC = "/root/reference/src"

import pybullet_shim  # noqa: E402
from gen_golden_epmc_from_reference import stream_uniforms  # noqa: E402
from lifelike_agility_and_play_b200.model.compile_model import load_model_blob  # noqa: E402
from oracle import oracle  # noqa: E402

SEED = 20240918
KEYS = ["prop", "prop_a", "percept_2d", "percept_1d", "percept_front", "percept_vec", "oppo_info", "oppo_info_cheat", "flag_info",
        "flag_info_cheat", "with_flag", "control_spd"]


class ScriptedRandom:
    """Hands the reference the same uniforms the engine would draw (pair gid 0)."""

    def __init__(self):
        self.episode, self.reset_k, self.push_draws, self.flag_draws = -1, 0, 0, 0
        self.push_slot, self.flag_slot, self.flag_reset_left = 0, 0, 0

    @staticmethod
    def caller():
        return sys._getframe(2).f_code.co_name

    def _reset_draw(self):
        u = stream_uniforms(SEED, 0, self.episode, 1, self.reset_k // 4)[self.reset_k % 4]
        self.reset_k += 1
        return u

    def uniform(self, lo=0.0, hi=1.0):
        who = self.caller()
        if who == "__init__":             # ChaseTagGameEnv.__init__ draws a throw-away friction (CTG:61)
            return lo + 0.5 * (hi - lo)
        if who == "reset":                # episodic_fix_spd (first draw of an episode, CTG:262) ... friction (CTG:277)
            if self.reset_k != 2 or self.episode < 0:     # k == 2 is the friction draw of the running reset
                self.episode += 1
                self.reset_k = self.push_draws = self.flag_draws = 0
                self.flag_reset_left = 2
            u = self._reset_draw()
        elif who == "randomize_force":    # theta, h, v (PR:89-93)
            u = stream_uniforms(SEED, 0, self.episode, 2, self.push_draws)[self.push_slot]
            self.push_slot += 1
            if self.push_slot == 3:
                self.push_slot, self.push_draws = 0, self.push_draws + 1
        elif who == "randomize_init_states":   # the four position draws (CTG:205-206)
            u = self._reset_draw()
        elif who == "_randomize_flag_pos":     # at reset: reset stream; after a switch: stream 4 (CTG:218-221)
            if self.flag_reset_left > 0:
                self.flag_reset_left -= 1
                u = self._reset_draw()
            else:
                u = stream_uniforms(SEED, 0, self.episode, 4, self.flag_draws)[self.flag_slot]
                self.flag_slot += 1
                if self.flag_slot == 2:
                    self.flag_slot, self.flag_draws = 0, self.flag_draws + 1
        else:
            raise RuntimeError("unexpected np.random.uniform caller " + who)
        return lo + u * (hi - lo)

    def rand(self):
        assert self.caller() == "randomize_init_states"                                  # CTG:210
        return self._reset_draw()

    def randint(self, lo, hi):
        assert self.caller() == "reset"                                                  # CTG:266
        return lo + int(np.floor(self._reset_draw() * (hi - lo)))


def main():
    assert os.path.isdir(REF_SRC), "reference tree not mounted"
    eng = oracle.make_engine(2, load_model_blob(), None, env_kind=2, kp=50.0, kd=0.5, max_tau=16.0, ground_friction=1.0, max_steps=1000,
                             push_interval_steps=499)
    FB = pybullet_shim.FakeBulletClient
    FB.oracle_engine, FB.boxes_block_rays, FB.pair_contacts = eng, True, True
    pybullet_shim.install()
    sys.path.insert(0, REF_SRC)
    import builtins
    real_print = builtins.print
    builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and (a[0].startswith("Current episodic") or a[0].startswith("Terminates"))) else real_print(*a, **k)
    if not hasattr(np, "float"):
        np.float = float                      # CTG:584 uses the alias numpy >= 1.24 removed
    from lifelike.sim_envs.pybullet_envs.create_pybullet_envs import create_chase_tag_game
    from lifelike.utils.constants import STATES_INFO_12_RUN_0
    init0 = copy.deepcopy(STATES_INFO_12_RUN_0)      # mutated by every reset (CTG:209-215)
    sr = ScriptedRandom()
    np.random.uniform, np.random.rand, np.random.randint = sr.uniform, sr.rand, sr.randint
    max_steps = 40
    env_config = {
        'arena_id': 'CTG', 'render': False, 'control_freq': 50.0,
        'prop_type': ['joint_pos', 'joint_vel', 'root_ang_vel_loc', 'root_lin_vel_loc', 'e_g'],
        'kp': 50.0, 'kd': 0.5, 'max_tau': 16, 'max_steps': max_steps, 'obs_randomization': {},
        'env_randomize_config': {
            'friction_range': [0.4, 3.0],
            'disturb_force_config': {'start_time': 0.5, 'interval_time': 1.0, 'duration_time': 0.2,
                                     'horizontal_force': [0, 50], 'vertical_force': [0, 10]}},
        'element_config': {'rand_cube': False, 'hurdle': False, 'hole': False},
    }
    env = create_chase_tag_game(**env_config)
    bc = env.bullet_client
    bc.bodies[env.flag_id].is_flag = True
    robots = env.legged_robots
    rec = {k: [] for k in ["episode", "action", "reward", "done", "state", "aux", "obs", "reset_obs", "reset_state", "reset_aux",
                           "tp_step", "tp_robot", "tp_state"]}
    arng = np.random.default_rng(13)

    def flat_obs(o):
        return np.concatenate([np.asarray(o[k], dtype=np.float64).reshape(-1) for k in KEYS])

    def flat_state(si):
        return np.concatenate([si["base_pos"], si["base_orn"], si["base_lin_vel"], si["base_ang_vel"], si["joint_pos"], si["joint_vel"]])

    def aux(i):
        fr = env.force_randomizer
        return np.array([env.counter, float(env.with_flag[i]), env.target_pos[0], env.target_pos[1], env.episodic_fix_spd,
                         float(env.oppo_visible[i]), float(env.switch_flag_at_this_frame), env.total_spds[i], env.max_spds[i], fr._count,
                         0.0, 0.0, 0.0, bc.bodies[robots[i].robot_id].foot_mu, sr.push_draws, sr.flag_draws, 0.0, 0.0], dtype=np.float64)

    def teleport(i, pos=None, yaw=None):
        si = robots[i].get_states_info()
        st = flat_state(si)
        if pos is not None:
            st[0:3] = pos
        if yaw is not None:
            base = np.asarray(init0["base_orn"], dtype=np.float64)
            from scipy.spatial.transform import Rotation as R
            st[3:7] = (R.from_euler("z", yaw) * R.from_quat(base)).as_quat()
        st[7:13] = 0.0
        st[13:25] = init0["joint_pos"]
        st[25:37] = 0.0
        st = st.astype(np.float32).astype(np.float64)
        robots[i].set_states_info({"base_pos": list(st[0:3]), "base_orn": list(st[3:7]), "base_lin_vel": list(st[7:10]),
                                   "base_ang_vel": list(st[10:13]), "joint_pos": list(st[13:25]), "joint_vel": list(st[25:37])})
        rec["tp_step"].append(len(rec["reward"])); rec["tp_robot"].append(i); rec["tp_state"].append(st)

    n_ep = 8
    for ep in range(n_ep):
        obs = env.reset()
        rec["reset_obs"].append(np.stack([flat_obs(o) for o in obs]))
        rec["reset_state"].append(np.stack([flat_state(r.get_states_info()) for r in robots]))
        rec["reset_aux"].append(np.stack([aux(0), aux(1)]))
        for t in range(max_steps + 5):
            scale = [0.15, 0.15]
            if ep == 3:
                scale = [0.7, 0.15]           # robot 0 flails until it falls: the episode ends
            if ep == 4:
                scale = [0.15, 0.7]           # robot 1 falling does not end the episode (CTG:462)
            flag = np.array(env.target_pos)
            if ep == 1 and t == 3:            # the robot that does not hold the flag walks into it: switch (CTG:573-579)
                i = 1 if env.with_flag[0] else 0
                teleport(i, pos=[flag[0] - 0.16, flag[1], 0.31], yaw=0.0)
            if ep == 1 and t == 9:            # ... and the new non-holder touches the re-placed flag: switch back
                i = 1 if env.with_flag[0] else 0
                teleport(i, pos=[flag[0], flag[1] - 0.14, 0.31], yaw=1.0)
            if ep == 2 and t == 5:            # robot 1 is dropped onto robot 0's flank: tag, done (CTG:464)
                p0 = np.array(robots[0].get_states_info()["base_pos"])
                teleport(1, pos=[p0[0], p0[1] + 0.22, p0[2]], yaw=0.3)
            if ep == 5 and t == 4:            # flag 3 cm in front of robot 0's head handle, robot 1 straight behind it
                teleport(0, pos=[flag[0] - 0.33, flag[1], 0.30], yaw=0.0)
                teleport(1, pos=[min(flag[0] + 1.2, 2.2), flag[1], 0.30], yaw=np.pi)
            if ep == 6 and t == 2:            # front feet pushed into the +x wall, the other robot into the -y wall
                teleport(0, pos=[2.30, 0.3, 0.31], yaw=0.0)
                teleport(1, pos=[-0.5, -2.30, 0.31], yaw=-np.pi / 2)
            a = [(scale[i] * arng.standard_normal(12)).astype(np.float32) for i in range(2)]
            o, r, d, info = env.step([{'A_LLC': a[i].astype(np.float64)} for i in range(2)])
            rec["episode"].append(ep); rec["action"].append(np.stack(a)); rec["reward"].append(np.array(r, dtype=np.float64)); rec["done"].append(bool(d))
            rec["state"].append(np.stack([flat_state(rb.get_states_info()) for rb in robots])); rec["aux"].append(np.stack([aux(0), aux(1)]))
            rec["obs"].append(np.stack([flat_obs(x) for x in o]))
            if d:
                break
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["obs"] = out["obs"].astype(np.float32)        # what the engine emits; fp64 -> fp32 rounding is 6e-8 relative
    out["reset_obs"] = out["reset_obs"].astype(np.float32)
    out["init_state"] = np.concatenate([init0["base_pos"], init0["base_orn"], init0["base_lin_vel"], init0["base_ang_vel"],
                                        init0["joint_pos"], init0["joint_vel"]])
    out["seed"] = SEED
    out["max_steps"] = max_steps
    path = os.path.join(ROOT, "tests", "golden", "sepmc_reference_golden.npz")
    np.savez_compressed(path, **out)
    real_print("wrote", path, "steps", len(rec["reward"]), "dones", int(np.sum(rec["done"])), "per-episode lengths",
               np.bincount(np.asarray(rec["episode"])), "switch steps", np.flatnonzero(out["aux"][:, 0, 6] > 0),
               "reward!=0", np.flatnonzero(np.abs(out["reward"]).sum(1) > 0), "invisible", np.flatnonzero(out["aux"][:, :, 5].min(1) < 1),
               "size", os.path.getsize(path))


if __name__ == "__main__":
    main()
