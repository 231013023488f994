"""Generate tests/golden/epmc_e{1,2,3}_reference_golden.npz: the UNMODIFIED reference PlayGroundEnv + BulletStatics with
element_id 1 (hurdles), 2 ("holes": bars to pass under) and 3 (cubes, easy) -- the corridor arenas of
max_game_elements/bullet_static_entities.py -- executed in this container on tests/golden/pybullet_shim.py.

The shim mirrors the static boxes the reference creates into the oracle before every stepSimulation (foot spheres collide
with them) and answers rayTestBatch with its own numpy slab test.  np.random is replaced by a scripted source handing out
the engine's Philox streams: 1 = reset (friction, yaw, command frequency), 2 = push randomiser, 3 = command (only the target
speed is drawn for elements != 0), 5 = terrain draws in the order the reference consumes them (wall width, wall gap, object
count, per-object sizes / spacings, target offset).  The frozen files pin, through the oracle's own sampling path: terrain
generation, target placement, the three perception arrays against the box list, the average-speed reward incl. the
reach bonus, termination.  Scenario teleports (reference LeggedRobot.set_states_info, fp32-representable): onto a box,
astride an obstacle edge, next to the target.

    python tests/golden/gen_golden_epmc_terrain_from_reference.py     (needs /root/reference; run from the repo root)"""
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

SEED = 20240919
KEYS = ["prop", "prop_a", "percep_2d", "percep_1d", "percep_front", "target"]
TERRAIN_CALLERS = ("_generate_random_width_walls", "_create_hurdles", "_generate_one_hurdle", "_create_holes", "_generate_one_hole",
                   "_create_cubes", "_generate_one_cube_set")


class ScriptedRandom:
    def __init__(self):
        self.episode, self.push_draws, self.cmd_draws, self.push_slot, self.terrain_k = -1, 0, 0, 0, 0

    @staticmethod
    def caller():
        return sys._getframe(2).f_code.co_name

    def _terrain(self):
        u = stream_uniforms(SEED, 0, self.episode, 5, self.terrain_k // 4)[self.terrain_k % 4]
        self.terrain_k += 1
        return u

    def uniform(self, lo=0.0, hi=1.0):
        who = self.caller()
        if who == "__init__":
            return lo + 0.5 * (hi - lo)
        if who == "reset":               # friction (PGE:209): first draw of a new episode
            self.episode += 1
            self.push_draws = self.cmd_draws = self.terrain_k = 0
            u = stream_uniforms(SEED, 0, self.episode, 1, 0)[0]
        elif who == "randomize_force":
            u = stream_uniforms(SEED, 0, self.episode, 2, self.push_draws)[self.push_slot]
            self.push_slot += 1
            if self.push_slot == 3:
                self.push_slot, self.push_draws = 0, self.push_draws + 1
        elif who == "step":              # elements != 0 draw only target_spd (PGE:316-317): slot 1 of the command draw
            u = stream_uniforms(SEED, 0, self.episode, 3, self.cmd_draws)[1]
            self.cmd_draws += 1
        elif who in TERRAIN_CALLERS:
            u = self._terrain()
        else:
            raise RuntimeError("unexpected np.random.uniform caller " + who)
        return lo + u * (hi - lo)

    def rand(self):
        assert self.caller() == "randomize_init_states"
        return stream_uniforms(SEED, 0, self.episode, 1, 0)[1]

    def randint(self, lo, hi=None):
        who = self.caller()
        if who == "reset":                                                               # cmd_vary_freq (PGE:223)
            return lo + int(np.floor(stream_uniforms(SEED, 0, self.episode, 1, 0)[2] * (hi - lo)))
        assert who in TERRAIN_CALLERS, who
        return lo + int(np.floor(self._terrain() * (hi - lo)))


def run(element, sr):
    FB = pybullet_shim.FakeBulletClient
    eng = oracle.make_engine(1, load_model_blob(), None, env_kind=1, element_id=element, kp=50.0, kd=0.5, max_tau=16.0, ground_friction=1.0)
    FB.oracle_engine, FB.boxes_block_rays, FB.terrain_boxes = eng, True, True
    from lifelike.sim_envs.pybullet_envs.create_pybullet_envs import create_playground_game
    from lifelike.utils.constants import STATES_INFO_12_RUN_0
    init0 = run.init0
    for k in init0:                       # every env instance starts from the pristine module-level dict (PGE:181-189 mutates it)
        STATES_INFO_12_RUN_0[k] = copy.deepcopy(init0[k])
    max_steps = 40
    env_config = {
        'arena_id': 'Playground', 'render': False, 'control_freq': 50.0,
        'prop_type': ['joint_pos', 'joint_vel', 'root_ang_vel_loc', 'root_lin_vel_loc', 'e_g'],
        'kp': 50.0, 'kd': 0.5, 'max_tau': 16, 'max_steps': max_steps, 'obs_randomization': {},
        'env_randomize_config': {
            'element_id': element, 'height_range': [0.0, 0.0], 'friction_range': [0.4, 3.0],
            'disturb_force_config': {'start_time': 0.5, 'interval_time': 1.0, 'duration_time': 0.2,
                                     'horizontal_force': [0, 50], 'vertical_force': [0, 10]},
            'cmd_vary_freq_range': [25, 40], 'target_spd_range': [0.5, 3.0], 'auxiliary_radius': 0.02,
            'hole_config': {'min_gap_height': 0.25, 'max_gap_height': 0.25}},
    }
    env = create_playground_game(**env_config)
    inner = env.env
    bc = inner.bullet_client
    robot = inner.legged_robot
    rec = {k: [] for k in ["episode", "action", "reward", "done", "state", "aux", "obs", "reset_obs", "reset_state", "reset_aux", "boxes", "nbox",
                           "tp_step", "tp_state"]}
    arng = np.random.default_rng(17 + element)

    def flat_obs(o):
        return np.concatenate([np.asarray(o[k], dtype=np.float64).reshape(-1) for k in KEYS])

    def flat_state(si):
        return np.concatenate([si["base_pos"], si["base_orn"], si["base_lin_vel"], si["base_ang_vel"], si["joint_pos"], si["joint_vel"]])

    def aux():
        fr = inner.force_randomizer
        return np.array([inner.counter, inner.cmd_vary_freq, inner._target_pos[0], inner._target_pos[1], inner.target_spd,
                         inner.target_angle, inner.last_pos_diff_len, inner.total_spd, inner.max_spd, fr._count,
                         fr._randomized_force[0], fr._randomized_force[1], fr._randomized_force[2],
                         bc.bodies[robot.robot_id].foot_mu, sr.push_draws, sr.cmd_draws, 0.0, inner.init_pos_diff_len], dtype=np.float64)

    def boxes():
        bx = [np.r_[o.state[0:3], o.box] for o in bc.bodies if o.kind == "static" and getattr(o, "box", None) is not None and np.any(o.box > 0)]
        out = np.zeros((36, 6)); out[:len(bx)] = bx
        return out, len(bx)

    def teleport(pos, yaw):
        from scipy.spatial.transform import Rotation as R
        st = flat_state(robot.get_states_info())
        st[0:3] = pos
        st[3:7] = (R.from_euler("z", yaw) * R.from_quat(np.asarray(init0["base_orn"], dtype=np.float64))).as_quat()
        st[7:13] = 0.0; st[13:25] = init0["joint_pos"]; st[25:37] = 0.0
        st = st.astype(np.float32).astype(np.float64)
        robot.set_states_info({"base_pos": list(st[0:3]), "base_orn": list(st[3:7]), "base_lin_vel": list(st[7:10]),
                               "base_ang_vel": list(st[10:13]), "joint_pos": list(st[13:25]), "joint_vel": list(st[25:37])})
        rec["tp_step"].append(len(rec["reward"])); rec["tp_state"].append(st)

    for ep in range(5):
        obs = env.reset(inter_kwargs={})[0]
        rec["reset_obs"].append(flat_obs(obs)); rec["reset_state"].append(flat_state(robot.get_states_info())); rec["reset_aux"].append(aux())
        b, nb = boxes()
        rec["boxes"].append(b); rec["nbox"].append(nb)
        first = b[2]                                     # first obstacle behind the two walls
        gap = 2 * (b[0, 1] - b[0, 4])
        for t in range(max_steps + 5):
            if ep == 1 and t == 2:                       # standing on / right above the first obstacle (bars: under it), facing +x
                z = 0.31 + (first[2] + first[5] if element != 2 else 0.0)
                teleport([first[0], 0.0, z], 0.0)
            if ep == 2 and t == 2:                       # front feet just before the obstacle's front edge, walking into it; at the wall
                teleport([first[0] - first[3] - 0.21, 0.5 * gap - 0.16, 0.31], 0.0)
            if ep == 3 and t == 4:                       # 0.6 m short of the target, then inside the 0.5 m reach radius
                teleport([inner._target_pos[0] - 0.6, 0.05, 0.33 + (0.25 if element == 3 else 0.0)], 0.0)
            if ep == 3 and t == 9:
                teleport([inner._target_pos[0] - 0.3, 0.05, 0.33 + (0.25 if element == 3 else 0.0)], 0.0)
            if ep == 4 and t == 3:                       # sideways in the corridor, looking at a wall
                teleport([first[0] + 0.9, -0.1, 0.31], np.pi / 2)
            scale = 0.15 if ep != 0 else 0.5
            a = (scale * arng.standard_normal(12)).astype(np.float32)
            o, r, d, info = env.step([{'A_LLC': a.astype(np.float64)}])
            rec["episode"].append(ep); rec["action"].append(a); rec["reward"].append(r[0]); rec["done"].append(bool(d))
            rec["state"].append(flat_state(robot.get_states_info())); rec["aux"].append(aux()); rec["obs"].append(flat_obs(o[0]))
            if d:
                break
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["obs"] = out["obs"].astype(np.float32); out["reset_obs"] = out["reset_obs"].astype(np.float32)
    out["init_state"] = np.concatenate([init0[k] for k in ("base_pos", "base_orn", "base_lin_vel", "base_ang_vel", "joint_pos", "joint_vel")])
    out["seed"] = SEED; out["max_steps"] = max_steps; out["element_id"] = element
    path = os.path.join(ROOT, "tests", "golden", "epmc_e%d_reference_golden.npz" % element)
    np.savez_compressed(path, **out)
    run.real_print("wrote", path, "steps", len(rec["reward"]), "per-episode lengths", np.bincount(np.asarray(rec["episode"])), "boxes", rec["nbox"],
                   "reach steps", np.flatnonzero(np.asarray(rec["reward"]) > 0.2), "size", os.path.getsize(path))
    eng.close()


def main():
    assert os.path.isdir(REF_SRC), "reference tree not mounted"
    # the first engine only exists so that install() finds one; run() swaps in one engine per element
    pybullet_shim.FakeBulletClient.oracle_engine = oracle.make_engine(1, load_model_blob(), None, env_kind=1, kp=50.0, kd=0.5, max_tau=16.0)
    pybullet_shim.install()
    sys.path.insert(0, REF_SRC)
    import builtins
    run.real_print = builtins.print
    builtins.print = lambda *a, **k: None if (a and isinstance(a[0], str) and (a[0].startswith("Current episodic") or a[0].startswith("Terminates"))) else run.real_print(*a, **k)
    from lifelike.utils.constants import STATES_INFO_12_RUN_0
    run.init0 = copy.deepcopy(STATES_INFO_12_RUN_0)
    for element in (1, 2, 3):
        sr = ScriptedRandom()
        np.random.uniform, np.random.rand, np.random.randint = sr.uniform, sr.rand, sr.randint
        run(element, sr)


if __name__ == "__main__":
    main()
