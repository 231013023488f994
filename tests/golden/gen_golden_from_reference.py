"""Generate tests/golden/pmc_reference_golden.npz by running the UNMODIFIED reference environment
(/root/reference/src/lifelike/...: create_tracking_game -> PrimitiveLevelEnv / LeggedRobot / MotionLib, with their
scipy arithmetic) in this container.  pybullet / gym / tleague are absent here, so they are replaced by
tests/golden/pybullet_shim.py, whose `stepSimulation` is the oracle's physics sub-step.

    python tests/golden/gen_golden_from_reference.py      (needs /root/reference; run from the repo root)

The fixture freezes, for three episodes on three small synthetic clips: the sampled (clip, time) of every reset, the
float32 actions, and the reference's obs dict / reward / done / env clock / robot state / kinematic state after every
step, plus MotionLib quantities (margin, max_steps, prioritized sampling probabilities after every `done`).
tests/test_golden_reference.py replays the actions through the oracle and requires agreement to ~1e-9 -- i.e. the
oracle's restatement of the reference's environment logic is pinned by the reference itself."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
REF_SRC = "/root/reference/src"

import pybullet_shim  # noqa: E402
from lifelike_agility_and_play_b200.mocap import MocapTable, synthetic_clip  # noqa: E402
from lifelike_agility_and_play_b200.model.compile_model import load_model_blob  # noqa: E402
from oracle import oracle  # noqa: E402


def main(obstacle=False):
    assert os.path.isdir(REF_SRC), "reference tree not mounted"
    rng = np.random.default_rng(42)
    lens = [150, 230, 410]
    clips = [synthetic_clip(n, rng) for n in lens]
    # make clip 1 turn sharply so that slerp / rotvec paths see large angles and w < 0 quaternions
    yaw = np.linspace(0, 3.5, lens[1])
    clips[1][:, 3:7] = np.stack([0 * yaw, 0 * yaw, np.sin(yaw / 2), np.cos(yaw / 2)], 1) * np.where(yaw > 2.0, -1, 1)[:, None]
    if obstacle:
        # jumps: the base rises above 0.5 m twice in clip 2 and once in clip 0 => hurdle plates at the apexes (utils/obstacle.py:16)
        t2 = np.arange(lens[2]); clips[2][:, 2] += 0.25 * np.exp(-((t2 - 60) / 14.0) ** 2) + 0.22 * np.exp(-((t2 - 215) / 14.0) ** 2)
        t0 = np.arange(lens[0]); clips[0][:, 2] += 0.2 * np.exp(-((t0 - 12) / 6.0) ** 2)
    offsets = np.zeros(4, np.int32); offsets[1:] = np.cumsum(lens)
    table = MocapTable(np.concatenate(clips, 0), offsets, 1.0 / 120.0, ["clip_%d.txt" % i for i in range(3)])
    tmp = tempfile.mkdtemp(prefix="llq_golden_")
    for i, c in enumerate(clips):
        with open(os.path.join(tmp, "clip_%d.txt" % i), "w") as f:
            json.dump({"FrameDuration": 1.0 / 120.0, "LegOrder": ["FR", "FL", "HR", "HL"], "Frames": c.tolist()}, f)
    # the JSON round trip is exact for float64 (repr), but re-read to be sure the table equals what MotionLib sees
    env_config = {
        'arena_id': 'LeggedRobotTracking', 'render': False, 'data_path': tmp, 'control_freq': 50.0,
        'prop_type': ['joint_pos', 'joint_vel', 'root_ang_vel_loc', 'root_lin_vel_loc', 'e_g'],
        'prioritized_sample_factor': 3.0, 'set_obstacle': bool(obstacle), 'obstacle_height': 0.2, 'kp': 50.0, 'kd': 0.5, 'max_tau': 18,
        'reward_weights': {'joint_pos': 0.3, 'joint_vel': 0.05, 'end_effector': 0.1, 'root_pose': 0.5, 'root_vel': 0.05},
    }
    eng = oracle.make_engine(1, load_model_blob(), table, kp=50.0, kd=0.5, max_tau=18.0)
    pybullet_shim.FakeBulletClient.oracle_engine = eng
    pybullet_shim.install()
    sys.path.insert(0, REF_SRC)
    import time as _time
    _time.sleep = lambda s: None                      # PLE:241-244 real-time throttle (documented deviation)
    from lifelike.sim_envs.pybullet_envs.create_pybullet_envs import create_tracking_game
    env = create_tracking_game(**env_config)
    inner = env.env
    ml = inner._motion_generator
    rec = {k: [] for k in ("episode", "clip", "time0", "action", "prop", "prop_a", "future", "reward", "done", "time",
                           "state", "kin", "prob", "reset_prop", "reset_future", "reset_state", "ob_id", "ob_hit")}
    np.random.seed(7)
    arng = np.random.default_rng(5)
    for ep in range(6):
        obs = env.reset(inter_kwargs={})[0]
        rec["clip"].append(int(inner.sampled_data_idx)); rec["time0"].append(float(inner.time))
        rec["reset_prop"].append(obs["prop"]); rec["reset_future"].append(obs["future"])
        si = inner._legged_robot.get_states_info()
        rec["reset_state"].append(np.concatenate([si["base_pos"], si["base_orn"], si["base_lin_vel"], si["base_ang_vel"], si["joint_pos"], si["joint_vel"]]))
        for t in range(400):
            a = (0.15 * arng.standard_normal(12)).astype(np.float32)
            o, r, d, info = env.step([a.astype(np.float64)])
            o = o[0]
            si = inner._legged_robot.get_states_info()
            ki = inner._legged_robot_kin.get_states_info()
            rec["episode"].append(ep); rec["action"].append(a); rec["prop"].append(o["prop"]); rec["prop_a"].append(o["prop_a"])
            rec["future"].append(o["future"]); rec["reward"].append(r[0]); rec["done"].append(bool(d)); rec["time"].append(inner.time)
            rec["state"].append(np.concatenate([si["base_pos"], si["base_orn"], si["base_lin_vel"], si["base_ang_vel"], si["joint_pos"], si["joint_vel"]]))
            rec["kin"].append(np.concatenate([ki["base_pos"], ki["base_orn"], ki["base_lin_vel"], ki["base_ang_vel"], ki["joint_pos"], ki["joint_vel"]]))
            rec["ob_id"].append(getattr(inner, "ob_id", 0) if inner._obstacle is not None else -1)
            rec["ob_hit"].append(len(inner._bullet_client.getContactPoints(bodyA=0)) > 0)
            if d:
                rec["prob"].append(np.array(inner._prioritized_sample_probability, dtype=np.float64))
                break
    out = {k: np.asarray(v) for k, v in rec.items()}
    out.update(frames=table.frames, offsets=table.offsets, frame_dt=table.frame_dt, margin=ml.margin,
               max_steps=np.asarray(ml.max_steps), num_env_steps=inner.num_env_steps)
    out["obstacle"] = bool(obstacle)
    path = os.path.join(ROOT, "tests", "golden", "pmc_obstacle_reference_golden.npz" if obstacle else "pmc_reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "steps", len(rec["reward"]), "episodes", len(rec["clip"]), "dones", int(np.sum(rec["done"])),
          "clips", rec["clip"], "obstacle hits", int(np.sum(rec["ob_hit"])), "ob ids", sorted(set(rec["ob_id"])), "size", os.path.getsize(path))


if __name__ == "__main__":
    main(obstacle="--obstacle" in sys.argv)
