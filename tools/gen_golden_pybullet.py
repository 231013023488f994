#!/usr/bin/env python
"""Dump golden trajectories from the UNMODIFIED reference PMC env on a host that has PyBullet (not this image).

    pip install pybullet==3.2.5 gym==0.21 scipy numpy      # on the PyBullet host (python 3.7-3.9)
    PYTHONPATH=<reference>/src python tools/gen_golden_pybullet.py --data <reference>/data/mocap_data --out pmc_pybullet_golden.npz

It seeds numpy, stubs `tleague.utils.logger` (only used for one log line, motion_lib.py:7,29), removes the real-time
sleep (primitive_level_env.py:241-244) and records, per policy step: action, obs dict, reward, done, env clock, dynamic
and kinematic robot state, foot positions, and per *sub-step* the dynamic state (by wrapping stepSimulation).  Replaying
the file in the style of tests/test_golden_reference.py closes the physics parity that is unpinned in this environment
(DESIGN.md 6): compare against oracle/libllq_cpu.so first, then re-tune llq_config (contact_erp, contact_breaking, ...)."""
import argparse
import sys
import types

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True)
    ap.add_argument("--out", default="pmc_pybullet_golden.npz")
    ap.add_argument("--episodes", type=int, default=8)
    ap.add_argument("--max-steps", type=int, default=200)
    a = ap.parse_args()
    tl, tlu, lg = types.ModuleType("tleague"), types.ModuleType("tleague.utils"), types.ModuleType("tleague.utils.logger")
    lg.log = lambda *x, **k: None
    tlu.logger, tl.utils = lg, tlu
    sys.modules.update({"tleague": tl, "tleague.utils": tlu, "tleague.utils.logger": lg})
    import time
    time.sleep = lambda s: None
    from lifelike.sim_envs.pybullet_envs.create_pybullet_envs import create_tracking_game
    cfg = {'arena_id': 'LeggedRobotTracking', 'render': False, 'data_path': a.data, 'control_freq': 50.0,
           'prop_type': ['joint_pos', 'joint_vel', 'root_ang_vel_loc', 'root_lin_vel_loc', 'e_g'],
           'prioritized_sample_factor': 3.0, 'set_obstacle': False, 'kp': 50.0, 'kd': 0.5, 'max_tau': 18,
           'reward_weights': {'joint_pos': 0.3, 'joint_vel': 0.05, 'end_effector': 0.1, 'root_pose': 0.5, 'root_vel': 0.05}}
    env = create_tracking_game(**cfg)
    inner = env.env
    bc, robot = inner._bullet_client, inner._legged_robot
    sub = []
    real_step = bc.stepSimulation

    def flat(si):
        return np.concatenate([si["base_pos"], si["base_orn"], si["base_lin_vel"], si["base_ang_vel"], si["joint_pos"], si["joint_vel"]])

    def wrapped():
        real_step()
        sub.append(flat(robot.get_states_info()))
    bc.stepSimulation = wrapped
    rec = {k: [] for k in ("episode", "clip", "time0", "action", "prop", "prop_a", "future", "reward", "done", "time", "state", "kin",
                           "substates", "feet")}
    np.random.seed(7)
    rng = np.random.default_rng(5)
    mu = np.array([.0124, -.011, -.0793, -.0125, -.0108, -.0806, .0402, -.0505, -.1956, -.0433, -.0515, -.2156])
    sg = np.array([.0853, .1525, .1747, .0847, .1503, .1766, .1025, .2023, .3701, .1021, .2035, .426])
    for ep in range(a.episodes):
        env.reset()
        rec["clip"].append(int(inner.sampled_data_idx)); rec["time0"].append(float(inner.time))
        for t in range(a.max_steps):
            act = np.clip(mu + sg * rng.standard_normal(12), -1, 1).astype(np.float32)
            del sub[:]
            o, r, d, _ = env.step([act.astype(np.float64)])
            o = o[0]
            robot.compute_end_effector_info()
            rec["episode"].append(ep); rec["action"].append(act); rec["prop"].append(o["prop"]); rec["prop_a"].append(o["prop_a"])
            rec["future"].append(o["future"]); rec["reward"].append(r[0]); rec["done"].append(bool(d)); rec["time"].append(inner.time)
            rec["state"].append(flat(robot.get_states_info())); rec["kin"].append(flat(inner._legged_robot_kin.get_states_info()))
            rec["substates"].append(np.array(sub)); rec["feet"].append(robot.end_effector_position.copy())
            if d:
                break
    np.savez_compressed(a.out, **{k: np.asarray(v) for k, v in rec.items()})
    print("wrote", a.out, len(rec["reward"]), "steps")


if __name__ == "__main__":
    main()
