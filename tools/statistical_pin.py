#!/usr/bin/env python
"""Statistical pin of the physics restatement: roll the reference's *shipped, Bullet-trained* PMC policy
(data/models/primitive_level.model, architecture of networks/legged_robot/pmc_net/pmc_net.py:33-178 with the policy_config of
test_scripts/primitive_level/test_primitive_level_env.py:39-57, argmax actions) on the shipped mocap clips inside this repo's
engine and compare (a) how well it tracks -- per-step reward, episode length relative to the clip, termination causes -- and
(b) the distribution of the observations it produces with the running mean / std the model accumulated in the real PyBullet
env (first four arrays of the model file).  A policy trained on Bullet only tracks, and only reproduces its own training
distribution, if the dynamics it meets here behave like Bullet's.

    python tools/statistical_pin.py [--engine oracle|cuda] [--policy host|device] [--envs 64] [--steps 600] [--out profiles/...json]
    python tools/statistical_pin.py --stage DIR      # pack model weights + mocap table into DIR (git-ignored scratch that travels
                                                     # with gpurun), then on the GPU box: --staged DIR --engine cuda --policy device

Needs /root/reference (model + mocap data) or a staged copy of the two; nothing under tests/ depends on it."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from lifelike_agility_and_play_b200 import _capi as capi  # noqa: E402
from lifelike_agility_and_play_b200.mocap import load_mocap, load_packed, save_packed  # noqa: E402
from lifelike_agility_and_play_b200.model.compile_model import load_model_blob  # noqa: E402
from lifelike_agility_and_play_b200.policy import PmcPolicy  # noqa: E402
from load_reference_model import load  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="oracle", choices=["oracle", "cuda"])
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--model", default="/root/reference/data/models/primitive_level.model")
    ap.add_argument("--data", default="/root/reference/data/mocap_data")
    ap.add_argument("--obstacle", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--cfg", default="{}", help="python dict of llq_config overrides, e.g. \"{'contact_erp': 0.2}\"")
    ap.add_argument("--policy", default="host", choices=["host", "device"], help="numpy forward or the CUDA policy kernel (llq_policy.cu)")
    ap.add_argument("--stage", default="", help="write weights.npz + mocap.npz to this directory and exit")
    ap.add_argument("--staged", default="", help="read weights.npz + mocap.npz from this directory instead of /root/reference")
    a = ap.parse_args()
    if a.staged:
        wz = np.load(os.path.join(a.staged, "weights.npz"))
        weights = [wz["w%d" % i] for i in range(28)]
        mocap = load_packed(os.path.join(a.staged, "mocap.npz"))
    else:
        weights = load(a.model).model
        mocap = load_mocap(a.data)
    if a.stage:
        os.makedirs(a.stage, exist_ok=True)
        np.savez(os.path.join(a.stage, "weights.npz"), **{"w%d" % i: np.asarray(w, np.float32) for i, w in enumerate(weights)})
        save_packed(mocap, os.path.join(a.stage, "mocap.npz"))
        print("staged", a.stage)
        return
    pol = PmcPolicy(weights)
    dev_pol = None
    if a.policy == "device":
        import torch
        from lifelike_agility_and_play_b200.policy import DevicePolicy
        dev_pol = DevicePolicy(weights, device=0)
        t_act = torch.zeros((a.envs, 12), device="cuda", dtype=torch.float32)
    if a.engine == "oracle":
        from oracle import oracle
        lib = oracle.load()
    else:
        lib = capi.load_cuda_library()
    eng = capi.VecEngine(lib, a.envs, load_model_blob(), mocap, seed=2024, auto_reset=0, kp=50.0, kd=0.5, max_tau=18.0,
                         prioritized_sample_factor=0.0, **eval(a.cfg))
    if a.obstacle:            # set_obstacle=True, obstacle_height=0.2 of test_primitive_level_env.py:32-33 (plate half extents PLE:184)
        from lifelike_agility_and_play_b200.mocap import obstacle_table
        tab, offs = obstacle_table(mocap)
        eng.load_obstacles(tab, offs, (0.025, 0.5, 0.2))
    obs = eng.reset()
    n = a.envs
    rew_sum, steps_alive = np.zeros(n), np.zeros(n, int)
    ep_len, ep_rew, ep_frac, ep_cause, ep_clip = [], [], [], [], []
    clip0, t0 = eng.get(capi.F_CLIP).copy(), eng.get(capi.F_TIME).copy()
    frames = np.diff(mocap.offsets)
    X = []
    for t in range(a.steps):
        if dev_pol is not None:
            t_obs = torch.from_numpy(np.ascontiguousarray(obs, np.float32)).cuda()
            dev_pol.forward(t_obs.data_ptr(), obs.shape[1], a.envs, t_act.data_ptr(), None, None)
            torch.cuda.synchronize()
            act = t_act.cpu().numpy()
        else:
            act = pol.act(obs)
        X.append(obs[:, :135].copy())
        prev_clip, prev_t0 = clip0.copy(), t0.copy()
        obs, r, d = eng.step(act.astype(np.float32))
        rew_sum += r; steps_alive += 1
        if d.any():
            st, kin = eng.get(capi.F_STATE).astype(np.float64), eng.get(capi.F_KIN_STATE).astype(np.float64)
            obs_r = eng.reset(d.astype(np.uint8))
            obs = np.where(d[:, None] != 0, obs_r, obs)
            clip_now, t_now = eng.get(capi.F_CLIP), eng.get(capi.F_TIME)
            for i in np.flatnonzero(d):
                x, y, z, w = st[i, 3:7] / np.linalg.norm(st[i, 3:7])
                r22 = 1 - 2 * (x * x + y * y)
                left_z = 2 * (x * z + y * w) * 2 * (x * y + z * w) - 2 * (y * z - x * w) * (1 - 2 * (y * y + z * z))
                fall = r22 < 0.5 or abs(left_z) > 0.7071
                dp = float(np.sum((st[i, 0:3] - kin[i, 0:3]) ** 2))
                qd = kin[i, 3:7] / np.linalg.norm(kin[i, 3:7])
                ang = 2 * np.arccos(min(1.0, abs(float(np.dot(qd, [x, y, z, w])))))
                ep_cause.append("fall" if fall else ("diff" if (dp > 1.0 or ang > 1.0) else "end"))
                ep_clip.append(int(prev_clip[i]))
                ep_len.append(int(steps_alive[i])); ep_rew.append(float(rew_sum[i] / steps_alive[i]))
                avail = (frames[prev_clip[i]] - mocap.margin() - 1) * mocap.frame_dt - prev_t0[i]      # seconds of clip left at reset
                ep_frac.append(float(min(1.0, steps_alive[i] * 0.02 / max(avail, 0.02))))
                rew_sum[i] = 0; steps_alive[i] = 0
                clip0[i], t0[i] = clip_now[i], t_now[i]
    X = np.concatenate(X)
    mean, std = X.mean(0), X.std(0)
    names = ["joint_pos"] * 12 + ["joint_vel"] * 12 + ["ang_vel_loc"] * 3 + ["lin_vel_loc"] * 3 + ["e_g"] * 3
    newest = slice(66, 99)                       # newest of the three stacked prop frames
    m_ref, s_ref = pol.prop_mean[newest], pol.prop_std[newest]
    z = (mean[newest] - m_ref) / s_ref
    rep = {
        "engine": a.engine, "policy": a.policy, "envs": n, "steps": a.steps, "env_steps": int(n * a.steps), "episodes_finished": len(ep_len),
        "mean_reward_per_step": float(np.mean(ep_rew)) if ep_rew else None,
        "median_episode_steps": float(np.median(ep_len)) if ep_len else None,
        "fraction_of_remaining_clip_survived_mean": float(np.mean(ep_frac)) if ep_frac else None,
        "episodes_reaching_clip_end_frac": float(np.mean(np.array(ep_frac) > 0.98)) if ep_frac else None,
        "termination_causes": {c: int(sum(1 for x in ep_cause if x == c)) for c in ("end", "fall", "diff")},
        "early_terminations_by_clip": {mocap.names[c]: int(sum(1 for x, y in zip(ep_clip, ep_cause) if x == c and y != "end"))
                                       for c in sorted(set(ep_clip)) if any(x == c and y != "end" for x, y in zip(ep_clip, ep_cause))},
        "episodes_by_clip": {mocap.names[c]: int(sum(1 for x in ep_clip if x == c)) for c in sorted(set(ep_clip))},
        "obs_mean_minus_model_mean_in_model_std": {k: [round(float(v), 3) for v in z[[i for i, nm in enumerate(names) if nm == k]]]
                                                  for k in dict.fromkeys(names)},
        "obs_std_over_model_std": {k: [round(float(v), 3) for v in (std[newest] / s_ref)[[i for i, nm in enumerate(names) if nm == k]]]
                                   for k in dict.fromkeys(names)},
        "action_mean": [round(float(v), 4) for v in X[:, 123:135].mean(0)], "action_mean_model": [round(float(v), 4) for v in pol.prop_mean[123:135]],
        "action_std": [round(float(v), 4) for v in X[:, 123:135].std(0)], "action_std_model": [round(float(v), 4) for v in pol.prop_std[123:135]],
    }
    print(json.dumps(rep, indent=1))
    if a.out:
        json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
