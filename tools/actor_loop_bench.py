#!/usr/bin/env python
"""On-device actor loop with the reference's SHIPPED policy and mocap clips (staged by tools/statistical_pin.py --stage DIR):
RolloutWorker = CUDA policy kernel (value head + Gaussian sampling) -> fused CUDA env step, records written into the trajectory
slab; reports env-steps/s and how many joint-limit / contact rows the solver handled per env and sub-step (the benchmark's
random-action workload keeps far more joints on their stops than a trained policy does).

    python tools/actor_loop_bench.py --staged .scratch/pin [--envs 4096] [--steps 512] [--deterministic]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifelike_agility_and_play_b200 import _capi as capi  # noqa: E402
from lifelike_agility_and_play_b200.mocap import load_packed  # noqa: E402
from lifelike_agility_and_play_b200.model.compile_model import load_model_blob  # noqa: E402
from lifelike_agility_and_play_b200.parallel import RolloutWorker, slab_records  # noqa: E402
from lifelike_agility_and_play_b200.policy import DevicePolicy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--staged", required=True)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--unroll", type=int, default=128)
    ap.add_argument("--deterministic", action="store_true")
    a = ap.parse_args()
    wz = np.load(os.path.join(a.staged, "weights.npz"))
    weights = [wz["w%d" % i] for i in range(28)]
    mocap = load_packed(os.path.join(a.staged, "mocap.npz"))
    eng = capi.VecEngine(capi.load_cuda_library(), a.envs, load_model_blob(), mocap, device=0, seed=7, auto_reset=1,
                         kp=50.0, kd=0.5, max_tau=18.0, prioritized_sample_factor=3.0)
    pol = DevicePolicy(weights, device=0)
    worker = RolloutWorker(eng, pol, a.unroll, "cuda:0", sample=not a.deterministic, seed=3)
    worker.start(eng.reset())

    def run(n):
        for _ in range(n):
            if worker.t == worker.T:
                slab = worker.finish_unroll()
                with torch.cuda.stream(worker.stream):
                    run.last = slab_records(slab, bootstrap_value=worker.bootstrap_value)      # lambda-returns + PMCInputs layout on the device
            worker.step()
    run.last = None
    run(2 * a.unroll)                                   # past the first episodes: steady mix of clips and phases
    worker.stream.synchronize()
    c0 = eng.counters().copy()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(worker.stream)
    run(a.steps)
    e1.record(worker.stream)
    worker.stream.synchronize()
    ms = e0.elapsed_time(e1)
    c = eng.counters() - c0
    rew = worker.buf[:max(worker.t, 1), :, 219].mean().item()
    print(json.dumps({"envs": a.envs, "steps": a.steps, "ms_per_step": ms / a.steps, "env_steps_per_s": a.envs * a.steps / (ms * 1e-3),
                      "sampled_actions": not a.deterministic, "episodes_finished": int(c[1]),
                      "contact_rows_per_env_substep": float(c[2]) / (a.envs * a.steps * 10), "limit_rows_per_env_substep": float(c[3]) / (a.envs * a.steps * 10),
                      "mean_reward_current_unroll": rew, "records_per_unroll": list(run.last.shape) if run.last is not None else None}))


if __name__ == "__main__":
    main()
