#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched PMC mocap-tracking rollout (BASELINE.json configs[1]).

One "step" = one policy step (10 physics sub-steps + mocap + observation + reward + termination + auto-reset)
of every environment of the batch.  Contract: see the task statement / DESIGN.md 7.

    python bench.py --gpus 1 --steps 512 --warmup 32            # our CUDA engine
    python bench.py --impl reference --steps 20 --warmup 3      # CPU arm (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus 8 ...                                    # env shards, one process per GPU
"""
import argparse
import json
import os

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"                  # NCCL's version banner goes to stdout: keep that to the one JSON line
os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # CPU arm: idle OpenMP threads must not spin away the container's CPU quota
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_ENV_STEP = 1676          # 157 words read + 262 words written, fp32
MU_A = np.array([.0124, -.011, -.0793, -.0125, -.0108, -.0806, .0402, -.0505, -.1956, -.0433, -.0515, -.2156], np.float32)
SIGMA_A = np.array([.0853, .1525, .1747, .0847, .1503, .1766, .1025, .2023, .3701, .1021, .2035, .426], np.float32)
TRAJ_WIDTH = 223                        # obs 207 | action 12 | reward | done | neglogp | value  
UNROLL = 128                            # example_pmc_train.sh:145
METRIC = {"pmc": "env-steps/sec PMC mocap-tracking", "epmc": "env-steps/sec EPMC playground",
          "sepmc": "env-steps/sec SEPMC chase-tag game (one env = one pair of robots, shipped empty arena)"}
WORKLOAD = {"pmc": "4096-env batched PMC mocap-tracking, flat ground, per GPU (BASELINE configs[1])",
            "epmc": "8192-env batched EPMC playground (BASELINE configs[2]), per GPU; --element 3 (default) = corridor with cube steps, 1 = hurdles, "
                    "2 = bars, 0 = the flat joystick arena example_epmc_train.sh ships; the reference has box terrain, no heightfield",
            "sepmc": "2-agent SEPMC chase-tag game, 4096 env-pairs (8192 robots) per GPU, arena of example_sepmc_train.sh (BASELINE configs[4])"}
# algorithmic bytes per env-step: PMC 157 words read + 262 written; EPMC without a terrain box list: 177 read + 991 written
# SEPMC per pair-step: 2 robots x (182 words read + 1052 written: state, history, aux, the 965-wide observation)
ALGO_BYTES = {"pmc": 1676, "epmc": 4672, "sepmc": 9872}
OBS_W = {"pmc": 207, "epmc": 916, "sepmc": 965}
ROBOTS_PER_ENV = {"pmc": 1, "epmc": 1, "sepmc": 2}
ELEMENT = [3]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU (BASELINE configs[1]: 4096)")
    ap.add_argument("--no-gather", action="store_true", help="N>1: skip the NCCL trajectory gather to rank 0")
    ap.add_argument("--block", type=int, default=0, help="CUDA block size override (32/64/128)")
    ap.add_argument("--cpu-envs", type=int, default=4096, help="CPU arm: environments per step (default: the same 4096-env batch as the GPU arm)")
    ap.add_argument("--element", type=int, default=3, help="EPMC element_id (0 flat joystick arena, 1 hurdles, 2 bars, 3 cubes)")
    ap.add_argument("--env", default="pmc", choices=["pmc", "epmc", "sepmc"],
                    help="pmc = BASELINE configs[1] (headline); epmc = configs[2] on the flat element-0 arena (8192 envs); "
                         "sepmc = configs[4] (4096 pairs; --envs counts robots)")
    a = ap.parse_args()
    ELEMENT[0] = a.element
    if a.env in ("epmc", "sepmc") and a.envs == 4096:
        a.envs = 8192
    return a


def synthetic_inputs(n_clips=66):
    from lifelike_agility_and_play_b200.model.compile_model import load_model_blob
    from lifelike_agility_and_play_b200.mocap import synthetic_mocap
    return load_model_blob(), synthetic_mocap(n_clips, seed=0)


def action_pool_np(n, count, seed):
    rng = np.random.default_rng(seed)
    a = MU_A + SIGMA_A * rng.standard_normal((count, n, 12)).astype(np.float32)
    return np.clip(a, -1.0, 1.0).astype(np.float32)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_engine(lib_or_none, n, env, **over):
    """Engine for the bench workload on the CUDA library (lib_or_none=None) or a given library (the oracle)."""
    from lifelike_agility_and_play_b200 import _capi as capi
    lib = lib_or_none if lib_or_none is not None else capi.load_cuda_library()
    blob, mocap = synthetic_inputs()
    if env == "epmc":
        from lifelike_agility_and_play_b200.sim_envs.playground_env import INIT_STATE_RUN_0, epmc_engine_config
        erc = {'element_id': ELEMENT[0], 'friction_range': [0.4, 3.0], 'cmd_vary_freq_range': [9999, 10000], 'target_spd_range': [0.5, 3.0],
               'hole_config': {'min_gap_height': 0.25, 'max_gap_height': 0.25},
               'disturb_force_config': {'start_time': 0.5, 'interval_time': 1.0, 'duration_time': 0.2, 'horizontal_force': [0, 50], 'vertical_force': [0, 10]}}
        cfg = epmc_engine_config(50.0, 50.0, 0.5, 16, 1000, erc)       # train_scripts/example_epmc_train.sh:88-117
        cfg.update(over)
        eng = capi.VecEngine(lib, n, blob, None, **cfg)
        eng.set_init_state(INIT_STATE_RUN_0)
        return eng
    if env == "sepmc":
        from lifelike_agility_and_play_b200.sim_envs.chase_tag_game_env import sepmc_engine_config
        from lifelike_agility_and_play_b200.sim_envs.playground_env import INIT_STATE_RUN_0
        erc = {'friction_range': [0.4, 3.0],
               'disturb_force_config': {'start_time': 0.5, 'interval_time': 1.0, 'duration_time': 0.2, 'horizontal_force': [0, 50], 'vertical_force': [0, 10]}}
        cfg = sepmc_engine_config(50.0, 50.0, 0.5, 16, 1000, erc)      # train_scripts/example_sepmc_train.sh:94-117
        cfg.update(over)
        eng = capi.VecEngine(lib, n, blob, None, **cfg)
        eng.set_init_state(INIT_STATE_RUN_0)
        return eng
    return capi.VecEngine(lib, n, blob, mocap, **over)


def _cpu_engine(n_envs, env, threads):
    from oracle import oracle
    eng = make_engine(oracle.load(), n_envs, env, seed=1234, auto_reset=1, num_threads=threads)
    eng.reset()
    return eng


def _cpu_quota():
    """CPUs this container may use: min(hardware threads, cgroup CPU quota).  The GPU boxes expose 128 hardware threads under
    a 16-CPU quota; an OpenMP team wider than the quota burns it in spin-waits and gets throttled (measured: 128 threads ->
    3 k env-steps/s, 16-32 threads -> 150-200 k; tools/cpu_arm_sweep.py)."""
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            return max(1, min(ncpu, int(np.ceil(float(q) / float(per))))), ncpu
    except Exception:
        pass
    try:
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            return max(1, min(ncpu, int(np.ceil(q / per)))), ncpu
    except Exception:
        pass
    return ncpu, ncpu


def _cpu_sustained(eng, pool, seconds, min_steps=3):
    """Steps for `seconds`; returns the median step time of the second half of the window (sustained, after the OpenMP team
    and the cores are up to speed and any CPU-quota burst is spent)."""
    hist, t_start = [], time.perf_counter()
    while len(hist) < min_steps or time.perf_counter() - t_start < seconds:
        t0 = time.perf_counter()
        eng.step(pool[len(hist) % len(pool)])
        hist.append(time.perf_counter() - t0)
    return float(np.median(hist[len(hist) // 2:]))


def time_cpu_arm(n_envs, steps, warmup, threads=0, env="pmc"):
    """Oracle port of the reference step on the host cores (kind 'port': the reference itself is Python over the
    pybullet wheel, which is not installable here -- DESIGN.md 6).  The CPU arm gets its best sustained configuration: a
    short auto-tune over (envs per call: whole batch or cache-sized blocks of 256; OpenMP threads: the container's CPU quota,
    twice that, or every hardware thread) picks the fastest, then `steps` calls of it are timed after a >= 1 s warm-up.
    Returns (env-steps/s, seconds, threads used, envs per call)."""
    quota, ncpu = _cpu_quota()
    rpe = ROBOTS_PER_ENV[env]
    ths = [threads] if threads > 0 else sorted({quota, min(ncpu, 2 * quota), ncpu})
    cands = [(ne, th) for ne in sorted({n_envs, min(256, n_envs)}) for th in ths]
    best = None
    for ne, th in cands:
        eng = _cpu_engine(ne, env, th)
        rate = (ne // rpe) / _cpu_sustained(eng, action_pool_np(ne, 8, 5678), 1.5)
        if best is None or rate > best[0]:
            best = (rate, ne, th)
        eng.close()
    _, ne, th = best
    eng = _cpu_engine(ne, env, th)
    pool = action_pool_np(ne, 8, 5678)
    _cpu_sustained(eng, pool, 1.0, min_steps=max(3, warmup))
    t0 = time.perf_counter()
    for k in range(steps):
        eng.step(pool[k % 8])
    dt = time.perf_counter() - t0
    eng.close()
    return (ne // rpe) * steps / dt, dt, th, ne


def run_reference(args, rank):
    if rank != 0:
        return
    val, dt, cores, n = time_cpu_arm(args.cpu_envs, args.steps, args.warmup, env=args.env)
    sample = ("%d envs x %d steps of the same workload after a >= 1 s warm-up, oracle/libllq_cpu.so, OpenMP over envs; "
              "(envs per call, threads) = (%d, %d) picked by a short auto-tune" % (n, args.steps, n, cores))
    line = {
        "impl": "reference", "metric": METRIC[args.env], "value": val, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.env] + "; CPU arm steps a %d-env sample" % n, "envs_per_step": n},
        "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "cpu_quota": _cpu_quota()[0], "hw_threads": _cpu_quota()[1]},
        "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from lifelike_agility_and_play_b200 import _capi as capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs                           # robots
    nu = n // ROBOTS_PER_ENV[args.env]      # env-steps per engine step (SEPMC: pairs)
    ow = OBS_W[args.env]
    traj_w = ow + 16                        # obs | action 12 | reward | done | neglogp | value
    eng = make_engine(None, n, args.env, device=local_rank, seed=1234, auto_reset=1, global_env_offset=rank * n)
    if args.block:
        eng.set_option("block", args.block)
    eng.reset()

    POOL = 16
    pool = torch.from_numpy(action_pool_np(n, POOL, 5678 + rank)).to(dev)
    do_gather = world > 1 and not args.no_gather
    slab = torch.zeros((UNROLL, n, traj_w), device=dev, dtype=torch.float32)         # [T, N_local, obs+16] send slab
    reward = torch.zeros((n,), device=dev, dtype=torch.float32)
    done = torch.zeros((n,), device=dev, dtype=torch.uint8)
    recv = None
    if do_gather and rank == 0:
        recv = [torch.empty_like(slab) for _ in range(world)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)          # > 126 MB L2
    # everything below runs on one explicit (non-default) stream: the engine launches on it, the CUDA events are
    # recorded on it (torch.cuda.Event only sees torch's current stream)
    bench_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(bench_stream)
    stream = bench_stream.cuda_stream
    assert stream != 0

    def one_step(i):
        t = i % UNROLL
        row = slab[t]
        # the fused kernel writes the observation straight into the trajectory slab (row stride 223 floats)
        eng.step_device(pool[i % POOL].data_ptr(), row.data_ptr(), reward.data_ptr(), done.data_ptr(), obs_ld=traj_w, stream=stream)
        row[:, ow:ow + 12] = pool[i % POOL]
        row[:, ow + 12] = reward
        row[:, ow + 13] = done

    def gather():
        dist.gather(slab, recv, dst=0)

    for i in range(args.warmup):
        one_step(i)
    if do_gather:
        gather()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    c0 = eng.counters()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    gev = []
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                       # L2 flush between timed steps (not timed)
        ev0[i].record()
        one_step(args.warmup + i)
        ev1[i].record()
        if do_gather and (i + 1) % UNROLL == 0:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gather(); b.record()
            gev.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    step_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    gather_ms = sum(a.elapsed_time(b) for a, b in gev)
    c1 = eng.counters()

    # hot (no flush, back-to-back) variant: what a resident rollout loop sees
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for i in range(args.steps):
        one_step(args.warmup + args.steps + i)
    h1.record()
    torch.cuda.synchronize()
    hot_ms = h0.elapsed_time(h1)
    sampler.stop()

    # dominant kernel alone (events inside the engine, on the launching stream), L2 flushed
    eng.set_option("profile", 1)
    ks, kr = [], []
    for i in range(min(args.steps, 64)):
        flush.fill_(i & 0xFF)
        one_step(i)
        torch.cuda.synchronize()
        a, b = eng.timing()
        ks.append(a); kr.append(b)
    eng.set_option("profile", 0)
    kern_ms = float(np.mean(ks))

    tot = torch.tensor([step_ms, gather_ms, hot_ms, kern_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    step_ms, gather_ms, hot_ms, kern_ms = [float(x) for x in tot.tolist()]
    total_ms = step_ms + gather_ms
    total_env_steps = nu * world * args.steps
    value = total_env_steps / (total_ms * 1e-3)

    # end-to-end through the public host API (numpy in, numpy out; H2D + D2H inside the timed region)
    # (actions come from pinned host memory, results are read back into pinned host memory: VecEngine.step_pinned)
    e2e_steps = min(args.steps, 128)
    host_pool = action_pool_np(n, 4, 999 + rank)
    act_p, obs_p, rew_p, done_p = eng.pinned_io()
    for i in range(4):
        act_p[...] = host_pool[i % 4]
        eng.step_pinned(act_p, obs_p, rew_p, done_p)
    eng.sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        act_p[...] = host_pool[i % 4]               # the policy's output lands in the pinned action buffer
        eng.step_pinned(act_p, obs_p, rew_p, done_p)
        _ = float(rew_p[0])                         # host reads the step's result
    eng.sync()
    e2e_s = time.perf_counter() - t0
    # the plain numpy API (pageable buffers, staging copies inside llq_step) for comparison
    out = (np.empty((n, ow), np.float32), np.empty((n,), np.float32), np.empty((n,), np.uint8))
    t1 = time.perf_counter()
    for i in range(32):
        eng.step(host_pool[i % 4], out=out)
    e2e_pageable = nu * 32 / (time.perf_counter() - t1)
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = nu * world * e2e_steps / float(e2e_t.item())

    # row f2: the whole actor loop on the device -- policy forward (csrc/llq_policy.cu, random weights of the shipped architecture)
    # reads the observation rows in place, writes the actions the next fused step consumes; no host round trip
    actor = None
    if args.env == "pmc" and world == 1:
        from lifelike_agility_and_play_b200.policy import DevicePolicy
        prng = np.random.default_rng(42)
        shapes = [(1, 135), (1, 135), (1, 72), (1, 72), (207, 256), (256,), (256, 256), (256,), (256, 1), (1,), (207, 256), (256,), (256, 256), (256,),
                  (256, 32), (32,), (32, 256), (135, 64), (64,), (32, 32), (32,), (96, 256), (256,), (256, 256), (256,), (256, 12), (12,), (1, 12)]
        wts = [(prng.standard_normal(sh) / np.sqrt(sh[0] if len(sh) == 2 and sh[0] > 1 else 1.0)).astype(np.float32) for sh in shapes]
        wts[1] = np.abs(wts[1]) + 0.5; wts[3] = np.abs(wts[3]) + 0.5
        wts[25] *= 0.05                                        # small actions, like a trained policy's
        pol = DevicePolicy(wts, device=local_rank)
        obs_t = torch.zeros((n, ow), device=dev, dtype=torch.float32)
        act_t = torch.zeros((n, 12), device=dev, dtype=torch.float32)
        eng.step_device(pool[0].data_ptr(), obs_t.data_ptr(), reward.data_ptr(), done.data_ptr(), obs_ld=ow, stream=stream)

        def actor_step():
            pol.forward(obs_t.data_ptr(), ow, n, act_t.data_ptr(), None, stream)
            eng.step_device(act_t.data_ptr(), obs_t.data_ptr(), reward.data_ptr(), done.data_ptr(), obs_ld=ow, stream=stream)
        for i in range(8):
            actor_step()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(args.steps):
            actor_step()
        a1.record()
        torch.cuda.synchronize()
        actor_ms = a0.elapsed_time(a1)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(64):
            pol.forward(obs_t.data_ptr(), ow, n, act_t.data_ptr(), None, stream)
        p1.record()
        torch.cuda.synchronize()
        val_t = torch.zeros((n,), device=dev, dtype=torch.float32)
        nlp_t = torch.zeros((n,), device=dev, dtype=torch.float32)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for i in range(64):
            pol.forward_ex(obs_t.data_ptr(), ow, n, act_t.data_ptr(), None, val_t.data_ptr(), nlp_t.data_ptr(), 1, i, stream)
        q1.record()
        torch.cuda.synchronize()
        actor = {"value": nu * args.steps / (actor_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": actor_ms / args.steps,
                 "policy_kernel_ms": p0.elapsed_time(p1) / 64, "policy_kernel_ms_with_value_head_and_sampling": q0.elapsed_time(q1) / 64, "policy": "PMC net 207-256-256-32 VQ(256) + 135/32-96-256-256-12, 3xTF32 mma.sync (fp32-level accuracy), random weights",
                 "note": "policy forward + fused env step, observations and actions stay in HBM (hot L2, no flush)"}
        pol.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = ALGO_BYTES[args.env] * nu / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("pmc_step_kernel_dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": METRIC[args.env], "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.env] + ("; sharded as in configs[3]" if world > 1 else ""),
                   "element_id": args.element if args.env == "epmc" else None, "envs_per_gpu": nu, "global_envs": nu * world, "robots_per_gpu": n, "substeps": 10, "solver_iters": 10,
                   "mocap": "66 synthetic clips, 229k frames" if args.env == "pmc" else None,
                   "auto_reset": True, "prioritized_sample_factor": 3.0 if args.env == "pmc" else None,
                   "actions": "N(mu_a, sigma_a) clipped +-1, device resident",
                   "l2": "flushed (256 MiB write) between timed steps; per-step CUDA events summed",
                   "parallelism": "env shards x%d%s" % (world, ", NCCL gather of [128,N,obs+16] slabs to rank 0 every 128 steps" if do_gather else "")},
        "value_hot_l2": nu * world * args.steps / (hot_ms * 1e-3),
        "value_no_gather": nu * world * args.steps / (step_ms * 1e-3),
        "gather_ms_total": gather_ms, "wall_s_timed_region": wall,
        "e2e": {"value": e2e_val, "unit": "env-steps/s", "h2d_bytes_per_step": n * 12 * 4, "d2h_bytes_per_step": n * (ow * 4 + 4 + 1),
                "steps": e2e_steps, "api": "VecEngine.step_pinned(numpy over page-locked memory) -> llq_step_ex(LLQ_IO_PINNED)",
                "value_pageable_numpy_api": e2e_pageable * world},
        "gpu_launches": int(c1[4] - c0[4]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "pmc_step_kernel<128,%d>" % {"pmc": 0, "epmc": 1 if args.element == 0 else 3, "sepmc": 2}[args.env], "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_env_step": ALGO_BYTES[args.env],
                     "peak_source": peak_src,
                     "note": "latency/issue bound by design (SURVEY 7): ~2e5 flop per 1.7 kB; see profiles/"},
        "clocks": sampler.summary(),
    }
    if actor is not None:
        line["on_device_actor_loop"] = actor
    if world == 1:
        cval, cdt, cores, cne = time_cpu_arm(args.cpu_envs, 64, 3, env=args.env)
        line["cpu_baseline"] = {"value": cval, "unit": "env-steps/s", "cores": cores, "kind": "port", "cpu_quota": _cpu_quota()[0], "hw_threads": _cpu_quota()[1],
                                "sample": "%d envs x 64 steps of the same workload on the host cores after a >= 1 s warm-up (oracle/libllq_cpu.so, OpenMP over envs; envs per call and threads auto-tuned)" % cne}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
