"""The oracle's restatement of the reference's *environment logic* against golden vectors produced by the reference's
own Python code (tests/golden/gen_golden_from_reference.py ran the unmodified PrimitiveLevelEnv / LeggedRobot /
MotionLib with pybullet replaced by a shim over the oracle physics).  Physics numerics themselves remain unpinned."""
import os

import numpy as np
import pytest

from lifelike_agility_and_play_b200 import _capi as capi
from lifelike_agility_and_play_b200.mocap import MocapTable

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pmc_reference_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _table(g):
    return MocapTable(g["frames"], g["offsets"].astype(np.int32), float(g["frame_dt"]), ["c%d" % i for i in range(len(g["offsets"]) - 1)])


def _replay(eng, g, check, with_ob=False):
    ep_of_step = g["episode"]
    step = 0
    n_ep = len(g["clip"])
    prob_i = 0
    for ep in range(n_ep):
        obs = eng.reset_to(int(g["clip"][ep]), float(g["time0"][ep]))
        check("reset_prop", obs[0, :99], g["reset_prop"][ep])
        check("reset_future", obs[0, 135:], g["reset_future"][ep])
        check("reset_state", eng.get(capi.F_STATE)[0], g["reset_state"][ep])
        assert np.all(obs[0, 99:135] == 0)
        while step < len(ep_of_step) and ep_of_step[step] == ep:
            o, r, d = eng.step(g["action"][step][None])
            check("prop", o[0, :99], g["prop"][step])
            check("prop_a", o[0, 99:135], g["prop_a"][step])
            check("future", o[0, 135:], g["future"][step])
            check("reward", r[0], g["reward"][step])
            check("state", eng.get(capi.F_STATE)[0], g["state"][step])
            check("kin", eng.get(capi.F_KIN_STATE)[0], g["kin"][step])
            assert abs(eng.get(capi.F_TIME)[0] - g["time"][step]) < 1e-12
            assert bool(d[0]) == bool(g["done"][step]), "done mismatch at step %d" % step
            if with_ob and g["ob_id"][step] >= 0:
                assert int(eng.get(capi.F_OB_ID)[0]) == int(g["ob_id"][step]), "active plate (PLE:262-268) at step %d" % step
            if d[0]:
                p = eng.get(capi.F_SAMPLE_PROB)
                assert np.allclose(p, g["prob"][prob_i], rtol=1e-6, atol=1e-9), "prioritized sampling probabilities (PLE:239-240)"
                prob_i += 1
            step += 1
    assert step == len(ep_of_step) and prob_i == len(g["prob"])


def test_motionlib_constants(gold, make_oracle):
    assert int(gold["margin"]) == 125 and int(gold["num_env_steps"]) == 10
    n = np.diff(gold["offsets"])
    assert np.allclose(gold["max_steps"], (n - 125) * (1 / 120.0) / 0.02)


def test_oracle_replays_reference_golden(gold, make_oracle):
    eng = make_oracle(1, mocap=_table(gold), kp=50.0, kd=0.5, max_tau=18.0, prioritized_sample_factor=3.0)
    worst = {}

    def check(name, got, want):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        # quaternion sign is not observable: compare up to sign
        if name in ("state", "kin", "reset_state"):
            if np.dot(got[3:7], want[3:7]) < 0:
                got = got.copy(); got[3:7] *= -1
        err = np.max(np.abs(got - want) / (1.0 + np.abs(want)))
        worst[name] = max(worst.get(name, 0.0), float(err))
    _replay(eng, gold, check)
    # the C-ABI returns float32: ~6e-8 relative quantisation; the arithmetic itself agrees to ~1e-12
    for k, v in worst.items():
        assert v < 5e-7, (k, v, worst)


GOLD_OB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pmc_obstacle_reference_golden.npz")


def _with_obstacles(eng, g):
    from lifelike_agility_and_play_b200.mocap import obstacle_table
    tab, offs = obstacle_table(_table(g))
    assert list(np.diff(offs)) == [1, 0, 2]                       # clip 0: one jump apex, clip 1: none, clip 2: two
    eng.load_obstacles(tab, offs, (0.025, 0.5, 0.2))             # PLE:184 with obstacle_height = 0.2 (test_primitive_level_env.py:34)
    return eng


def test_oracle_replays_reference_obstacle_golden(make_oracle):
    """set_obstacle=True (example_pmc_train.sh:74): plate placement at the jump apexes, the 0.5 s hand-over rule, and
    'touching the plate ends the episode' -- replayed against the unmodified reference env."""
    g = np.load(GOLD_OB)
    assert bool(g["obstacle"]) and g["ob_hit"].sum() >= 2 and set(g["ob_id"]) >= {0, 1}
    eng = _with_obstacles(make_oracle(1, mocap=_table(g), kp=50.0, kd=0.5, max_tau=18.0, prioritized_sample_factor=3.0), g)
    worst = {}

    def check(name, got, want):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        if name in ("state", "kin", "reset_state") and np.dot(got[3:7], want[3:7]) < 0:
            got = got.copy(); got[3:7] *= -1
        worst[name] = max(worst.get(name, 0.0), float(np.max(np.abs(got - want) / (1.0 + np.abs(want)))))
    _replay(eng, g, check, with_ob=True)
    assert max(worst.values()) < 5e-7, worst
    # every obstacle contact of the reference run ended its episode
    assert np.all(g["done"][g["ob_hit"]])


@pytest.mark.gpu
def test_cuda_obstacle_matches_oracle(make_cuda, make_oracle):
    g = np.load(GOLD_OB)
    n = 256
    gpu = _with_obstacles(make_cuda(n, mocap=_table(g), seed=8), g)
    cpu = _with_obstacles(make_oracle(n, mocap=_table(g), seed=8), g)
    gpu.reset(); cpu.reset()
    rng = np.random.default_rng(0)
    hits = 0
    for t in range(40):
        a = (0.15 * rng.standard_normal((n, 12))).astype(np.float32)
        for f in (capi.F_STATE, capi.F_WARMSTART, capi.F_OBS, capi.F_TIME, capi.F_CLIP, capi.F_REWARD_SUM, capi.F_OB_ID):
            gpu.set(f, cpu.get(f))
        og, rg, dg = gpu.step(a); oc, rc, dc = cpu.step(a)
        margin = cpu.get(capi.F_DECISION_MARGIN)
        assert np.array_equal(gpu.get(capi.F_OB_ID), cpu.get(capi.F_OB_ID))
        assert (dg == dc)[margin > 2.5e-4].mean() > 0.995
        hits += int(dc.sum())
        m = dc.astype(np.uint8)
        if m.any():
            cpu.reset(m); gpu.reset(m)
            assert np.all(gpu.get(capi.F_OB_ID)[m == 1] == 0)
    assert hits > 0


@pytest.mark.gpu
def test_cuda_replays_reference_golden(gold, make_cuda):
    """Same replay through the CUDA engine, teacher-free (open loop within each episode): fp32 drift is allowed to grow
    with the episode, so only the reset observations, the clocks, the kinematic targets and the first steps are tight."""
    eng = make_cuda(1, mocap=_table(gold), kp=50.0, kd=0.5, max_tau=18.0, prioritized_sample_factor=3.0)
    ep_of_step = gold["episode"]
    step = 0
    for ep in range(len(gold["clip"])):
        obs = eng.reset_to(int(gold["clip"][ep]), float(gold["time0"][ep]))
        assert np.allclose(obs[0, :99], gold["reset_prop"][ep], rtol=1e-4, atol=1e-4)
        assert np.allclose(obs[0, 135:], gold["reset_future"][ep], rtol=1e-4, atol=1e-4)
        k = 0
        while step < len(ep_of_step) and ep_of_step[step] == ep:
            o, r, d = eng.step(gold["action"][step][None])
            assert abs(eng.get(capi.F_TIME)[0] - gold["time"][step]) < 1e-12
            kin = eng.get(capi.F_KIN_STATE)[0].astype(np.float64)
            want = gold["kin"][step].copy()
            if np.dot(kin[3:7], want[3:7]) < 0:
                want[3:7] *= -1
            assert np.allclose(kin, want, rtol=2e-4, atol=2e-4), "mocap target at step %d" % step
            assert np.allclose(o[0, 99:135], gold["prop_a"][step], atol=1e-6)
            if k < 3:
                assert np.allclose(o[0, :99], gold["prop"][step], rtol=1e-3, atol=1e-3)
                assert abs(r[0] - gold["reward"][step]) < 1e-3
            k += 1
            step += 1
            if gold["done"][step - 1]:
                break
