"""SEPMC (ChaseTagGameEnv, shipped empty arena) -- the oracle against golden vectors produced by the reference's own code
(tests/golden/gen_golden_sepmc_from_reference.py: unmodified ChaseTagGameEnv / PushRandomizer / BulletStaticsV4 / LeggedRobot
on the pybullet shim, consuming the engine's Philox streams).  The replay goes through the oracle's own pair reset()/step()
sampling path: reset draws, two-robot push schedule, observations of both agents, visibility, flag switches, tag, rewards
and termination are pinned; scenario teleports recorded in the file are re-applied."""
import ctypes as C
import os

import numpy as np
import pytest

from lifelike_agility_and_play_b200 import _capi as capi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sepmc_reference_golden.npz")
SEPMC_CFG = dict(env_kind=2, kp=50.0, kd=0.5, max_tau=16.0, ground_friction=1.0, friction_lo=0.4, friction_hi=3.0,
                 push_h_lo=0.0, push_h_hi=50.0, push_v_lo=0.0, push_v_hi=10.0, push_start_count=-250, push_interval_steps=499,
                 push_duration_steps=100, push_enabled=1)
# obs layout (CTG:101-114)
SEG = dict(prop=(0, 99), prop_a=(99, 135), percept_2d=(135, 460), percept_1d=(460, 588), percept_front=(588, 913), percept_vec=(913, 918),
           oppo_info=(918, 933), oppo_info_cheat=(933, 948), flag_info=(948, 955), flag_info_cheat=(955, 962), with_flag=(962, 964),
           control_spd=(964, 965))
EXACT_AUX = [0, 1, 5, 6, 9, 14, 15]         # counter, with_flag, visible, switch, push count, push draws, flag draws
CONT_AUX = [2, 3, 4, 7, 8, 13]              # flag xy, speed command, speed stats, friction


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def relerr(a, b):
    return float(np.max(np.abs(a - b) / (1 + np.abs(b))))


def align(st, want):
    """q and -q are the same orientation"""
    ws = np.array(want, dtype=np.float64)
    for i in range(len(ws)):
        if np.dot(st[i, 3:7], ws[i, 3:7]) < 0:
            ws[i, 3:7] *= -1
    return ws


def teleports(g):
    tp = {}
    for s, r, st in zip(g["tp_step"], g["tp_robot"], g["tp_state"]):
        tp.setdefault(int(s), []).append((int(r), st))
    return tp


def test_oracle_replays_reference_sepmc_golden(gold, oracle_lib, blob):
    g = gold
    eng = capi.VecEngine(oracle_lib, 2, blob, None, seed=int(g["seed"]), max_steps=int(g["max_steps"]), **SEPMC_CFG)
    assert eng.obs_dim == 965
    eng.set_init_state(g["init_state"])
    lib = oracle_lib.lib
    lib.llq_oracle_set_state64.restype = C.c_int
    tp = teleports(g)
    step, worst = 0, 0.0
    for ep in range(len(g["reset_obs"])):
        obs = eng.reset()
        err = relerr(obs.astype(np.float64), g["reset_obs"][ep])
        assert err < 5e-7, ("reset obs", ep, err, np.argwhere(np.abs(obs - g["reset_obs"][ep]) > 1e-6)[:5])
        aux = eng.get(capi.F_AUX)
        assert np.array_equal(aux[:, EXACT_AUX], g["reset_aux"][ep][:, EXACT_AUX]), (ep, aux[:, EXACT_AUX], g["reset_aux"][ep][:, EXACT_AUX])
        assert np.allclose(aux[:, CONT_AUX], g["reset_aux"][ep][:, CONT_AUX], rtol=1e-6, atol=1e-9)
        assert relerr(eng.get(capi.F_STATE).astype(np.float64), align(eng.get(capi.F_STATE), g["reset_state"][ep])) < 5e-7
        while step < len(g["episode"]) and g["episode"][step] == ep:
            for r, st in tp.get(step, []):
                st = np.ascontiguousarray(st, dtype=np.float64)
                assert lib.llq_oracle_set_state64(eng._h, r, st.ctypes.data_as(C.c_void_p)) == 0
            o, r, d = eng.step(g["action"][step])
            err = relerr(o.astype(np.float64), g["obs"][step])
            # bit-level differences in the PD torque (numpy vs C evaluation order) are amplified by the contact solver once the
            # 50 N per-sub-step random pushes start (sub-step count > 0): identical to fp32 rounding before, 1e-4 inside the window
            tol = 5e-7 if g["aux"][step][0, 9] <= 0 else 1e-4
            if g["aux"][step][0, 9] <= 0:
                worst = max(worst, err)
            assert err < tol, ("obs", step, err, np.argwhere(np.abs(o - g["obs"][step]) > 1e-6)[:5])
            assert np.allclose(r, g["reward"][step], atol=1e-7), ("reward", step, r, g["reward"][step])
            assert bool(d[0]) == bool(d[1]) == bool(g["done"][step]), ("done", step)
            aux = eng.get(capi.F_AUX)
            assert np.array_equal(aux[:, EXACT_AUX], g["aux"][step][:, EXACT_AUX]), ("counters", step, aux[:, EXACT_AUX], g["aux"][step][:, EXACT_AUX])
            assert np.allclose(aux[:, CONT_AUX], g["aux"][step][:, CONT_AUX], rtol=2 * tol, atol=1e-9), ("aux", step)
            st = eng.get(capi.F_STATE).astype(np.float64); ws = align(st, g["state"][step])
            assert relerr(st, ws) < tol, ("state", step)
            step += 1
    assert step == len(g["episode"])
    # the file exercises every event class
    assert (g["aux"][:, 0, 6] > 0).sum() >= 2 and (np.abs(g["reward"]).sum(1) > 0).sum() >= 3 and (g["aux"][:, :, 5].min(1) < 1).any()
    print("oracle vs reference SEPMC golden: worst rel obs err %.2e" % worst)
    eng.close()
