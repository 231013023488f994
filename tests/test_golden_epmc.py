"""EPMC (PlayGroundEnv, element_id 0) -- the oracle against golden vectors produced by the reference's own code
(tests/golden/gen_golden_epmc_from_reference.py: unmodified PlayGroundEnv / PushRandomizer / BulletStatics / LeggedRobot on
the pybullet shim, consuming the engine's Philox streams).  The replay goes through the oracle's own reset()/step()
*sampling* path, so friction / yaw / command / push draws, the push schedule and the joystick logic are all pinned."""
import os

import numpy as np
import pytest

from lifelike_agility_and_play_b200 import _capi as capi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "epmc_reference_golden.npz")
EPMC_CFG = dict(env_kind=1, kp=50.0, kd=0.5, max_tau=16.0, ground_friction=1.0, cmd_freq_lo=25, cmd_freq_hi=40,
                friction_lo=0.4, friction_hi=3.0, push_h_lo=0.0, push_h_hi=50.0, push_v_lo=0.0, push_v_hi=10.0,
                target_spd_lo=0.5, target_spd_hi=3.0, push_start_count=-250, push_interval_steps=499, push_duration_steps=100,
                push_enabled=1)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _obs_of(g, i):
    return np.concatenate([g[k][i] for k in ("prop", "prop_a", "percep_2d", "percep_1d", "percep_front", "target")])


def _replay(eng, g, tol_obs, tol_state, check_state=True):
    eng.set_init_state(g["init_state"])
    step = 0
    worst = 0.0
    for ep in range(len(g["reset_obs"])):
        obs = eng.reset()
        want = g["reset_obs"][ep]
        err = np.max(np.abs(obs[0] - want) / (1 + np.abs(want)))
        assert err < tol_obs, ("reset obs", ep, err)
        aux = eng.get(capi.F_AUX)[0]
        assert np.allclose(aux[[0, 1, 9, 14, 15]], g["reset_aux"][ep][[0, 1, 9, 14, 15]])          # counters: exact
        assert np.allclose(aux[[2, 3, 4, 6, 10, 11, 12, 13]], g["reset_aux"][ep][[2, 3, 4, 6, 10, 11, 12, 13]], rtol=1e-6, atol=1e-9)
        while step < len(g["episode"]) and g["episode"][step] == ep:
            o, r, d = eng.step(g["action"][step][None])
            want = _obs_of(g, step)
            err = np.max(np.abs(o[0] - want) / (1 + np.abs(want)))
            worst = max(worst, err)
            assert err < tol_obs, ("obs", step, err, int(np.argmax(np.abs(o[0] - want))))
            assert abs(r[0] - g["reward"][step]) < tol_obs * 1e-1 + 1e-7, ("reward", step, r[0], g["reward"][step])
            assert bool(d[0]) == bool(g["done"][step]), ("done", step)
            aux = eng.get(capi.F_AUX)[0]
            assert np.allclose(aux[[0, 1, 9, 14, 15]], g["aux"][step][[0, 1, 9, 14, 15]]), ("counters", step, aux, g["aux"][step])
            if check_state:
                st = eng.get(capi.F_STATE)[0].astype(np.float64); ws = g["state"][step].copy()
                if np.dot(st[3:7], ws[3:7]) < 0:
                    ws[3:7] *= -1
                assert np.max(np.abs(st - ws) / (1 + np.abs(ws))) < tol_state, ("state", step)
                assert np.allclose(aux[[2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13]], g["aux"][step][[2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13]], rtol=1e-5, atol=1e-7)
            step += 1
    assert step == len(g["episode"])
    return worst


def test_oracle_replays_reference_epmc_golden(gold, oracle_lib, blob):
    eng = capi.VecEngine(oracle_lib, 1, blob, None, seed=int(gold["seed"]), max_steps=int(gold["max_steps"]), **EPMC_CFG)
    assert eng.obs_dim == 916
    worst = _replay(eng, gold, 5e-7, 5e-7)
    print("oracle vs reference EPMC golden: worst rel obs err %.2e" % worst)
    eng.close()


def test_epmc_push_schedule_k7(oracle_lib, blob, gold):
    """silent for 250 sub-steps, first window = counts 1..99, then every 499 sub-steps a resample + 100 sub-steps."""
    eng = capi.VecEngine(oracle_lib, 1, blob, None, seed=3, substeps=1, max_steps=5000, **EPMC_CFG)
    eng.set_init_state(gold["init_state"])
    eng.reset()
    counts, draws = [], []
    for t in range(1400):
        eng.step(np.zeros((1, 12), np.float32))
        a = eng.get(capi.F_AUX)[0]
        counts.append(int(a[9])); draws.append(int(a[14]))
    counts = np.array(counts); draws = np.array(draws)
    assert counts[0] == -249 and counts[249] == 0 and counts[250] == 1
    assert draws[0] == 1 and draws[250 + 497] == 1 and draws[250 + 498] == 2          # count reaches 499 -> resample, count := 0
    assert counts[250 + 498] == 0 and counts[250 + 499] == 1


@pytest.mark.gpu
def test_cuda_replays_reference_epmc_golden(gold, built, blob):
    """Open loop in fp32: trajectories separate chaotically once the robot flails, so the continuous quantities are only
    compared over the first steps of every episode; resets (incl. the sampled friction / yaw / command frequency / push
    force) are compared tightly.  Teacher-forced parity is in tests/test_parity_epmc_gpu.py."""
    g = gold
    eng = capi.VecEngine(capi.load_cuda_library(), 1, blob, None, seed=int(g["seed"]), max_steps=int(g["max_steps"]), **EPMC_CFG)
    assert eng.obs_dim == 916
    eng.set_init_state(g["init_state"])
    first = {int(ep): int(np.argmax(g["episode"] == ep)) for ep in np.unique(g["episode"])}
    for ep in range(len(g["reset_obs"])):
        if ep > 0:   # the replay of the previous episode was truncated: carry over what persists across resets (target_spd)
            aux = eng.get(capi.F_AUX); aux[0, 4] = g["aux"][first[ep] - 1][4]; eng.set(capi.F_AUX, aux)
        obs = eng.reset()
        want = g["reset_obs"][ep]
        assert np.max(np.abs(obs[0] - want) / (1 + np.abs(want))) < 1e-5
        aux = eng.get(capi.F_AUX)[0]
        assert np.allclose(aux[[0, 1, 9, 14, 15]], g["reset_aux"][ep][[0, 1, 9, 14, 15]])
        assert np.allclose(aux[[2, 3, 4, 6, 10, 11, 12, 13]], g["reset_aux"][ep][[2, 3, 4, 6, 10, 11, 12, 13]], rtol=1e-5, atol=1e-6)
        for t in range(8):
            step = first[ep] + t
            o, r, d = eng.step(g["action"][step][None])
            want = _obs_of(g, step)
            assert np.max(np.abs(o[0] - want) / (1 + np.abs(want))) < 1e-3, ("obs", step)
            assert abs(r[0] - g["reward"][step]) < 1e-6 and bool(d[0]) == bool(g["done"][step])
            aux = eng.get(capi.F_AUX)[0]
            assert np.allclose(aux[[0, 1, 9, 14, 15]], g["aux"][step][[0, 1, 9, 14, 15]])
            assert np.allclose(aux[[2, 3, 4, 5, 10, 11, 12, 13]], g["aux"][step][[2, 3, 4, 5, 10, 11, 12, 13]], rtol=1e-4, atol=1e-4)
    eng.close()


# ------------------------------------------------------------------------------------------------------------------------
# elements 1-3 (hurdles, bars, cubes): tests/golden/gen_golden_epmc_terrain_from_reference.py
def terrain_gold(element):
    return np.load(os.path.join(os.path.dirname(GOLD), "epmc_e%d_reference_golden.npz" % element))


def terrain_cfg(g):
    cfg = dict(EPMC_CFG)
    cfg.update(element_id=int(g["element_id"]), max_steps=int(g["max_steps"]), hole_gap_lo=0.25, hole_gap_hi=0.25,
               wall_width_lo=0.02, wall_width_hi=0.5, wall_gap_lo=1.0, wall_gap_hi=20.0,
               auxiliary_radius=0.02)             # env_randomize_config['auxiliary_radius'] of the generator (example_epmc_train.sh:112)
    return cfg


T_EXACT = [0, 1, 9, 14, 15]                 # counter, cmd_vary_freq, push count, push draws, command draws
T_CONT = [2, 3, 4, 6, 7, 8, 13, 17]         # target xy, target speed, last / init distance, speed stats, friction


@pytest.mark.parametrize("element", [1, 2, 3])
def test_oracle_replays_reference_epmc_terrain_golden(element, oracle_lib, blob):
    """Terrain generation (Philox stream 5), target placement, perception against the box list, foot contacts with boxes,
    the average-speed reward with the reach bonus and termination -- through the oracle's own reset()/step() path."""
    import ctypes as C
    g = terrain_gold(element)
    eng = capi.VecEngine(oracle_lib, 1, blob, None, seed=int(g["seed"]), **terrain_cfg(g))
    eng.set_init_state(g["init_state"])
    lib = oracle_lib.lib
    lib.llq_oracle_set_state64.restype = C.c_int
    tp = {int(s): st for s, st in zip(g["tp_step"], g["tp_state"])}
    step = 0
    for ep in range(len(g["reset_obs"])):
        obs = eng.reset()
        nb = int(eng.get(capi.F_NBOX)[0])
        assert nb == int(g["nbox"][ep])
        bx = eng.get(capi.F_BOXES)[0].reshape(36, 6)
        assert np.allclose(bx[:nb], g["boxes"][ep][:nb], rtol=1e-6, atol=1e-6), ("boxes", ep)
        err = np.max(np.abs(obs[0] - g["reset_obs"][ep]) / (1 + np.abs(g["reset_obs"][ep])))
        assert err < 5e-7, ("reset obs", ep, err, np.argwhere(np.abs(obs[0] - g["reset_obs"][ep]) > 1e-6)[:5])
        aux = eng.get(capi.F_AUX)[0]
        assert np.array_equal(aux[T_EXACT], g["reset_aux"][ep][T_EXACT])
        assert np.allclose(aux[T_CONT], g["reset_aux"][ep][T_CONT], rtol=1e-6, atol=1e-9)
        while step < len(g["episode"]) and g["episode"][step] == ep:
            if step in tp:
                st = np.ascontiguousarray(tp[step], dtype=np.float64)
                assert lib.llq_oracle_set_state64(eng._h, 0, st.ctypes.data_as(C.c_void_p)) == 0
            o, r, d = eng.step(g["action"][step][None])
            tol = 5e-7 if g["aux"][step][9] <= 0 else 1e-4          # inside a push window 1-ulp torque differences get amplified
            err = np.max(np.abs(o[0] - g["obs"][step]) / (1 + np.abs(g["obs"][step])))
            assert err < tol, ("obs", step, err, np.argwhere(np.abs(o[0] - g["obs"][step]) > 1e-6)[:5])
            assert abs(r[0] - g["reward"][step]) < 2 * tol, ("reward", step, r[0], g["reward"][step])
            assert bool(d[0]) == bool(g["done"][step]), ("done", step)
            aux = eng.get(capi.F_AUX)[0]
            assert np.array_equal(aux[T_EXACT], g["aux"][step][T_EXACT]), ("counters", step)
            assert np.allclose(aux[T_CONT], g["aux"][step][T_CONT], rtol=2 * tol, atol=1e-7), ("aux", step, aux[T_CONT], g["aux"][step][T_CONT])
            st = eng.get(capi.F_STATE)[0].astype(np.float64); ws = g["state"][step].copy()
            if np.dot(st[3:7], ws[3:7]) < 0:
                ws[3:7] *= -1
            assert np.max(np.abs(st - ws) / (1 + np.abs(ws))) < tol, ("state", step)
            step += 1
    assert step == len(g["episode"])
    assert (g["reward"] > 0.2).any()                 # the reach bonus occurs in the file
    eng.close()
