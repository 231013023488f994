"""Knee-wheel ground contacts (llq_config.knee_contacts; include/llq.h): the robot's knee wheels (link_*W of max.urdf) carry
it when it kneels.  Why they are in the engine: tools/statistical_pin.py."""
import ctypes as C

import numpy as np
import pytest

from lifelike_agility_and_play_b200 import _capi as capi

KNEEL_Q = np.array([0.0, -1.35, 2.6] * 4, np.float32)      # thighs swung back, shanks folded: the knees are the lowest points
WHEEL_R = np.array([0.028, 0.028, 0.036, 0.036])


def _kneel(eng):
    eng.reset()
    st = eng.get(capi.F_STATE)
    st[:, 0:3] = [0.0, 0.0, 0.16]; st[:, 3:7] = [0, 0, 0, 1]; st[:, 7:13] = 0
    st[:, 13:25] = KNEEL_Q; st[:, 25:37] = 0
    eng.set(capi.F_STATE, st); eng.set(capi.F_WARMSTART, np.zeros((st.shape[0], 32), np.float32))


def _wheel_heights(oracle_lib, eng_h, state37):
    out = np.zeros((32, 3)); n = C.c_int32(0)
    s64 = np.ascontiguousarray(state37, dtype=np.float64)
    oracle_lib.lib.llq_oracle_proxy_positions.restype = C.c_int
    assert oracle_lib.lib.llq_oracle_proxy_positions(eng_h, s64.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 32, C.byref(n)) == 0
    return out[[10, 13, 16, 19], 2]          # proxy table order: ..., hip, foot, wheel per leg (kind 1 rows)


def test_knee_wheels_carry_a_kneeling_robot(oracle_lib, blob, small_mocap):
    res = {}
    for knee in (1, 0):
        eng = capi.VecEngine(oracle_lib, 1, blob, small_mocap, seed=1, auto_reset=0, knee_contacts=knee)
        _kneel(eng)
        for t in range(80):
            eng.step(np.zeros((1, 12), np.float32))
        st = eng.get(capi.F_STATE)[0]
        res[knee] = (float(st[2]), _wheel_heights(oracle_lib, eng._h, st), eng.get(capi.F_WARMSTART)[0].copy())
        eng.close()
    z1, w1, warm1 = res[1]
    z0, w0, warm0 = res[0]
    assert np.all(w1 > WHEEL_R - 0.003) and z1 > 0.0, (z1, w1)            # resting on the wheels (and feet)
    assert np.any(warm1[4:8] > 0)                                          # a remembered impulse that belongs to a knee wheel (spheres 4-7)
    assert np.all(w0 < -0.05) and z0 < -0.1 and np.all(warm0[4:] == 0)      # feet only: the knees pass through the floor


@pytest.mark.gpu
def test_cuda_knee_contacts_match_the_oracle(built, blob, small_mocap, oracle_lib):
    n = 256
    gpu = capi.VecEngine(capi.load_cuda_library(), n, blob, small_mocap, seed=3, auto_reset=0)
    cpu = capi.VecEngine(oracle_lib, n, blob, small_mocap, seed=3, auto_reset=0)
    _kneel(cpu)
    rng = np.random.default_rng(0)
    st = cpu.get(capi.F_STATE)
    st[:, 2] += rng.uniform(-0.03, 0.03, n).astype(np.float32)             # some start above, some already in contact
    st[:, 13:25] += 0.15 * rng.standard_normal((n, 12)).astype(np.float32)
    cpu.set(capi.F_STATE, st)
    gpu.reset()
    worst, knee_steps = 0.0, 0
    from test_parity_gpu import blockrel
    for t in range(40):
        for f in (capi.F_STATE, capi.F_WARMSTART, capi.F_OBS, capi.F_TIME, capi.F_CLIP):
            gpu.set(f, cpu.get(f))
        a = (0.05 * rng.standard_normal((n, 12))).astype(np.float32)
        og, rg, dg = gpu.step(a); oc, rc, dc = cpu.step(a)
        wg, wc = gpu.get(capi.F_WARMSTART), cpu.get(capi.F_WARMSTART)
        e = blockrel(gpu.get(capi.F_STATE), cpu.get(capi.F_STATE))
        ok = e < 1e-4
        worst = max(worst, float(np.percentile(e, 99)))
        knee_steps += int((wc[:, 4:8] > 0).any(1).sum())
        assert ok.mean() > 0.98, (t, ok.mean())
        assert np.array_equal(wg[ok] > 0, wc[ok] > 0) or ((wg[ok] > 0) != (wc[ok] > 0)).mean() < 0.01   # same spheres in contact
    assert knee_steps > 1000                     # the comparison really ran on knee contacts
    print("knee-contact parity: 99th percentile of the state error %.1e over %d env-steps with a wheel contact" % (worst, knee_steps))
    gpu.close(); cpu.close()
