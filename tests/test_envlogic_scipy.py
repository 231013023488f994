"""T1: the oracle's observation / reward / mocap arithmetic against an independent numpy + scipy transliteration of
primitive_level_env.py:247-426 and motion_lib.py:88-166 (scipy is the library the reference itself uses)."""
import numpy as np
from scipy.spatial.transform import Rotation as R, Slerp

from lifelike_agility_and_play_b200 import _capi as capi
from helpers import foot_positions


def ref_mocap_state(fc, fn, frac, dt):
    """motion_lib.py:88-166."""
    pos = fc[0:3] + frac * (fn[0:3] - fc[0:3])
    orn = Slerp([0, 1], R.from_quat([fc[3:7], fn[3:7]]))(frac).as_quat()
    lin = (fn[0:3] - fc[0:3]) / dt
    rv = (R.from_quat(fn[3:7]) * R.from_quat(fc[3:7]).inv()).as_rotvec()
    angle = np.sqrt(np.sum(rv ** 2)); axis = rv / (angle + 1e-8)
    ang = axis * angle / dt
    jp = fc[7:] + frac * (fn[7:] - fc[7:]); jv = (fn[7:] - fc[7:]) / dt
    return np.concatenate([pos, orn, lin, ang, jp, jv])


def ref_future(mocap, clip, frame_id, frac, base_pos, base_orn):
    """motion_lib.py:75-86 + primitive_level_env.py:299-317."""
    dt = mocap.frame_dt
    frames = mocap.clip(clip)[frame_id:frame_id + 122]
    rb = R.from_quat(base_orn)
    out = []
    for tf in (1. / 30., 1. / 15., 1. / 3., 1.):
        t = dt * frac + tf
        fid = int(np.floor(t / dt)); ff = t / dt - fid
        s = ref_mocap_state(frames[fid], frames[fid + 1], ff, dt)
        rv = R.from_quat((rb.inv() * R.from_quat(s[3:7])).as_quat()).as_rotvec()
        angle = np.sqrt(np.sum(rv ** 2)); axis = rv / (angle + 1e-8)
        out += list(rb.inv().apply(s[0:3] - base_pos)) + list(axis * angle) + list(s[13:25])
    return np.array(out)


def ref_prop(st):
    rb = R.from_quat(st[3:7])
    return np.concatenate([st[13:25], st[25:37], rb.inv().apply(st[10:13]), rb.inv().apply(st[7:10]), rb.as_matrix()[2, :]])


def ref_reward(model, dyn, kin, w=(0.3, 0.05, 0.1, 0.5, 0.05)):
    w = np.array(w) / np.sum(w)
    r_jp = np.exp(-1.0 * np.sum((dyn[13:25] - kin[13:25]) ** 2))
    r_jv = np.exp(-0.1 * np.sum((dyn[25:37] - kin[25:37]) ** 2))
    r_ee = np.exp(-40.0 * np.sum((foot_positions(model, dyn) - foot_positions(model, kin)) ** 2))
    rv = R.from_quat((R.from_quat(kin[3:7]) * R.from_quat(dyn[3:7]).inv()).as_quat()).as_rotvec()
    angle = np.sqrt(np.sum(rv ** 2))
    r_pose = np.exp(-20.0 * np.sum((dyn[0:3] - kin[0:3]) ** 2) - 10.0 * angle ** 2)
    r_vel = np.exp(-2 * np.sum((dyn[7:10] - kin[7:10]) ** 2) - 0.2 * np.sum((dyn[10:13] - kin[10:13]) ** 2))
    return float(w @ np.array([r_jp, r_jv, r_ee, r_pose, r_vel])), angle


def test_reset_state_and_obs_match_scipy(make_oracle, small_mocap):
    n = 32
    eng = make_oracle(n, seed=12)
    obs = eng.reset()
    clip, t0 = eng.get(capi.F_CLIP), eng.get(capi.F_TIME)
    st = eng.get(capi.F_STATE).astype(np.float64)
    dt = small_mocap.frame_dt
    for i in range(n):
        nf = len(small_mocap.clip(clip[i]))
        assert 0 <= t0[i] < dt * (nf - 125 - 1)                                     # ML:50-51
        fid = int(np.floor(t0[i] / dt)); frac = (t0[i] - fid * dt) / dt           # ML:52-53
        want = ref_mocap_state(small_mocap.clip(clip[i])[fid], small_mocap.clip(clip[i])[fid + 1], frac, dt)
        got = st[i].copy()
        if np.dot(got[3:7], want[3:7]) < 0:
            got[3:7] *= -1
        assert np.allclose(got, want, rtol=2e-6, atol=2e-6)
        p = ref_prop(want)
        assert np.allclose(obs[i, 0:33], p, rtol=2e-6, atol=2e-6) and np.allclose(obs[i, 66:99], p, rtol=2e-6, atol=2e-6)
        fut = ref_future(small_mocap, clip[i], fid, frac, want[0:3], want[3:7])
        assert np.allclose(obs[i, 135:], fut, rtol=3e-6, atol=3e-6)


def test_step_reward_done_match_scipy(model, make_oracle, small_mocap):
    n = 24
    eng = make_oracle(n, seed=2)
    eng.reset()
    rng = np.random.default_rng(3)
    dt = small_mocap.frame_dt
    for t in range(12):
        t_before = eng.get(capi.F_TIME).copy()
        a = (0.2 * rng.standard_normal((n, 12))).astype(np.float32)
        obs, rew, done = eng.step(a)
        dyn = eng.get(capi.F_STATE).astype(np.float64); kin = eng.get(capi.F_KIN_STATE).astype(np.float64)
        clip = eng.get(capi.F_CLIP)
        for i in range(n):
            # mocap clock lags by one sub-step (PLE:208-210): the frame cursor uses the time before the last increment
            tl = t_before[i]
            for _ in range(9):
                tl += 0.002
            fid = int(np.floor(tl / dt)); frac = (tl - fid * dt) / dt
            want_kin = ref_mocap_state(small_mocap.clip(clip[i])[fid], small_mocap.clip(clip[i])[fid + 1], frac, dt)
            k = kin[i].copy()
            if np.dot(k[3:7], want_kin[3:7]) < 0:
                k[3:7] *= -1
            assert np.allclose(k, want_kin, rtol=3e-6, atol=3e-6)
            r, angle = ref_reward(model, dyn[i], kin[i])
            assert abs(r - rew[i]) < 2e-5
            assert np.allclose(obs[i, 66:99], ref_prop(dyn[i]), rtol=1e-5, atol=1e-5)
            assert np.allclose(obs[i, 135:], ref_future(small_mocap, clip[i], fid, frac, dyn[i, 0:3], dyn[i, 3:7]), rtol=1e-4, atol=2e-5)
            rot = R.from_quat(dyn[i, 3:7]).as_matrix()
            fwd, up = rot[:, 0], rot[:, 2]
            left_z = up[0] * fwd[1] - up[1] * fwd[0]
            fall = abs(left_z) > np.sin(np.pi / 4) or up[2] < 0.5                                    # LR:171-178, K4
            ended = fid >= len(small_mocap.clip(clip[i])) - 125 - 1                                   # ML:168-172
            diff = abs(angle) > 1.0 or np.sum((dyn[i, 0:3] - kin[i, 0:3]) ** 2) > 1.0                # PLE:319-335
            assert bool(done[i]) == bool(fall or ended or diff)
        m = done.astype(np.uint8)
        if m.any():
            eng.reset(m)


def test_reward_is_one_when_tracking_is_perfect(model, make_oracle):
    """K1: every term is exp(0) and the weights are renormalised to 1."""
    kin = np.zeros(37); kin[2] = 0.33; kin[6] = 1.0
    kin[13:25] = [-0.03, -0.78, 1.69] * 2 + [-0.03, -0.73, 1.57] * 2
    r, _ = ref_reward(model, kin, kin, w=(0.6, 0.05, 0.1, 0.15, 0.1))
    assert abs(r - 1.0) < 1e-12


def test_prioritized_sampling_rule(make_oracle):
    """PLE:235-240 with the batched tie rule: the highest finished env index owning a clip wins its slot."""
    n = 8
    eng = make_oracle(n, seed=0, prioritized_sample_factor=3.0)
    eng.reset_to(np.array([0, 0, 1, 1, 2, 2, 3, 3]), np.full(n, 0.1))
    st = eng.get(capi.F_STATE); st[2:6, 2] += 5.0        # envs 2..5 are thrown away from the reference => done (pos err)
    eng.set(capi.F_STATE, st)
    eng.set(capi.F_REWARD_SUM, np.arange(n, dtype=np.float32))
    o, r, d = eng.step(np.zeros((n, 12), np.float32))
    assert list(d) == [0, 0, 1, 1, 1, 1, 0, 0]
    avg = eng.get(capi.F_AVG_REWARD); prob = eng.get(capi.F_SAMPLE_PROB)
    rs = eng.get(capi.F_REWARD_SUM)
    from lifelike_agility_and_play_b200.mocap import synthetic_mocap
    mc = synthetic_mocap(6, seed=3, min_frames=380, max_frames=700)
    ms = (np.diff(mc.offsets) - 125) * mc.frame_dt / 0.02
    want = np.zeros(6); want[1] = rs[3] / ms[1]; want[2] = rs[5] / ms[2]
    assert np.allclose(avg, want, rtol=1e-6)
    p = (1 - want) ** 3; p /= p.sum()
    assert np.allclose(prob, p, rtol=1e-9)
