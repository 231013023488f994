"""Accuracy of the engine's own sin/cos (csrc/llq_math.cuh: llq_sincosf), re-evaluated here in float32 with numpy from the
constants in the header: three-term Cody-Waite reduction by pi/2 + degree-7 / degree-8 kernels must stay within 1.5 ulp of
float64 over the range the kernels use it on (joint angles, yaw, half rotation angles: |x| << 100 rad) -- i.e. the accuracy
class of the CUDA library's sincosf, whose Payne-Hanek slow path the kernel no longer carries."""
import os
import re

import numpy as np

f32 = np.float32
HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lifelike_agility_and_play_b200", "csrc", "llq_math.cuh")


def _constants():
    src = open(HDR).read()
    body = src[src.index("LLQ_DI void llq_sincosf("):]
    body = body[:body.index("\n}\n")]
    nums = [float(x.rstrip("f")) for x in re.findall(r"(?<![\w.])-?\d+\.\d+(?:e[+-]?\d+)?f", body)]
    return body, nums


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _sincos(x, k):
    two_over_pi, c1, c2, c3, s3, s2, s1, d4, d3, d2, d1, one = [np.full_like(x, v) for v in k]
    j = np.rint((x * two_over_pi).astype(f32)).astype(f32)
    a = _fma(j, c1, x); a = _fma(j, c2, a); a = _fma(j, c3, a)
    s = (a * a).astype(f32)
    t = _fma(_fma(s3, s, s2), s, s1)
    sa = _fma((t * s).astype(f32), a, a)
    u = _fma(_fma(_fma(d4, s, d3), s, d2), s, d1)
    ca = _fma(u, s, one)
    q = j.astype(np.int64)
    S = np.where(q & 1, ca, sa); C = np.where(q & 1, sa, ca)
    return np.where(q & 2, -S, S), np.where((q + 1) & 2, -C, C)


def test_llq_sincosf_is_within_one_and_a_half_ulp():
    body, k = _constants()
    assert len(k) == 12 and abs(k[0] - 2 / np.pi) < 1e-7 and abs(k[1] + k[2] + k[3] + np.pi / 2) < 1e-15, k
    assert "rintf" in body and "fmaf(j," in body
    rng = np.random.default_rng(0)
    for lo, hi in ((-np.pi / 4, np.pi / 4), (-8.0, 8.0), (-100.0, 100.0), (-1e-3, 1e-3)):
        x = rng.uniform(lo, hi, 1_000_000).astype(f32)
        S, C = _sincos(x, [f32(v) for v in k])
        xs = x.astype(np.float64)
        for got, ref in ((S, np.sin(xs)), (C, np.cos(xs))):
            err = np.abs(got.astype(np.float64) - ref)
            ulp = np.spacing(np.abs(ref).astype(f32)).astype(np.float64)
            assert err.max() < 8e-8 and (err / ulp).max() < 1.5, (lo, hi, err.max(), (err / ulp).max())
    # exact at the points the physics cares about
    S, C = _sincos(np.array([0.0], f32), [f32(v) for v in k])
    assert S[0] == 0.0 and C[0] == 1.0
