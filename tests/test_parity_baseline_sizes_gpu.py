"""CUDA engine vs CPU oracle, teacher-forced, at exactly the batch sizes of BASELINE.json's single-GPU configurations:
4096 PMC envs on the bench's 66-clip table (configs[1]), 8192 EPMC envs in the cube corridor (configs[2]), 8192 SEPMC robots =
4096 chase-tag pairs (configs[4]).  Each engine pair is aged first (auto-reset rollouts with the benchmark's action distribution,
so that episodes of mixed age, resting contacts, joint-limit rows and -- EPMC / SEPMC -- push windows are present), then >= 3
policy steps are compared from identical (state, contact memory, history, clocks, bookkeeping) with two criteria:

  * per env-step (the bar of tests/test_parity_gpu.py): block-relative error of observation blocks and state below 1e-4.  In the aged
    batch 0.5-2 % of the env-steps have a joint resting ON its stop (|q - limit| ~ 1e-8 rad: Bullet's limit row exists only while the
    stop is violated, the ERP term parks the joint exactly there) or a sphere at the contact threshold: fp32 and fp64 take different
    branches of the step there and the joint velocity differs by O(1).  Those env-steps (oracle decision margin below MARGIN_OK =
    2e-5 rad / m) are counted and reported; of the others at most 0.1 % (PMC) / 2.5 % (EPMC, SEPMC: shipped friction range up to 3.0,
    99th percentile below 1e-3) may exceed 1e-4, and for PMC every deviating env-step must lie within NEAR_BRANCH of a branch;
  * per element: |cuda - oracle| <= 1e-4 |oracle| + 1e-3 for >= 99.9 % of ALL compared numbers (observation entries + 37 state
    entries of every env); the share within + 1e-4 absolute is printed beside it.
"""
import os

import numpy as np
import pytest

from lifelike_agility_and_play_b200 import _capi as capi
from test_parity_gpu import MARGIN_OK, MU_A, SIGMA_A, TOL, blockrel

pytestmark = pytest.mark.gpu
STEPS = int(os.environ.get("LLQ_BASELINE_PARITY_STEPS", 3))


def _actions(rng, n, scale=1.0):
    return np.clip(MU_A + scale * SIGMA_A * rng.standard_normal((n, 12)).astype(np.float32), -1, 1).astype(np.float32)


def _elementwise(pairs):
    a = np.concatenate([np.asarray(x, np.float64).ravel() for x, _ in pairs])
    b = np.concatenate([np.asarray(y, np.float64).ravel() for _, y in pairs])
    d = np.abs(a - b)
    return float((d <= 1e-4 * np.abs(b) + 1e-3).mean()), float((d <= 1e-4 * np.abs(b) + 1e-4).mean()), a.size


def _compare(gpu, cpu, n, fields, segments, rng, pre_roll, label):
    """Both engines run with auto_reset = 1 and are stepped together, the CUDA engine teacher-forced from the oracle before every step
    (resets are bit-exact in their sampling, so finished envs restart identically on both sides); the first `pre_roll` steps only age the
    batch, the last STEPS are compared."""
    cpu.reset(); gpu.reset()
    pmc = cpu.cfg.env_kind == capi.ENV_PMC
    E, M, DD, pairs = [], [], [], []
    for t in range(pre_roll + STEPS):
        a = _actions(rng, n)
        for f in fields:
            gpu.set(f, cpu.get(f))
        if pmc:
            gpu.set(capi.F_AVG_REWARD, cpu.get(capi.F_AVG_REWARD))   # prioritized-sampling table of the aged batch
        og, rg, dg = gpu.step(a)
        oc, rc, dc = cpu.step(a)
        if t < pre_roll:
            continue
        same_ep = gpu.get(capi.F_EPISODE_ID) == cpu.get(capi.F_EPISODE_ID)    # an env whose `done` differed restarted on one side only
        sg, sc = gpu.get(capi.F_STATE), cpu.get(capi.F_STATE)
        e = np.maximum.reduce([blockrel(og[:, lo:hi], oc[:, lo:hi]) for lo, hi in segments] + [blockrel(sg, sc)])
        E.append(np.where(same_ep, e, 0.0)); M.append(cpu.get(capi.F_DECISION_MARGIN)); DD.append((dg != dc) | ~same_ep)
        pairs += [(og[same_ep], oc[same_ep]), (sg[same_ep], sc[same_ep])]
    e, m, dd = np.concatenate(E), np.concatenate(M), np.concatenate(DD)
    p3, p4, cnt = _elementwise(pairs)
    cg, cc = gpu.counters(), cpu.counters()
    amb = m <= MARGIN_OK
    print("%s: %d envs, %d ageing + %d compared policy steps; env-step err 50/99/99.9/max = %.1e %.1e %.1e %.1e; %d above 1e-4 "
          "(%d of them among the %d fp32-ambiguous env-steps with margin <= 2e-5), %d done mismatches; "
          "elements within rtol 1e-4 + atol 1e-3: %.5f, + atol 1e-4: %.5f (of %d); rows solved cuda / oracle: contact %d / %d, limit %d / %d; "
          "episodes finished %d / %d" % (
              label, n, pre_roll, STEPS, np.percentile(e, 50), np.percentile(e, 99), np.percentile(e, 99.9), e.max(), int((e >= TOL).sum()),
              int(((e >= TOL) & amb).sum()), int(amb.sum()), int(dd.sum()),
              p3, p4, cnt, cg[2], cc[2], cg[3], cc[3], cg[1], cc[1]))
    return e, m, dd, p3, p4


def test_pmc_4096_envs_on_the_bench_table(built, blob, oracle_lib):
    from bench import synthetic_inputs
    _, mocap = synthetic_inputs()
    n = 4096
    gpu = capi.VecEngine(capi.load_cuda_library(), n, blob, mocap, seed=1234, auto_reset=1)
    cpu = capi.VecEngine(oracle_lib, n, blob, mocap, seed=1234, auto_reset=1)
    fields = (capi.F_STATE, capi.F_WARMSTART, capi.F_OBS, capi.F_TIME, capi.F_CLIP, capi.F_REWARD_SUM, capi.F_EPISODE_ID)
    e, m, dd, p3, p4 = _compare(gpu, cpu, n, fields, [(0, 99), (135, 207)], np.random.default_rng(11), 40, "PMC configs[1]")
    bad = (e >= TOL) | dd
    assert (bad & (m > MARGIN_OK)).mean() <= 1e-3 and bad.mean() <= 1e-2
    assert np.all(m[bad] < 5e-4), "a deviation > 1e-4 away from any branch of the step: margins %s" % m[bad]
    assert p3 >= 0.999 and p4 >= 0.999
    gpu.close(); cpu.close()


def test_epmc_8192_envs_cube_corridor(built, blob, oracle_lib):
    from bench import make_engine
    n = 8192
    gpu = make_engine(None, n, "epmc", seed=1234, auto_reset=1)
    cpu = make_engine(oracle_lib, n, "epmc", seed=1234, auto_reset=1)
    fields = (capi.F_STATE, capi.F_WARMSTART, capi.F_OBS, capi.F_TIME, capi.F_AUX, capi.F_EPISODE_ID, capi.F_REWARD_SUM)
    e, m, dd, p3, p4 = _compare(gpu, cpu, n, fields, [(0, 135), (135, 460), (460, 588), (588, 913), (913, 916)], np.random.default_rng(12), 12,
                                "EPMC configs[2], element 3, friction range [0.4, 3.0]")
    assert np.array_equal(gpu.get(capi.F_NBOX), cpu.get(capi.F_NBOX))
    bad = (e >= TOL) | dd
    # shipped friction range (foot friction up to 3.0 on ground 1.0): Bullet's 10-sweep Gauss-Seidel is not contractive there and
    # amplifies fp32 rounding (tests/test_parity_epmc_gpu.py measures it per sub-step); in the aged batch 1-2 % of the env-steps land
    # between 1e-4 and 1e-3, none of the non-ambiguous ones far beyond
    assert (bad & (m > MARGIN_OK)).mean() <= 2.5e-2 and np.percentile(e, 99) < 1e-3 and dd.mean() <= 1e-3
    assert p3 >= 0.999 and p4 >= 0.999
    gpu.close(); cpu.close()


def test_sepmc_4096_pairs(built, blob, oracle_lib):
    from bench import make_engine
    n = 8192
    gpu = make_engine(None, n, "sepmc", seed=1234, auto_reset=1)
    cpu = make_engine(oracle_lib, n, "sepmc", seed=1234, auto_reset=1)
    fields = (capi.F_STATE, capi.F_WARMSTART, capi.F_OBS, capi.F_TIME, capi.F_AUX, capi.F_EPISODE_ID, capi.F_REWARD_SUM)
    segs = [(0, 135), (135, 460), (460, 588), (588, 913), (913, 965)]
    e, m, dd, p3, p4 = _compare(gpu, cpu, n, fields, segs, np.random.default_rng(13), 12, "SEPMC configs[4], 4096 pairs")
    bad = (e >= TOL) | dd
    # shipped friction range (foot friction up to 3.0 on ground 1.0): Bullet's 10-sweep Gauss-Seidel is not contractive there and
    # amplifies fp32 rounding (tests/test_parity_epmc_gpu.py measures it per sub-step); in the aged batch 1-2 % of the env-steps land
    # between 1e-4 and 1e-3, none of the non-ambiguous ones far beyond
    assert (bad & (m > MARGIN_OK)).mean() <= 2.5e-2 and np.percentile(e, 99) < 1e-3 and dd.mean() <= 1e-3
    assert p3 >= 0.999 and p4 >= 0.999
    gpu.close(); cpu.close()
