import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from lifelike_agility_and_play_b200 import _capi as capi
from lifelike_agility_and_play_b200.model.compile_model import load_model_blob
from oracle import oracle
from test_golden_epmc import EPMC_CFG, GOLD
from test_parity_gpu import MU_A, SIGMA_A, blockrel
np.set_printoptions(precision=5, suppress=True, linewidth=220)
blob = load_model_blob(); g = np.load(GOLD)
n = 1024
cfg = dict(EPMC_CFG); cfg.update(solver_iters=int(os.environ.get("ITERS", 10)), cmd_freq_lo=3, cmd_freq_hi=9, max_steps=40 * int(os.environ.get("SUB1", "1") == "1" and 10 or 1), substeps=int(os.environ.get("SUB1", "1") == "1" and 1 or 10))
gpu = capi.VecEngine(capi.load_cuda_library(), n, blob, None, seed=7, **cfg); cpu = capi.VecEngine(oracle.load(), n, blob, None, seed=7, **cfg)
for e in (gpu, cpu): e.set_init_state(g["init_state"])
gpu.reset(); cpu.reset()
rng = np.random.default_rng(3)
rows = []
for t in range(int(os.environ.get("T", 12))):
    a = np.clip(MU_A + SIGMA_A * rng.standard_normal((n, 12)).astype(np.float32), -1, 1).astype(np.float32)
    for f in (capi.F_STATE, capi.F_WARMSTART, capi.F_OBS, capi.F_TIME, capi.F_AUX, capi.F_EPISODE_ID, capi.F_REWARD_SUM):
        gpu.set(f, cpu.get(f))
    s0 = cpu.get(capi.F_STATE).copy(); w0 = cpu.get(capi.F_WARMSTART).copy()
    og, rg, dg = gpu.step(a); oc, rc, dc = cpu.step(a)
    sg, sc = gpu.get(capi.F_STATE), cpu.get(capi.F_STATE)
    es = blockrel(sg, sc); e1 = blockrel(og[:, :135], oc[:, :135]); e2 = blockrel(og[:, 135:], oc[:, 135:]); er = np.abs(rg - rc) / np.maximum(1e-4, np.abs(rc)) * 0.1
    m = cpu.get(capi.F_DECISION_MARGIN); aux = cpu.get(capi.F_AUX)
    e = np.maximum.reduce([es, e1, e2, er])
    for i in np.where((e > 1e-4) | (dg != dc))[0]:
        d = np.abs(sg[i].astype(np.float64) - sc[i]); j = int(np.argmax(d))
        rows.append((float(e[i]), t, int(i), float(es[i]), float(e1[i]), float(e2[i]), float(er[i]), float(m[i]), j, float(sg[i, j]), float(sc[i, j]), int(dg[i]), int(dc[i]),
                     float(s0[i, 2]), float(np.abs(s0[i, 25:]).max()), int((w0[i] > 0).sum()), int((cpu.get(capi.F_WARMSTART)[i] > 0).sum()), float(aux[i, 13]), int(aux[i, 9])))
    mm = dc.astype(np.uint8)
    if mm.any(): cpu.reset(mm); gpu.reset(mm)
rows.sort(reverse=True)
print("n outliers", len(rows))
print("err step env | es e_prop e_percep e_rew | margin comp gpu cpu | dg dc | z0 max|qd0| contacts0 contacts1 mu pushcount")
for r in rows[:40]:
    print("%.2e %2d %4d | %.1e %.1e %.1e %.1e | %.1e %2d %9.4f %9.4f | %d %d | %.3f %6.2f %d %d %.2f %d" % r)
