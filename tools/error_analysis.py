"""Teacher-forced error distribution CUDA vs oracle; optional variant libraries (run under gpurun)."""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lifelike_agility_and_play_b200 import _capi as capi
from lifelike_agility_and_play_b200.model.compile_model import load_model_blob
from lifelike_agility_and_play_b200.mocap import synthetic_mocap
from oracle import oracle

np.set_printoptions(precision=6, suppress=True, linewidth=220)
MU_A = np.array([.0124, -.011, -.0793, -.0125, -.0108, -.0806, .0402, -.0505, -.1956, -.0433, -.0515, -.2156], np.float32)
SIGMA_A = np.array([.0853, .1525, .1747, .0847, .1503, .1766, .1025, .2023, .3701, .1021, .2035, .426], np.float32)
blob = load_model_blob(); mocap = synthetic_mocap(6, seed=3, min_frames=380, max_frames=700)
n, steps = 2048, int(os.environ.get("STEPS", 8))
libs = [("default", capi.CUDA_LIB_PATH)] + [(os.path.basename(p), p) for p in sorted(glob.glob(os.path.join(os.path.dirname(capi.CUDA_LIB_PATH), "variants", "*.so")))]


def blockrel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b), axis=1) / np.maximum(1.0, np.max(np.abs(b), axis=1))


for name, path in libs:
    lib = capi.LlqLibrary(path)
    gpu = capi.VecEngine(lib, n, blob, mocap, seed=5)
    cpu = oracle.make_engine(n, blob, mocap, seed=5)
    gpu.reset(); cpu.reset()
    rng = np.random.default_rng(0)
    errs, margins, info = [], [], []
    for t in range(steps):
        a = np.clip(MU_A + SIGMA_A * rng.standard_normal((n, 12)).astype(np.float32), -1, 1).astype(np.float32)
        for f in (capi.F_STATE, capi.F_WARMSTART, capi.F_OBS, capi.F_TIME, capi.F_CLIP, capi.F_REWARD_SUM):
            gpu.set(f, cpu.get(f))
        s0 = cpu.get(capi.F_STATE).copy()
        og, rg, dg = gpu.step(a); oc, rc, dc = cpu.step(a)
        sg, sc = gpu.get(capi.F_STATE), cpu.get(capi.F_STATE)
        e = np.maximum.reduce([blockrel(og[:, :99], oc[:, :99]), blockrel(og[:, 135:], oc[:, 135:]), blockrel(sg, sc)])
        m = cpu.get(capi.F_DECISION_MARGIN)
        errs.append(e); margins.append(m)
        w = np.argsort(-e)[:3]
        for i in w:
            d = np.abs(sg[i].astype(np.float64) - sc[i])
            j = int(np.argmax(d))
            info.append((float(e[i]), float(m[i]), t, int(i), j, float(sg[i, j]), float(sc[i, j]), float(np.abs(sc[i, 25:]).max()),
                         int((cpu.get(capi.F_WARMSTART)[i] > 0).sum())))
        mm = dc.astype(np.uint8)
        if mm.any():
            cpu.reset(mm); gpu.reset(mm)
    e = np.concatenate(errs); m = np.concatenate(margins)
    safe = m > 2e-5
    print("== %s: %d samples; safe %d; percentiles of rel err (safe) 50/90/99/99.9/max: %s ; unsafe max %.3g" % (
        name, e.size, safe.sum(), ["%.2e" % np.percentile(e[safe], p) for p in (50, 90, 99, 99.9, 100)], e[~safe].max() if (~safe).any() else 0))
    for margin_thr in (1e-5, 1e-4, 1e-3):
        s2 = m > margin_thr
        print("   margin>%g: n=%d max=%.2e  #>1e-4: %d" % (margin_thr, s2.sum(), e[s2].max(), (e[s2] > 1e-4).sum()))
    info.sort(reverse=True)
    for it in info[:8]:
        print("   err %.2e margin %.2e step %d env %d comp %d gpu %.6f cpu %.6f max|qd| %.2f contacts %d" % it)
    gpu.close(); cpu.close()
