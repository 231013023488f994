"""Debug helper: compare CUDA engine against the oracle at sub-step granularity (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lifelike_agility_and_play_b200 import _capi as capi
from lifelike_agility_and_play_b200.model.compile_model import load_model_blob
from lifelike_agility_and_play_b200.mocap import synthetic_mocap
from oracle import oracle

np.set_printoptions(precision=6, suppress=True, linewidth=200)
blob = load_model_blob()
mocap = synthetic_mocap(4, seed=7, min_frames=400, max_frames=600)
n = int(os.environ.get("N", 16))
for substeps in (1, 10):
    gpu = capi.VecEngine(capi.load_cuda_library(), n, blob, mocap, seed=11, substeps=substeps)
    cpu = oracle.make_engine(n, blob, mocap, seed=11, substeps=substeps)
    og, oc = gpu.reset(), cpu.reset()
    print("substeps", substeps, "clip eq", np.array_equal(gpu.get(capi.F_CLIP), cpu.get(capi.F_CLIP)),
          "time diff", np.abs(gpu.get(capi.F_TIME) - cpu.get(capi.F_TIME)).max())
    print(" reset obs maxdiff", np.abs(og - oc).max(), "state maxdiff", np.abs(gpu.get(capi.F_STATE) - cpu.get(capi.F_STATE)).max())
    rng = np.random.default_rng(0)
    for t in range(60):
        a = (0.1 * rng.standard_normal((n, 12))).astype(np.float32)
        gpu.set(capi.F_STATE, cpu.get(capi.F_STATE)); gpu.set(capi.F_WARMSTART, cpu.get(capi.F_WARMSTART))
        gpu.set(capi.F_OBS, cpu.get(capi.F_OBS)); gpu.set(capi.F_TIME, cpu.get(capi.F_TIME))
        og, rg, dg = gpu.step(a)
        oc, rc, dc = cpu.step(a)
        sg, sc = gpu.get(capi.F_STATE), cpu.get(capi.F_STATE)
        ds = np.abs(sg - sc)
        if t % 10 == 0 or ds.max() > 1e-3:
            i = np.unravel_index(np.argmax(ds), ds.shape)
            print(" t", t, "state maxdiff %.3g at %s" % (ds.max(), i), "obs %.3g" % np.abs(og - oc).max(), "rew %.3g" % np.abs(rg - rc).max(),
                  "done", int(dg.sum()), int(dc.sum()), "warm %.3g" % np.abs(gpu.get(capi.F_WARMSTART) - cpu.get(capi.F_WARMSTART)).max(),
                  "ncontacts", (cpu.get(capi.F_WARMSTART) > 0).sum())
            if ds.max() > 1e-2:
                e = i[0]
                print("  gpu", sg[e]); print("  cpu", sc[e]); break
    print(" counters gpu", gpu.counters(), "cpu", cpu.counters())
