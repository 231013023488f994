"""CPU-arm diagnostic: env-steps/s of the oracle port for (envs per call, OpenMP threads, OpenMP environment) on this host.
Usage: python tools/cpu_arm_sweep.py            (parent: spawns one child per OpenMP environment)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import bench
    ncpu = os.cpu_count()
    out = []
    for ne in (256, 1024, 4096):
        for th in sorted({ncpu, ncpu // 2, ncpu // 4, 32}):
            if th < 1:
                continue
            eng = bench._cpu_engine(ne, "pmc", th)
            pool = bench.action_pool_np(ne, 8, 5678)
            ts, t0 = [], time.perf_counter()
            while time.perf_counter() - t0 < 2.0:
                t1 = time.perf_counter(); eng.step(pool[len(ts) % 8]); ts.append(time.perf_counter() - t1)
            eng.close()
            ts = np.array(ts)
            out.append({"envs": ne, "threads": th, "steps": len(ts), "best_kenvs_s": round(ne / ts.min() / 1e3, 1),
                        "median_kenvs_s": round(ne / np.median(ts) / 1e3, 1), "last10_kenvs_s": round(ne / np.median(ts[-10:]) / 1e3, 1)})
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    subprocess.run("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|^CPU\\(s\\)'; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc", shell=True)
    for name, env in (("default", {}), ("bind", {"OMP_PROC_BIND": "close", "OMP_PLACES": "cores"}), ("passive", {"OMP_WAIT_POLICY": "passive"})):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print("==", name, env)
        try:
            for row in json.loads(r.stdout.strip().splitlines()[-1]):
                print("  ", row)
        except Exception:
            print(r.stdout[-2000:], r.stderr[-2000:])
