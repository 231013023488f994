"""Throughput vs number of environments per GPU, and variant-library timing (run under gpurun)."""
import sys, os, glob, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lifelike_agility_and_play_b200 import _capi as capi
from bench import synthetic_inputs, action_pool_np

blob, mocap = synthetic_inputs()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
libs = [("default", capi.CUDA_LIB_PATH)] + [(os.path.basename(p), p) for p in sorted(glob.glob(os.path.join(os.path.dirname(capi.CUDA_LIB_PATH), "variants", "libllq_cuda*.so")))]
envs = [int(x) for x in os.environ.get("ENVS", "1024,4096,16384,65536").split(",")]
blocks = [int(x) for x in os.environ.get("BLOCKS", "32").split(",")]
out = []
for name, path in libs:
    lib = capi.LlqLibrary(path)
    for n in envs:
        for blk in blocks:
            eng = capi.VecEngine(lib, n, blob, mocap, seed=1234, auto_reset=1, **({'knee_contacts': int(os.environ['KNEE'])} if 'KNEE' in os.environ else {}))
            eng.reset()
            pool = torch.from_numpy(action_pool_np(n, 4, 5678)).to(dev)
            obs = torch.empty((n, 207), device=dev); rew = torch.empty((n,), device=dev); done = torch.empty((n,), device=dev, dtype=torch.uint8)
            K = int(os.environ.get("K", 100))
            for i in range(int(os.environ.get("WARM", 20))):      # joint-limit activity takes ~100 steps to reach its steady state
                eng.step_device(pool[i % 4].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), stream=stream.cuda_stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K):
                eng.step_device(pool[i % 4].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), stream=stream.cuda_stream)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            out.append({"lib": name, "envs": n, "block": blk, "ms_per_step": ms, "env_steps_per_s": n / ms * 1e3})
            print(out[-1], flush=True)
            eng.close()
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
