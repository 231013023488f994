"""Step-kernel time with and without knee-wheel contacts on the bench workload (4096 PMC envs, N(mu,sigma) actions)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
for knee in (1, 0):
    eng = bench.make_engine(None, 4096, "pmc", device=0, seed=1234, auto_reset=1, knee_contacts=knee)
    eng.reset()
    pool = torch.from_numpy(bench.action_pool_np(4096, 16, 5678)).cuda()
    obs = torch.zeros((4096, 207), device="cuda"); r = torch.zeros(4096, device="cuda"); d = torch.zeros(4096, device="cuda", dtype=torch.uint8)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    for i in range(300):
        eng.step_device(pool[i % 16].data_ptr(), obs.data_ptr(), r.data_ptr(), d.data_ptr(), obs_ld=207, stream=st.cuda_stream)
    torch.cuda.synchronize()
    c0 = eng.counters()
    eng.set_option("profile", 1)
    ks = []
    for i in range(64):
        eng.step_device(pool[i % 16].data_ptr(), obs.data_ptr(), r.data_ptr(), d.data_ptr(), obs_ld=207, stream=st.cuda_stream)
        torch.cuda.synchronize()
        ks.append(eng.timing()[0])
    c1 = eng.counters()
    print("knee", knee, "kernel ms %.4f" % np.mean(ks), "contact rows/step", (c1[2] - c0[2]) / 64, "limit rows/step", (c1[3] - c0[3]) / 64, "dones/step", (c1[1] - c0[1]) / 64)
    eng.close()
