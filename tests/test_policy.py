"""Host reference of the PMC policy forward (lifelike_agility_and_play_b200/policy.py; pmc_net.py:33-58,99-178)."""
import numpy as np

from lifelike_agility_and_play_b200.policy import PmcPolicy

SHAPES = [(1, 135), (1, 135), (1, 72), (1, 72), (207, 256), (256,), (256, 256), (256,), (256, 1), (1,), (207, 256), (256,), (256, 256), (256,),
          (256, 32), (32,), (32, 256), (135, 64), (64,), (32, 32), (32,), (96, 256), (256,), (256, 256), (256,), (256, 12), (12,), (1, 12)]


def random_weights(seed=0):
    rng = np.random.default_rng(seed)
    w = [(rng.standard_normal(s) / np.sqrt(s[0] if len(s) == 2 and s[0] > 1 else 1.0)).astype(np.float32) for s in SHAPES]
    w[1] = np.abs(w[1]) + 0.1; w[3] = np.abs(w[3]) + 0.1        # running std > 0
    return w


def test_policy_forward_structure():
    pol = PmcPolicy(random_weights())
    rng = np.random.default_rng(1)
    obs = rng.standard_normal((50, 207)).astype(np.float32) * 3
    a, idx = pol.act(obs, return_code=True)
    assert a.shape == (50, 12) and idx.shape == (50,) and idx.min() >= 0 and idx.max() < 256
    p, f = pol.normalise(obs)
    assert np.abs(p).max() <= 5.0 and np.abs(f).max() <= 5.0                         # pmc_net.py:133,136
    z, idx2 = pol.encode(p, f)
    d = ((z[:, :, None] - pol.codebook[None]) ** 2).sum(1)                            # brute-force nearest code
    assert np.array_equal(idx2, d.argmin(1))
    # the action only depends on the observation through (normalised prop, code): same prop + same code -> same action
    obs2 = obs.copy(); obs2[:, 135:] += 1e-4
    a2, idx3 = pol.act(obs2, return_code=True)
    same = idx3 == idx
    assert same.mean() > 0.8 and np.allclose(a2[same], a[same], atol=1e-5)
    # batch independence
    assert np.allclose(pol.act(obs[7:8]), a[7:8], atol=1e-5)


def test_policy_library_exports():
    import ctypes
    import os
    from lifelike_agility_and_play_b200 import policy
    import __graft_entry__
    __graft_entry__.build()
    assert os.path.exists(policy.POLICY_LIB_PATH)
    lib = ctypes.CDLL(policy.POLICY_LIB_PATH)
    for name in policy.POLICY_EXPORTS:
        assert hasattr(lib, name), name
    assert policy.pack_weights(random_weights()).size == policy.N_WEIGHTS == 358647
    hdr = open(os.path.join(os.path.dirname(policy.POLICY_LIB_PATH), "..", "..", "include", "llq_policy.h")).read()
    for name in policy.POLICY_EXPORTS:
        assert name + "(" in hdr


import pytest  # noqa: E402


@pytest.mark.gpu
def test_device_policy_matches_host_forward():
    """CUDA forward vs the numpy restatement on random weights and observations: same code for (almost) every row, actions
    within 1e-4 relative where the code agrees (a different nearest code is only possible at fp32 ties of the distance)."""
    import torch
    from lifelike_agility_and_play_b200.policy import DevicePolicy
    w = random_weights(3)
    host = PmcPolicy(w)
    dev = DevicePolicy(w, device=0)
    rng = np.random.default_rng(5)
    for n, ld in ((4096, 223), (77, 207), (1, 207)):
        obs = (2.0 * rng.standard_normal((n, ld))).astype(np.float32)
        t_obs = torch.from_numpy(obs).cuda()
        t_act = torch.zeros((n, 12), device="cuda", dtype=torch.float32)
        t_code = torch.zeros((n,), device="cuda", dtype=torch.int32)
        dev.forward(t_obs.data_ptr(), ld, n, t_act.data_ptr(), t_code.data_ptr(), None)
        torch.cuda.synchronize()
        a_ref, c_ref = host.act(obs[:, :207], return_code=True)
        code = t_code.cpu().numpy(); act = t_act.cpu().numpy()
        same = code == c_ref
        assert same.mean() >= (0.999 if n >= 1000 else 1.0), same.mean()
        if not same.all():
            # every mismatch must be a tie at fp32 resolution: in fp64, the code the kernel picked is as near as the host's
            # pick to within the rounding of a 256-wide fp32 encoder (|z|^2 cancels; the gap is compared with the distance scale)
            p, f = host.normalise(obs[:, :207])
            z, _ = host.encode(p, f)
            z = z.astype(np.float64)[~same]; cb = host.codebook.astype(np.float64)
            d = ((z[:, :, None] - cb[None]) ** 2).sum(1)
            rows = np.arange(z.shape[0])
            gap = np.abs(d[rows, code[~same]] - d[rows, c_ref[~same]]) / (1.0 + d.min(1))
            assert gap.max() < 2e-5, (gap.max(), int((~same).sum()))
        err = np.abs(act[same] - a_ref[same]).max() / (1.0 + np.abs(a_ref).max())
        assert err < 1e-4, err
        # rollout entry: value head + sampled actions with their -log p
        t_val = torch.zeros((n,), device="cuda", dtype=torch.float32)
        t_nlp = torch.zeros((n,), device="cuda", dtype=torch.float32)
        t_smp = torch.zeros((n, 12), device="cuda", dtype=torch.float32)
        dev.forward_ex(t_obs.data_ptr(), ld, n, t_smp.data_ptr(), None, t_val.data_ptr(), t_nlp.data_ptr(), seed=7, counter=3)
        torch.cuda.synchronize()
        v_ref = host.value(obs[:, :207])
        assert np.abs(t_val.cpu().numpy() - v_ref).max() < 1e-4 * (1.0 + np.abs(v_ref).max())
        smp, nlp = t_smp.cpu().numpy(), t_nlp.cpu().numpy()
        assert np.abs(nlp[same] - host.neglogp(smp, act)[same]).max() < 1e-3          # consistent with (sample, mean, logstd)
        eps = ((smp - act) / np.exp(host.logstd.reshape(-1)))[same]
        if n >= 4096:
            assert abs(eps.mean()) < 0.02 and abs(eps.std() - 1.0) < 0.02 and np.abs(eps).max() < 6.5
            assert abs(np.corrcoef(eps[:, 0], eps[:, 1])[0, 1]) < 0.05
        t_smp2 = torch.zeros_like(t_smp)
        dev.forward_ex(t_obs.data_ptr(), ld, n, t_smp2.data_ptr(), None, None, t_nlp.data_ptr(), seed=7, counter=3)
        t_smp3 = torch.zeros_like(t_smp)
        dev.forward_ex(t_obs.data_ptr(), ld, n, t_smp3.data_ptr(), None, None, t_nlp.data_ptr(), seed=7, counter=4)
        torch.cuda.synchronize()
        assert torch.equal(t_smp2, t_smp) and not torch.equal(t_smp3, t_smp)           # keyed by (seed, counter), reproducible
    dev.close()
