"""N>1 host path on CPU: two gloo ranks shard the envs (oracle engines standing in for the per-GPU engines), fill
trajectory slabs and gather them to the learner rank; the result must equal one process stepping the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T, N_PER, WORLD = 5, 6, 2


def _actions(t, n_global):
    rng = np.random.default_rng(100 + t)
    return (0.1 * rng.standard_normal((n_global, 12))).astype(np.float32)


def _rollout(rank, world, n_per, slab):
    from lifelike_agility_and_play_b200.model.compile_model import load_model_blob
    from lifelike_agility_and_play_b200.mocap import synthetic_mocap
    from lifelike_agility_and_play_b200.parallel import shard_offset
    from oracle import oracle
    eng = oracle.make_engine(n_per, load_model_blob(), synthetic_mocap(6, seed=3, min_frames=380, max_frames=700),
                             seed=99, global_env_offset=shard_offset(rank, n_per if world > 1 else 0), num_threads=1)
    eng.reset()
    for t in range(T):
        a = _actions(t, N_PER * WORLD)[rank * n_per:(rank + 1) * n_per] if world > 1 else _actions(t, N_PER * WORLD)
        obs, rew, done = eng.step(a)
        slab.record(t, torch.from_numpy(a), torch.from_numpy(rew), torch.from_numpy(done.astype(np.float32)), obs=torch.from_numpy(obs))
    eng.close()


def _worker(rank, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from lifelike_agility_and_play_b200.parallel import TrajectorySlab
    slab = TrajectorySlab(T, N_PER, "cpu")
    _rollout(rank, WORLD, N_PER, slab)
    out = slab.gather_to_learner(dst=0)
    if rank == 0:
        q.put(torch.cat(out, dim=1).numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_gather_equals_single_process(built):
    from lifelike_agility_and_play_b200.parallel import TrajectorySlab, TRAJ_WIDTH
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = TrajectorySlab(T, N_PER * WORLD, "cpu")
    _rollout(0, 1, N_PER * WORLD, ref)
    assert gathered.shape == (T, N_PER * WORLD, TRAJ_WIDTH)
    assert np.array_equal(gathered, ref.buf.numpy()), "sharded rollout differs from the single-process rollout"
    # learner side: the gathered slab turns into one TLeague-format unroll per global env (parallel/unroll.py), identical to the
    # single-process ones
    from lifelike_agility_and_play_b200.parallel import RECORD_WIDTH, slab_to_unrolls
    ua, ub = slab_to_unrolls(torch.from_numpy(gathered), "k"), slab_to_unrolls(ref.buf, "k")
    assert len(ua) == N_PER * WORLD and ua[0][1].shape == (T * RECORD_WIDTH,)
    assert all(np.array_equal(x[1], y[1]) and x[3] == y[3] for x, y in zip(ua, ub))


def _xch_worker(rank, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from lifelike_agility_and_play_b200.parallel import TrajectoryExchange
    x = TrajectoryExchange(T, N_PER, 7, "cpu")
    got = []
    for u in range(3):                                # three unrolls: both slabs are reused
        x.slab()[...] = torch.arange(T * N_PER * 7, dtype=torch.float32).reshape(T, N_PER, 7) + 1000 * u + 100000 * rank
        b = x.hand_over()
        g = x.gathered(b)
        if rank == 0:
            got.append(g.clone().numpy())
        else:
            assert g is None
    if rank == 0:
        q.put(np.stack(got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_trajectory_exchange_hands_every_unroll_to_the_learner(built):
    """The designed hand-over (parallel/trajectory.py: ping-pong slabs, grouped send / recv) on two gloo ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_xch_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == (3, WORLD, T, N_PER, 7)
    base = np.arange(T * N_PER * 7, dtype=np.float32).reshape(T, N_PER, 7)
    for u in range(3):
        for r in range(WORLD):
            assert np.array_equal(got[u, r], base + 1000 * u + 100000 * r)
