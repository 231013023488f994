"""Multi-GPU plumbing of the rollout: env shards and the trajectory gather to the learner rank.

Environments are independent, so the data path has no collective: rank r owns the contiguous block
[r*N, (r+1)*N) of global env ids (RNG streams are keyed by global id, hence results do not depend on the world size).
The only exchange is the hand-over of a finished unroll -- a [T, N_local, 223] fp32 slab whose record mirrors the
reference's PMCInputs (networks/legged_robot/pmc_net/pmc_net_data.py:7-16; actor->learner push, distill_actor.py:164-167)
-- to the learner rank, one gather per unroll (NCCL over NVLink on GPUs, gloo in the CPU tests).  The fused step kernel
writes the observation part of each record directly into the slab (llq_step_ex obs_ld = 223), so there is no staging copy
between stepping and the send buffer.
"""
import torch
import torch.distributed as dist

OBS_DIM, ACT_DIM = 207, 12
COL_ACTION, COL_REWARD, COL_DONE, COL_NEGLOGP, COL_VALUE = 207, 219, 220, 221, 222
TRAJ_WIDTH = 223


def shard_offset(rank, envs_per_rank):
    """Global id of this rank's env 0 (-> llq_config.global_env_offset)."""
    return int(rank) * int(envs_per_rank)


class TrajectorySlab:
    def __init__(self, unroll, n_envs, device):
        self.unroll, self.n = int(unroll), int(n_envs)
        self.buf = torch.zeros((self.unroll, self.n, TRAJ_WIDTH), dtype=torch.float32, device=device)

    def row(self, t):
        return self.buf[t % self.unroll]

    def record(self, t, action, reward, done, obs=None):
        """Fill the non-observation columns of record t (obs is written by the kernel unless given)."""
        r = self.row(t)
        if obs is not None:
            r[:, :OBS_DIM] = obs
        r[:, COL_ACTION:COL_ACTION + ACT_DIM] = action
        r[:, COL_REWARD] = reward
        r[:, COL_DONE] = done

    def gather_to_learner(self, dst=0, recv=None):
        """All ranks call this once per unroll; returns the list of per-rank slabs on `dst`, None elsewhere."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return [self.buf]
        if dist.get_rank() == dst:
            if recv is None:
                recv = [torch.empty_like(self.buf) for _ in range(dist.get_world_size())]
            dist.gather(self.buf, recv, dst=dst)
            return recv
        dist.gather(self.buf, None, dst=dst)
        return None
