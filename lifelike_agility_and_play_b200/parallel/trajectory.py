"""Multi-GPU plumbing of the rollout: env shards and the trajectory gather to the learner rank.

Environments are independent, so the data path has no collective: rank r owns the contiguous block
[r*N, (r+1)*N) of global env ids (RNG streams are keyed by global id, hence results do not depend on the world size).
The only exchange is the hand-over of a finished unroll -- a [T, N_local, 223] fp32 slab whose record mirrors the
reference's PMCInputs (networks/legged_robot/pmc_net/pmc_net_data.py:7-16; actor->learner push, distill_actor.py:164-167)
-- to the learner rank, one gather per unroll (NCCL over NVLink on GPUs, gloo in the CPU tests).  The fused step kernel
writes the record (observation, and with the "record" option action / reward / done) directly into the slab (llq_step_ex
obs_ld = 223), so there is no staging copy between stepping and the send buffer.  `TrajectoryExchange` is the designed
hand-over: two slabs ping-pong, the finished one travels as grouped point-to-point sends / receives on a side
stream while the next unroll is stepped into the other.
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


class TrajectoryExchange:
    """Double-buffered hand-over of finished unrolls to the learner rank, overlapped with stepping.

    Rank r steps its envs into ``slab()`` (a ``[T, N_local, width]`` device tensor the fused kernel writes in place).  When the
    unroll is complete ``hand_over()`` posts the transfer of that slab on a side stream -- one grouped
    ``batch_isend_irecv`` (ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd underneath): every non-learner rank sends its
    slab, the learner posts one receive per sender into ``[world, T, N_local, width]`` and copies its own slab device-to-device
    -- and flips to the other slab, so the 128 steps of unroll k+1 run while unroll k is on the NVLinks.  The stepping stream
    only waits for a transfer when it is about to overwrite that slab again, one whole unroll later.
    CPU tensors (gloo, the world_size-2 tests) take the same path without streams.
    """

    def __init__(self, unroll, n_envs, width, device, dst=0, group=None):
        self.T, self.n, self.width, self.dst, self.group = int(unroll), int(n_envs), int(width), int(dst), group
        self.dev = torch.device(device)
        self.cuda = self.dev.type == "cuda"
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.slabs = [torch.zeros((self.T, self.n, self.width), dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.recv = None
        if self.rank == self.dst and self.world > 1:
            self.recv = [torch.empty((self.world, self.T, self.n, self.width), dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.cur = 0
        self.posted = [False, False]
        self.bytes_per_rank = self.T * self.n * self.width * 4
        if self.cuda:
            self.side = torch.cuda.Stream(self.dev)
            self.ready = [torch.cuda.Event() for _ in range(2)]
            self.sent = [torch.cuda.Event() for _ in range(2)]

    def slab(self):
        """The slab the current unroll is written into."""
        return self.slabs[self.cur]

    def _post(self, b):
        if self.world == 1:
            return
        ops = []
        if self.rank == self.dst:
            for r in range(self.world):
                if r == self.dst:
                    self.recv[b][r].copy_(self.slabs[b], non_blocking=True)
                else:
                    ops.append(dist.P2POp(dist.irecv, self.recv[b][r], r, self.group))
        else:
            ops.append(dist.P2POp(dist.isend, self.slabs[b], self.dst, self.group))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()                      # NCCL: orders the side stream behind the transfer (the host does not block); gloo: blocks

    def hand_over(self):
        """The current slab is complete on the caller's current stream: start its transfer, continue in the other slab.
        Returns the index of the slab now in flight (pass it to ``gathered`` on the learner rank)."""
        b = self.cur
        if self.cuda:
            cur = torch.cuda.current_stream(self.dev)
            self.ready[b].record(cur)
            self.side.wait_event(self.ready[b])
            with torch.cuda.stream(self.side):
                self._post(b)
                self.sent[b].record(self.side)
        else:
            self._post(b)
        self.posted[b] = True
        self.cur ^= 1
        if self.cuda and self.posted[self.cur]:
            torch.cuda.current_stream(self.dev).wait_event(self.sent[self.cur])   # do not overwrite a slab that is still being sent
        return b

    def wait(self, b):
        """Make the caller's current stream wait for transfer b (CPU: already complete)."""
        if self.cuda and self.posted[b]:
            torch.cuda.current_stream(self.dev).wait_event(self.sent[b])

    def gathered(self, b):
        """Learner rank: ``[world, T, N_local, width]`` of unroll b after ``wait(b)``; a single rank gets its own slab; None elsewhere."""
        self.wait(b)
        if self.world == 1:
            return self.slabs[b].unsqueeze(0)
        return self.recv[b] if self.rank == self.dst else None
