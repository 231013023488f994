"""On-device actor loop producing learner-ready unrolls (rows f2 + f3 joined to the hot path).

Replaces, for one GPU's block of environments, the reference actor's per-env loop `obs -> agent.step -> env.step -> queue`
(learning/actors/distill_actor.py:205-270) and its unroll assembly (`:120-167`): the policy kernel (csrc/llq_policy.cu) reads
observation row t of the trajectory slab in place, the fused env step (llq_step_ex, LLQ_IO_DEVICE) consumes its actions and
writes the *next* observation straight into row t+1 of the slab, so record t = (obs_t, a_t, r_t, done_t) is aligned the way
`PMCInputs` wants it (X = the observation the action was computed from) without any staging copy.  Nothing synchronises with
the host inside an unroll.  Two slabs ping-pong: `finish_unroll()` hands back the `[T, N, 223]` view of the finished one (it stays
valid for the whole next unroll, so the NCCL gather / unroll conversion overlaps the stepping) and carries observation T over
to row 0 of the other.
"""
import torch

from .trajectory import ACT_DIM, COL_ACTION, COL_DONE, COL_NEGLOGP, COL_REWARD, COL_VALUE, OBS_DIM, TRAJ_WIDTH


class RolloutWorker:
    def __init__(self, engine, policy, unroll, device, sample=True, seed=0):
        """`engine`: a `_capi.VecEngine` on the CUDA library with auto_reset=1 (PMC, 207-wide observations);
        `policy`: a `policy.DevicePolicy` on the same device; `unroll`: T (128 in example_pmc_train.sh:145);
        `sample`: draw the actions from the Gaussian head and record -log p (training rollouts) instead of the mean (evaluation)."""
        if engine.obs_dim != OBS_DIM:
            raise ValueError("RolloutWorker drives the PMC env (207-wide observations)")
        self.eng, self.pol, self.T, self.n = engine, policy, int(unroll), engine.n
        self.dev = torch.device(device)
        self.bufs = [torch.zeros((self.T + 1, self.n, TRAJ_WIDTH), dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.buf = self.bufs[0]
        self.act = torch.zeros((self.n, ACT_DIM), dtype=torch.float32, device=self.dev)
        self.rew = torch.zeros((self.n,), dtype=torch.float32, device=self.dev)
        self.done = torch.zeros((self.n,), dtype=torch.uint8, device=self.dev)
        self.val = torch.zeros((self.n,), dtype=torch.float32, device=self.dev)
        self.nlp = torch.zeros((self.n,), dtype=torch.float32, device=self.dev)
        self.boots = [torch.zeros((self.n,), dtype=torch.float32, device=self.dev) for _ in range(2)]   # V(obs_T) per slab
        self._scratch = torch.zeros((self.n, ACT_DIM), dtype=torch.float32, device=self.dev)
        self.bootstrap_value = self.boots[0]
        self.sample, self.seed, self.calls = bool(sample), int(seed), 0
        self.row_gid0 = int(engine.cfg.global_env_offset)      # noise keyed by the global env id: equal seeds on two shards still differ
        engine.set_option("record", 2)                          # a_t | r_t | done_t go to the slab row BEFORE the one receiving obs_{t+1}
        # everything this worker launches (kernels through the C-ABI and torch's column copies) is ordered on ONE side stream: a
        # NULL stream would mean "the engine's own non-blocking stream" to llq_step_ex and would not order with torch's work
        self.stream = torch.cuda.Stream(self.dev)
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        self.t = 0
        self.launches = 0

    def start(self, first_obs):
        """`first_obs` [N, 207] (host or device): the observation `engine.reset()` returned."""
        first = torch.as_tensor(first_obs, dtype=torch.float32).to(self.dev)
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream):
            self.buf[0, :, :OBS_DIM] = first
        self.t = 0

    def step(self):
        """One policy forward + one fused env step; fills record t.  Asynchronous on the worker's stream."""
        assert self.t < self.T, "unroll is full: call finish_unroll()"
        s = self.stream.cuda_stream
        row, nxt = self.buf[self.t], self.buf[self.t + 1]
        # Every column of record t is written by the two kernels themselves, no copies: the policy kernel reads observation t in
        # place and writes V / -log p into the value / neglogp columns of row t (row stride 223); the fused step (record option 2)
        # writes a_t | r_t | done_t into row t and observation t+1 into row t+1.
        fsz = 4
        self.pol.forward_rec(row.data_ptr(), TRAJ_WIDTH, self.n, self.act.data_ptr(), row.data_ptr() + COL_VALUE * fsz,
                             (row.data_ptr() + COL_NEGLOGP * fsz) if self.sample else None, TRAJ_WIDTH, self.seed, self.calls,
                             self.row_gid0, s)
        self.eng.step_device(self.act.data_ptr(), nxt.data_ptr(), self.rew.data_ptr(), self.done.data_ptr(), obs_ld=TRAJ_WIDTH, stream=s)
        self.calls += 1
        self.t += 1
        self.launches += 3            # policy, step, reset kernels; nothing else runs inside an unroll

    def finish_unroll(self):
        """Copy-free `[T, N, 223]` view of the finished records, valid until the end of the NEXT unroll; stepping continues
        in the other slab, whose row 0 receives observation T.  `self.bootstrap_value` [N] = V(observation T) for this slab."""
        assert self.t == self.T
        done_buf = self.buf
        idx = 0 if done_buf is self.bufs[0] else 1
        self.buf = self.bufs[1 - idx]
        with torch.cuda.stream(self.stream):
            # V of the observation that follows the last record: the bootstrap of the lambda-return (unroll.slab_records)
            self.pol.forward_ex(done_buf[self.T].data_ptr(), TRAJ_WIDTH, self.n, self._scratch.data_ptr(), None,
                                self.boots[idx].data_ptr(), None, 0, 0, self.stream.cuda_stream)
            self.buf[0, :, :OBS_DIM] = done_buf[self.T, :, :OBS_DIM]
        self.bootstrap_value = self.boots[idx]
        self.launches += 1
        self.t = 0
        return done_buf[:self.T]

    def wait(self):
        """Make torch's current stream wait for everything queued so far (call before reading a finished slab there)."""
        torch.cuda.current_stream(self.dev).wait_stream(self.stream)
