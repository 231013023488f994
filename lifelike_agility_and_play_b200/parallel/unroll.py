"""TLeague-format unrolls from a gathered trajectory slab.

The reference's actor hands a finished unroll of ONE environment to the learner as the tuple
`(model_key, flat fp32 array, infos, shapes)` (learning/actors/distill_actor.py:164-167): every time step's record is
flattened leaf by leaf, all steps are concatenated into one 1-D array, and `shapes` holds the leaf shapes of one step so
that the learner's data server can restore the structure.  For the PMC policy-gradient learner the record is `PMCInputs`
(networks/legged_robot/pmc_net/pmc_net_data.py:7-16; placeholders pmc_net.py:60-96):

    X = OrderedDict(prop (99,), prop_a (36,), future (72,))   observation *before* the action (PLE:117-124)
    A (12,)   neglogp ()   discount ()   r (1,)   R (1,)   V (1,)   flatparam (24,) = mean | logstd

`discount` is gamma while the episode runs and 0 on the step that ended it; `R` is the lambda-return the reference learner
also builds for its ppo2 loss (pmc_net.py:213-224: `multistep_forward_view(reward, discounts, vpred[1:], lambda_)`):

    R_t = r_t + discount_t * ((1 - lam) * V_{t+1} + lam * R_{t+1}),     R_T = V_T (bootstrap)

with gamma = lam = 0.95 in the shipped training script (train_scripts/example_pmc_train.sh:21-22).

Here the rollout engine produces `[T, N, 223]` slabs (parallel/trajectory.py: obs 207 | action 12 | reward | done | neglogp |
value); this module turns a slab into the N per-environment tuples.  The return recursion runs as torch ops on whatever
device the slab lives on (a scan over T vectorised over the N environments); the flattening is a host-side reshape.
"""
from collections import OrderedDict

import numpy as np
import torch

from .trajectory import ACT_DIM, COL_ACTION, COL_DONE, COL_NEGLOGP, COL_REWARD, COL_VALUE, OBS_DIM, TRAJ_WIDTH

OBS_LEAVES = OrderedDict([("prop", 99), ("prop_a", 36), ("future", 72)])        # PLE:117-124
# leaf order of a flattened PMCInputs record (namedtuple order, the observation dict in key-insertion order)
RECORD_SHAPES = ((99,), (36,), (72,), (ACT_DIM,), (), (), (1,), (1,), (1,), (2 * ACT_DIM,))
RECORD_WIDTH = OBS_DIM + ACT_DIM + 5 + 2 * ACT_DIM                              # 248 floats per time step
GAMMA, LAM, LOGSTD_INIT = 0.95, 0.95, -2.0                                      # example_pmc_train.sh:21-22, pmc_net_data.py:93


def lambda_returns(reward, discount, value, bootstrap_value, lam=LAM):
    """R_t = r_t + discount_t * ((1-lam) V_{t+1} + lam R_{t+1}) over the leading (time) axis; all inputs `[T, N]`,
    `bootstrap_value` `[N]` = V of the observation that follows the slab's last step."""
    T = reward.shape[0]
    out = torch.empty_like(reward)
    nxt_v, nxt_r = bootstrap_value, bootstrap_value
    for t in range(T - 1, -1, -1):
        nxt_r = reward[t] + discount[t] * ((1.0 - lam) * nxt_v + lam * nxt_r)
        out[t] = nxt_r
        nxt_v = value[t]
    return out


def slab_records(slab, bootstrap_value=None, gamma=GAMMA, lam=LAM, flatparam=None, logstd=LOGSTD_INIT):
    """`[T, N, 223]` slab -> `[N, T, 248]` float32 records in PMCInputs leaf order (on the slab's device)."""
    assert slab.dim() == 3 and slab.shape[2] == TRAJ_WIDTH, "expected a [T, N, %d] trajectory slab" % TRAJ_WIDTH
    T, N, _ = slab.shape
    r, done, v = slab[:, :, COL_REWARD], slab[:, :, COL_DONE], slab[:, :, COL_VALUE]
    discount = gamma * (1.0 - done)
    if bootstrap_value is None:
        bootstrap_value = v[-1]
    ret = lambda_returns(r, discount, v, bootstrap_value.to(slab.dtype), lam)
    rec = torch.empty((N, T, RECORD_WIDTH), dtype=torch.float32, device=slab.device)
    c = OBS_DIM + ACT_DIM
    rec[:, :, :c] = slab[:, :, :c].transpose(0, 1)
    rec[:, :, c + 0] = slab[:, :, COL_NEGLOGP].t()
    rec[:, :, c + 1] = discount.t()
    rec[:, :, c + 2] = r.t()
    rec[:, :, c + 3] = ret.t()
    rec[:, :, c + 4] = v.t()
    if flatparam is None:       # deterministic head: mean = the action taken, log-std at its initial value
        rec[:, :, c + 5:c + 5 + ACT_DIM] = slab[:, :, COL_ACTION:COL_ACTION + ACT_DIM].transpose(0, 1)
        rec[:, :, c + 5 + ACT_DIM:] = logstd
    else:
        rec[:, :, c + 5:] = flatparam.transpose(0, 1)
    return rec


def slab_to_unrolls(slab, model_key, infos=None, **kw):
    """The N tuples `(model_key, flat array, infos, shapes)` the reference actor would have pushed, one per environment
    (distill_actor.py:164-167).  `infos[i]` = list of the `info` dicts of env i's episodes that ended inside the slab."""
    rec = slab_records(slab, **kw).cpu().numpy()
    n = rec.shape[0]
    return [(model_key, rec[i].reshape(-1), list(infos[i]) if infos is not None else [], RECORD_SHAPES) for i in range(n)]


def unflatten_unroll(flat, shapes=RECORD_SHAPES):
    """Inverse of the flattening (what the learner's data server does with `shapes`): list over time of leaf lists."""
    sizes = [int(np.prod(s)) if len(s) else 1 for s in shapes]
    w = sum(sizes)
    assert flat.size % w == 0
    steps = flat.reshape(-1, w)
    out = []
    for row in steps:
        leaves, o = [], 0
        for s, k in zip(shapes, sizes):
            leaves.append(row[o:o + k].reshape(s))
            o += k
        out.append(leaves)
    return out
