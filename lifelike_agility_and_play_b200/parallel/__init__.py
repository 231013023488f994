from .trajectory import TRAJ_WIDTH, TrajectoryExchange, TrajectorySlab, shard_offset  # noqa: F401
from .unroll import RECORD_SHAPES, RECORD_WIDTH, lambda_returns, slab_records, slab_to_unrolls, unflatten_unroll  # noqa: F401
from .rollout import RolloutWorker  # noqa: F401
