from .trajectory import TRAJ_WIDTH, TrajectorySlab, shard_offset  # noqa: F401
