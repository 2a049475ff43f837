from .batched_env import BatchedHumanoidEnv  # noqa: F401
from .jvrc_walk import JvrcWalkEnv  # noqa: F401
