from .storage import BatchData  # noqa: F401
from .policies import FF_V, Gaussian_FF_Actor  # noqa: F401
from .workers import DeviceRolloutWorker  # noqa: F401
from .optim import FusedClipAdam  # noqa: F401
from .ppo import PPO  # noqa: F401
