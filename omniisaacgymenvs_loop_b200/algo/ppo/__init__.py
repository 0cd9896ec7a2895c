"""The loopz PPO learner (reference `omniisaacgymenvs/algo/ppo`) on the C-ABI kernels of csrc/ppo_loopz.cu."""
from . import dagger, module, ppo, storage  # noqa: F401
from .ppo import PPO  # noqa: F401
from .storage import ObsStorage, RolloutStorage  # noqa: F401
