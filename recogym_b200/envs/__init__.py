from .configuration import Configuration
from .context import Context, DefaultContext
from .observation import Observation
from .session import Session, OrganicSessions
from .reco_env_v1 import RecoEnv1, env_args, env_1_args
from .reco_env_v0 import RecoEnv0, env_0_args, DeviceEnv0
from .device_env import DeviceEnv, BatchState, rollout_log
from .features.time import TimeGenerator, DefaultTimeGenerator, NormalTimeGenerator
