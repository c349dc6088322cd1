"""wurm_b200: B200-native batched snake environments behind the API of oscarknagg/wurm's
`wurm.envs.SingleSnake` / `wurm.envs.MultiSnake` (hand-written sm_100a CUDA behind a C ABI)."""
from .envs import SingleSnake, MultiSnake, SimpleGridworld  # noqa: F401
from .host_io import HostStepper  # noqa: F401,E402
from .graph import GraphedStepper  # noqa: F401,E402
