from .emulator import InProcessEngine, node_state, unwrap_spec  # noqa: F401
from .dist_engine import DistEngine, init_process_group  # noqa: F401
