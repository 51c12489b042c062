from .learner import COINNLearner  # noqa: F401
from .reducer import COINNReducer  # noqa: F401
from .nodes import COINNLocal, COINNRemote  # noqa: F401
