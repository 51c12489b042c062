from .basetrainer import NNTrainer  # noqa: F401
