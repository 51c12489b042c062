from .local import COINNLocal  # noqa: F401
from .remote import COINNRemote, EmptyDataHandle, check  # noqa: F401
