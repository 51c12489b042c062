from .imageutils import Image  # noqa: F401
from . import imageutils, plotter  # noqa: F401
