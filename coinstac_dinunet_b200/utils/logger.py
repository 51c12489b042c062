"""Console logging helpers gated by a verbosity flag (ref utils/logger.py:4-24)."""
import math as _math

__all__ = ['error', 'warn', 'info', 'success', 'lazy_debug']


def _emit(prefix, msg, enabled):
    if enabled:
        print(f"{prefix}{msg}", flush=True)


def error(msg, debug=True):
    _emit("####   [Error!]   ####: ", msg, debug)


def warn(msg, debug=True):
    _emit("---  [Warning!]  ---: ", msg, debug)


def info(msg, debug=True):
    _emit("", msg, debug)


def success(msg, debug=True):
    _emit("***  [Success!] ***: ", msg, debug)


def lazy_debug(x, add=1):
    """Log-sparse cadence: true for 1,2,4,6,9,12,15,18,20,24,....

    The period grows like ``log(x)`` so long runs log (and plot) progressively less.
    """
    period = int(_math.log(x + 1) + add)
    return x % period == 0
