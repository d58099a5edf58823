"""Logging and execution-time profiling of the hot path (``cashocs/log.py``).

The reference wraps exactly four operations with ``@log.profile_execution_time(action)`` at TRACE level:
assembling a system (``cashocs/_utils/linalg.py:107``), solving a linear system (``:483``), the a-priori mesh
test and the intersection test (``cashocs/geometry/mesh_testing.py:83,160``).  The same decorator exists here;
because the wrapped functions only *enqueue* kernels, the elapsed time is taken with CUDA events on the current
stream (a synchronisation per call), so profiling is off unless the log level is TRACE -- the hot path never
synchronises for logging otherwise.  ``timings()`` returns the accumulated (calls, milliseconds) per action.
"""

from __future__ import annotations

import collections
import functools
import logging
import os
from typing import Any, Callable, TypeVar

TRACE = 5
DEBUG, INFO, WARNING, ERROR, CRITICAL = logging.DEBUG, logging.INFO, logging.WARNING, logging.ERROR, logging.CRITICAL
logging.addLevelName(TRACE, "TRACE")

_logger = logging.getLogger("cashocs_b200")
if not _logger.handlers:
    _handler = logging.StreamHandler()
    _handler.setFormatter(logging.Formatter("cashocs_b200 - %(levelname)s: %(message)s"))
    _logger.addHandler(_handler)
_logger.setLevel(TRACE if os.environ.get("CASHOCS_B200_TRACE") == "1" else logging.WARNING)

_timings: dict[str, list] = collections.defaultdict(lambda: [0, 0.0])
T = TypeVar("T")


def set_log_level(level: int) -> None:
    """``cashocs.log.set_log_level``."""
    _logger.setLevel(level)


def log(level: int, message: str) -> None:
    _logger.log(level, message)


def trace(message: str) -> None:
    _logger.log(TRACE, message)


def debug(message: str) -> None:
    _logger.debug(message)


def info(message: str) -> None:
    _logger.info(message)


def warning(message: str) -> None:
    _logger.warning(message)


def timings(reset: bool = False) -> dict[str, tuple[int, float]]:
    """Accumulated (number of calls, device milliseconds) per profiled action."""
    out = {k: (v[0], v[1]) for k, v in _timings.items()}
    if reset:
        _timings.clear()
    return out


def profile_execution_time(action: str, level: int = TRACE) -> Callable[[Callable[..., T]], Callable[..., T]]:
    """Profiles the execution time of a function (``cashocs/log.py:361-389``), measured on the device."""

    def decorator(func: Callable[..., T]) -> Callable[..., T]:
        @functools.wraps(func)
        def wrapper(*args: Any, **kwargs: Any) -> T:
            if not _logger.isEnabledFor(level):
                return func(*args, **kwargs)
            import torch

            if not torch.cuda.is_available():
                return func(*args, **kwargs)
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            result = func(*args, **kwargs)
            stop.record()
            stop.synchronize()
            ms = start.elapsed_time(stop)
            t = _timings[action]
            t[0] += 1
            t[1] += ms
            _logger.log(level, f"Elapsed time for {action}: {ms:.3f} ms.")
            return result

        return wrapper

    return decorator
