"""A *minimal* stand-in for the handful of ``ray`` calls that rayfed driver scripts make.

Scripts written for the reference start with ``ray.init(address='local')``, end with
``ray.shutdown()`` and occasionally call ``ray.get`` / ``ray.wait`` on the local futures
inside a FedObject or ``ray.get_actor(<proxy name>)``.  None of that needs Ray here — the
party runtime is in-process — so a one-line change keeps such scripts running::

    import rayfed_b200.compat.ray as ray      # instead of: import ray

This is NOT a Ray re-implementation (no ``ray.remote``, no cluster, no object store); use
``fed.remote`` for tasks and actors.
"""
from __future__ import annotations

import logging
from typing import List, Optional

from ..core import get as _get
from ..core import get_runtime, shutdown_runtime
from ..core import wait as _wait
from ..exceptions import FedGetTimeoutError as GetTimeoutError  # noqa: F401  (ray.exceptions.GetTimeoutError)
from ..exceptions import FedTaskError as RayTaskError  # noqa: F401

logger = logging.getLogger(__name__)
__version__ = "0.0.0+rayfed_b200.compat"
_initialized = False


def init(address: Optional[str] = None, **kwargs):
    """No cluster to start: the party runtime comes up with ``fed.init``."""
    global _initialized
    _initialized = True
    if kwargs:
        logger.debug(f"ray.init options ignored by the in-process runtime: {sorted(kwargs)}")
    return None


def is_initialized() -> bool:
    return _initialized


def shutdown():
    global _initialized
    _initialized = False
    shutdown_runtime(True)


def get(refs, timeout: Optional[float] = None):
    return _get(refs, timeout)


def wait(refs: List, num_returns: int = 1, timeout: Optional[float] = None):
    return _wait(refs, num_returns=num_returns, timeout=timeout)


def get_actor(name: str):
    """Named lookup in this party's runtime (proxy hosts are registered under the reference's names)."""
    return get_runtime().get_named(name)


class exceptions:  # noqa: N801  (mirrors the module path ray.exceptions.*)
    GetTimeoutError = GetTimeoutError
    RayTaskError = RayTaskError
    RayError = Exception
