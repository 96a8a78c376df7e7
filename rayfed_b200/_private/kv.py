"""In-process KV with the reference's key namespacing.

Parity: fed/_private/compatible_utils.py:68-74,106-188.  The reference parks
the pickled configs in Ray's GCS so that proxy *actor processes* can read them.
Here every component of a party lives in the controller process, so the store
is a dict guarded by a lock — but the public shape (``kv.put/get/delete/reset``,
``RAYFED#{job}#{key}`` keys, ``_init_internal_kv`` / ``_clear_internal_kv``) is
unchanged.  The backing dict is process-global: two jobs in one process see
disjoint key spaces, exactly like two jobs on one Ray cluster.
"""
from __future__ import annotations

import threading
from typing import Dict, Optional

from . import constants

_STORE: Dict[bytes, bytes] = {}
_STORE_LOCK = threading.Lock()


def wrap_kv_key(job_name: str, key: str) -> str:
    """``key`` → ``RAYFED#{job}#{key}`` (str in, str out; bytes handled too)."""
    if isinstance(key, bytes):
        key = key.decode("utf-8")
    return constants.RAYFED_JOB_KV_DATA_KEY_FMT.format(job_name, key)


def _b(x) -> bytes:
    return x if isinstance(x, bytes) else str(x).encode("utf-8")


class InternalKv:
    """KV façade bound to one job name."""

    def __init__(self, job_name: str) -> None:
        self._job_name = job_name

    def initialize(self) -> None:
        return None

    def put(self, k: str, v: bytes) -> bool:
        with _STORE_LOCK:
            _STORE[_b(wrap_kv_key(self._job_name, k))] = v
        return True

    def get(self, k: str) -> Optional[bytes]:
        with _STORE_LOCK:
            return _STORE.get(_b(wrap_kv_key(self._job_name, k)))

    def delete(self, k: str) -> int:
        with _STORE_LOCK:
            return 1 if _STORE.pop(_b(wrap_kv_key(self._job_name, k)), None) is not None else 0

    def reset(self) -> None:
        return None

    def _ping(self) -> str:
        return "pong"


kv: Optional[InternalKv] = None
_init_lock = threading.Lock()


def _init_internal_kv(job_name: str) -> bool:
    """Create the module-level ``kv`` (idempotent). Returns True if created."""
    global kv
    with _init_lock:
        if kv is None:
            kv = InternalKv(job_name)
            kv.initialize()
            return True
        return False


def _clear_internal_kv() -> None:
    """Delete this job's config keys and drop the façade."""
    global kv
    with _init_lock:
        if kv is not None:
            kv.delete(constants.KEY_OF_CLUSTER_CONFIG)
            kv.delete(constants.KEY_OF_JOB_CONFIG)
            kv.reset()
            kv = None
    from .. import config as fed_config

    fed_config._reset_config_cache()


def raw_get(wrapped_key) -> Optional[bytes]:
    """Look up an already-namespaced key (what ``ray internal_kv`` exposed)."""
    with _STORE_LOCK:
        return _STORE.get(_b(wrapped_key))


def raw_keys():
    with _STORE_LOCK:
        return list(_STORE.keys())


# -- migration conveniences (the reference keeps these next to the KV in compatible_utils.py) ------
def init_ray(address: str = None, **kwargs):
    """The reference's scripts call ``compatible_utils.init_ray(address='local')`` before
    ``fed.init`` (compatible_utils.py:51-57).  There is no Ray to start here: accepted, ignored."""
    return None


def _compare_version_strings(version1: str, version2: str) -> bool:
    """True iff ``version1`` > ``version2`` (numeric, component-wise; missing components = 0)."""
    def parts(v):
        out = []
        for p in str(v).split("."):
            digits = "".join(ch for ch in p if ch.isdigit())
            out.append(int(digits) if digits else 0)
        return out

    a, b = parts(version1), parts(version2)
    n = max(len(a), len(b))
    a += [0] * (n - len(a))
    b += [0] * (n - len(b))
    return a > b
