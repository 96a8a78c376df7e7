"""Utilities (parity: fed/utils.py).

``resolve_dependencies`` (fed/utils.py:48-83), logging set-up
(fed/utils.py:99-146), TLS helpers (:149-163), ``is_cython`` (:166-179),
``dict2tuple`` (:182-195), address validation (:198-239), ``start_command``
(:242-254), ``get_package_version`` (:29-45).  Observability helpers new to
this framework live in sibling modules (``tracing``, ``clocks``, ``checkpoint``).
"""
from __future__ import annotations

import logging
import re
import subprocess
import sys
from typing import Dict, Optional

from ..core import is_ref, is_refs
from ..fed_object import FedObject
from ..tree_util import has_containers, tree_flatten, tree_unflatten

logger = logging.getLogger(__name__)

is_ray_object_refs = is_refs  # reference name


def get_package_version(package_name: str) -> Optional[str]:
    """Installed version of ``package_name`` or None."""
    try:
        from importlib import metadata

        return metadata.version(package_name)
    except Exception:  # noqa: BLE001
        return None


_RECV = None


def _recv_fn():
    global _RECV
    if _RECV is None:  # late import: barriers imports this module
        from ..proxy.barriers import recv

        _RECV = recv
    return _RECV


def resolve_dependencies(current_party, current_fed_task_id, *args, **kwargs):
    """Swap every ``FedObject`` leaf of ``(args, kwargs)`` for a local future.

    * owned here → its `Ref`;
    * owned elsewhere → the `Ref` of a receive keyed
      ``(obj.id, current_fed_task_id)``, cached on the handle so an object is
      received at most once per party.
    Top-level futures gate the task in the scheduler; nested ones reach the
    task body as futures and are read with ``fed.get`` (same as the reference).
    """
    recv = _recv_fn()

    if not has_containers(args, kwargs):
        # hot path: flat argument list, no pytree walk.  Several remote arguments (a fan-in: the
        # aggregator of N parties) are claimed with one call into the transport.
        pending = [a for a in args if type(a) is FedObject and a._node_party != current_party and a._ref is None]
        if len(pending) > 1 and not kwargs:
            from ..proxy.barriers import recv_many

            seen, items, objs = set(), [], []
            for a in pending:
                if id(a) not in seen:
                    seen.add(id(a))
                    items.append((a._node_party, a.get_fed_task_id()))
                    objs.append(a)
            for a, ref in zip(objs, recv_many(current_party, items, current_fed_task_id)):
                a._cache_ref(ref)

        def resolve(arg):
            if type(arg) is not FedObject:
                return arg
            if arg.get_party() == current_party:
                return arg.get_ref()
            ref = arg.get_ref()
            if ref is None:
                ref = recv(current_party, arg.get_party(), arg.get_fed_task_id(), current_fed_task_id)
                arg._cache_ref(ref)
            return ref

        return tuple(resolve(a) for a in args), {k: resolve(v) for k, v in kwargs.items()} if kwargs else {}

    leaves, spec = tree_flatten((args, kwargs))
    for i, arg in enumerate(leaves):
        if type(arg) is not FedObject:
            continue
        if arg.get_party() == current_party:
            leaves[i] = arg.get_ref()
            continue
        ref = arg.get_ref()
        if ref is None:
            ref = recv(current_party, arg.get_party(), arg.get_fed_task_id(), current_fed_task_id)
            arg._cache_ref(ref)
        leaves[i] = ref
    return tree_unflatten(leaves, spec)


# -- logging -------------------------------------------------------------------
class _ContextFilter(logging.Filter):
    def __init__(self, key: str, value: str):
        super().__init__(key)
        self._key = key
        self._value = value

    def filter(self, record) -> bool:
        if not hasattr(record, self._key):
            setattr(record, self._key, self._value)
        return True


def setup_logger(logging_level, logging_format, date_format, log_dir=None, party_val=None, job_name=None):
    """Route the root logger through one stream handler whose records carry
    ``party`` and ``jobname`` (format: constants.RAYFED_LOG_FMT)."""
    root = logging.getLogger()
    if isinstance(logging_level, str):
        logging_level = logging.getLevelName(logging_level.upper())
    root.setLevel(logging_level)
    for h in list(root.handlers):
        root.removeHandler(h)
    handler = logging.StreamHandler(stream=sys.stderr)
    handler.setLevel(logging_level)
    handler.setFormatter(logging.Formatter(fmt=logging_format, datefmt=date_format))
    handler.addFilter(_ContextFilter("party", party_val))
    handler.addFilter(_ContextFilter("jobname", job_name))
    root.addHandler(handler)


# -- TLS -------------------------------------------------------------------------
def tls_enabled(tls_config) -> bool:
    return bool(tls_config)


def load_cert_config(cert_config: Dict):
    """→ ``(ca_cert | None, private_key, cert_chain)`` as bytes."""

    def read(path):
        with open(path, "rb") as f:
            return f.read()

    ca = read(cert_config["ca_cert"]) if "ca_cert" in cert_config else None
    return ca, read(cert_config["key"]), read(cert_config["cert"])


# -- misc -----------------------------------------------------------------------
def is_cython(obj) -> bool:
    """True for Cython functions / bound methods."""

    def check(x):
        return type(x).__name__ == "cython_function_or_method"

    return check(obj) or (hasattr(obj, "__func__") and check(obj.__func__))


def dict2tuple(dic):
    """``{'k': 'v'}`` → ``(('k', 'v'),)``; tuples and None pass through."""
    if dic is None or isinstance(dic, tuple):
        return dic
    if isinstance(dic, dict):
        return tuple(dic.items())
    logger.warning(f"Unable to convert type {type(dic)} to tuple, skip converting {dic}.")
    return dic


_ADDRESS_RULES = (
    re.compile(r"^\d{1,3}(\.\d{1,3}){3}:\d+$"),  # ip:port
    re.compile(r"^[a-zA-Z0-9.-]+:\d+$"),  # hostname:port
    re.compile(r"^https?://.+"),  # http(s) link (L7 gateway)
    re.compile(r"^unix:.+"),  # unix domain socket (native transport only)
)


def validate_address(address: str) -> None:
    if address is None:
        raise ValueError("The address shouldn't be None.")
    if address in ("local", "localhost"):
        return
    if isinstance(address, str) and any(r.match(address) for r in _ADDRESS_RULES):
        return
    raise ValueError(
        "The address format is invalid. It should be in one of the following formats:\n"
        "- `local` for starting a new cluster, or `localhost` for connecting a local cluster.\n"
        "- 'ip:port' format, where the IP needs to follow the IP address specifications and the port is a number.\n"
        "- 'hostname:port' format, where the hostname is a string and the port is a number.\n"
        "- An HTTPS or HTTP link starting with 'https://' or 'http://'."
    )


def validate_addresses(addresses: dict):
    for address in addresses.values():
        assert isinstance(address, str) and address, f"Address should be string but got {address}."
        validate_address(address)


def start_command(command: str, timeout=60):
    """Run a shell command; raise if it wrote to stderr."""
    proc = subprocess.Popen(command, shell=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    output, error = proc.communicate(timeout=timeout)
    if len(error) != 0:
        raise RuntimeError(f"Failed to start command [{command}], the error is:\n {error.decode()}")
    return output
