"""Cross-party (de)serialization for the host path.

* ``dumps`` — fast path for exact builtin scalars (C pickle), cloudpickle for
  everything else (same wire format as the reference: a pickle stream,
  fed/proxy/grpc/grpc_proxy.py:202).
* ``loads`` — plain unpickle, or a restricted unpickler when the job config
  carries ``serializing_allowed_list``.

Parity for the allow-list: fed/_private/serialization_utils.py:24-83 —
``{module: [names] | None}``; ``None`` or ``"*"`` admits the whole module; a
trailing ``.*`` on a module key admits sub-modules.  The reference monkey-patches
``cloudpickle.loads`` process-wide (serialization_utils.py:83); here the policy
is an explicit object handed to the receiver, so nothing global is patched.
"""
from __future__ import annotations

import io
import pickle
from typing import Dict, Iterable, Optional

import cloudpickle

_FAST_TYPES = (type(None), bool, int, float, str, bytes)


def dumps(obj) -> bytes:
    if type(obj) in _FAST_TYPES:
        return pickle.dumps(obj, protocol=5)
    return cloudpickle.dumps(obj, protocol=5)


class _RestrictedUnpickler(pickle.Unpickler):
    def __init__(self, file, allowed: Dict[str, Optional[Iterable[str]]]):
        super().__init__(file)
        self._allowed = allowed

    def _module_rule(self, module: str):
        if module in self._allowed:
            return True, self._allowed[module]
        # "pkg.*" admits pkg and every sub-module
        parts = module.split(".")
        for i in range(len(parts), 0, -1):
            key = ".".join(parts[:i]) + ".*"
            if key in self._allowed:
                return True, self._allowed[key]
        return False, None

    def find_class(self, module, name):
        found, names = self._module_rule(module)
        if found and (names is None or names == "*" or name in names or "*" in names):
            return super().find_class(module, name)
        if module == "builtins" and name in ("Exception", "BaseException"):
            # not whitelisted either: fall through to the error below
            pass
        raise pickle.UnpicklingError(f"{module}.{name} is not allowed by serializing_allowed_list.")


def restricted_loads(data: bytes, allowed: Dict[str, Optional[Iterable[str]]]):
    if isinstance(data, str):
        raise TypeError("Can't load pickle from unicode string")
    return _RestrictedUnpickler(io.BytesIO(data), allowed).load()


class Deserializer:
    """Callable unpickler bound to an optional allow-list."""

    # What the runtime itself must always be able to ship (error replacement).
    _ALWAYS = {
        "rayfed_b200.exceptions": None,
        "fed.exceptions": None,
    }

    def __init__(self, allowed: Optional[Dict[str, Optional[Iterable[str]]]] = None):
        if allowed is not None:
            merged = dict(self._ALWAYS)
            merged.update(allowed)
            allowed = merged
        self._allowed = allowed

    @property
    def allowed(self):
        return self._allowed

    def __call__(self, data: bytes):
        if self._allowed is None:
            return pickle.loads(data)
        return restricted_loads(data, self._allowed)


def _apply_loads_function_with_whitelist(allowed=None) -> Deserializer:
    """Name kept from the reference; returns the policy object instead of
    patching ``cloudpickle`` globally."""
    return Deserializer(allowed)
