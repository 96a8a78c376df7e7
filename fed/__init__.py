"""``import fed`` compatibility shim: code written against ray-project/rayfed
keeps working unchanged.  Every ``fed.*`` module path of the reference that has
a counterpart here is aliased to it (``fed.api``, ``fed.config``,
``fed.exceptions``, ``fed.fed_object``, ``fed.tree_util``, ``fed.utils``,
``fed.cleanup``, ``fed.proxy.barriers``, ``fed.proxy.base_proxy``,
``fed.proxy.grpc.grpc_proxy``, ``fed._private.*``)."""
import importlib
import sys

import rayfed_b200 as _impl
from rayfed_b200 import *  # noqa: F401,F403
from rayfed_b200 import __all__, __version__  # noqa: F401

_ALIASES = {
    "fed.api": "rayfed_b200.api",
    "fed.config": "rayfed_b200.config",
    "fed.exceptions": "rayfed_b200.exceptions",
    "fed.fed_object": "rayfed_b200.fed_object",
    "fed.tree_util": "rayfed_b200.tree_util",
    "fed.utils": "rayfed_b200.utils",
    "fed.cleanup": "rayfed_b200.cleanup",
    "fed.proxy": "rayfed_b200.proxy",
    "fed.proxy.barriers": "rayfed_b200.proxy.barriers",
    "fed.proxy.base_proxy": "rayfed_b200.proxy.base_proxy",
    "fed.proxy.grpc": "rayfed_b200.proxy.grpc",
    "fed.proxy.grpc.grpc_proxy": "rayfed_b200.proxy.grpc.grpc_proxy",
    "fed.proxy.grpc.grpc_options": "rayfed_b200.proxy.grpc.grpc_options",
    "fed._private": "rayfed_b200._private",
    "fed._private.constants": "rayfed_b200._private.constants",
    "fed._private.global_context": "rayfed_b200._private.global_context",
    "fed._private.fed_actor": "rayfed_b200._private.fed_actor",
    "fed._private.fed_call_holder": "rayfed_b200._private.fed_call_holder",
    "fed._private.message_queue": "rayfed_b200._private.message_queue",
    "fed._private.serialization_utils": "rayfed_b200._private.serialization",
    "fed._private.compatible_utils": "rayfed_b200._private.kv",
    "fed.fl": "rayfed_b200.fl",
    "fed.ops": "rayfed_b200.ops",
}
for _alias, _target in _ALIASES.items():
    _mod = importlib.import_module(_target)
    sys.modules[_alias] = _mod
    _parent, _, _leaf = _alias.rpartition(".")
    if _parent in sys.modules and _parent != "fed":
        setattr(sys.modules[_parent], _leaf, _mod) if not hasattr(sys.modules[_parent], _leaf) else None
for _name in ("api", "config", "exceptions", "fed_object", "tree_util", "utils", "cleanup", "proxy", "fl", "ops"):
    globals()[_name] = sys.modules[f"fed.{_name}"]
_private = sys.modules["fed._private"]
