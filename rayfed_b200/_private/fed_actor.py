"""Fed actors (parity: fed/_private/fed_actor.py:26-145).

``FedActorHandle`` exists in every party; only the owner party instantiates
the class, on its own ordered mailbox lane of the party runtime (and its own
CUDA stream when the party has a device).  Attribute access validates the
method name against the class body in *every* party (so a typo fails
everywhere, fed_actor.py:48-49) and returns a ``FedActorMethod`` bound to a
fresh call holder.
"""
from __future__ import annotations

import logging

from ..core import get_runtime
from .fed_call_holder import FedCallHolder

logger = logging.getLogger(__name__)

# ``options`` keys understood by the local runtime; anything else (num_cpus,
# resources, ...) is accepted for source compatibility and ignored.
_ACTOR_OPTION_KEYS = ("max_concurrency", "name")


class FedActorHandle:
    def __init__(self, fed_class_task_id, addresses, cls, party, node_party, options) -> None:
        self._fed_class_task_id = fed_class_task_id
        self._addresses = addresses
        self._body = cls
        self._party = party
        self._node_party = node_party
        self._options = options or {}
        self._actor_handle = None

    def __getattr__(self, method_name: str):
        # Raises AttributeError for unknown methods — in every party.
        getattr(self._body, method_name)
        # actor-level options apply to every method (reference: fed_actor.py:67,76)
        method = FedActorMethod(self._addresses, self._party, self._node_party, self, method_name, self._options)
        if not method_name.startswith("_"):
            # `actor.method` is looked up once: later accesses find the bound method in the instance
            # dict and never come back here (`.options()` hands out a copy, so sharing it is safe)
            self.__dict__[method_name] = method
        return method

    def _execute_impl(self, cls_args, cls_kwargs):
        """Instantiate — owner party only."""
        if self._node_party == self._party:
            opts = self._options
            self._actor_handle = get_runtime().create_actor(
                self._body, cls_args, cls_kwargs,
                name=opts.get("name", "") or "",
                max_concurrency=opts.get("max_concurrency", 1) or 1,
            )
            return self._actor_handle._ready_ref
        return None

    def _execute_remote_method(self, method_name, options, _args, _kwargs):
        num_returns = (options or {}).get("num_returns", 1) or 1
        return self._actor_handle.submit_method(method_name, _args, _kwargs, num_returns)


class FedActorMethod:
    __slots__ = ("_addresses", "_party", "_node_party", "_fed_actor_handle", "_method_name", "_options",
                 "_fed_call_holder")

    def __init__(self, addresses, party, node_party, fed_actor_handle, method_name, options=None) -> None:
        self._addresses = addresses
        self._party = party
        self._node_party = node_party
        self._fed_actor_handle = fed_actor_handle
        self._method_name = method_name
        self._options = dict(options) if options else {}
        self._fed_call_holder = FedCallHolder(node_party, self._execute_impl, self._options)

    def remote(self, *args, **kwargs):
        return self._fed_call_holder.internal_remote(*args, **kwargs)

    def options(self, **options):
        """A method bound with call options (``num_returns=...``).  Returns a NEW bound method: the
        one cached on the actor handle stays as it is (in the reference every attribute access
        makes a fresh object, fed_actor.py:62-80, so options never leak between calls either)."""
        return FedActorMethod(self._addresses, self._party, self._node_party, self._fed_actor_handle,
                              self._method_name, options)

    def _execute_impl(self, args, kwargs):
        return self._fed_actor_handle._execute_remote_method(self._method_name, self._options, args, kwargs)
