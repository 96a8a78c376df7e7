"""Party-local runtime: the native replacement for what the reference borrowed
from Ray core (tasks, actors, futures).  See :mod:`rayfed_b200.core.runtime`."""
from .runtime import (  # noqa: F401
    ActorHandle,
    PartyRuntime,
    Ref,
    get,
    get_runtime,
    is_ref,
    is_refs,
    native,
    shutdown_runtime,
    wait,
)
