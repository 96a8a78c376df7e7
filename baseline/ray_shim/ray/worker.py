from ._private.worker import _global_node, global_worker  # noqa: F401
