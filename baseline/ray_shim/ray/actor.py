from . import ActorClass, ActorHandle, ActorMethod  # noqa: F401
