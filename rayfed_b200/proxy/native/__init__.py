"""Native (C++ framed-socket) host transport for parties sharing a box."""
from .socket_proxy import NativeReceiverProxy, NativeSenderProxy, NativeSenderReceiverProxy  # noqa: F401
