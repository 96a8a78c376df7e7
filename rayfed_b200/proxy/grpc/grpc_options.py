"""Default gRPC channel options (parity: fed/proxy/grpc/grpc_options.py:19-83):
retries on UNAVAILABLE (5 attempts, 5 s → 30 s, ×2), 500 MiB message caps,
``so_reuseport`` off so a port clash is an error rather than silent sharing."""
import json

_GRPC_SERVICE = "GrpcService"

_DEFAULT_GRPC_RETRY_POLICY = {
    "maxAttempts": 5,
    "initialBackoff": "5s",
    "maxBackoff": "30s",
    "backoffMultiplier": 2,
    "retryableStatusCodes": ["UNAVAILABLE"],
}

_DEFAULT_GRPC_MAX_SEND_MESSAGE_LENGTH = 500 * 1024 * 1024
_DEFAULT_GRPC_MAX_RECEIVE_MESSAGE_LENGTH = 500 * 1024 * 1024


def _service_config(retry_policy) -> str:
    return json.dumps({"methodConfig": [{"name": [{"service": _GRPC_SERVICE}], "retryPolicy": retry_policy}]})


_DEFAULT_GRPC_CHANNEL_OPTIONS = {
    "grpc.enable_retries": 1,
    "grpc.so_reuseport": 0,
    "grpc.max_send_message_length": _DEFAULT_GRPC_MAX_SEND_MESSAGE_LENGTH,
    "grpc.max_receive_message_length": _DEFAULT_GRPC_MAX_RECEIVE_MESSAGE_LENGTH,
    "grpc.service_config": _service_config(_DEFAULT_GRPC_RETRY_POLICY),
}


def get_grpc_options(retry_policy=None, max_send_message_length=None, max_receive_message_length=None):
    """List-of-tuples form of the defaults with optional overrides."""
    return [
        ("grpc.max_send_message_length", max_send_message_length or _DEFAULT_GRPC_MAX_SEND_MESSAGE_LENGTH),
        ("grpc.max_receive_message_length", max_receive_message_length or _DEFAULT_GRPC_MAX_RECEIVE_MESSAGE_LENGTH),
        ("grpc.enable_retries", 1),
        ("grpc.service_config", _service_config(retry_policy or _DEFAULT_GRPC_RETRY_POLICY)),
        ("grpc.so_reuseport", 0),
    ]
