"""Shared constants (parity: fed/_private/constants.py:16-44).

Key names and the KV namespacing are kept identical to the reference so that
tooling written against it (e.g. tests that inspect ``RAYFED#job#key``) keeps
working; the values are served by an in-process store instead of Ray's GCS.
"""

KEY_OF_CLUSTER_CONFIG = "CLUSTER_CONFIG"
KEY_OF_JOB_CONFIG = "JOB_CONFIG"
KEY_OF_GRPC_METADATA = "GRPC_METADATA"

KEY_OF_CLUSTER_ADDRESSES = "CLUSTER_ADDRESSES"
KEY_OF_CURRENT_PARTY_NAME = "CURRENT_PARTY_NAME"
KEY_OF_TLS_CONFIG = "TLS_CONFIG"
KEY_OF_CROSS_SILO_COMM_CONFIG_DICT = "CROSS_SILO_COMM_CONFIG_DICT"

RAYFED_LOG_FMT = (
    "%(asctime)s.%(msecs)03d %(levelname)s %(filename)s:%(lineno)s"
    " [%(party)s] -- [%(jobname)s] %(message)s"
)
RAYFED_DATE_FMT = "%Y-%m-%d %H:%M:%S"

RAYFED_DEFAULT_JOB_NAME = "Anonymous_job"
RAYFED_JOB_KV_DATA_KEY_FMT = "RAYFED#{}#{}"

RAYFED_DEFAULT_SENDER_PROXY_ACTOR_NAME = "SenderProxyActor"
RAYFED_DEFAULT_RECEIVER_PROXY_ACTOR_NAME = "ReceiverProxyActor"
RAYFED_DEFAULT_SENDER_RECEIVER_PROXY_ACTOR_NAME = "SenderReceiverProxyActor"

# Transport selection (``config['cross_silo_comm']['transport']``).
TRANSPORT_AUTO = "auto"
TRANSPORT_NATIVE = "native"  # C++ framed sockets (+ NVLink for CUDA tensors)
TRANSPORT_GRPC = "grpc"  # wire-compatible with the reference, TLS capable
TRANSPORT_NVLINK = "nvlink"  # native + mandatory device path
