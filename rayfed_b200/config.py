"""Job / cluster configuration (parity: fed/config.py:15-195).

``fed.init`` stores a pickled *cluster config* (addresses, current party, TLS)
and *job config* (the raw ``cross_silo_comm`` dict) in the KV store under
``RAYFED#{job}#CLUSTER_CONFIG|JOB_CONFIG``; any component reads them back
through :func:`get_cluster_config` / :func:`get_job_config`, which cache per
process (fed/config.py:46-75).  ``from_dict`` silently drops unknown keys
(fed/config.py:156-161) — kept, since reference user configs rely on it.
"""
from __future__ import annotations

import json
import pickle
from dataclasses import dataclass, fields
from typing import Dict, List, Optional

from ._private import constants as fed_constants
from ._private import kv as fed_kv


class ClusterConfig:
    """Read-only view of the pickled cluster config."""

    def __init__(self, raw_bytes: bytes) -> None:
        self._data = pickle.loads(raw_bytes)

    @property
    def cluster_addresses(self):
        return self._data[fed_constants.KEY_OF_CLUSTER_ADDRESSES]

    @property
    def current_party(self):
        return self._data[fed_constants.KEY_OF_CURRENT_PARTY_NAME]

    @property
    def tls_config(self):
        return self._data[fed_constants.KEY_OF_TLS_CONFIG]


class JobConfig:
    def __init__(self, raw_bytes: Optional[bytes]) -> None:
        self._data = {} if raw_bytes is None else pickle.loads(raw_bytes)

    @property
    def cross_silo_comm_config_dict(self) -> Dict:
        return self._data.get(fed_constants.KEY_OF_CROSS_SILO_COMM_CONFIG_DICT, {})


_cluster_config: Optional[ClusterConfig] = None
_job_config: Optional[JobConfig] = None


def get_cluster_config(job_name: str = None) -> ClusterConfig:
    global _cluster_config
    if _cluster_config is None:
        assert job_name is not None, "Initializing internal kv need to provide job_name."
        fed_kv._init_internal_kv(job_name)
        raw = fed_kv.kv.get(fed_constants.KEY_OF_CLUSTER_CONFIG)
        assert raw is not None, "fed.init() has not stored a cluster config for this job."
        _cluster_config = ClusterConfig(raw)
    return _cluster_config


def get_job_config(job_name: str = None) -> JobConfig:
    global _job_config
    if _job_config is None:
        assert job_name is not None, "Initializing internal kv need to provide job_name."
        fed_kv._init_internal_kv(job_name)
        _job_config = JobConfig(fed_kv.kv.get(fed_constants.KEY_OF_JOB_CONFIG))
    return _job_config


def _reset_config_cache() -> None:
    global _cluster_config, _job_config
    _cluster_config = None
    _job_config = None


@dataclass
class CrossSiloMessageConfig:
    """Options shared by every cross-party transport.

    The first twelve fields are the reference's (fed/config.py:125-136). The
    remaining ones configure the B200-native transports and are ignored by gRPC.

    Attributes:
        proxy_max_restarts: restarts allowed for the sender proxy host.
        timeout_in_ms: per-message timeout (send ack / proxy readiness).
        messages_max_size_in_bytes: cap for a single host-path message.
        exit_on_sending_failure: on a failed send, SIGINT self → failure handler
            → exit code 1.
        continue_waiting_for_data_sending_on_error: keep draining sends at
            shutdown even after an error.
        serializing_allowed_list: ``{module: [names] | None}`` unpickle allow-list.
        send_resource_label / recv_resource_label: accepted for compatibility
            (there is no cluster scheduler to place proxies on).
        http_header: extra metadata sent with each gRPC request.
        max_concurrency: concurrency of the proxy hosts.
        expose_error_trace: ship the original exception text to peers.
        use_global_proxy: share proxy names across jobs, or per-job names.
        transport: ``auto`` | ``native`` | ``grpc`` | ``nvlink``.
        retry_max_attempts / retry_initial_backoff_s / retry_max_backoff_s /
            retry_backoff_multiplier: connect-retry policy of the native
            transport (defaults mirror fed/proxy/grpc/grpc_options.py:19-25).
        device_arena_bytes: per (src→dst) receive-arena size in the symmetric heap.
        device_wire_dtype: optional on-the-wire cast fused into the push kernel
            (``None`` | ``bf16`` | ``fp8_e4m3``).
        device_min_bytes: CUDA tensors smaller than this ride the host frame.
        device_zero_copy: hand received tensors out as views of the arena (default True);
            ``False`` always copies them out with the unpack kernel.
        device_zero_copy_min_bytes: blocks smaller than this are copied out (default 64 KiB).
        device_wait_timeout_ms: patience of the watchdog behind a parked device receive
            (default: 10 x ``timeout_in_ms``, at least 10 min — a receive legitimately lasts as
            long as the producer's device work).
    """

    proxy_max_restarts: int = None
    timeout_in_ms: int = 60000
    messages_max_size_in_bytes: int = None
    exit_on_sending_failure: Optional[bool] = False
    continue_waiting_for_data_sending_on_error: Optional[bool] = False
    serializing_allowed_list: Optional[Dict[str, str]] = None
    send_resource_label: Optional[Dict[str, str]] = None
    recv_resource_label: Optional[Dict[str, str]] = None
    http_header: Optional[Dict[str, str]] = None
    max_concurrency: Optional[int] = None
    expose_error_trace: Optional[bool] = False
    use_global_proxy: Optional[bool] = True
    # ---- B200-native additions ------------------------------------------------
    transport: Optional[str] = "auto"
    retry_max_attempts: Optional[int] = None
    retry_initial_backoff_s: Optional[float] = None
    retry_max_backoff_s: Optional[float] = None
    retry_backoff_multiplier: Optional[float] = None
    device_arena_bytes: Optional[int] = None
    device_wire_dtype: Optional[str] = None
    device_min_bytes: Optional[int] = None
    device_zero_copy: Optional[bool] = None
    device_zero_copy_min_bytes: Optional[int] = None
    device_wait_timeout_ms: Optional[int] = None

    def __json__(self):
        return json.dumps(self.__dict__)

    @classmethod
    def from_json(cls, json_str):
        return cls(**json.loads(json_str))

    @classmethod
    def from_dict(cls, data: Dict) -> "CrossSiloMessageConfig":
        data = data or {}
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in data.items() if k in known})


@dataclass
class GrpcCrossSiloMessageConfig(CrossSiloMessageConfig):
    """gRPC-only knobs (parity: fed/config.py:164-195).

    Attributes:
        grpc_channel_options: list of ``(name, value)`` channel args; overrides
            ``messages_max_size_in_bytes`` when it names the same option.
        grpc_retry_policy: dict placed in the channel's service config.
    """

    grpc_channel_options: List = None
    grpc_retry_policy: Dict[str, str] = None
