"""Native host transport: Python face of ``fedcore::SocketTransport``.

Same contract as the reference's gRPC pair (fed/proxy/grpc/grpc_proxy.py) —
one acknowledged message per object keyed (upstream, downstream), job-name
check (417), 4xx ⇒ failed send, per-message timeout, connect retry/backoff —
but frames travel over C++ sockets and land directly in the C++ `Mailbox`:

* ``send`` serialises, hands the bytes to the C++ writer, and returns the ack
  `Ref` immediately (nothing blocks, nothing is awaited on a Python loop);
* ``get_data`` returns the Mailbox `Ref`; decoding (unpickle / allow-list) runs
  lazily on the thread that consumes the value.

Retry defaults mirror fed/proxy/grpc/grpc_options.py:19-25 (5 attempts, 5 s →
30 s, ×2) so asynchronous party start-up behaves like the reference.
"""
from __future__ import annotations

import logging
from typing import Dict

from ..._private.serialization import Deserializer, dumps
from ...config import CrossSiloMessageConfig
from ...core import native
from ...exceptions import FedRemoteError, FedTaskError
from ..base_proxy import ReceiverProxy, SenderProxy, SenderReceiverProxy

logger = logging.getLogger(__name__)

_DEFAULT_RETRY = dict(max_attempts=5, initial_backoff_s=5.0, max_backoff_s=30.0, multiplier=2.0)


def _retry_from_config(cfg: CrossSiloMessageConfig) -> dict:
    r = dict(_DEFAULT_RETRY)
    if cfg.retry_max_attempts is not None:
        r["max_attempts"] = int(cfg.retry_max_attempts)
    if cfg.retry_initial_backoff_s is not None:
        r["initial_backoff_s"] = float(cfg.retry_initial_backoff_s)
    if cfg.retry_max_backoff_s is not None:
        r["max_backoff_s"] = float(cfg.retry_max_backoff_s)
    if cfg.retry_backoff_multiplier is not None:
        r["multiplier"] = float(cfg.retry_backoff_multiplier)
    return r


def _ack_decoder(data: bytes, is_error: bool):
    """Python flavour of the ack decoding (the transport's default is the native one)."""
    if is_error:
        raise RuntimeError(data.decode("utf-8", "replace"))
    return True


class _PayloadDecoder:
    """Bytes → value on the consumer thread; remote errors become the Ref's error."""

    def __init__(self, deserializer: Deserializer):
        self._loads = deserializer

    def __call__(self, data: bytes, is_error: bool):
        obj = self._loads(data)
        if isinstance(obj, FedRemoteError):
            raise FedTaskError(obj, "", "recv")
        return obj


class NativeSenderProxy(SenderProxy):
    REFS_RAISE_REMOTE_ERRORS = True

    def __init__(self, addresses: Dict, party: str, job_name: str, tls_config: Dict, proxy_config: Dict = None):
        cfg = CrossSiloMessageConfig.from_dict(proxy_config)
        super().__init__(addresses, party, job_name, tls_config, cfg)
        nat = native()
        self._transport = nat.SocketTransport(party, job_name, nat.Mailbox())
        retry = _retry_from_config(cfg)
        timeout_s = (cfg.timeout_in_ms or 60000) / 1000.0
        for name, addr in addresses.items():
            self._transport.add_peer(
                name, addr, retry["max_attempts"], retry["initial_backoff_s"], retry["max_backoff_s"],
                retry["multiplier"], timeout_s,
                # proxy_max_restarts (fed/proxy/barriers.py:301-307): un-acknowledged sends are
                # retried N times on a re-established connection instead of failing
                int(cfg.proxy_max_restarts or 0),
            )
        self._max_bytes = cfg.messages_max_size_in_bytes

    def is_ready(self):
        return True

    def get_proxy_config(self, dest_party=None):
        return self._proxy_config

    def send(self, dest_party, data, upstream_seq_id, downstream_seq_id):
        payload = dumps(data)
        if self._max_bytes is not None and len(payload) > self._max_bytes:
            raise RuntimeError(
                f"message of {len(payload)} bytes exceeds messages_max_size_in_bytes={self._max_bytes}"
            )
        # the ack future decodes natively (True / RuntimeError(text)): no Python decoder object, so
        # transport threads complete, forward and drop it without the GIL
        return self._transport.send(
            dest_party, str(upstream_seq_id), str(downstream_seq_id), payload, isinstance(data, FedRemoteError)
        )

    # -- owner -> all-parties broadcast (fed.get): serialise ONCE, N-1 native sends ------------
    def can_broadcast(self) -> bool:
        return True

    def broadcast(self, dest_parties, data, upstream_seq_id, downstream_seq_id):
        """Same object to every party in ``dest_parties``; returns {dest: ack Ref}.  The
        reference pickles and issues one RPC per destination (fed/api.py:564-579)."""
        payload = dumps(data)
        if self._max_bytes is not None and len(payload) > self._max_bytes:
            raise RuntimeError(
                f"message of {len(payload)} bytes exceeds messages_max_size_in_bytes={self._max_bytes}"
            )
        dest_parties = list(dest_parties)
        acks = self._transport.send_many(dest_parties, str(upstream_seq_id), str(downstream_seq_id), payload,
                                         isinstance(data, FedRemoteError))
        return dict(zip(dest_parties, acks))

    def drain(self, timeout: float = 60.0) -> bool:
        return self._transport.drain(timeout)

    def get_stats(self):
        return dict(self._transport.stats())

    def stop(self):
        self._transport.stop()


class NativeReceiverProxy(ReceiverProxy):
    REFS_RAISE_REMOTE_ERRORS = True

    def __init__(self, listening_address: str, party: str, job_name: str, tls_config: Dict,
                 proxy_config: Dict = None):
        cfg = CrossSiloMessageConfig.from_dict(proxy_config)
        super().__init__(listening_address, party, job_name, tls_config, cfg)
        nat = native()
        self._mailbox = nat.Mailbox()
        self._transport = nat.SocketTransport(party, job_name, self._mailbox)
        if cfg.messages_max_size_in_bytes:
            # enforced on RECEIVE, before any allocation (the reference: grpc.max_receive_message_length)
            self._transport.set_max_message_bytes(int(cfg.messages_max_size_in_bytes))
        self._decoder = _PayloadDecoder(Deserializer(cfg.serializing_allowed_list))
        self._state = None

    def start(self):
        err = self._transport.listen(self._listening_address)
        if err:
            self._state = (
                False,
                f"Failed to listen on {self._listening_address}: {err}",
            )
        else:
            self._state = (True, None)

    def is_ready(self):
        return self._state if self._state is not None else (False, "receiver proxy was not started")

    def get_proxy_config(self):
        return self._proxy_config

    def get_data(self, src_party, upstream_seq_id, curr_seq_id):
        ref = self._mailbox.claim(str(upstream_seq_id), str(curr_seq_id))
        ref.set_decoder(self._decoder)
        return ref

    def get_data_many(self, items, curr_seq_id):
        """``items``: [(src_party, upstream_seq_id)] — the remote arguments of one task; one native
        call claims them all."""
        return self._mailbox.claim_many([str(up) for _, up in items], str(curr_seq_id), self._decoder)

    @property
    def mailbox(self):
        return self._mailbox

    def get_stats(self):
        s = dict(self._transport.stats())
        s.update({f"mailbox_{k}": v for k, v in self._mailbox.stats().items()})
        return s

    def stop(self):
        self._transport.stop()
        self._mailbox.clear()


class NativeSenderReceiverProxy(SenderReceiverProxy):
    """Both directions behind one object (``receiver_sender_proxy_cls``)."""

    REFS_RAISE_REMOTE_ERRORS = True

    def __init__(self, addresses: Dict, self_party: str, tls_config: Dict, proxy_config: Dict = None):
        cfg = CrossSiloMessageConfig.from_dict(proxy_config)
        super().__init__(addresses, self_party, tls_config, cfg)
        from ..._private.global_context import get_global_context

        ctx = get_global_context()
        job = ctx.get_job_name() if ctx is not None else ""
        self._recv = NativeReceiverProxy(addresses[self_party], self_party, job, tls_config, proxy_config)
        self._send = NativeSenderProxy(addresses, self_party, job, tls_config, proxy_config)

    def start(self):
        self._recv.start()

    def is_ready(self):
        return self._recv.is_ready()

    def get_data(self, src_party, upstream_seq_id, curr_seq_id):
        return self._recv.get_data(src_party, upstream_seq_id, curr_seq_id)

    def send(self, dest_party, data, upstream_seq_id, downstream_seq_id):
        return self._send.send(dest_party, data, upstream_seq_id, downstream_seq_id)

    def get_proxy_config(self):
        return self._proxy_config

    def get_stats(self):
        s = self._send.get_stats()
        r = self._recv.get_stats()
        s["receive_op_count"] = r["receive_op_count"]
        s["bytes_received"] = r["bytes_received"]
        return s

    def drain(self, timeout: float = 60.0) -> bool:
        return self._send.drain(timeout)

    def stop(self):
        self._send.stop()
        self._recv.stop()
