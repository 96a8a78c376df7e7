"""Transport plug-in contracts (parity: fed/proxy/base_proxy.py:21-106).

A transport is supplied to ``fed.init`` as classes implementing these ABCs:
``sender_proxy_cls`` + ``receiver_proxy_cls``, or one combined
``receiver_sender_proxy_cls``.  Methods may be plain or ``async``; a plain
method may return a native `Ref` (completion future) instead of blocking.

Constructor signature (all three): ``(addresses, party, job_name, tls_config,
proxy_config=None)`` — ``proxy_config`` arrives as the raw ``cross_silo_comm``
dict, exactly as in the reference (fed/proxy/barriers.py:136-143).
"""
from __future__ import annotations

import abc
from typing import Dict

from ..config import CrossSiloMessageConfig


class SenderProxy(abc.ABC):
    def __init__(self, addresses: Dict, party: str, job_name: str, tls_config: Dict,
                 proxy_config: CrossSiloMessageConfig = None) -> None:
        self._addresses = addresses
        self._party = party
        self._tls_config = tls_config
        self._proxy_config = proxy_config
        self._job_name = job_name

    async def is_ready(self):
        return True

    async def get_proxy_config(self, dest_party=None):
        return self._proxy_config

    @abc.abstractmethod
    def send(self, dest_party, data, upstream_seq_id, downstream_seq_id):
        """Deliver ``data`` to ``dest_party`` under key (upstream, downstream).
        Return truthy on success (or a `Ref` / awaitable that resolves to it)."""

    def stop(self):
        """Release sockets / channels. Optional."""


class ReceiverProxy(abc.ABC):
    def __init__(self, listening_address: str, party: str, job_name: str, tls_config: Dict,
                 proxy_config: CrossSiloMessageConfig = None) -> None:
        self._listening_address = listening_address
        self._party = party
        self._tls_config = tls_config
        self._proxy_config = proxy_config
        self._job_name = job_name

    @abc.abstractmethod
    def start(self):
        """Start serving (may block for async servers; see ProxyHost)."""

    @abc.abstractmethod
    def is_ready(self):
        """Return ``(ok: bool, message: str | None)``."""

    @abc.abstractmethod
    def get_data(self, src_party, upstream_seq_id, curr_seq_id):
        """Value of the message keyed (upstream, curr) — blocking, awaitable, or a `Ref`."""

    async def get_proxy_config(self):
        return self._proxy_config

    def stop(self):
        """Stop serving. Optional."""


class SenderReceiverProxy(abc.ABC):
    def __init__(self, addresses: Dict, self_party: str, tls_config: Dict,
                 proxy_config: CrossSiloMessageConfig = None) -> None:
        self._addresses = addresses
        self._party = self_party
        self._tls_config = tls_config
        self._proxy_config = proxy_config

    @abc.abstractmethod
    def start(self):
        pass

    @abc.abstractmethod
    def get_data(self, src_party, upstream_seq_id, curr_seq_id):
        pass

    @abc.abstractmethod
    def send(self, dest_party, data, upstream_seq_id, downstream_seq_id):
        pass

    async def is_ready(self):
        return True

    async def get_proxy_config(self):
        return self._proxy_config

    def stop(self):
        pass
