"""gRPC transport (parity: fed/proxy/grpc/grpc_proxy.py).

Kept as the **off-box fallback** and as the measured *reference-protocol*
baseline: the bytes on the wire are what the reference sends — a cloudpickle
blob inside unary ``GrpcService.SendData`` with string seq ids and the job name
(grpc_proxy.py:193-220) — and the receiver keeps the reference's rendezvous
semantics (whichever of SendData/get_data comes first creates the entry,
:272-293 vs :309-342), the 417 job-name refusal (:310-320), 4xx ⇒ send failure
(:179-190), TLS / mutual TLS (:128-139,:357-375), ``http_header`` → gRPC
metadata (:113,:156-168), option precedence defaults <
``messages_max_size_in_bytes`` < ``grpc_channel_options`` <
``grpc_retry_policy`` (:64-99).

Implementation differences: the rendezvous table is the native `Mailbox`
(arrival completes a `Ref`; no asyncio.Event pair, no lock held across awaits),
and stubs come from :mod:`.fed_pb` (runtime-built descriptors).
"""
from __future__ import annotations

import asyncio
import copy
import json
import logging
import threading
from typing import Dict

import grpc

from ... import utils as fed_utils
from ..._private.serialization import Deserializer, dumps
from ...config import CrossSiloMessageConfig, GrpcCrossSiloMessageConfig
from ...core import native
from ..base_proxy import ReceiverProxy, SenderProxy
from . import fed_pb
from .grpc_options import _DEFAULT_GRPC_CHANNEL_OPTIONS, _GRPC_SERVICE

logger = logging.getLogger(__name__)


def parse_grpc_options(proxy_config: CrossSiloMessageConfig) -> dict:
    """Project config fields onto gRPC channel-arg names (later wins)."""
    opts = {}
    if proxy_config is None:
        return opts
    if isinstance(proxy_config, CrossSiloMessageConfig) and proxy_config.messages_max_size_in_bytes is not None:
        opts["grpc.max_send_message_length"] = proxy_config.messages_max_size_in_bytes
        opts["grpc.max_receive_message_length"] = proxy_config.messages_max_size_in_bytes
    if isinstance(proxy_config, GrpcCrossSiloMessageConfig):
        if proxy_config.grpc_channel_options is not None:
            opts.update(proxy_config.grpc_channel_options)
        if proxy_config.grpc_retry_policy is not None:
            opts["grpc.service_config"] = json.dumps(
                {"methodConfig": [{"name": [{"service": _GRPC_SERVICE}],
                                   "retryPolicy": proxy_config.grpc_retry_policy}]}
            )
    return opts


class GrpcSenderProxy(SenderProxy):
    def __init__(self, cluster: Dict, party: str, job_name: str, tls_config: Dict, proxy_config: Dict = None):
        proxy_config = GrpcCrossSiloMessageConfig.from_dict(proxy_config)
        super().__init__(cluster, party, job_name, tls_config, proxy_config)
        self._grpc_metadata = proxy_config.http_header or {}
        self._grpc_options = copy.deepcopy(_DEFAULT_GRPC_CHANNEL_OPTIONS)
        self._grpc_options.update(parse_grpc_options(self._proxy_config))
        self._stubs = {}  # dest party -> stub (channel is reused)
        self._channels = {}

    def _stub_for(self, dest_party):
        stub = self._stubs.get(dest_party)
        if stub is not None:
            return stub
        dest_addr = self._addresses[dest_party]
        _, channel_options = self.get_grpc_config_by_party(dest_party)
        if fed_utils.tls_enabled(self._tls_config):
            ca_cert, private_key, cert_chain = fed_utils.load_cert_config(self._tls_config)
            credentials = grpc.ssl_channel_credentials(
                certificate_chain=cert_chain, private_key=private_key, root_certificates=ca_cert
            )
            channel = grpc.aio.secure_channel(dest_addr, credentials, options=channel_options)
        else:
            channel = grpc.aio.insecure_channel(dest_addr, options=channel_options)
        stub = fed_pb.GrpcServiceStub(channel)
        self._channels[dest_party] = channel
        self._stubs[dest_party] = stub
        return stub

    async def send(self, dest_party, data, upstream_seq_id, downstream_seq_id):
        grpc_metadata, _ = self.get_grpc_config_by_party(dest_party)
        timeout = self._proxy_config.timeout_in_ms / 1000
        response = await send_data_grpc(
            data=data,
            stub=self._stub_for(dest_party),
            upstream_seq_id=upstream_seq_id,
            downstream_seq_id=downstream_seq_id,
            job_name=self._job_name,
            timeout=timeout,
            metadata=grpc_metadata,
        )
        self.handle_response_error(response)
        return response.result

    def get_grpc_config_by_party(self, dest_party):
        """(metadata, channel options) for a destination. Like the reference this
        is job-wide today; the per-party hook is kept for API parity."""
        grpc_metadata = dict(self._grpc_metadata)
        grpc_options = dict(self._grpc_options)
        cfg = self._proxy_config
        if cfg is not None:
            if cfg.http_header is not None:
                grpc_metadata.update(dict(cfg.http_header))
            grpc_options.update(parse_grpc_options(cfg))
        return grpc_metadata, fed_utils.dict2tuple(grpc_options)

    async def get_proxy_config(self, dest_party=None):
        if dest_party is None:
            grpc_options = fed_utils.dict2tuple(self._grpc_options)
        else:
            _, grpc_options = self.get_grpc_config_by_party(dest_party)
        proxy_config = dict(self._proxy_config.__dict__)
        proxy_config["grpc_options"] = grpc_options
        return proxy_config

    def handle_response_error(self, response):
        if response.code == 200:
            return
        if 400 <= response.code < 500:
            # delivered, but refused by the peer: still a sending failure
            logger.warning(
                f"Request was successfully sent but got error response, "
                f"code: {response.code}, message: {response.result}."
            )
            raise RuntimeError(response.result)

    async def stop(self):
        for ch in self._channels.values():
            try:
                await ch.close()
            except Exception:  # noqa: BLE001
                pass
        self._channels.clear()
        self._stubs.clear()


async def send_data_grpc(data, stub, upstream_seq_id, downstream_seq_id, timeout, job_name, metadata=None):
    payload = dumps(data)
    request = fed_pb.SendDataRequest(
        data=payload,
        upstream_seq_id=str(upstream_seq_id),
        downstream_seq_id=str(downstream_seq_id),
        job_name=job_name,
    )
    response = await stub.SendData(request, metadata=fed_utils.dict2tuple(metadata), timeout=timeout)
    logger.debug(
        f"Received data response from seq_id {upstream_seq_id} to {downstream_seq_id}, "
        f"code: {response.code}, result: {response.result}."
    )
    return response


class GrpcReceiverProxy(ReceiverProxy):
    def __init__(self, listen_addr: str, party: str, job_name: str, tls_config: Dict, proxy_config: Dict = None):
        proxy_config = GrpcCrossSiloMessageConfig.from_dict(proxy_config)
        super().__init__(listen_addr, party, job_name, tls_config, proxy_config)
        self._grpc_options = copy.deepcopy(_DEFAULT_GRPC_CHANNEL_OPTIONS)
        self._grpc_options.update(parse_grpc_options(self._proxy_config))
        self._mailbox = native().Mailbox()
        self._loads = Deserializer(proxy_config.serializing_allowed_list)
        self._ready = threading.Event()
        self._ready_result = None
        self._server = None

    async def start(self):
        port = self._listening_address[self._listening_address.rindex(":") + 1:]
        try:
            await _run_grpc_server(
                port, self._mailbox, self._party, self._job_name, self._tls_config,
                fed_utils.dict2tuple(self._grpc_options), self._on_ready, self._hold_server,
            )
        except RuntimeError as err:
            msg = (
                f"Grpc server failed to listen to port: {port}"
                f" Try another port by setting `listen_addr` into `cluster` config"
                f" when calling `fed.init`. Grpc error msg: {err}"
            )
            self._on_ready((False, msg))

    def _hold_server(self, server):
        self._server = server

    def _on_ready(self, result):
        self._ready_result = result
        self._ready.set()

    async def is_ready(self):
        while not self._ready.is_set():
            await asyncio.sleep(0.005)
        return self._ready_result

    def get_data(self, src_party, upstream_seq_id, curr_seq_id):
        logger.debug(f"Getting data for {curr_seq_id} from {upstream_seq_id}")
        ref = self._mailbox.claim(str(upstream_seq_id), str(curr_seq_id))
        loads = self._loads

        def decode(data: bytes, is_error: bool):
            return loads(data)

        ref.set_decoder(decode)
        return ref

    async def get_proxy_config(self):
        proxy_config = dict(self._proxy_config.__dict__)
        proxy_config["grpc_options"] = fed_utils.dict2tuple(self._grpc_options)
        return proxy_config

    async def stop(self):
        if self._server is not None:
            await self._server.stop(0.1)
            self._server = None


class SendDataService(fed_pb.GrpcServiceServicer):
    def __init__(self, mailbox, party, job_name):
        self._mailbox = mailbox
        self._party = party
        self._job_name = job_name

    async def SendData(self, request, context):
        job_name = request.job_name
        if job_name != self._job_name:
            logger.warning(
                f"Receive data from job {job_name}, ignore it. The reason may be that the "
                "ReceiverProxy is listening on the same address with that job."
            )
            return fed_pb.SendDataResponse(
                code=417, result=f"JobName mis-match, expected {self._job_name}, got {job_name}."
            )
        self._mailbox.deliver_raw(request.upstream_seq_id, request.downstream_seq_id, request.data, False)
        return fed_pb.SendDataResponse(code=200, result="OK")


async def _run_grpc_server(port, mailbox, party, job_name, tls_config=None, grpc_options=None, on_ready=None,
                           hold=None):
    logger.info(f"ReceiveProxy binding port {port}, options: {grpc_options}...")
    server = grpc.aio.server(options=grpc_options)
    fed_pb.add_GrpcServiceServicer_to_server(SendDataService(mailbox, party, job_name), server)
    if fed_utils.tls_enabled(tls_config):
        ca_cert, private_key, cert_chain = fed_utils.load_cert_config(tls_config)
        credentials = grpc.ssl_server_credentials(
            [(private_key, cert_chain)],
            root_certificates=ca_cert,
            require_client_auth=ca_cert is not None,
        )
        server.add_secure_port(f"[::]:{port}", credentials)
    else:
        server.add_insecure_port(f"[::]:{port}")
    await server.start()
    if hold is not None:
        hold(server)
    logger.info(f"Successfully start Grpc service with{'out' if not fed_utils.tls_enabled(tls_config) else ''} credentials.")
    if on_ready is not None:
        on_ready((True, None))
    await server.wait_for_termination()
