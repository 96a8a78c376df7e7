"""Transport-level behaviour (mirrors fed/tests/test_transport_proxy.py,
test_transport_proxy_tls.py, test_enable_tls_across_parties.py,
test_grpc_options_on_proxies.py, test_retry_policy.py,
multi-jobs/test_ignore_other_job_msg.py, test_setup_proxy_actor.py)."""
import json
import pickle

import pytest

import rayfed_b200 as fed
from party_utils import comm_config, free_ports, make_addresses, run_parties
from rayfed_b200._private import constants, global_context, kv
from rayfed_b200.core import get, get_runtime
from rayfed_b200.proxy import barriers
from rayfed_b200.proxy.barriers import _start_receiver_proxy, _start_sender_proxy, send


def _bare_context(job, party, addresses, comm=None, tls=None):
    """What fed.init does before starting proxies — lets us drive barriers directly."""
    global_context.init_global_context(party, job, False, False)
    global_context.get_global_context().get_cleanup_manager().start()
    kv._init_internal_kv(job)
    kv.kv.put(constants.KEY_OF_CLUSTER_CONFIG, pickle.dumps({
        constants.KEY_OF_CLUSTER_ADDRESSES: addresses, constants.KEY_OF_CURRENT_PARTY_NAME: party,
        constants.KEY_OF_TLS_CONFIG: tls or {}}))
    kv.kv.put(constants.KEY_OF_JOB_CONFIG, pickle.dumps({constants.KEY_OF_CROSS_SILO_COMM_CONFIG_DICT: comm or {}}))


def _teardown():
    barriers._stop_proxies()
    global_context.clear_global_context()
    kv._clear_internal_kv()
    from rayfed_b200.core import shutdown_runtime

    shutdown_runtime()


def _proxies(transport):
    if transport == "grpc":
        from rayfed_b200.proxy.grpc.grpc_proxy import GrpcReceiverProxy, GrpcSenderProxy

        return GrpcSenderProxy, GrpcReceiverProxy
    from rayfed_b200.proxy.native.socket_proxy import NativeReceiverProxy, NativeSenderProxy

    return NativeSenderProxy, NativeReceiverProxy


@pytest.mark.parametrize("transport", ["native", "grpc"])
def test_n_to_1_transport_loopback(transport):
    """N interleaved sends / get_data on keys (i, i+1), party sending to itself."""
    party = "test_party"
    addresses = make_addresses([party])
    comm = comm_config(transport)
    _bare_context("test_n_to_1", party, addresses, comm)
    sender_cls, receiver_cls = _proxies(transport)
    try:
        _start_receiver_proxy(addresses, party, "info", proxy_cls=receiver_cls, proxy_config=comm)
        _start_sender_proxy(addresses, party, "info", proxy_cls=sender_cls, proxy_config=comm)
        receiver = get_runtime().get_named(barriers.receiver_proxy_actor_name())
        sent, got = [], []
        for i in range(10):
            sent.append(send(party, f"data-{i}", i, i + 1))
            got.append(receiver.get_data(party, i, i + 1))
        assert all(get(sent, timeout=20))
        assert get(got, timeout=20) == [f"data-{i}" for i in range(10)]
        assert receiver._get_stats()["receive_op_count"] == 10
    finally:
        _teardown()


def test_grpc_metadata_and_job_name_reach_the_server():
    import asyncio
    import threading

    import grpc

    from rayfed_b200.proxy.grpc import fed_pb
    from rayfed_b200.proxy.grpc.grpc_proxy import GrpcSenderProxy

    seen = {}

    class Servicer(fed_pb.GrpcServiceServicer):
        async def SendData(self, request, context):
            seen["job"] = request.job_name
            seen["meta"] = dict(context.invocation_metadata())
            seen["key"] = (request.upstream_seq_id, request.downstream_seq_id)
            return fed_pb.SendDataResponse(code=200, result="OK")

    port = free_ports(1)[0]
    loop = asyncio.new_event_loop()
    started = threading.Event()
    holder = {}

    async def serve():
        server = grpc.aio.server()
        fed_pb.add_GrpcServiceServicer_to_server(Servicer(), server)
        server.add_insecure_port(f"127.0.0.1:{port}")
        await server.start()
        holder["server"] = server
        started.set()
        await server.wait_for_termination()

    t = threading.Thread(target=lambda: loop.run_until_complete(serve()), daemon=True)
    t.start()
    assert started.wait(10)
    party = "test_party"
    addresses = {party: f"127.0.0.1:{port}"}
    comm = {"http_header": {"key": "value", "token": "abc"}}
    _bare_context("test_meta_job", party, addresses, comm)
    try:
        _start_sender_proxy(addresses, party, "info", proxy_cls=GrpcSenderProxy, proxy_config=comm)
        assert get(send(party, "data", 0, 1), timeout=20)
        assert seen["job"] == "test_meta_job"
        assert seen["meta"]["key"] == "value" and seen["meta"]["token"] == "abc"
        assert seen["key"] == ("0", "1")
    finally:
        _teardown()
        try:  # best-effort stop of the fake server (a daemon thread owns its loop)
            asyncio.run_coroutine_threadsafe(holder["server"].stop(None), loop).result(2)
        except Exception:  # noqa: BLE001
            pass


@pytest.mark.parametrize("transport", ["native", "grpc"])
def test_other_job_messages_are_refused_with_417(transport):
    party = "test_party"
    addresses = make_addresses([party])
    comm = comm_config(transport)
    _bare_context("job_a", party, addresses, comm)
    sender_cls, receiver_cls = _proxies(transport)
    try:
        _start_receiver_proxy(addresses, party, "info", proxy_cls=receiver_cls, proxy_config=comm)
        if transport == "grpc":
            import grpc

            from rayfed_b200.proxy.grpc import fed_pb

            with grpc.insecure_channel(addresses[party]) as ch:
                stub = fed_pb.GrpcServiceStub(ch)
                resp = stub.SendData(fed_pb.SendDataRequest(data=b"x", upstream_seq_id="1",
                                                            downstream_seq_id="2", job_name="job_b"))
            assert resp.code == 417 and "JobName mis-match" in resp.result
        else:
            foreign = sender_cls(addresses, party, "job_b", {}, comm)
            ack = foreign.send(party, "x", 1, 2)
            err = ack.exception(10)
            assert err is not None and "JobName mis-match" in str(err)
            foreign.stop()
    finally:
        _teardown()


def test_grpc_option_precedence():
    from rayfed_b200.proxy.grpc.grpc_proxy import GrpcSenderProxy

    retry = {"maxAttempts": 4, "initialBackoff": "5s", "maxBackoff": "5s", "backoffMultiplier": 1,
             "retryableStatusCodes": ["UNAVAILABLE"]}
    p = GrpcSenderProxy({"a": "127.0.0.1:1"}, "a", "job", {}, {
        "messages_max_size_in_bytes": 100,
        "grpc_channel_options": [("grpc.max_send_message_length", 200), ("grpc.default_authority", "x")],
        "grpc_retry_policy": retry,
    })
    _, opts = p.get_grpc_config_by_party("a")
    opts = dict(opts)
    assert opts["grpc.max_send_message_length"] == 200  # channel options beat messages_max_size_in_bytes
    assert opts["grpc.max_receive_message_length"] == 100
    assert opts["grpc.default_authority"] == "x"
    assert opts["grpc.so_reuseport"] == 0 and opts["grpc.enable_retries"] == 1
    policy = json.loads(opts["grpc.service_config"])["methodConfig"][0]["retryPolicy"]
    assert policy == retry
    default = dict(GrpcSenderProxy({"a": "127.0.0.1:1"}, "a", "job", {}, {}).get_grpc_config_by_party("a")[1])
    assert default["grpc.max_send_message_length"] == 500 * 1024 * 1024
    assert json.loads(default["grpc.service_config"])["methodConfig"][0]["retryPolicy"]["maxAttempts"] == 5


def test_proxy_config_visible_through_host():
    addresses = make_addresses(["alice"])
    fed.init(addresses=addresses, party="alice", logging_level="warning",
             config={"cross_silo_comm": {"transport": "grpc", "timeout_in_ms": 4321, "http_header": {"k": "v"},
                                         "send_resource_label": {"x": 1}}})
    host = get_runtime().get_named(barriers.sender_proxy_actor_name())
    cfg = host._get_proxy_config()
    assert cfg["timeout_in_ms"] == 4321 and cfg["http_header"] == {"k": "v"}
    assert dict(cfg["grpc_options"])["grpc.enable_retries"] == 1
    fed.shutdown()


def _tls_party(party, addresses, tls):
    import rayfed_b200 as fed

    fed.init(addresses=addresses, party=party, tls_config=tls, logging_level="warning",
             config={"cross_silo_comm": comm_config("auto")})
    from rayfed_b200.proxy.grpc.grpc_proxy import GrpcSenderProxy

    assert isinstance(get_runtime().get_named(barriers.sender_proxy_actor_name()).proxy, GrpcSenderProxy)

    @fed.remote
    def f(x):
        return x + 1

    assert fed.get(f.party("bob").remote(f.party("alice").remote(1))) == 3
    fed.shutdown()


def test_mutual_tls_across_parties(tmp_path):
    from rayfed_b200.tool.generate_tls_certs import generate_self_signed_tls_certs

    tls = generate_self_signed_tls_certs(str(tmp_path))
    addresses = {p: a.replace("127.0.0.1", "localhost") for p, a in make_addresses(["alice", "bob"]).items()}
    run_parties(_tls_party, ["alice", "bob"], args=(addresses, tls))


def test_custom_proxy_classes_are_pluggable():
    """A user transport written against the ABCs (in-memory loopback)."""
    from rayfed_b200.proxy.base_proxy import ReceiverProxy, SenderProxy

    box = {}

    class MemSender(SenderProxy):
        def send(self, dest_party, data, upstream_seq_id, downstream_seq_id):
            box[(str(upstream_seq_id), str(downstream_seq_id))] = data
            return True

    class MemReceiver(ReceiverProxy):
        def start(self):
            pass

        def is_ready(self):
            return True, None

        def get_data(self, src_party, upstream_seq_id, curr_seq_id):
            return box[(str(upstream_seq_id), str(curr_seq_id))]

    addresses = make_addresses(["alice"])
    fed.init(addresses=addresses, party="alice", sender_proxy_cls=MemSender, receiver_proxy_cls=MemReceiver,
             logging_level="warning")
    assert get(send("alice", {"x": 1}, "u", "d"), timeout=5) is True
    assert get(fed.recv("alice", "alice", "u", "d"), timeout=5) == {"x": 1}
    fed.shutdown()


def test_combined_sender_receiver_proxy():
    from rayfed_b200.proxy.native.socket_proxy import NativeSenderReceiverProxy

    addresses = make_addresses(["alice"])
    fed.init(addresses=addresses, party="alice", receiver_sender_proxy_cls=NativeSenderReceiverProxy,
             logging_level="warning", config={"cross_silo_comm": comm_config()})
    assert barriers.sender_proxy_actor_name() == barriers.receiver_proxy_actor_name() == "SenderReceiverProxyActor"
    assert get(send("alice", 123, 5, 6), timeout=10)
    assert get(fed.recv("alice", "alice", 5, 6), timeout=10) == 123
    stats = get_runtime().get_named("SenderReceiverProxyActor")._get_stats()
    assert stats["send_op_count"] == 1 and stats["receive_op_count"] == 1
    fed.shutdown()
