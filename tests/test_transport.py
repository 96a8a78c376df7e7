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


def test_native_receiver_bounds_attacker_controlled_lengths():
    """a 40-byte header announcing a huge payload must neither abort the party
    (std::length_error escaping a std::thread) nor make it allocate; the connection is dropped
    and the receiver keeps serving."""
    import socket
    import struct

    from rayfed_b200.core import native

    MAGIC, FRAME_DATA = 0x42444546, 1
    nat = native()
    port = free_ports(1)[0]
    mb = nat.Mailbox()
    srv = nat.SocketTransport("bob", "job", mb)
    assert srv.listen(f"127.0.0.1:{port}") == ""
    srv.set_max_message_bytes(1 << 20)
    assert srv.max_message_bytes() == 1 << 20
    try:
        # FrameHeader {u32 magic, u16 type, u16 flags, u32 code, u32 job_len, u32 up_len, u32 down_len,
        #              u64 msg_id, u64 data_len}: the magic is what a real client sends
        cli = nat.SocketTransport("alice", "job", nat.Mailbox())
        cli.add_peer("bob", f"127.0.0.1:{port}", 3, 0.05, 0.2, 1.5, 5.0)
        ok = cli.send("bob", "1", "2", b"x" * 100, False, None)
        assert ok.result(5) is not None
        big = cli.send("bob", "3", "4", b"y" * (2 << 20), False, None)  # over the receiver's cap
        with pytest.raises(Exception):
            big.result(5)
        # a raw, well-formed header (magic, type=data) with absurd lengths: the exact reproducer of the finding
        s = socket.create_connection(("127.0.0.1", port))
        s.sendall(struct.pack("<IBBHQIIIiQ", MAGIC, FRAME_DATA, 0, 0, 1, 3, 1, 1, 0, 0xFFFFFFFFFFFFFFF0))
        s.settimeout(2)
        try:
            s.recv(4096)  # a 413 ack, then the server closes
        except OSError:
            pass
        s.close()
        # the server is still alive and serving
        again = cli.send("bob", "5", "6", b"z", False, None)
        assert again.result(5) is not None
        assert mb.claim("5", "6").result(2) == b"z"
        assert srv.stats()["rejected"] >= 1
        cli.stop()
    finally:
        srv.stop()


class _Evil:
    def __reduce__(self):
        import os

        return (os.system, ("echo pwned > /tmp/rayfed_b200_pwned",))


def test_nvlink_receiver_applies_the_allow_list_to_the_skeleton():
    """the skeleton inside a ('fedcuda', k, skeleton) frame must go through the same
    restricted unpickler as any other payload (fed/_private/serialization_utils.py:24-63)."""
    import os

    from rayfed_b200.exceptions import FedTaskError
    from rayfed_b200.proxy.nvlink.nvlink_proxy import NvlinkReceiverProxy

    marker = "/tmp/rayfed_b200_pwned"
    if os.path.exists(marker):
        os.unlink(marker)
    port = free_ports(1)[0]
    cfg = comm_config("native", serializing_allowed_list={"numpy": None})
    rx = NvlinkReceiverProxy(f"127.0.0.1:{port}", "bob", "job", {}, cfg)
    rx.start()
    try:
        evil = pickle.dumps(("fedcuda", 0, pickle.dumps(_Evil())))
        rx.mailbox.deliver_raw("1", "2", evil, False)
        ref = rx.get_data("alice", "1", "2")
        with pytest.raises((pickle.UnpicklingError, FedTaskError, Exception)) as ei:
            ref.result(5)
        assert "not allowed" in str(ei.value) or "Unpickling" in type(ei.value).__name__
        assert not os.path.exists(marker), "the allow-list was bypassed: code ran on the receiver"
        # an allowed payload in the same envelope still decodes
        ok = pickle.dumps(("fedcuda", 0, pickle.dumps({"a": 1})))
        rx.mailbox.deliver_raw("3", "4", ok, False)
        assert rx.get_data("alice", "3", "4").result(5) == {"a": 1}
        # the batch entry point (the remote arguments of one task) decodes the same frames, one
        # decoder per source party
        rx.mailbox.deliver_raw("5#0", "9", pickle.dumps(("fedcuda", 0, pickle.dumps([1, 2]))), False)
        refs = rx.get_data_many([("alice", "5#0"), ("carol", "6#0")], "9")
        rx.mailbox.deliver_raw("6#0", "9", pickle.dumps("plain"), False)
        assert [r.result(5) for r in refs] == [[1, 2], "plain"]
        assert rx._decoder_for("alice") is rx._decoder_for("alice")
        assert rx._decoder_for("alice") is not rx._decoder_for("carol")
    finally:
        rx.stop()


@pytest.mark.parametrize("max_resends", [0, 2])
def test_proxy_max_restarts_resends_unacknowledged_messages(max_resends, monkeypatch):
    """`proxy_max_restarts` (fed/proxy/barriers.py:301-307: the sender proxy is restarted and its
    in-flight sends retried N times).  Here: a message written on a connection that dies before
    the acknowledgement is re-sent on the re-established connection — or fails the send when the
    option is off."""
    import socket
    import threading

    from rayfed_b200.core import native

    # plain socket data path: with the shared-memory offer the first frame on the wire would be the
    # handshake, not the message this test wants to lose
    monkeypatch.setenv("RAYFED_B200_SHM", "0")
    nat = native()
    port = free_ports(1)[0]
    swallowed = threading.Event()
    lst = socket.socket()
    lst.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    lst.bind(("127.0.0.1", port))
    lst.listen(1)

    def black_hole():
        c, _ = lst.accept()
        c.recv(65536)  # take the frame, never acknowledge, die
        c.close()
        lst.close()
        swallowed.set()

    threading.Thread(target=black_hole, daemon=True).start()
    cli = nat.SocketTransport("alice", "job", nat.Mailbox())
    cli.add_peer("bob", f"127.0.0.1:{port}", 40, 0.05, 0.2, 1.5, 10.0, max_resends)
    ack = cli.send("bob", "1", "2", b"payload", False, None)
    assert swallowed.wait(5)
    mb = nat.Mailbox()
    srv = nat.SocketTransport("bob", "job", mb)  # the "restarted" receiver on the same address
    assert srv.listen(f"127.0.0.1:{port}") == ""
    try:
        if max_resends:
            assert ack.result(10) is not None
            assert mb.claim("1", "2").result(5) == b"payload"
        else:
            with pytest.raises(Exception):
                ack.result(10)
    finally:
        cli.stop()
        srv.stop()


def test_a_peer_that_corrupts_its_ring_counters_is_dropped_not_followed(tmp_path):
    """The producer side of a shared-memory ring is another process: counters it scribbles over must
    end the connection, never turn into a copy past the ring (csrc/core/shm_conn.h).  A hand-made
    client performs the hello handshake with its own segment, then publishes an impossible `head`."""
    import mmap
    import os
    import socket
    import struct
    import time

    from rayfed_b200.core import native

    if not os.path.isdir("/dev/shm"):
        pytest.skip("no /dev/shm")
    nat = native()
    port = free_ports(1)[0]
    mb = nat.Mailbox()
    rx = nat.SocketTransport("bob", "job", mb)
    assert not rx.listen(f"127.0.0.1:{port}")
    cap = 64 * 1024
    name = f"/rfb.test.{os.getpid()}"
    path = "/dev/shm" + name
    hdr = struct.Struct("<IBBHQIIIiQ")                      # FrameHeader, packed
    try:
        with open(path, "wb") as f:
            f.truncate(4096 + 2 * cap)
        fd = os.open(path, os.O_RDWR)
        seg = mmap.mmap(fd, 4096 + 2 * cap)
        os.close(fd)
        struct.pack_into("<II", seg, 0, 0x52424653, cap)     # magic "SFBR", capacity; counters start at 0
        s = socket.create_connection(("127.0.0.1", port))
        s.sendall(hdr.pack(0x42444546, 3, 0, 0, 0, 0, 0, 0, 0, len(name)) + name.encode())
        ack = s.recv(hdr.size, socket.MSG_WAITALL)
        assert hdr.unpack(ack)[1] == 2 and hdr.unpack(ack)[8] == 200     # the server switched to the rings
        # a well-formed frame head that announces a body bigger than the ring ...
        body = 200 * 1024
        frame = hdr.pack(0x42444546, 1, 0, 0, 7, 3, 1, 1, 0, body) + b"job" + b"1" + b"2"
        seg[4096:4096 + len(frame)] = frame
        struct.pack_into("<Q", seg, 64, len(frame))          # c2s.head: publish it
        deadline = time.time() + 3
        while struct.unpack_from("<Q", seg, 128)[0] < len(frame) and time.time() < deadline:
            time.sleep(0.005)                                 # c2s.tail: the server took the head and waits for the body
        assert struct.unpack_from("<Q", seg, 128)[0] == len(frame)
        # ... then a counter that claims a terabyte is available: following it would copy 200 KiB
        # out of a 64 KiB ring, past the end of the mapping
        struct.pack_into("<Q", seg, 64, 1 << 40)
        s.settimeout(5.0)
        t0 = time.perf_counter()
        assert s.recv(1) == b""                               # dropped (EOF), within a reader slice or two
        assert time.perf_counter() - t0 < 3.0
        assert rx.stats()["receive_op_count"] == 0            # ... and nothing was "received" from beyond the ring
        s.close()
        seg.close()
        # the party is alive and serves well-behaved peers
        tx = nat.SocketTransport("alice", "job", nat.Mailbox())
        tx.add_peer("bob", f"127.0.0.1:{port}", 5, 0.05, 0.2, 1.5, 10.0)
        assert tx.send("bob", "1", "2", b"fine").result(10) is True
        assert mb.claim("1", "2").result(5) == b"fine"
        tx.stop()
    finally:
        if os.path.exists(path):
            os.unlink(path)
        rx.stop()
