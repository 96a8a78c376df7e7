"""Unit tests of the small building blocks (no party processes)."""
import pickle
import threading
import time

import pytest

from rayfed_b200 import core
from rayfed_b200._private.message_queue import MessageQueueManager
from rayfed_b200._private.serialization import Deserializer, dumps, restricted_loads
from rayfed_b200.cleanup import CleanupManager
from rayfed_b200.config import CrossSiloMessageConfig, GrpcCrossSiloMessageConfig
from rayfed_b200.exceptions import FedRemoteError, FedTaskError
from rayfed_b200.fed_object import FedObject
from rayfed_b200.proxy import barriers


def test_message_queue_drains_in_order_and_stops():
    seen = []
    q = MessageQueueManager(lambda m: seen.append(m) or True, thread_name="t")
    q.start()
    for i in range(50):
        q.append(i)
    q.stop(wait_for_sending=True)
    assert seen == list(range(50)) and not q.is_started()


def test_message_queue_abort_skips_pending():
    gate = threading.Event()
    seen = []

    def handler(m):
        gate.wait(2)
        seen.append(m)
        return True

    q = MessageQueueManager(handler, thread_name="t")
    q.start()
    for i in range(5):
        q.append(i)
    time.sleep(0.05)
    stopper = threading.Thread(target=lambda: q.stop(wait_for_sending=False))
    stopper.start()
    time.sleep(0.05)
    gate.set()
    stopper.join(5)
    assert seen == [0]  # the in-flight message finishes, the rest is abandoned


def test_message_queue_handler_false_ends_loop():
    seen = []
    q = MessageQueueManager(lambda m: seen.append(m) or m < 2, thread_name="t")
    q.start()
    for i in range(5):
        q.append(i)
    time.sleep(0.2)
    assert seen == [0, 1, 2] and not q.is_started()


def test_cleanup_manager_retires_successes_without_queueing():
    cm = CleanupManager("alice", lambda: True)
    cm.start()
    refs = [core.Ref() for _ in range(10)]
    for r in refs:
        cm.push_to_sending(r, "bob", 1, 2)
    assert cm.pending_sends() == 10
    for r in refs:
        r.set_value(True)
    assert cm.pending_sends() == 0 and cm.get_last_sending_error() is None
    cm.stop(wait_for_sending=True)


def test_cleanup_manager_records_link_failure_and_waits_for_drain():
    cm = CleanupManager("alice", lambda: True)
    cm.start()
    bad = core.Ref()
    cm.push_to_sending(bad, "bob", 1, 2)
    slow = core.Ref()
    cm.push_to_sending(slow, "bob", 3, 4)
    bad.set_error(RuntimeError("link down"))
    threading.Timer(0.2, lambda: slow.set_value(True)).start()
    t0 = time.time()
    cm.stop(wait_for_sending=True)
    assert time.time() - t0 >= 0.15  # waited for the unresolved send
    assert isinstance(cm.get_last_sending_error(), RuntimeError)
    assert cm.pending_sends() == 0


def test_serialization_fast_path_and_cloudpickle():
    assert pickle.loads(dumps(None)) is None and pickle.loads(dumps("x")) == "x"
    f = pickle.loads(dumps(lambda x: x + 1))  # lambdas need cloudpickle
    assert f(1) == 2


def test_restricted_loads_rules():
    import numpy

    blob = dumps(numpy.arange(3))
    assert restricted_loads(blob, {"numpy.*": None, "numpy": None}).tolist() == [0, 1, 2]
    with pytest.raises(pickle.UnpicklingError):
        restricted_loads(blob, {"math": None})
    # the runtime's own error type is always admitted so failures can propagate
    d = Deserializer({"math": None})
    err = d(dumps(FedRemoteError("alice", None)))
    assert isinstance(err, FedRemoteError) and err.src_party == "alice"
    with pytest.raises(TypeError):
        restricted_loads("not-bytes", {})


def test_config_json_roundtrip_and_defaults():
    c = CrossSiloMessageConfig(timeout_in_ms=5, http_header={"a": "b"})
    c2 = CrossSiloMessageConfig.from_json(c.__json__())
    assert c2.timeout_in_ms == 5 and c2.http_header == {"a": "b"} and c2.use_global_proxy is True
    assert CrossSiloMessageConfig().timeout_in_ms == 60000 and CrossSiloMessageConfig().transport == "auto"
    g = GrpcCrossSiloMessageConfig.from_dict({"grpc_channel_options": [("grpc.enable_retries", 1)]})
    assert g.grpc_channel_options == [("grpc.enable_retries", 1)]


def test_fed_object_ids_and_send_bookkeeping():
    o = FedObject("alice", 7, None, 2)
    assert o.get_party() == "alice" and o.get_fed_task_id() == "7#2" and o.get_ray_object_ref() is None
    assert not o._was_sending_or_sent_to_party("bob")
    o._mark_is_sending_to_party("bob")
    assert o._was_sending_or_sent_to_party("bob") and not o._was_sending_or_sent_to_party("carol")
    r = core.Ref()
    o._cache_ray_object_ref(r)
    assert o.get_ref() is r


def test_proxy_actor_naming_rules():
    barriers.set_proxy_actor_name("job", True)
    assert (barriers.sender_proxy_actor_name(), barriers.receiver_proxy_actor_name()) == (
        "SenderProxyActor", "ReceiverProxyActor")
    barriers.set_proxy_actor_name("job", False)
    assert barriers.sender_proxy_actor_name() == "SenderProxyActor_job"
    barriers.set_proxy_actor_name("job", False, True)
    assert barriers.sender_proxy_actor_name() == barriers.receiver_proxy_actor_name() == "SenderReceiverProxyActor_job"
    barriers.set_proxy_actor_name("job", True)
    d = {}
    barriers.add_two_dim_dict(d, "a", "b", 1)
    assert barriers.key_exists_in_two_dim_dict(d, "a", "b") and barriers.get_from_two_dim_dict(d, "a", "b") == 1
    assert barriers.pop_from_two_dim_dict(d, "a", "b") == 1 and not barriers.key_exists_in_two_dim_dict(d, "a", "b")


def test_task_error_pickles_and_keeps_cause():
    e = FedTaskError.from_exception(ValueError("boom"), "task")
    e2 = pickle.loads(pickle.dumps(e))
    assert isinstance(e2.cause, ValueError) and "boom" in str(e2) and e2.task_name == "task"
    r = FedRemoteError("alice", ValueError("x"))
    assert pickle.loads(pickle.dumps(r)).src_party == "alice"


def test_socket_transport_unix_socket_and_retry_policy(tmp_path):
    nat = core.native()
    path = str(tmp_path / "s.sock")
    mb = nat.Mailbox()
    server = nat.SocketTransport("bob", "job", mb)
    client = nat.SocketTransport("alice", "job", nat.Mailbox())
    # the peer is not listening yet: connect retries (4 x 50 ms) cover a late start
    client.add_peer("bob", "unix:" + path, 8, 0.05, 0.05, 1.0, 5.0)
    ack = client.send("bob", "1", "2", b"late", False, lambda b, e: (_ for _ in ()).throw(RuntimeError(b)) if e else True)
    time.sleep(0.12)
    assert server.listen("unix:" + path) == ""
    assert ack.result(5) is True
    assert mb.claim("1", "2").result(5) == b"late"
    # a peer that never appears fails the send after max_attempts
    client.add_peer("carol", "unix:" + str(tmp_path / "nope.sock"), 2, 0.02, 0.02, 1.0, 5.0)
    bad = client.send("carol", "1", "2", b"x", False, None)
    assert bad.exception(5) is not None
    assert client.drain(2.0)
    client.stop()
    server.stop()


def test_top_level_send_recv_exports():
    import rayfed_b200 as fed
    from party_utils import comm_config, make_addresses

    fed.init(addresses=make_addresses(["alice"]), party="alice", logging_level="warning",
             config={"cross_silo_comm": comm_config()})
    assert core.get(fed.send("alice", {"k": [1, 2]}, "u", "d"), timeout=10)
    assert core.get(fed.recv("alice", "alice", "u", "d"), timeout=10) == {"k": [1, 2]}
    fed.shutdown()


def test_migration_shims_for_reference_style_scripts():
    """`import rayfed_b200.compat.ray as ray` + `compatible_utils.init_ray(...)`: the Ray calls a
    rayfed driver script makes keep working without Ray."""
    import fed._private.compatible_utils as compatible_utils
    import rayfed_b200.compat.ray as ray
    from rayfed_b200.core import get_runtime

    assert compatible_utils.init_ray(address="local", resources={"127.0.0.1": 2}) is None
    assert compatible_utils._compare_version_strings("2.10.0", "2.9.3") is True
    assert compatible_utils._compare_version_strings("2.0.0", "2.0.0") is False
    assert compatible_utils._compare_version_strings("1.13.0", "2.0.0") is False
    ray.init(address="local", num_cpus=2)
    assert ray.is_initialized()
    rt = get_runtime()
    r = rt.submit(lambda x: x * 2, (21,))
    assert ray.get(r) == 42
    ready, rest = ray.wait([r], num_returns=1, timeout=1)
    assert len(ready) == 1 and not rest
    slow = rt.submit(__import__("time").sleep, (0.5,))
    with pytest.raises(ray.exceptions.GetTimeoutError):
        ray.get(slow, timeout=0.01)
    ray.get(slow)
