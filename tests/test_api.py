"""Single-party API behaviour (mirrors fed/tests/test_api.py, test_repeat_init.py,
test_listening_address.py, test_internal_kv.py, multi-jobs/test_multi_proxy_actor.py)."""
import pickle

import pytest

import rayfed_b200 as fed
from party_utils import comm_config, make_addresses
from rayfed_b200 import config as fed_config
from rayfed_b200._private import constants, kv
from rayfed_b200._private.global_context import get_global_context
from rayfed_b200.core import get_runtime
from rayfed_b200.proxy import barriers


def _init(party="alice", **kw):
    addresses = make_addresses(["alice"])
    fed.init(addresses=addresses, party=party, logging_level="warning",
             config={"cross_silo_comm": comm_config(**kw.pop("comm", {}))}, **kw)
    return addresses


def test_init_stores_cluster_and_job_config():
    addresses = _init(comm={"timeout_in_ms": 1234})
    cluster = fed_config.get_cluster_config()
    assert cluster.cluster_addresses == addresses
    assert cluster.current_party == "alice"
    assert cluster.tls_config == {}
    assert fed_config.get_job_config().cross_silo_comm_config_dict["timeout_in_ms"] == 1234
    fed.shutdown()


def test_remote_without_party_raises():
    _init()

    @fed.remote
    def f():
        return 1

    with pytest.raises(ValueError):
        f.remote()

    @fed.remote
    class A:
        def m(self):
            return 1

    with pytest.raises(ValueError):
        A.remote()
    fed.shutdown()


def test_remote_rejects_non_callables():
    with pytest.raises(TypeError):
        fed.remote(num_returns=2)(42)


def test_options_replace_not_merge_and_num_returns():
    _init()

    @fed.remote(num_returns=2)
    def two():
        return 1, 2

    a, b = two.party("alice").remote()
    assert fed.get([a, b]) == [1, 2]
    assert a.get_fed_task_id().endswith("#0") and b.get_fed_task_id().endswith("#1")

    @fed.remote
    def one():
        return 1, 2

    x = one.party("alice").options(num_returns=2).options().remote()  # options() replaced -> 1 return
    assert fed.get(x) == (1, 2)
    fed.shutdown()


def test_unknown_actor_method_raises_attribute_error():
    _init()

    @fed.remote
    class A:
        def m(self):
            return 1

    a = A.party("alice").remote()
    with pytest.raises(AttributeError):
        a.nope
    assert fed.get(a.m.remote()) == 1
    fed.shutdown()


def test_seq_ids_task_actor_get():
    _init()
    ctx = get_global_context()

    @fed.remote
    def f():
        return 1

    @fed.remote
    class A:
        def m(self):
            return 2

    o = f.party("alice").remote()
    assert ctx.current_seq_id() == 1
    a = A.party("alice").remote()  # two ids
    assert ctx.current_seq_id() == 3
    m = a.m.remote()
    assert ctx.current_seq_id() == 4
    assert fed.get(o) == 1  # +1
    assert ctx.current_seq_id() == 5
    assert fed.get(m.get_ref()) == 2  # raw future: +0
    assert ctx.current_seq_id() == 5
    fed.shutdown()


def test_shutdown_is_idempotent_and_safe_without_init():
    fed.shutdown()
    _init()
    fed.shutdown()
    fed.shutdown()
    assert get_global_context() is None


def test_repeat_init_and_seq_ids_restart():
    ids = []
    for _ in range(2):
        _init()

        @fed.remote
        def f():
            return 7

        o = f.party("alice").remote()
        ids.append(o.get_fed_task_id())
        assert fed.get(o) == 7
        fed.shutdown()
    assert ids[0] == ids[1] == "1#0"


def test_internal_kv_namespacing_and_reset():
    _init(job_name="jobx")
    assert kv.kv.put("key", b"val") is True
    assert kv.kv.get("key") == b"val"
    assert kv.raw_get(b"RAYFED#jobx#key") == b"val"
    raw = kv.raw_get("RAYFED#jobx#" + constants.KEY_OF_CLUSTER_CONFIG)
    assert pickle.loads(raw)[constants.KEY_OF_CURRENT_PARTY_NAME] == "alice"
    fed.shutdown()
    assert kv.kv is None
    assert kv.raw_get("RAYFED#jobx#" + constants.KEY_OF_CLUSTER_CONFIG) is None


def test_per_job_proxy_names_when_not_global():
    _init(job_name="j1", comm={"use_global_proxy": False})
    assert barriers.sender_proxy_actor_name() == "SenderProxyActor_j1"
    assert barriers.receiver_proxy_actor_name() == "ReceiverProxyActor_j1"
    rt = get_runtime()
    assert rt.has_named("SenderProxyActor_j1") and rt.has_named("ReceiverProxyActor_j1")
    fed.shutdown()
    _init(job_name="j2")
    assert barriers.sender_proxy_actor_name() == "SenderProxyActor"
    fed.shutdown()


@pytest.mark.parametrize("transport", ["native", "grpc"])
def test_listening_port_in_use_is_reported(transport):
    import socket

    s = socket.socket()
    s.bind(("0.0.0.0", 0))
    s.listen(1)
    port = s.getsockname()[1]
    try:
        with pytest.raises(AssertionError) as e:
            fed.init(addresses={"alice": f"127.0.0.1:{port}"}, party="alice", logging_level="warning",
                     config={"cross_silo_comm": comm_config(transport)})
        assert str(port) in str(e.value)
    finally:
        s.close()
        fed.shutdown()


def test_kill_only_acts_in_owner_party_and_fails_later_calls():
    from rayfed_b200.exceptions import FedActorDiedError

    _init()

    @fed.remote
    class A:
        def m(self):
            return 1

    a = A.party("alice").remote()
    assert fed.get(a.m.remote()) == 1
    fed.kill(a)
    with pytest.raises(FedActorDiedError):
        fed.get(a.m.remote())
    fed.shutdown()


def test_init_validates_arguments():
    with pytest.raises(AssertionError):
        fed.init(addresses=None, party="alice")
    with pytest.raises(AssertionError):
        fed.init(addresses={"bob": "127.0.0.1:1"}, party="alice")
    with pytest.raises(ValueError):
        fed.init(addresses={"alice": "not-an-address"}, party="alice")
    with pytest.raises(AssertionError):
        fed.init(addresses=make_addresses(["alice"]), party="alice", tls_config={"cert": "x"})
    fed.shutdown()


def test_config_from_dict_drops_unknown_keys():
    from rayfed_b200.config import CrossSiloMessageConfig, GrpcCrossSiloMessageConfig

    c = CrossSiloMessageConfig.from_dict({"timeout_ms": 5, "timeout_in_ms": 7, "bogus": 1})
    assert c.timeout_in_ms == 7 and not hasattr(c, "bogus")
    g = GrpcCrossSiloMessageConfig.from_dict({"grpc_retry_policy": {"maxAttempts": 2}, "http_header": {"a": "b"}})
    assert g.grpc_retry_policy == {"maxAttempts": 2} and g.http_header == {"a": "b"}


def test_init_picks_the_spin_policy_from_cpus_and_parties_per_host(monkeypatch):
    """Adaptive spinning needs CPUs to burn: `fed.init` gives every party on this host an equal
    share of half the usable CPUs and leaves long spins off below two spinners per party
    (csrc/core/ref.h; measured on the B200 host: profiles/README.md R2.2b)."""
    from rayfed_b200.api import _local_party_count
    from rayfed_b200.core import native

    assert _local_party_count({"a": "127.0.0.1:1", "b": "localhost:2", "c": "10.0.0.5:3"}, "a") == 2
    assert _local_party_count({"a": "10.0.0.4:1", "b": "10.0.0.4:2", "c": "10.0.0.5:3"}, "c") == 1
    assert _local_party_count({"a": "[::1]:1", "b": "127.0.0.1:2"}, "b") == 2

    import os

    if "RAYFED_B200_ADAPTIVE_SPIN_US" in os.environ or "RAYFED_B200_MAX_SPINNERS" in os.environ:
        pytest.skip("policy overridden by the environment (read once per process by the native module)")
    nat = native()
    before = nat.spin_info()
    assert before["adaptive_cap_us"] == 0            # nothing configured outside a job
    _init()
    during = nat.spin_info()
    share = during["usable_cpus"] // 2               # one party on this host
    if share >= 2:
        assert during["adaptive_cap_us"] == 500 and during["max_long_spinners"] == min(8, share)
    else:
        assert during["adaptive_cap_us"] == 0
    fed.shutdown()
    assert nat.spin_info()["adaptive_cap_us"] == 0
    # many parties on a small host: off
    nat.set_spin_policy_auto(during["usable_cpus"])
    assert nat.spin_info()["adaptive_cap_us"] == 0
    nat.set_spin_policy(-1, -1)
