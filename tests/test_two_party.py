"""2-party integration behaviour, one OS process per party (mirrors
fed/tests/test_basic_pass_fed_objects.py, test_fed_get.py, test_cache_fed_objects.py,
test_options.py, test_pass_fed_objects_in_containers_*.py, test_ping_others.py,
test_async_startup_2_clusters.py) — parametrised over both host transports."""
import time

import pytest

from party_utils import comm_config, make_addresses, run_parties

TRANSPORTS = ["native", "grpc"]
PARTIES = ["alice", "bob"]


def _init(party, addresses, transport, **cfg):
    import rayfed_b200 as fed

    fed.init(addresses=addresses, party=party, logging_level="warning",
             config={"cross_silo_comm": comm_config(transport), **cfg})
    return fed


# -- README demo --------------------------------------------------------------------
def _readme_demo(party, addresses, transport):
    fed = _init(party, addresses, transport)

    @fed.remote
    class MyActor:
        def __init__(self, value):
            self.value = value

        def inc(self, num):
            self.value += num
            return self.value

    @fed.remote
    def aggregate(a, b):
        return a + b

    alice = MyActor.party("alice").remote(1)
    bob = MyActor.party("bob").remote(2)
    total = aggregate.party("bob").remote(alice.inc.remote(1), bob.inc.remote(1))
    assert fed.get(total) == 5
    fed.shutdown()


@pytest.mark.parametrize("transport", TRANSPORTS)
def test_readme_demo(transport):
    run_parties(_readme_demo, PARTIES, args=(make_addresses(PARTIES), transport))


# -- FedObject as actor constructor argument ---------------------------------------------
def _actor_ctor_arg(party, addresses, transport, owner):
    fed = _init(party, addresses, transport)

    @fed.remote
    def f():
        return 100

    @fed.remote
    class My:
        def __init__(self, value):
            self._value = value

        def get_value(self):
            return self._value

    o = f.party("alice").remote()
    my = My.party(owner).remote(o)
    assert fed.get(my.get_value.remote()) == 100
    assert fed.get(o) == 100
    fed.shutdown()


@pytest.mark.parametrize("owner", ["alice", "bob"])
def test_fed_object_as_actor_ctor_arg(owner):
    run_parties(_actor_ctor_arg, PARTIES, args=(make_addresses(PARTIES), "native", owner))


# -- FedAvg-like loop with fed.get in the middle ---------------------------------------------
def _fed_get_loop(party, addresses, transport):
    if party == "alice":
        time.sleep(0.5)  # asynchronous start-up is covered by connect retries
    fed = _init(party, addresses, transport)

    @fed.remote
    class MyModel:
        def __init__(self, party, step_length):
            self._step_length = step_length
            self._weights = 0

        def train(self):
            self._weights += self._step_length
            return self._weights

        def get_weights(self):
            return self._weights

        def set_weights(self, new_weights):
            self._weights = new_weights
            return new_weights

    @fed.remote
    def mean(x, y):
        return (x + y) / 2

    alice_model = MyModel.party("alice").remote("alice", 2)
    bob_model = MyModel.party("bob").remote("bob", 4)
    means = []
    for _ in range(3):
        w1, w2 = alice_model.train.remote(), bob_model.train.remote()
        new_weights = mean.party("alice").remote(w1, w2)
        means.append(fed.get(new_weights))
        alice_model.set_weights.remote(new_weights)
        bob_model.set_weights.remote(new_weights)
    assert means == [3, 6, 9]
    assert fed.get([alice_model.get_weights.remote(), bob_model.get_weights.remote()]) == [9, 9]
    fed.shutdown()


@pytest.mark.parametrize("transport", TRANSPORTS)
def test_fed_get_in_loop(transport):
    run_parties(_fed_get_loop, PARTIES, args=(make_addresses(PARTIES), transport))


# -- each (object, destination) is pushed once; each remote object received once --------------
def _cache(party, addresses, transport):
    fed = _init(party, addresses, transport)
    from rayfed_b200.core import get_runtime
    from rayfed_b200.proxy.barriers import receiver_proxy_actor_name, sender_proxy_actor_name

    @fed.remote
    def f():
        return "hello"

    @fed.remote
    def g(x, index):
        return x + str(index)

    o = f.party("alice").remote()
    o1 = g.party("bob").remote(o, 1)
    o2 = g.party("bob").remote(o, 2)
    assert fed.get([o, o1, o2]) == ["hello", "hello1", "hello2"]
    rt = get_runtime()
    if party == "bob":
        assert rt.get_named(receiver_proxy_actor_name())._get_stats()["receive_op_count"] == 1
    if party == "alice":
        assert rt.get_named(sender_proxy_actor_name())._get_stats()["send_op_count"] == 1
    fed.shutdown()


@pytest.mark.parametrize("transport", TRANSPORTS)
def test_cache_fed_objects(transport):
    run_parties(_cache, PARTIES, args=(make_addresses(PARTIES), transport))


# -- num_returns for functions and actor methods across parties ---------------------------------
def _options(party, addresses, transport):
    fed = _init(party, addresses, transport)

    @fed.remote
    class Foo:
        def bar(self):
            return "a", "b"

    @fed.remote
    def bar(x):
        return x, x + "2"

    foo = Foo.party("alice").remote()
    a, b = foo.bar.options(num_returns=2).remote()
    c, d = bar.party("bob").options(num_returns=2).remote(a)
    assert fed.get([a, b, c, d]) == ["a", "b", "a", "a2"]
    fed.shutdown()


def test_num_returns_across_parties():
    run_parties(_options, PARTIES, args=(make_addresses(PARTIES), "native"))


# -- FedObjects nested in containers: tasks receive futures and read them with fed.get ------------
def _containers(party, addresses, transport, use_actor):
    fed = _init(party, addresses, transport)

    @fed.remote
    def foo(i):
        return f"foo-{i}"

    @fed.remote
    def bar(li):
        import rayfed_b200 as fed_in_task

        assert fed_in_task.get(li[1][0]) == "foo-2"
        return [fed_in_task.get(li[0]), fed_in_task.get(li[1][0]), li[2]]

    @fed.remote
    class Bar:
        def run(self, li):
            import rayfed_b200 as fed_in_task

            return [fed_in_task.get(li[0]), fed_in_task.get(li[1][0]), li[2]]

    o1, o2 = foo.party("alice").remote(1), foo.party("bob").remote(2)
    li = [o1, [o2], "plain"]
    if use_actor:
        out = Bar.party("bob").remote().run.remote(li)
    else:
        out = bar.party("bob").remote(li)
    assert fed.get(out) == ["foo-1", "foo-2", "plain"]
    fed.shutdown()


@pytest.mark.parametrize("use_actor", [False, True])
def test_fed_objects_in_containers(use_actor):
    run_parties(_containers, PARTIES, args=(make_addresses(PARTIES), "native", use_actor))


# -- init barrier ----------------------------------------------------------------------------------
def _ping(party, addresses, transport):
    if party == "bob":
        time.sleep(1.0)
    fed = _init(party, addresses, transport, barrier_on_initializing=True)

    @fed.remote
    def f():
        return "up"

    # a collective keeps both parties alive until each has passed its barrier
    assert fed.get([f.party("alice").remote(), f.party("bob").remote()]) == ["up", "up"]
    fed.shutdown()


@pytest.mark.parametrize("transport", TRANSPORTS)
def test_barrier_on_initializing_waits_for_late_party(transport):
    run_parties(_ping, PARTIES, args=(make_addresses(PARTIES), transport))


def _ping_fails(party, addresses):
    import rayfed_b200 as fed
    from rayfed_b200.proxy.barriers import ping_others

    fed.init(addresses=addresses, party=party, logging_level="error",
             config={"cross_silo_comm": comm_config("native", retry_max_attempts=1, timeout_in_ms=500)})
    try:
        ping_others(addresses, party, max_retries=2)
    except RuntimeError as e:
        assert "Failed to wait for parties" in str(e)
    else:
        raise AssertionError("ping_others should have failed")
    fed.shutdown()


def test_ping_others_gives_up():
    run_parties(_ping_fails, ["alice"], args=(make_addresses(PARTIES),))


# -- three parties: fed.get broadcasts to everyone ----------------------------------------------------
def _three(party, addresses):
    fed = _init(party, addresses, "native")

    @fed.remote
    def f(x):
        return x * 2

    @fed.remote
    def add(*xs):
        return sum(xs)

    parts = [f.party(p).remote(i + 1) for i, p in enumerate(addresses)]
    total = add.party("carol").remote(*parts)
    assert fed.get(total) == 12
    assert fed.get(parts) == [2, 4, 6]
    fed.shutdown()


def test_three_parties():
    parties = ["alice", "bob", "carol"]
    run_parties(_three, parties, args=(make_addresses(parties),))


# -- N parties (parametrise N in {2, 4, 8}): the reference's FedAvg-like flow
#    (fed/tests/test_fed_get.py:50-85) generalised — per-party actors train, one party averages,
#    every party installs the mean; checked in EVERY controller ---------------------------------
def _n_party_fedavg(party, addresses, transport):
    fed = _init(party, addresses, transport)
    parties = list(addresses)
    n = len(parties)

    @fed.remote
    class Trainer:
        def __init__(self, rank):
            self.w = float(rank)
            self.seen = []

        def train(self):
            self.w += 1.0
            return self.w

        def set_weights(self, w):
            self.w = w
            self.seen.append(w)
            return w

        def history(self):
            return self.seen

    @fed.remote
    def mean(*ws):
        return sum(ws) / len(ws)

    trainers = {p: Trainer.party(p).remote(i) for i, p in enumerate(parties)}
    server = parties[-1]
    expected = []
    cur = [float(i) for i in range(n)]
    for rnd in range(3):
        ws = [trainers[p].train.remote() for p in parties]
        avg = mean.party(server).remote(*ws)
        acks = [trainers[p].set_weights.remote(avg) for p in parties]
        cur = [c + 1.0 for c in cur]
        m = sum(cur) / n
        cur = [m] * n
        expected.append(m)
        got = fed.get(acks)
        assert got == [m] * n, (rnd, got, m)
    hist = fed.get([trainers[p].history.remote() for p in parties])
    assert hist == [expected] * n, hist
    fed.shutdown()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_n_party_fedavg_rounds(n):
    parties = [f"p{i}" for i in range(n)]
    run_parties(_n_party_fedavg, parties, args=(make_addresses(parties), "native"), timeout=150)


def test_four_party_fedavg_rounds_over_grpc():
    parties = [f"p{i}" for i in range(4)]
    run_parties(_n_party_fedavg, parties, args=(make_addresses(parties), "grpc"), timeout=150)


# -- fed.aggregate: cross-party mean as one DAG-level call (DAG path on CPU) ----------------------
def _aggregate_job(party, addresses):
    import numpy as np
    import torch

    fed = _init(party, addresses, "native")
    parties = list(addresses)
    n = len(parties)

    @fed.remote
    class Trainer:
        def __init__(self, rank):
            self.rank = rank
            self.w = None

        def weights(self):
            r = float(self.rank)
            return {"w": torch.full((3, 2), r), "b": np.full(4, r + 1.0), "steps": r * 10, "half": torch.full((2,), r).half()}

        def set_weights(self, w):
            self.w = w
            return True

        def check(self, expect_w, expect_b, expect_steps):
            w = self.w
            assert torch.allclose(w["w"], torch.full((3, 2), expect_w)) and w["w"].dtype == torch.float32
            assert np.allclose(w["b"], expect_b) and abs(w["steps"] - expect_steps) < 1e-9
            assert w["half"].dtype == torch.float16
            return True

    trainers = {p: Trainer.party(p).remote(i) for i, p in enumerate(parties)}
    before = fed._private.global_context.get_global_context().current_seq_id()
    new = fed.aggregate([trainers[p].weights.remote() for p in parties])
    after = fed._private.global_context.get_global_context().current_seq_id()
    assert after - before == n + 1 + (n - 1)  # n weights() calls + mean + (n-1) redistributions, in EVERY party
    assert [o.get_party() for o in new] == parties
    for p, w in zip(parties, new):
        trainers[p].set_weights.remote(w)
    m = (n - 1) / 2.0
    assert all(fed.get([trainers[p].check.remote(m, m + 1.0, m * 10) for p in parties]))

    # weighted, inputs given in a different order than the party list
    ws = [float(i + 1) for i in range(n)]
    tot = sum(ws)
    objs = [trainers[p].weights.remote() for p in reversed(parties)]
    new = fed.aggregate(objs, weights=[w / tot for w in reversed(ws)], fused=False)
    assert [o.get_party() for o in new] == list(reversed(parties))
    exp = sum(w * i for i, w in enumerate(ws)) / tot
    got = fed.get(new[0])
    assert torch.allclose(got["w"], torch.full((3, 2), exp))

    with pytest.raises(ValueError):
        fed.aggregate([trainers[parties[0]].weights.remote()])  # not one object per party
    with pytest.raises(RuntimeError):
        fed.aggregate([trainers[p].weights.remote() for p in parties], fused=True)  # no device channel here
    fed.shutdown()


@pytest.mark.parametrize("n", [2, 4])
def test_fed_aggregate_dag_path(n):
    parties = [f"p{i}" for i in range(n)]
    run_parties(_aggregate_job, parties, args=(make_addresses(parties),), timeout=150)
