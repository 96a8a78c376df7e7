"""The native party runtime: futures, dependency gating, actor mailboxes, kill."""
import threading
import time

import pytest

from rayfed_b200 import core
from rayfed_b200.exceptions import FedActorDiedError, FedGetTimeoutError, FedTaskError


@pytest.fixture()
def rt():
    r = core.PartyRuntime(pool_threads=4)
    yield r
    r.shutdown()


def test_task_and_dependency_gating(rt):
    gate = core.Ref()
    a = rt.submit(lambda x: x + 1, (gate,))
    b = rt.submit(lambda x, y=0: x * y, (a,), {"y": a})
    assert not a.done()
    gate.set_value(2)
    assert core.get(b, timeout=5) == 9


def test_error_flows_along_edges(rt):
    def boom():
        raise ValueError("bad")

    a = rt.submit(boom)
    b = rt.submit(lambda x: x, (a,))
    with pytest.raises(FedTaskError) as e:
        core.get(b, timeout=5)
    assert isinstance(e.value.cause, ValueError) and "bad" in str(e.value.cause)
    assert "boom" in e.value.traceback_str


def test_num_returns(rt):
    x, y = rt.submit(lambda: (1, 2), num_returns=2)
    assert core.get([x, y], timeout=5) == [1, 2]
    bad = rt.submit(lambda: (1, 2, 3), num_returns=2)
    with pytest.raises(FedTaskError):
        core.get(bad[0], timeout=5)


def test_actor_mailbox_is_ordered_and_stateful(rt):
    class Acc:
        def __init__(self, v):
            self.v = v
            self.log = []

        def add(self, d):
            self.v += d
            self.log.append(d)
            return self.v

    gate = core.Ref()
    h = rt.create_actor(Acc, (10,))
    r1 = h.submit_method("add", (gate,), {})
    r2 = h.submit_method("add", (5,), {})
    time.sleep(0.05)
    assert not r2.done()  # FIFO: blocked behind r1's unresolved input
    gate.set_value(1)
    assert core.get([r1, r2], timeout=5) == [11, 16]
    assert h.instance.log == [1, 5]


def test_actor_kill_fails_pending_and_future_calls(rt):
    class Slow:
        def work(self, ev):
            ev.wait(2)
            return "done"

    ev = threading.Event()
    h = rt.create_actor(Slow)
    first = h.submit_method("work", (ev,), {})
    queued = h.submit_method("work", (ev,), {})
    time.sleep(0.05)
    h.kill()
    ev.set()
    with pytest.raises(FedActorDiedError):
        core.get(queued, timeout=5)
    with pytest.raises(FedActorDiedError):
        core.get(h.submit_method("work", (ev,), {}), timeout=5)
    assert core.get(first, timeout=5) == "done"  # the running call finishes


def test_get_timeout_and_wait(rt):
    never = core.Ref()
    with pytest.raises(FedGetTimeoutError):
        core.get(never, timeout=0.05)
    done = rt.submit(lambda: 1)
    ready, rest = core.wait([never, done], num_returns=1, timeout=2)
    assert ready == [done] and rest == [never]


def test_named_registry(rt):
    rt.register_named("x", object())
    with pytest.raises(ValueError):
        rt.register_named("x", object())
    assert rt.has_named("x")
    with pytest.raises(ValueError):
        rt.get_named("nope")


def test_mailbox_rendezvous_is_order_independent():
    nat = core.native()
    mb = nat.Mailbox()
    early = mb.claim("1", "2")  # consumer first
    mb.deliver_raw("1", "2", b"a")
    assert early.result(1) == b"a"
    mb.deliver_raw("3", "4", b"b")  # producer first
    assert mb.claim("3", "4").result(1) == b"b"
    assert mb.pending() == 0
    # never-claimed duplicates (the init ping) are overwritten, not accumulated
    mb.deliver_raw("ping", "ping", b"x")
    mb.deliver_raw("ping", "ping", b"y")
    assert mb.pending() == 1


def test_seq_counter():
    c = core.native().SeqCounter()
    assert [c.next() for _ in range(3)] == [1, 2, 3]
    c.reset(0)
    assert c.next() == 1


# -- device ordering in the C++ scheduler (csrc/core/device_hooks.h), with the recording fake ----
def test_executor_orders_consumer_streams_after_producer_events():
    from rayfed_b200.core import native

    nat = native()
    nat.install_fake_device_hooks()
    nat.fake_device_log(True)
    try:
        ex = nat.Executor(2, None)
        la = ex.create_lane("a", 1, lambda: 11)  # thread_init returns the lane's stream handle
        lb = ex.create_lane("b", 1, lambda: 22)
        r1 = ex.submit(la, lambda: 5, (), None, 1)[0]
        r2 = ex.submit(lb, lambda x: x + 1, (r1,), None, 1)[0]
        assert r2.result(5) == 6
        log = nat.fake_device_log()
        assert r1.cuda_event and r1.cuda_stream == 11 and r2.cuda_stream == 22
        assert f"record e{r1.cuda_event} s11" in log
        assert f"wait s22 e{r1.cuda_event}" in log  # the consumer's stream waits, not its thread
        assert log.index(f"wait s22 e{r1.cuda_event}") < log.index(f"record e{r2.cuda_event} s22")
        # same stream: already ordered, no wait is issued
        r3 = ex.submit(lb, lambda x: x, (r2,), None, 1)[0]
        assert r3.result(5) == 6
        assert f"wait s22 e{r2.cuda_event}" not in nat.fake_device_log()
        # num_returns > 1: every output carries the one event recorded after the body
        a, b = ex.submit(la, lambda: (1, 2), (), None, 2)
        assert (a.result(5), b.result(5)) == (1, 2) and a.cuda_event == b.cuda_event != 0
        # pool tasks pick the stream up after set_lane_thread_init (fed.init sets the device late)
        ex.set_lane_thread_init(0, lambda: 33)
        r4 = ex.submit(0, lambda x, y: x + y, (a, b), None, 1)[0]
        assert r4.result(5) == 3 and r4.cuda_stream == 33
        assert f"wait s33 e{a.cuda_event}" in nat.fake_device_log()
        # a host read synchronises the event (GIL released)
        r4.device_sync()
        assert f"sync e{r4.cuda_event}" in nat.fake_device_log()
        ex.shutdown(True, 2.0)
    finally:
        nat.set_device_hooks(None)


def test_ref_adopts_the_event_of_a_device_value():
    from rayfed_b200.core import native
    from rayfed_b200.core.runtime import DeviceValue, is_device_value

    nat = native()
    dv = DeviceValue({"k": 1}, 77, 5, None)
    assert is_device_value(dv) and not is_device_value({"k": 1})
    r = nat.Ref()
    r.set_value(dv)
    assert r.result(1) == {"k": 1} and r.cuda_event == 77 and r.cuda_stream == 5
    # a decoder may return one too (the NVLink receiver does)
    r2 = nat.Ref()
    r2.set_decoder(lambda data, is_err: DeviceValue(data.decode(), 88, 6, None))
    r2.set_raw(b"payload", False)
    assert r2.result(1) == "payload" and r2.cuda_event == 88
    plain = nat.Ref()
    plain.set_value(3)
    assert plain.cuda_event == 0
    plain.device_sync()  # no event, no hooks: a no-op


def test_actor_lane_waits_for_a_fan_in_on_one_latch(rt):
    """An actor method with many outstanding inputs (the aggregator of N parties) runs exactly when
    the LAST one resolves, whatever the order and the thread of the completions, and sees every
    value (csrc/core/executor.h::wait_inputs: one countdown latch instead of N sequential waits)."""
    import random

    class Agg:
        def total(self, *xs):
            return sum(xs)

    h = rt.create_actor(Agg)
    for round_no in range(20):
        deps = [core.Ref() for _ in range(9)]
        deps[3].set_value(100)                                  # some inputs are already there
        out = h.submit_method("total", tuple(deps) + (deps[0],), {})   # the same input twice
        order = [i for i in range(9) if i != 3]
        random.Random(round_no).shuffle(order)
        last = order.pop()

        def complete(ids):
            for i in ids:
                deps[i].set_value(i)

        ts = [threading.Thread(target=complete, args=(order[k::3],)) for k in range(3)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        time.sleep(0.002)
        assert not out.done()                                   # one input is still missing
        deps[last].set_value(last)
        assert out.result(5) == sum(i for i in range(9) if i != 3) + 100 + 0
    # an input that fails: the error flows along the edge, later calls still run
    bad, ok = core.Ref(), core.Ref()
    failing = h.submit_method("total", (bad, ok, 1), {})
    after = h.submit_method("total", (1, 2), {})
    ok.set_value(5)
    bad.set_error(ValueError("boom"))
    with pytest.raises(Exception, match="boom"):
        failing.result(5)
    assert after.result(5) == 3
