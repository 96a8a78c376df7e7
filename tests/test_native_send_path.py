"""The GIL-free send bookkeeping: native ack decoding, `Ref.forward_to`, `SendTracker`,
`SocketTransport.send_many` (csrc/core/ref.h, send_tracker.h, module.cpp).

The reference watches every send with one `ray.get` per message on a polling thread
(fed/cleanup.py:150-152); here an acknowledged send must not touch the interpreter at all.
"""
import socket
import threading
import time

import pytest

from rayfed_b200.core import native


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_ack_mode_decodes_without_a_python_decoder():
    n = native()
    ok = n.Ref()
    ok.set_ack_mode()
    ok.set_raw(b"OK", False)
    assert ok.result(1) is True and ok.exception(1) is None
    bad = n.Ref()
    bad.set_ack_mode()
    bad.set_raw(b"JobName mis-match", True)
    assert bad.is_error()
    with pytest.raises(RuntimeError, match="mis-match"):
        bad.result(1)


def test_forward_to_hands_over_raw_value_and_error_outcomes():
    n = native()
    # raw outcome, completed later from a foreign thread: the destination decodes it itself
    src, dst = n.Ref(), n.Ref()
    src.set_ack_mode()
    src.forward_to(dst)
    assert not dst.done()
    t = threading.Thread(target=lambda: src.set_raw(b"OK", False))
    t.start()
    t.join()
    assert dst.result(1) is True
    # a Python decoder travels with the forward
    src, dst = n.Ref(), n.Ref()
    src.set_decoder(lambda data, is_error: data.decode().upper())
    src.forward_to(dst)
    src.set_raw(b"abc", False)
    assert dst.result(1) == "ABC" and src.result(1) == "ABC"
    # value and error outcomes, source already resolved
    src, dst = n.Ref(), n.Ref()
    src.set_value({"k": 1})
    src.forward_to(dst)
    assert dst.result(1) == {"k": 1}
    src, dst = n.Ref(), n.Ref()
    src.set_error(ValueError("boom"))
    src.forward_to(dst)
    with pytest.raises(ValueError, match="boom"):
        dst.result(1)
    # callbacks of the destination fire once
    seen = []
    src, dst = n.Ref(), n.Ref()
    dst.add_done_callback(lambda r: seen.append(r.result()))
    src.forward_to(dst)
    src.set_value(7)
    assert seen == [7]


def test_send_tracker_retires_acknowledged_sends_natively():
    n = native()
    calls = []
    tracker = n.SendTracker(lambda ref, pack, is_error: calls.append((pack, is_error)))
    acks = []
    for i in range(50):
        r = n.Ref()
        r.set_ack_mode()
        tracker.watch(r, ("pack", i), False)
        acks.append(r)
    assert tracker.pending() == 50
    t = threading.Thread(target=lambda: [r.set_raw(b"OK", False) for r in acks])
    t.start()
    t.join()
    assert tracker.pending() == 0 and tracker.fast_retired() == 50 and tracker.slow_settled() == 0
    assert calls == []
    assert tracker.wait_idle(0.1)
    tracker.bury()


def test_send_tracker_slow_path_failures_and_python_outcomes():
    n = native()
    calls = []

    def handler(ref, pack, is_error):
        calls.append((pack, is_error, ref.exception() is None))
        tracker.retire(is_error)

    tracker = n.SendTracker(handler)
    failed = n.Ref()
    failed.set_ack_mode()
    tracker.watch(failed, ("to-bob", 1, 2), False)
    value = n.Ref()
    tracker.watch(value, ("generic-proxy", 3, 4), False)
    err = n.Ref()
    err.set_ack_mode()
    tracker.watch(err, ("error-msg", 5, 6), True)
    assert tracker.pending() == 3 and tracker.pending_errors() == 1
    assert not tracker.wait_idle(0.01)
    failed.set_raw(b"connection lost", True)
    value.set_value(True)
    assert tracker.pending() == 1
    tracker.set_draining(True)
    threading.Timer(0.05, lambda: err.set_raw(b"OK", False)).start()
    assert tracker.wait_idle(2.0, True) and tracker.wait_idle(2.0)
    assert calls == [(("to-bob", 1, 2), False, False), (("generic-proxy", 3, 4), False, True)]
    assert tracker.fast_retired() == 1 and tracker.slow_settled() == 2
    assert not tracker.settling()


def test_send_many_one_payload_to_several_parties():
    n = native()
    ports = [_free_port() for _ in range(3)]
    names = ["bob", "carol", "dave"]
    receivers = []
    for name, port in zip(names, ports):
        mb = n.Mailbox()
        t = n.SocketTransport(name, "job", mb)
        assert not t.listen(f"127.0.0.1:{port}")
        receivers.append((t, mb))
    sender = n.SocketTransport("alice", "job", n.Mailbox())
    for name, port in zip(names, ports):
        sender.add_peer(name, f"127.0.0.1:{port}", 5, 0.05, 0.5, 2.0, 10.0)
    try:
        acks = sender.send_many(names, "7", "9", b"payload")
        assert [a.result(10) for a in acks] == [True, True, True]
        for _, mb in receivers:
            assert mb.claim("7", "9").result(5) == b"payload"
        with pytest.raises(KeyError):
            sender.send_many(["bob", "nobody"], "8", "9", b"x")
    finally:
        sender.stop()
        for t, _ in receivers:
            t.stop()


def test_spin_info_reports_the_policy():
    info = native().spin_info()
    assert info["usable_cpus"] >= 1 and info["base_spin_us"] >= 0
    assert info["adaptive_cap_us"] >= 0 and info["max_long_spinners"] >= 0
    if info["usable_cpus"] < 32:
        import os

        if "RAYFED_B200_ADAPTIVE_SPIN_US" not in os.environ:
            assert info["adaptive_cap_us"] == 0   # small machines never spin long by default


def test_cleanup_manager_counts_through_the_tracker():
    from rayfed_b200.cleanup import CleanupManager

    n = native()
    cm = CleanupManager("alice", lambda: True)
    cm.start()
    try:
        refs = []
        for i in range(20):
            r = n.Ref()
            r.set_ack_mode()
            cm.push_to_sending(r, "bob", i, i + 1)
            refs.append(r)
        assert cm.pending_sends() == 20
        for r in refs:
            r.set_raw(b"OK", False)
        deadline = time.monotonic() + 2
        while cm.pending_sends() and time.monotonic() < deadline:
            time.sleep(0.001)
        assert cm.pending_sends() == 0
    finally:
        cm.stop(wait_for_sending=True)


def test_claim_many_matches_single_claims():
    n = native()
    mb = n.Mailbox()
    mb.deliver_obj("1#0", "9", "early", False)          # arrived before anybody asked
    refs = mb.claim_many(["1#0", "2#0", "3#0"], "9", None)
    assert refs[0].result(1) == "early" and not refs[1].done() and not refs[2].done()
    mb.deliver_obj("3#0", "9", "late", False)
    assert refs[2].result(1) == "late"
    # a second claim of a pending key shares the future; decoders are attached to all
    again = mb.claim_many(["2#0"], "9", lambda data, is_error: data.decode() + "!")
    mb2 = n.Mailbox()
    r = mb2.claim_many(["a"], "b", lambda data, is_error: data.decode() + "!")[0]
    r.set_raw(b"x", False)
    assert r.result(1) == "x!"
    mb.deliver_obj("2#0", "9", 5, False)
    assert refs[1].result(1) == 5 and again[0].result(1) == 5
    assert mb.pending() == 0


def test_actor_methods_are_cached_and_options_do_not_leak():
    from rayfed_b200._private.fed_actor import FedActorHandle

    class Body:
        def work(self):
            return 1

    h = FedActorHandle(1, {"alice": "127.0.0.1:1", "bob": "127.0.0.1:2"}, Body, "alice", "bob", {})
    m = h.work
    assert h.work is m                                   # looked up once
    two = m.options(num_returns=2)
    assert two is not m and two._options == {"num_returns": 2}
    assert m._options == {} and h.work._options == {}    # the cached method is untouched
    with pytest.raises(AttributeError):
        h.no_such_method


def test_doorbell_page_wakes_several_sleeping_receivers_with_one_call(tmp_path):
    """Opt-in path (RAYFED_B200_DOORBELL_PAGE=1): the receivers of one sender sleep on ONE futex
    word of the sender's anonymous doorbell page, each with its own wake bit; a broadcast rings
    them together.  Runs in a subprocess because the switch is read once per process."""
    import os
    import subprocess
    import sys
    import textwrap

    code = textwrap.dedent("""
        import socket, time
        from rayfed_b200.core import native
        def fp():
            s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p
        n = native(); names = ["b", "c", "d"]; ports = [fp() for _ in names]; rxs = []
        for nm, p in zip(names, ports):
            mb = n.Mailbox(); t = n.SocketTransport(nm, "job", mb); assert not t.listen(f"127.0.0.1:{p}"); rxs.append((t, mb))
        tx = n.SocketTransport("a", "job", n.Mailbox())
        for nm, p in zip(names, ports):
            tx.add_peer(nm, f"127.0.0.1:{p}", 5, 0.05, 0.5, 2.0, 10.0)
        slow = 0
        for i in range(40):
            time.sleep(0.01)                       # every reader is asleep again
            t0 = time.perf_counter()
            acks = tx.send_many(names, str(i), "k", b"x")
            for _, mb in rxs:
                assert mb.claim(str(i), "k").result(5) == b"x"
            slow += time.perf_counter() - t0 > 0.09
            assert [a.result(5) for a in acks] == [True] * 3
        # unicast goes through the page too (exact wake: one bit)
        time.sleep(0.01)
        assert tx.send("c", "u", "k", b"y").result(5) is True and rxs[1][1].claim("u", "k").result(5) == b"y"
        assert tx.stats()["paged_conns"] == 3 and all(t.stats()["paged_conns"] == 1 for t, _ in rxs)
        assert slow <= 2, slow                    # a missed wake-up costs a 100 ms slice, every time
        tx.stop(); [t.stop() for t, _ in rxs]
        print("OK")
    """)
    env = dict(os.environ, RAYFED_B200_DOORBELL_PAGE="1",
               PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
