"""Fault injection (mirrors fed/tests/test_cross_silo_error.py,
test_exit_on_failure_sending.py, serializations_tests/test_unpickle_with_whitelist.py)."""
import multiprocessing
import os

import pytest

from party_utils import comm_config, make_addresses, run_parties

PARTIES = ["alice", "bob"]


class MyError(Exception):
    pass


def _init(party, addresses, transport="native", handler=None, barrier=False, **comm):
    import rayfed_b200 as fed

    fed.init(addresses=addresses, party=party, logging_level=os.environ.get("FED_TEST_LOG", "error"),
             config={"cross_silo_comm": comm_config(transport, **comm),
                     # both parties are up before anything can fail: the tests are about error
                     # propagation, not about start-up races
                     "barrier_on_initializing": barrier},
             sending_failure_handler=handler)
    return fed


def _task_error(party, addresses, transport, use_actor, expose):
    fed = _init(party, addresses, transport, expose_error_trace=expose)
    from rayfed_b200 import FedRemoteError

    @fed.remote
    def error_func():
        raise MyError("Test normal task Error")

    @fed.remote
    class My:
        def error_func(self):
            raise MyError("Test actor task Error")

    o = My.party("alice").remote().error_func.remote() if use_actor else error_func.party("alice").remote()
    with pytest.raises(Exception) as e:
        fed.get(o)
    word = "actor task Error" if use_actor else "normal task Error"
    if party == "bob":
        assert isinstance(e.value.cause, FedRemoteError)
        assert "RemoteError occurred at alice" in str(e.value.cause)
        if expose:
            assert word in str(e.value.cause)
        else:
            assert "caused by" not in str(e.value.cause)
    else:
        assert isinstance(e.value.cause, MyError) and word in str(e.value.cause)
    fed.shutdown()


@pytest.mark.parametrize("transport", ["native", "grpc"])
@pytest.mark.parametrize("use_actor", [False, True])
def test_remote_error_reaches_both_parties(transport, use_actor):
    run_parties(_task_error, PARTIES, args=(make_addresses(PARTIES), transport, use_actor, True))


def test_error_trace_hidden_by_default():
    run_parties(_task_error, PARTIES, args=(make_addresses(PARTIES), "native", False, False))


def _error_into_downstream_task(party, addresses):
    fed = _init(party, addresses)
    from rayfed_b200 import FedRemoteError

    @fed.remote
    def error_func():
        raise MyError("boom")

    @fed.remote
    def consume(x):
        return x

    a = error_func.party("alice").remote()
    b = consume.party("bob").remote(a)  # bob's task never runs: its input is an error
    with pytest.raises(Exception) as e:
        fed.get(b)
    if party == "bob":
        assert isinstance(e.value.cause, FedRemoteError)
    fed.shutdown()


def test_error_propagates_through_downstream_task():
    run_parties(_error_into_downstream_task, PARTIES, args=(make_addresses(PARTIES),))


def _exit_on_failure_both(party, addresses):
    fed = _init(party, addresses, exit_on_sending_failure=True)

    @fed.remote
    def error_func():
        raise MyError("boom")

    @fed.remote
    def normal_func(a):
        return a

    try:
        a = error_func.party("alice").remote()
        b = normal_func.party("bob").remote(a)
        fed.get(b)
    finally:
        fed.shutdown()


def test_exit_on_sending_failure_exits_both_parties_with_1():
    run_parties(_exit_on_failure_both, PARTIES, args=(make_addresses(PARTIES),), expect=1)


def _peer_never_started(party, addresses, q):
    import time

    def handler(error):
        q.put("failure handler")

    fed = _init(party, addresses, handler=handler, barrier=False, exit_on_sending_failure=True,
                retry_max_attempts=2, retry_initial_backoff_s=0.1, retry_max_backoff_s=0.1)

    @fed.remote
    def f():
        return 1

    @fed.remote
    class My:
        def __init__(self, v):
            self.v = v

    o = f.party("alice").remote()
    My.party("bob").remote(o)  # bob does not exist: the push fails after the retries
    time.sleep(60)
    raise AssertionError("the SIGINT of the failure path should have ended this process")


def test_exit_when_peer_unreachable_runs_failure_handler():
    ctx = multiprocessing.get_context("spawn")
    q = ctx.Queue()
    run_parties(_peer_never_started, ["alice"], args=(make_addresses(PARTIES), q), expect=1, timeout=60)
    assert q.get(timeout=5) == "failure handler"


def _shutdown_once_after_error(party, addresses):
    import time

    fed = _init(party, addresses)
    from rayfed_b200._private.global_context import get_global_context

    @fed.remote
    def error_func():
        raise MyError("boom")

    @fed.remote
    def foo(e):
        return e

    a = error_func.party("alice").remote()
    foo.party("bob").remote(a)
    time.sleep(1.0)
    if party == "alice":
        assert get_global_context().get_cleanup_manager().get_last_sending_error() is not None
    fed.shutdown()
    fed.shutdown()


def test_sender_records_error_and_shutdown_is_clean():
    run_parties(_shutdown_once_after_error, PARTIES, args=(make_addresses(PARTIES),))


# -- restricted unpickling ---------------------------------------------------------------
class LocalThing:
    def __init__(self, v):
        self.v = v


def _whitelist(party, addresses, transport):
    import pickle

    import numpy

    fed = _init(party, addresses, transport,
                serializing_allowed_list={"numpy.core.numeric": ["*"], "numpy": None, "numpy.*": None})

    @fed.remote
    def make_array():
        return numpy.arange(4)

    @fed.remote
    def make_local():
        return LocalThing(3)

    @fed.remote
    def consume(x):
        return 1

    ok = consume.party("bob").remote(make_array.party("alice").remote())
    assert fed.get(ok) == 1
    bad = consume.party("bob").remote(make_local.party("alice").remote())
    if party == "bob":
        with pytest.raises(Exception) as e:
            fed.get(bad)
        assert isinstance(e.value, pickle.UnpicklingError) or "not allowed" in str(e.value)
    fed.shutdown()


@pytest.mark.parametrize("transport", ["native", "grpc"])
def test_serializing_allowed_list(transport):
    run_parties(_whitelist, PARTIES, args=(make_addresses(PARTIES), transport))


# -- controller divergence: the reference hangs silently; here the stalled fed.get says why ----------
def _diverged(party, addresses, log_path):
    import logging

    os.environ["RAYFED_B200_STALL_WARN_S"] = "0.5"
    import rayfed_b200 as fed
    from rayfed_b200 import api as fed_api

    fed_api._STALL_WARN_S = 0.5
    fed.init(addresses=addresses, party=party, logging_level="warning",
             config={"cross_silo_comm": comm_config("native"), "barrier_on_initializing": True})
    records = []

    class Grab(logging.Handler):
        def emit(self, rec):
            records.append(rec.getMessage())

    logging.getLogger("rayfed_b200.api").addHandler(Grab())

    @fed.remote
    def f(x):
        return x + 1

    if party == "alice":
        f.party("alice").remote(0)  # BUG in alice's copy of the script: one extra fed call
    a = f.party("alice").remote(1)  # seq ids now differ between the two controllers
    import threading

    out = []
    t = threading.Thread(target=lambda: out.append(fed.get(a)), daemon=True)
    t.start()
    t.join(2.5)
    if party == "bob":
        # bob waits for (1#0 -> 2) but alice sent (2#0 -> 3): never satisfiable
        assert t.is_alive()
        text = "\n".join(records)
        with open(log_path, "w") as fh:
            fh.write(text)
        assert "has been blocked" in text and "diverged" in text and "('2#0', '3')" in text, text
    else:
        import time

        deadline = time.time() + 40  # stay reachable until bob has made his observation
        while not os.path.exists(log_path) and time.time() < deadline:
            time.sleep(0.05)
    os._exit(0)  # the job is wedged by construction; leave without the drain


def test_diverged_controllers_are_diagnosed(tmp_path):
    addresses = make_addresses(["alice", "bob"])
    run_parties(_diverged, ["alice", "bob"], args=(addresses, str(tmp_path / "bob.log")), timeout=60)
