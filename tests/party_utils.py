"""Helpers for multi-party tests: one OS process per party, exit-code asserts
(same shape as the reference's suite, fed/tests/test_fed_get.py:88-95)."""
import multiprocessing
import os
import socket
import sys

_CTX = multiprocessing.get_context("spawn")

FAST_RETRY = {
    "retry_max_attempts": 40,
    "retry_initial_backoff_s": 0.05,
    "retry_max_backoff_s": 0.5,
    "retry_backoff_multiplier": 1.5,
}


def free_ports(n):
    socks, ports = [], []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        socks.append(s)
        ports.append(s.getsockname()[1])
    for s in socks:
        s.close()
    return ports


def make_addresses(parties):
    ports = free_ports(len(parties))
    return {p: f"127.0.0.1:{port}" for p, port in zip(parties, ports)}


FAST_GRPC_RETRY = {
    "maxAttempts": 5,
    "initialBackoff": "0.1s",
    "maxBackoff": "1s",
    "backoffMultiplier": 2,
    "retryableStatusCodes": ["UNAVAILABLE"],
}


def comm_config(transport="native", **extra):
    cfg = dict(FAST_RETRY)
    cfg["transport"] = transport
    if transport in ("grpc", "auto"):
        cfg["grpc_retry_policy"] = dict(FAST_GRPC_RETRY)  # default backs off 5 s .. 30 s
    cfg.update(extra)
    return cfg


def _child_main(target, watchdog_s, party, args, kwargs):
    # a party that is about to be killed for hanging first dumps every thread's Python stack to
    # stderr (pytest shows it on failure) — the only way to learn anything from a rare flake
    import faulthandler

    faulthandler.dump_traceback_later(max(1.0, watchdog_s), exit=False)
    target(party, *args, **kwargs)


def run_parties(target, parties, args=(), timeout=90, expect=0, kwargs=None):
    """Run ``target(party, *args)`` in one process per party; assert exit codes."""
    procs = [_CTX.Process(target=_child_main, args=(target, timeout - 3.0, p, tuple(args), kwargs or {}))
             for p in parties]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
    codes = []
    for p in procs:
        if p.is_alive():
            p.kill()
            p.join()
            codes.append("timeout")
        else:
            codes.append(p.exitcode)
    expected = expect if isinstance(expect, (list, tuple)) else [expect] * len(procs)
    assert codes == list(expected), f"party exit codes {codes}, expected {list(expected)}"
    return codes
