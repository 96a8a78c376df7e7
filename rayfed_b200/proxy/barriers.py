"""Barriers: proxy hosts and the ``send`` / ``recv`` / ``ping_others`` primitives.

Parity: fed/proxy/barriers.py (proxy actors :113-240,:339-412; ``_start_*``
:248-330,:415-459; naming :31-85; ``send`` :462-488; ``recv`` :491-494;
``ping_others`` :497-523).

In the reference the proxies are Ray *actors* in separate processes.  Here a
proxy lives in the controller process inside a **ProxyHost** that keeps the
same surface (``is_ready / send / get_data / _get_stats / _get_proxy_config``),
is registered under the same names in the party runtime's named registry
(``ray.get_actor`` analogue), and bridges three kinds of transport methods:

* plain methods returning a value,
* plain methods returning a native `Ref` (the native transports — completion
  is signalled by the C++ thread, no Python thread is parked), and
* ``async def`` methods (gRPC), run on the host's private asyncio loop.

``send`` keeps the reference's key trick: the *future* of the produced value is
handed over, and the push is released by the scheduler the moment the producer
finishes (fed/proxy/barriers.py:478-483) — the driver never blocks.
"""
from __future__ import annotations

import asyncio
import inspect
import itertools
import logging
import threading
import time
from typing import Dict, Optional

from .. import config as fed_config
from .._private import constants
from .._private.global_context import get_global_context
from ..core import get_runtime, is_ref, native
from ..core.runtime import DeviceValue, is_device_value
from ..exceptions import FedRemoteError, FedTaskError
from .. import _tracing
from .base_proxy import ReceiverProxy, SenderProxy, SenderReceiverProxy  # noqa: F401  (re-exported)

logger = logging.getLogger(__name__)

_SENDER_PROXY_ACTOR_NAME = None
_RECEIVER_PROXY_ACTOR_NAME = None
_SENDER_RECEIVER_PROXY_ACTOR_NAME = None

_DEFAULT_SENDER_PROXY_ACTOR_NAME = constants.RAYFED_DEFAULT_SENDER_PROXY_ACTOR_NAME
_DEFAULT_RECEIVER_PROXY_ACTOR_NAME = constants.RAYFED_DEFAULT_RECEIVER_PROXY_ACTOR_NAME
_DEFAULT_SENDER_RECEIVER_PROXY_ACTOR_NAME = constants.RAYFED_DEFAULT_SENDER_RECEIVER_PROXY_ACTOR_NAME


def sender_proxy_actor_name() -> str:
    return _SENDER_PROXY_ACTOR_NAME or _DEFAULT_SENDER_PROXY_ACTOR_NAME


def receiver_proxy_actor_name() -> str:
    return _RECEIVER_PROXY_ACTOR_NAME or _DEFAULT_RECEIVER_PROXY_ACTOR_NAME


def set_sender_proxy_actor_name(name: str):
    global _SENDER_PROXY_ACTOR_NAME
    _SENDER_PROXY_ACTOR_NAME = name


def set_receiver_proxy_actor_name(name: str):
    global _RECEIVER_PROXY_ACTOR_NAME
    _RECEIVER_PROXY_ACTOR_NAME = name


def set_proxy_actor_name(job_name: str, use_global_proxy: bool, sender_recvr_proxy: bool = False):
    """Global names, or per-job ``{name}_{job}`` when ``use_global_proxy`` is off.
    With a combined proxy both names point at the one host."""
    sender = (
        _DEFAULT_SENDER_RECEIVER_PROXY_ACTOR_NAME if sender_recvr_proxy else _DEFAULT_SENDER_PROXY_ACTOR_NAME
    )
    receiver = (
        _DEFAULT_SENDER_RECEIVER_PROXY_ACTOR_NAME if sender_recvr_proxy else _DEFAULT_RECEIVER_PROXY_ACTOR_NAME
    )
    if not use_global_proxy:
        sender = f"{sender}_{job_name}"
        receiver = f"{receiver}_{job_name}"
    set_sender_proxy_actor_name(sender)
    set_receiver_proxy_actor_name(receiver)


# -- tiny helpers kept for transports written against the reference -----------
def key_exists_in_two_dim_dict(the_dict, key_a, key_b) -> bool:
    return key_a in the_dict and key_b in the_dict[key_a]


def add_two_dim_dict(the_dict, key_a, key_b, val):
    the_dict.setdefault(key_a, {})[key_b] = val


def get_from_two_dim_dict(the_dict, key_a, key_b):
    return the_dict[key_a][key_b]


def pop_from_two_dim_dict(the_dict, key_a, key_b):
    return the_dict[key_a].pop(key_b)


# ---------------------------------------------------------------------------
class _LoopThread:
    """A private asyncio loop for transports with ``async def`` methods."""

    def __init__(self, name: str):
        self._loop = asyncio.new_event_loop()
        self._thread = threading.Thread(target=self._run, name=name, daemon=True)
        self._thread.start()

    def _run(self):
        asyncio.set_event_loop(self._loop)
        self._loop.run_forever()

    @property
    def loop(self):
        return self._loop

    def submit(self, coro):
        return asyncio.run_coroutine_threadsafe(coro, self._loop)

    def stop(self):
        if self._loop.is_running():
            self._loop.call_soon_threadsafe(self._loop.stop)
        self._thread.join(timeout=2)


def _needs_loop(proxy) -> bool:
    for name in ("send", "get_data", "start"):
        fn = getattr(proxy, name, None)
        if fn is not None and inspect.iscoroutinefunction(fn):
            return True
    return False


class ProxyHost:
    """In-process stand-in for a proxy actor."""

    def __init__(self, name: str, party: str, job_name: str, proxy, role: str):
        self._name = name
        self._party = party
        self._job_name = job_name
        self._proxy_instance = proxy
        self._role = role
        # itertools.count.__next__ is atomic under the GIL: lock-free op counters
        self._send_ops = itertools.count(1) if role in ("sender", "sender_receiver") else None
        self._recv_ops = itertools.count(1) if role in ("receiver", "sender_receiver") else None
        self._send_seen = 0
        self._recv_seen = 0
        self._native_refs = bool(getattr(proxy, "REFS_RAISE_REMOTE_ERRORS", False))
        self._device_aware = bool(getattr(proxy, "DEVICE_AWARE", False))
        self._loop: Optional[_LoopThread] = _LoopThread(f"{name}-loop") if _needs_loop(proxy) else None
        self._start_future = None

    # -- plumbing --------------------------------------------------------------
    @property
    def name(self):
        return self._name

    @property
    def proxy(self):
        return self._proxy_instance

    def _resolve(self, out, timeout=None):
        """Block for the outcome of a proxy call of any flavour."""
        if inspect.isawaitable(out):
            return self._loop_or_raise().submit(_await(out)).result(timeout)
        if is_ref(out):
            return out.result(-1.0 if timeout is None else timeout)
        return out

    def _loop_or_raise(self) -> _LoopThread:
        if self._loop is None:
            self._loop = _LoopThread(f"{self._name}-loop")
        return self._loop

    def _into_ref(self, out, result_ref, on_value=None):
        """Complete ``result_ref`` with the outcome of a proxy call (any flavour)."""

        def finish_value(v):
            if on_value is not None:
                try:
                    on_value(v, result_ref)
                    return
                except BaseException as e:  # noqa: BLE001
                    result_ref.set_error(e)
                    return
            result_ref.set_value(v)

        if inspect.isawaitable(out):
            fut = self._loop_or_raise().submit(_await(out))

            def done(f):
                try:
                    finish_value(f.result())
                except BaseException as e:  # noqa: BLE001
                    result_ref.set_error(e)

            fut.add_done_callback(done)
        elif is_ref(out):

            def chained(r):
                e = r.exception()
                if e is not None:
                    result_ref.set_error(e)
                else:
                    finish_value(r.result())

            out.add_done_callback(chained)
        else:
            finish_value(out)

    # -- actor-like surface ------------------------------------------------------
    def start(self):
        """Start serving. Async servers run for ever on the loop thread."""
        start = getattr(self._proxy_instance, "start", None)
        if start is None:
            return
        out = start()
        if inspect.isawaitable(out):
            self._start_future = self._loop_or_raise().submit(_await(out))

    def is_ready(self, timeout: Optional[float] = None):
        out = self._proxy_instance.is_ready()
        try:
            return self._resolve(out, timeout)
        except (TimeoutError, asyncio.TimeoutError):
            from ..exceptions import FedGetTimeoutError

            raise FedGetTimeoutError(f"proxy {self._name} was not ready within {timeout}s") from None

    def send(self, dest_party, data, upstream_seq_id, downstream_seq_id, is_error=False):
        """Issue the push; returns a `Ref` that resolves truthy on success.  A
        transport failure resolves to False (generic proxies, like
        fed/proxy/barriers.py:170-172) or raises from the ack future (native
        proxies) — the CleanupManager treats both as a failed send."""
        self._send_seen = next(self._send_ops)
        try:
            if is_device_value(data) and not self._device_aware:
                # host-only transport: the value must be complete before pickling
                if data.owner is not None:
                    data.owner.device_sync()
                data = data.value
            out = self._proxy_instance.send(dest_party, data, upstream_seq_id, downstream_seq_id)
        except KeyError:
            raise AssertionError(
                f"Failed to find {dest_party} in addresses {self._proxy_instance._addresses}.") from None
        except BaseException as e:  # noqa: BLE001
            logger.error(f"Failed to {'send error' if is_error else 'send'} to {dest_party}: {e!r}")
            result = native().Ref()
            result.set_value(False)
            return result
        if self._native_refs and is_ref(out):
            return out  # the C++ ack future: True, or raises the transport error
        result = native().Ref()
        inner = native().Ref()
        self._into_ref(out, inner)

        def finish(r):
            e = r.exception()
            if e is not None:
                logger.error(
                    f"Failed to {'send error' if is_error else 'send data'} to {dest_party}, "
                    f"upstream_seq_id: {upstream_seq_id}, downstream_seq_id: {downstream_seq_id}: {e}"
                )
                result.set_value(False)
            else:
                v = r.result()
                result.set_value(True if v is None else v)

        inner.add_done_callback(finish)
        return result

    def get_data(self, src_party, upstream_seq_id, curr_seq_id):
        """`Ref` of the message keyed (upstream, curr). A delivered
        ``FedRemoteError`` becomes the Ref's *error*, wrapped like a failed
        local task (parity: fed/proxy/barriers.py:227-233)."""
        self._recv_seen = next(self._recv_ops)
        out = self._proxy_instance.get_data(src_party, upstream_seq_id, curr_seq_id)
        if self._native_refs and is_ref(out):
            return out  # native transports already complete the Ref correctly
        result = native().Ref()

        def on_value(v, ref):
            if isinstance(v, FedRemoteError):
                ref.set_error(FedTaskError(v, "", "recv"))
            else:
                ref.set_value(v)

        self._into_ref(out, result, on_value)
        return result

    def get_data_many(self, items, curr_seq_id):
        """Batch flavour of `get_data` for the remote arguments of one task ([(src, upstream id)]);
        None when the proxy has no batch entry point (the caller falls back to single claims)."""
        many = getattr(self._proxy_instance, "get_data_many", None) if self._native_refs else None
        if many is None:
            return None
        for _ in items:
            self._recv_seen = next(self._recv_ops)
        return many(items, curr_seq_id)

    def _get_stats(self):
        stats = {}
        if self._send_ops is not None:
            stats["send_op_count"] = self._send_seen
        if self._recv_ops is not None:
            stats["receive_op_count"] = self._recv_seen
        extra = getattr(self._proxy_instance, "get_stats", None)
        if callable(extra):
            try:
                stats.update({f"transport_{k}": v for k, v in extra().items()})
            except Exception:  # noqa: BLE001
                pass
        return stats

    def _get_proxy_config(self, dest_party=None):
        fn = self._proxy_instance.get_proxy_config
        try:
            out = fn(dest_party) if dest_party is not None else fn()
        except TypeError:
            out = fn()
        return self._resolve(out)

    def stop(self):
        stop = getattr(self._proxy_instance, "stop", None)
        if callable(stop):
            try:
                self._resolve(stop(), 5)
            except Exception:  # noqa: BLE001
                logger.debug("proxy stop failed", exc_info=True)
        if self._loop is not None:
            self._loop.stop()
            self._loop = None


async def _await(x):
    return await x


# Class names kept so `from fed.proxy.barriers import SenderProxyActor` resolves.
class SenderProxyActor(ProxyHost):
    def __init__(self, addresses: Dict, party: str, job_name: str, tls_config: Dict = None,
                 logging_level: str = None, proxy_cls=None, name: str = None):
        cfg = fed_config.get_job_config(job_name).cross_silo_comm_config_dict
        proxy = proxy_cls(addresses, party, job_name, tls_config, cfg)
        super().__init__(name or sender_proxy_actor_name(), party, job_name, proxy, "sender")


class ReceiverProxyActor(ProxyHost):
    def __init__(self, listening_address: str, party: str, job_name: str, logging_level: str,
                 tls_config=None, proxy_cls=None, name: str = None):
        cfg = fed_config.get_job_config(job_name).cross_silo_comm_config_dict
        proxy = proxy_cls(listening_address, party, job_name, tls_config, cfg)
        super().__init__(name or receiver_proxy_actor_name(), party, job_name, proxy, "receiver")


class SenderReceiverProxyActor(ProxyHost):
    def __init__(self, addresses: Dict, party: str, job_name: str, tls_config: Dict = None,
                 logging_level: str = None, proxy_cls=None, name: str = None):
        cfg = fed_config.get_job_config(job_name).cross_silo_comm_config_dict
        proxy = proxy_cls(addresses, party, tls_config, cfg)
        if not hasattr(proxy, "_addresses"):
            proxy._addresses = addresses
        super().__init__(name or sender_proxy_actor_name(), party, job_name, proxy, "sender_receiver")


def _job_name() -> str:
    ctx = get_global_context()
    return ctx.get_job_name() if ctx is not None else constants.RAYFED_DEFAULT_JOB_NAME


_HOSTS = {"sender": None, "receiver": None}  # hot-path cache of the two proxy hosts


def _register(host: ProxyHost, *names):
    rt = get_runtime()
    for n in dict.fromkeys(names):
        rt.register_named(n, host)
        if n == sender_proxy_actor_name():
            _HOSTS["sender"] = host
        if n == receiver_proxy_actor_name():
            _HOSTS["receiver"] = host


def _start_receiver_proxy(addresses: Dict, party: str, logging_level: str, tls_config=None, proxy_cls=None,
                          proxy_config: Optional[Dict] = None, ready_timeout_second: int = 60):
    host = ReceiverProxyActor(
        listening_address=addresses[party], party=party, job_name=_job_name(), logging_level=logging_level,
        tls_config=tls_config, proxy_cls=proxy_cls,
    )
    _register(host, receiver_proxy_actor_name())
    host.start()
    server_state = host.is_ready(timeout=ready_timeout_second)
    assert server_state[0], server_state[1]
    logger.info("Succeeded to create receiver proxy actor.")
    return host


def _start_sender_proxy(addresses: Dict, party: str, logging_level: str, tls_config: Dict = None, proxy_cls=None,
                        proxy_config: Optional[Dict] = None, ready_timeout_second: int = 60):
    host = SenderProxyActor(
        addresses, party, job_name=_job_name(), tls_config=tls_config, logging_level=logging_level,
        proxy_cls=proxy_cls,
    )
    _register(host, sender_proxy_actor_name())
    assert host.is_ready(timeout=ready_timeout_second)
    logger.info("SenderProxyActor has successfully created.")
    return host


def _start_sender_receiver_proxy(addresses: Dict, party: str, logging_level: str, tls_config: Dict = None,
                                 proxy_cls=None, proxy_config: Optional[Dict] = None,
                                 ready_timeout_second: int = 60):
    host = SenderReceiverProxyActor(
        addresses, party, job_name=_job_name(), tls_config=tls_config, logging_level=logging_level,
        proxy_cls=proxy_cls,
    )
    _register(host, sender_proxy_actor_name(), receiver_proxy_actor_name())
    host.start()
    server_state = host.is_ready(timeout=ready_timeout_second)
    if isinstance(server_state, tuple):
        assert server_state[0], server_state[1]
    else:
        assert server_state
    logger.info("Succeeded to create receiver proxy actor.")
    return host


def _stop_proxies(drain: bool = False, drain_timeout: float = 30.0):
    """Stop both proxy hosts.  ``drain``: first let the sender flush what is still queued towards the
    peers (frames whose acknowledgement nobody waits for — device metadata, capability records —
    would otherwise be dropped by ``transport.stop()``)."""
    _HOSTS["sender"] = _HOSTS["receiver"] = None
    rt = get_runtime(create=False)
    if rt is None:
        return
    seen = set()
    for n in (sender_proxy_actor_name(), receiver_proxy_actor_name()):
        host = rt.unregister_named(n)
        if host is not None and id(host) not in seen:
            seen.add(id(host))
            if drain:
                fn = getattr(host.proxy, "drain", None)
                if callable(fn):
                    try:
                        fn(drain_timeout)
                    except Exception:  # noqa: BLE001
                        logger.debug("proxy drain failed", exc_info=True)
            host.stop()


# ---------------------------------------------------------------------------
def _copy_outcome(src, dst):
    e = src.exception()
    if e is not None:
        dst.set_error(e)
    else:
        dst.set_value(src.result())


def _value_for_send(ref):
    """The produced value, together with its device-completion event when it has one: a
    device-aware transport orders its push stream after the event instead of blocking."""
    v = ref.result()
    ev = ref.cuda_event
    return DeviceValue(v, ev, ref.cuda_stream, ref) if ev else v


def send(dest_party, data, upstream_seq_id, downstream_seq_id, is_error=False, check_sending=True):
    """Push ``data`` (a value or the `Ref` of one) to ``dest_party``.

    Returns the `Ref` of the send outcome.  When ``data`` is an unresolved
    `Ref` the push is issued by a completion callback on the thread that
    finishes the producer (no polling, no extra thread hop); if the producer
    *failed*, the outcome Ref carries that ``FedTaskError`` — which is how the
    CleanupManager tells "producer failed" from "link failed".
    """
    host: ProxyHost = _HOSTS["sender"] or get_runtime().get_named(sender_proxy_actor_name())
    if is_ref(data):
        if data.done():
            e = data.exception()
            if e is not None:
                result = native().Ref()
                result.set_error(e)
            else:
                result = host.send(dest_party, _value_for_send(data), upstream_seq_id, downstream_seq_id, is_error)
        else:
            result = native().Ref()

            def on_ready(r):
                e = r.exception()
                if e is not None:
                    result.set_error(e)
                    return
                try:
                    inner = host.send(dest_party, _value_for_send(r), upstream_seq_id, downstream_seq_id, is_error)
                except BaseException as ex:  # noqa: BLE001
                    result.set_error(ex)
                    return
                inner.forward_to(result)  # native continuation: an ack never takes the GIL

            data.add_done_callback(on_ready)
    else:
        result = host.send(dest_party, data, upstream_seq_id, downstream_seq_id, is_error)

    if _tracing._TRACER is not None:
        _tracing.trace_ref("send", result, seq_id=f"{upstream_seq_id}->{downstream_seq_id}", dest=dest_party,
                           is_error=bool(is_error))
    if check_sending:
        ctx = get_global_context()
        if ctx is not None:
            ctx.get_cleanup_manager().push_to_sending(
                result, dest_party, upstream_seq_id, downstream_seq_id, is_error
            )
    return result


def send_broadcast(dest_parties, data, upstream_seq_id, downstream_seq_id):
    """Owner -> all-other-parties push of one value (what ``fed.get`` does).  When the
    transport can multicast (NVLink channel with an NVLS mapping) every big CUDA tensor is
    written ONCE and the switch replicates it (K1b); otherwise this is N-1 ``send``s, like
    the reference (fed/api.py:564-579).  Returns {dest: outcome Ref}."""
    host: ProxyHost = _HOSTS["sender"] or get_runtime().get_named(sender_proxy_actor_name())
    proxy = host.proxy
    can = getattr(proxy, "can_broadcast", None)
    dest_parties = list(dest_parties)
    if not dest_parties:
        return {}
    if can is None or not can() or len(dest_parties) != len(proxy._addresses) - 1:
        return {d: send(d, data, upstream_seq_id, downstream_seq_id) for d in dest_parties}

    results = {d: native().Ref() for d in dest_parties}

    def issue(value):
        try:
            if is_device_value(value) and not host._device_aware:
                if value.owner is not None:
                    value.owner.device_sync()
                value = value.value
            for _ in dest_parties:
                host._send_seen = next(host._send_ops)
            acks = proxy.broadcast(dest_parties, value, upstream_seq_id, downstream_seq_id)
        except BaseException as e:  # noqa: BLE001
            for r in results.values():
                r.set_error(e)
            return
        for d, ack in acks.items():
            ack.forward_to(results[d])

    if is_ref(data):
        if data.done():
            e = data.exception()
            if e is not None:
                for r in results.values():
                    r.set_error(e)
            else:
                issue(_value_for_send(data))
        else:
            def on_ready(r):
                e = r.exception()
                if e is not None:
                    for res in results.values():
                        res.set_error(e)
                else:
                    issue(_value_for_send(r))

            data.add_done_callback(on_ready)
    else:
        issue(data)
    if _tracing._TRACER is not None:
        for d, r in results.items():
            _tracing.trace_ref("send", r, seq_id=f"{upstream_seq_id}->{downstream_seq_id}", dest=d, is_error=False,
                               broadcast=True)
    ctx = get_global_context()
    if ctx is not None:
        for d, r in results.items():
            ctx.get_cleanup_manager().push_to_sending(r, d, upstream_seq_id, downstream_seq_id, False)
    return results


def recv(party: str, src_party: str, upstream_seq_id, curr_seq_id):
    """`Ref` that resolves when the message keyed (upstream, curr) has arrived."""
    assert party, "Party can not be None."
    host: ProxyHost = _HOSTS["receiver"] or get_runtime().get_named(receiver_proxy_actor_name())
    ref = host.get_data(src_party, upstream_seq_id, curr_seq_id)
    if _tracing._TRACER is not None:
        _tracing.trace_ref("recv", ref, seq_id=f"{upstream_seq_id}->{curr_seq_id}", src=src_party)
    return ref


def recv_many(party: str, items, curr_seq_id):
    """`recv` for several (src_party, upstream_seq_id) pairs that feed one task: one call into the
    transport instead of one per argument.  Returns the list of `Ref`s."""
    host: ProxyHost = _HOSTS["receiver"] or get_runtime().get_named(receiver_proxy_actor_name())
    refs = host.get_data_many(items, curr_seq_id) if _tracing._TRACER is None else None
    if refs is None:
        return [recv(party, src, up, curr_seq_id) for src, up in items]
    return refs


def pending_summary(limit: int = 16):
    """(waiting, unclaimed) message keys of this party's receiver — ``waiting``: a consumer asked,
    nothing arrived yet; ``unclaimed``: arrived, nobody asked.  Used by the stall diagnostic of
    ``fed.get`` (controller-divergence detection)."""
    try:
        host = _HOSTS["receiver"] or get_runtime().get_named(receiver_proxy_actor_name())
    except Exception:  # noqa: BLE001
        return [], []
    mailbox = getattr(host.proxy, "_mailbox", None)
    if mailbox is None or not hasattr(mailbox, "snapshot"):
        return [], []
    waiting, unclaimed = mailbox.snapshot(limit)
    unclaimed = [k for k in unclaimed if k != ("ping", "ping") and not k[0].startswith("__")]
    return list(waiting), unclaimed


def ping_others(addresses: Dict[str, Dict], self_party: str, max_retries=3600):
    """Init barrier: ping every peer until each one has acknowledged.

    Same contract as fed/proxy/barriers.py:497-523 (raises ``RuntimeError``
    after ``max_retries`` rounds) but event-driven: a round waits on the ack
    futures (≤1 s) and the inter-round sleep backs off from 50 ms to 1 s instead
    of a fixed 1 s + 2 s cadence.
    """
    others = [p for p in addresses if p != self_party]
    tried = 0
    pause = 0.05
    from ..core import wait as wait_refs

    while tried < max_retries and others:
        logger.info(f"Try ping {others} at {tried} attemp, up to {max_retries} attemps.")
        tried += 1
        pings = {o: send(o, b"data", "ping", "ping", check_sending=False) for o in others}
        wait_refs(list(pings.values()), num_returns=len(pings), timeout=1)
        still = []
        for o, r in pings.items():
            ok = False
            if r.done():
                try:
                    ok = bool(r.result())
                except Exception:  # noqa: BLE001
                    ok = False
            if not ok:
                still.append(o)
        others = still
        if others:
            time.sleep(pause)
            pause = min(pause * 2, 1.0)
    if others:
        raise RuntimeError(f"Failed to wait for parties: {others} to start, abort `fed.init`.")
    return True
