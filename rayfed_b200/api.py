"""Public API: ``init / shutdown / remote / get / kill``.

Parity: fed/api.py (``init`` :67-296, ``shutdown``/``_shutdown`` :299-360,
``FedRemoteFunction`` :384-416, ``FedRemoteClass`` :419-448, ``remote``
:452-528, ``get`` :531-608, ``kill`` :611-623, SIGINT handler :56-64).

Every party runs the same script (multi-controller).  ``init`` brings up the
party's runtime and its transport; there is no Ray cluster to attach to.  New
relative to the reference:

* ``config["cross_silo_comm"]["transport"]``: ``auto`` (default; native framed
  sockets, plus the NVLink device path when the party has a CUDA device and
  peers share the box), ``native``, ``nvlink``, or ``grpc`` (wire-compatible
  with the reference; chosen automatically when ``tls_config`` is given).
* ``config["device"]``: the CUDA device of this party (e.g. ``"cuda:3"``).
"""
from __future__ import annotations

import functools
import inspect
import logging
import os
import pickle
import signal
import sys
import threading
from typing import Any, Callable, Dict, List, Union

from . import _tracing
from . import config as fed_config
from . import utils as fed_utils
from ._private import constants
from ._private import kv as fed_kv
from ._private.fed_actor import FedActorHandle
from ._private.fed_call_holder import FedCallHolder
from ._private.global_context import clear_global_context, get_global_context, init_global_context
from .config import CrossSiloMessageConfig
from .core import get as _local_get
from .core import get_runtime, is_refs, shutdown_runtime
from .exceptions import FedGetTimeoutError, FedRemoteError, FedTaskError
from .fed_object import FedObject
from .proxy import barriers
from .proxy.barriers import (
    _start_receiver_proxy,
    _start_sender_proxy,
    _start_sender_receiver_proxy,
    ping_others,
    recv,
    send,
    set_proxy_actor_name,
)
from .proxy.base_proxy import ReceiverProxy, SenderProxy, SenderReceiverProxy
from .utils import setup_logger

logger = logging.getLogger(__name__)

original_sigint = signal.getsignal(signal.SIGINT)


def _signal_handler(signum, frame):
    if signum == signal.SIGINT:
        signal.signal(signal.SIGINT, original_sigint)
        logger.warning(
            "Stop signal received (e.g. via SIGINT/Ctrl+C), try to shutdown fed. "
            "Press CTRL+C (or send SIGINT/SIGKILL/SIGTERM) to skip."
        )
        _shutdown(intended=False)


def _transport_choice(cross_silo_comm_dict: Dict) -> str:
    choice = (cross_silo_comm_dict.get("transport") or constants.TRANSPORT_AUTO).lower()
    if choice not in (
        constants.TRANSPORT_AUTO, constants.TRANSPORT_NATIVE, constants.TRANSPORT_GRPC, constants.TRANSPORT_NVLINK
    ):
        raise ValueError(f"Unknown cross_silo_comm.transport '{choice}'.")
    return choice


def _pick_transport(cross_silo_comm_dict: Dict, tls_config: Dict) -> str:
    choice = _transport_choice(cross_silo_comm_dict)
    if choice == constants.TRANSPORT_AUTO:
        # TLS and L7-gateway style addresses are only spoken by the gRPC transport.
        return constants.TRANSPORT_GRPC if tls_config else constants.TRANSPORT_NATIVE
    return choice


def init(
    addresses: Dict = None,
    party: str = None,
    config: Dict = {},
    tls_config: Dict = None,
    logging_level: str = "info",
    sender_proxy_cls: SenderProxy = None,
    receiver_proxy_cls: ReceiverProxy = None,
    receiver_sender_proxy_cls: SenderReceiverProxy = None,
    job_name: str = None,
    sending_failure_handler: Callable[[Exception], None] = None,
):
    """Initialise this party.

    Args:
        addresses: ``{party: "ip:port" | "host:port" | "http(s)://..." }`` —
            where each party can be reached by the others.
        party: this controller's party name; must be a key of ``addresses``.
        config: ``{"cross_silo_comm": {...}, "barrier_on_initializing": bool,
            "device": "cuda:N"}``; see :class:`CrossSiloMessageConfig`.
        tls_config: ``{"ca_cert", "cert", "key"}`` file paths (mutual TLS when
            ``ca_cert`` is present); selects the gRPC transport.
        logging_level: debug / info / warning / error / critical.
        sender_proxy_cls, receiver_proxy_cls, receiver_sender_proxy_cls:
            custom transport classes (see :mod:`rayfed_b200.proxy.base_proxy`).
        job_name: must be identical in all parties; messages of other jobs are
            refused. Defaults to ``"Anonymous_job"``.
        sending_failure_handler: called with the last sending error when a
            cross-party send failed and ``exit_on_sending_failure`` is set.
    """
    assert addresses, "Addresses should be provided."
    assert party, "Party should be provided."
    assert party in addresses, f"Party {party} is not in the addresses {addresses}."

    if job_name is None:
        job_name = constants.RAYFED_DEFAULT_JOB_NAME

    fed_utils.validate_addresses(addresses)

    config = config or {}
    cross_silo_comm_dict = config.get("cross_silo_comm", {}) or {}
    cross_silo_comm_config = CrossSiloMessageConfig.from_dict(cross_silo_comm_dict)

    init_global_context(
        current_party=party,
        job_name=job_name,
        exit_on_sending_failure=cross_silo_comm_config.exit_on_sending_failure,
        continue_waiting_for_data_sending_on_error=cross_silo_comm_config.continue_waiting_for_data_sending_on_error,
        sending_failure_handler=sending_failure_handler,
    )

    tls_config = {} if tls_config is None else tls_config
    if tls_config:
        assert "cert" in tls_config and "key" in tls_config, "Cert or key are not in tls_config."

    _configure_spinning(addresses, party)
    fed_kv._init_internal_kv(job_name)
    cluster_config = {
        constants.KEY_OF_CLUSTER_ADDRESSES: addresses,
        constants.KEY_OF_CURRENT_PARTY_NAME: party,
        constants.KEY_OF_TLS_CONFIG: tls_config,
    }
    fed_kv.kv.put(constants.KEY_OF_CLUSTER_CONFIG, pickle.dumps(cluster_config))
    job_config = {constants.KEY_OF_CROSS_SILO_COMM_CONFIG_DICT: cross_silo_comm_dict}
    fed_kv.kv.put(constants.KEY_OF_JOB_CONFIG, pickle.dumps(job_config))

    setup_logger(
        logging_level=logging_level,
        logging_format=constants.RAYFED_LOG_FMT,
        date_format=constants.RAYFED_DATE_FMT,
        party_val=_get_party(job_name),
        job_name=job_name,
    )
    trace_prefix = config.get("trace") or os.environ.get("RAYFED_B200_TRACE")
    if trace_prefix:
        _tracing.enable(party, str(trace_prefix))
    metrics_port = config.get("metrics_port", os.environ.get("RAYFED_B200_METRICS_PORT"))
    if metrics_port not in (None, ""):
        from .utils import metrics as _metrics

        _metrics.start_exporter(int(metrics_port), party=party, job=job_name,
                                addr=str(config.get("metrics_addr", "127.0.0.1")))
    logger.info(f"Started rayfed_b200 with {cluster_config}")

    if threading.current_thread() is threading.main_thread():
        signal.signal(signal.SIGINT, _signal_handler)

    runtime = get_runtime()
    device = config.get("device")
    if device is not None:
        runtime.set_device(str(device))

    get_global_context().get_cleanup_manager().start(
        exit_on_sending_failure=cross_silo_comm_config.exit_on_sending_failure,
        expose_error_trace=cross_silo_comm_config.expose_error_trace,
    )

    ready_timeout = cross_silo_comm_config.timeout_in_ms / 1000
    use_global_proxy = cross_silo_comm_dict.get("use_global_proxy", True)
    if receiver_sender_proxy_cls is not None:
        set_proxy_actor_name(job_name, use_global_proxy, True)
        _start_sender_receiver_proxy(
            addresses=addresses, party=party, logging_level=logging_level, tls_config=tls_config,
            proxy_cls=receiver_sender_proxy_cls, proxy_config=cross_silo_comm_dict,
            ready_timeout_second=ready_timeout,
        )
    else:
        transport = _pick_transport(cross_silo_comm_dict, tls_config)
        if receiver_proxy_cls is None:
            if transport == constants.TRANSPORT_GRPC:
                from .proxy.grpc.grpc_proxy import GrpcReceiverProxy as receiver_proxy_cls
            elif _device_path_wanted(cross_silo_comm_dict, tls_config, runtime):
                from .proxy.nvlink.nvlink_proxy import NvlinkReceiverProxy as receiver_proxy_cls
            else:
                from .proxy.native.socket_proxy import NativeReceiverProxy as receiver_proxy_cls
        set_proxy_actor_name(job_name, use_global_proxy)
        _start_receiver_proxy(
            addresses=addresses, party=party, logging_level=logging_level, tls_config=tls_config,
            proxy_cls=receiver_proxy_cls, proxy_config=cross_silo_comm_dict, ready_timeout_second=ready_timeout,
        )
        if sender_proxy_cls is None:
            if transport == constants.TRANSPORT_GRPC:
                from .proxy.grpc.grpc_proxy import GrpcSenderProxy as sender_proxy_cls
            elif _device_path_wanted(cross_silo_comm_dict, tls_config, runtime):
                from .proxy.nvlink.nvlink_proxy import NvlinkSenderProxy as sender_proxy_cls
            else:
                from .proxy.native.socket_proxy import NativeSenderProxy as sender_proxy_cls
        _start_sender_proxy(
            addresses=addresses, party=party, logging_level=logging_level, tls_config=tls_config,
            proxy_cls=sender_proxy_cls, proxy_config=cross_silo_comm_dict, ready_timeout_second=ready_timeout,
        )

    if receiver_sender_proxy_cls is None and _device_path_negotiable(cross_silo_comm_dict, tls_config):
        _maybe_start_device_channel(addresses, party, cross_silo_comm_config, runtime, job_name)

    if config.get("barrier_on_initializing", False):
        ping_others(addresses=addresses, self_party=party, max_retries=3600)


def _local_party_count(addresses: Dict, party: str) -> int:
    """Parties of this job whose address is on this party's host (they share its CPUs)."""
    def host_of(addr: str) -> str:
        addr = str(addr)
        if addr.startswith("unix:") or "/" in addr:
            return "localhost"
        host = addr.rsplit(":", 1)[0].strip("[]").lower()
        return "localhost" if host in ("127.0.0.1", "::1", "localhost", "0.0.0.0", "") else host

    mine = host_of(addresses.get(party, ""))
    return max(1, sum(1 for a in addresses.values() if host_of(a) == mine))


def _configure_spinning(addresses: Dict, party: str) -> None:
    """Adaptive spinning (csrc/core/ref.h) needs CPUs to burn: give every party on this host an
    equal share of half the usable CPUs, and leave it off below two spinners per party.
    `RAYFED_B200_ADAPTIVE_SPIN_US` / `RAYFED_B200_MAX_SPINNERS` override."""
    try:
        from .core import native

        native().set_spin_policy_auto(_local_party_count(addresses, party))
        logger.debug(f"spin policy: {native().spin_info()}")
    except Exception:  # noqa: BLE001  (a tuning knob must never fail init)
        logger.debug("could not configure adaptive spinning", exc_info=True)


def _host_identity() -> str:
    """Same string in every process of one OS instance (containers on a box share the boot id)."""
    import socket

    boot = ""
    try:
        with open("/proc/sys/kernel/random/boot_id") as f:
            boot = f.read().strip()
    except OSError:
        pass
    return f"{boot}/{socket.gethostname()}" if not os.environ.get("RAYFED_B200_HOST_ID") else os.environ[
        "RAYFED_B200_HOST_ID"]


def _maybe_start_device_channel(addresses, party, cfg: CrossSiloMessageConfig, runtime, job_name: str):
    """Negotiate, then bring up, the symmetric heap + DeviceChannel.

    Collective: every party whose transport choice is ``auto``/``nvlink`` reaches this point in
    ``init`` — whether or not it has a device — and all-gathers one capability record over the host
    transport that was just started (connect retries cover parties that start late).  The device
    path is used only if EVERY party advertises a CUDA device, the NVLink proxies and the same
    host; otherwise all of them stay on the host path (a CPU-only or off-box peer cannot map a
    CUDA IPC/VMM handle).  Reference behaviour being replaced: one gRPC channel per peer,
    fed/proxy/grpc/grpc_proxy.py:116-141."""
    import json

    if len(addresses) < 2:
        return
    rt = get_runtime()
    recv_host = rt.get_named(barriers.receiver_proxy_actor_name())
    send_host = rt.get_named(barriers.sender_proxy_actor_name())
    from .proxy.native.socket_proxy import NativeReceiverProxy, NativeSenderProxy

    if not (isinstance(recv_host.proxy, NativeReceiverProxy) and isinstance(send_host.proxy, NativeSenderProxy)):
        return  # user-supplied proxies: their transport, their rules
    start_unpackers = getattr(recv_host.proxy, "start_unpackers", None)
    has_device = runtime.device is not None and str(runtime.device).startswith("cuda")
    from .parallel.group import FedExchange, PartyGroup

    dev_index = 0
    if has_device:
        import torch

        dev = torch.device(runtime.device)
        dev_index = dev.index or 0
    mine = {"device": bool(has_device and start_unpackers is not None), "host": _host_identity()}
    blob = json.dumps(mine).encode()
    # Publish my capabilities to every peer with plain proxy sends (not fed sends: the op counters
    # and the send-checking queue stay untouched).  A party WITHOUT a device stops here — it never
    # blocks in init for a negotiation it has no stake in; parties with a device wait for everyone.
    for p in addresses:
        if p != party:
            try:
                send_host.proxy.send(p, blob, f"__caps__/{party}", "__caps__")
            except Exception as e:  # noqa: BLE001
                logger.debug(f"capability publish to {p} failed: {e!r}")
    if not mine["device"]:
        return
    wait_s = (cfg.timeout_in_ms or 60000) / 1000.0 * 2
    caps = []
    try:
        for p in addresses:
            if p == party:
                caps.append(mine)
            else:
                raw = recv_host.proxy.get_data(p, f"__caps__/{p}", "__caps__").result(wait_s)
                caps.append(json.loads(raw))
    except Exception as e:  # noqa: BLE001  (a peer that was configured host-only never answers)
        logger.warning(f"device-path negotiation failed ({e!r}); staying on the host transport.")
        return
    group = PartyGroup(list(addresses.keys()), party, FedExchange(), device_index=dev_index)
    agreed = all(c.get("device") for c in caps) and len({c.get("host") for c in caps}) == 1
    if not agreed:
        if has_device:
            logger.info("NVLink device path not used: "
                        + ("a peer has no CUDA device" if not all(c.get("device") for c in caps)
                           else "the parties are not on one host") + "; CUDA tensors ride the host transport.")
        if _transport_choice({"transport": cfg.transport}) == constants.TRANSPORT_NVLINK:
            raise RuntimeError("cross_silo_comm.transport='nvlink' needs every party on this box with a CUDA device")
        return
    import torch

    from .proxy.nvlink import nvlink_proxy
    from .proxy.nvlink.channel import DeviceChannel

    torch.cuda.set_device(torch.device(runtime.device))
    channel = DeviceChannel(
        group,
        region_bytes=cfg.device_arena_bytes or (256 << 20),
        wire_dtype=cfg.device_wire_dtype,
        timeout_s=(cfg.timeout_in_ms or 60000) / 1000.0,
        job_name=job_name,
        zero_copy=True if cfg.device_zero_copy is None else bool(cfg.device_zero_copy),
        zero_copy_min_bytes=cfg.device_zero_copy_min_bytes if cfg.device_zero_copy_min_bytes is not None else 64 << 10,
        consumer_streams=runtime.device_streams,
        wait_timeout_s=(cfg.device_wait_timeout_ms / 1000.0) if cfg.device_wait_timeout_ms else None,
    )
    nvlink_proxy.attach_channel(channel)
    start_unpackers(channel)
    logger.info(
        f"NVLink device channel up: {group.size} parties, {channel.region_bytes >> 20} MiB/region, "
        f"backend={channel.buf.backend}, multicast={channel.can_multicast}"
    )


def _device_path_negotiable(cross_silo_comm_dict: Dict, tls_config) -> bool:
    """``auto`` (without TLS) and ``nvlink`` negotiate the device path; ``native`` and ``grpc`` are
    host-only by definition."""
    choice = _transport_choice(cross_silo_comm_dict)
    if choice == constants.TRANSPORT_NVLINK:
        return True
    return choice == constants.TRANSPORT_AUTO and not tls_config


def _device_path_wanted(cross_silo_comm_dict: Dict, tls_config, runtime) -> bool:
    """Use the NVLink proxies (a superset of the native ones: without a channel they send plain
    host frames) when this party has a CUDA device and the device path is negotiable."""
    return (_device_path_negotiable(cross_silo_comm_dict, tls_config) and runtime.device is not None
            and str(runtime.device).startswith("cuda"))


def shutdown():
    """Shut this party down (idempotent; safe before/without ``init``)."""
    global_context = get_global_context()
    if global_context is not None and global_context.acquire_shutdown_flag():
        _shutdown(True)


def _shutdown(intended=True):
    """``intended=False`` is the failure path: run the failure handler, do not
    wait for pending sends, and exit(1) when ``exit_on_sending_failure``."""
    global_context = get_global_context()
    if global_context is None:
        return  # never initialised, or already cleaned

    if intended:
        logger.info("Shutdowning rayfed_b200 intendedly...")
    else:
        logger.warning("Shutdowning rayfed_b200 unintendedly...")
    last_sending_error = global_context.get_cleanup_manager().get_last_sending_error()
    last_received_error = global_context.get_last_recevied_error()
    if last_sending_error is not None:
        logger.error(f"Cross-silo sending error occured. {last_sending_error}")

    wait_for_sending = True
    if (
        last_sending_error is not None or last_received_error is not None
    ) and not global_context.get_continue_waiting_for_data_sending_on_error():
        wait_for_sending = False
    logger.info(f'{"Wait" if wait_for_sending else "No wait"} for data sending.')

    exit_on_sending_failure = global_context.get_exit_on_sending_failure()
    if not intended:
        failure_handler = global_context.get_sending_failure_handler()
        if failure_handler is not None:
            logger.info(f"Executing failure handler {failure_handler} ...")
            failure_handler(last_sending_error)

    job_name_at_exit = global_context.get_job_name()
    fed_kv._clear_internal_kv()
    clear_global_context(wait_for_sending=wait_for_sending)
    coll = sys.modules.get("rayfed_b200.collective")
    if coll is not None:
        coll.reset(job_name_at_exit)
    _teardown_runtime(wait_for_sending)
    metrics_mod = sys.modules.get("rayfed_b200.utils.metrics")
    if metrics_mod is not None:
        metrics_mod.stop_exporter()
    try:
        from .core import native

        native().set_spin_policy(-1, -1)  # back to "not configured" (off)
    except Exception:  # noqa: BLE001
        pass
    trace_file = _tracing.disable(dump=True)
    if trace_file:
        logger.info(f"Wrote the fed timeline to {trace_file}.")
    logger.info("Shutdowned rayfed_b200.")

    if not intended and exit_on_sending_failure:
        logger.critical("Exit now due to the previous error.")
        sys.exit(1)


def _teardown_runtime(wait_for_sending: bool):
    """What ``ray.shutdown()`` did for the reference: stop proxies, then the
    scheduler.  Pending acks were already drained by the CleanupManager."""
    try:
        barriers._stop_proxies(drain=wait_for_sending)
        try:
            from .proxy.nvlink import nvlink_proxy

            ch = nvlink_proxy.current_channel()
            if ch is not None:
                ch.synchronize()
                nvlink_proxy.detach_channel()
        except ImportError:
            pass
    finally:
        shutdown_runtime(wait=True)


def _get_addresses(job_name: str = None):
    return fed_config.get_cluster_config(job_name).cluster_addresses


def _get_party(job_name: str = None):
    return fed_config.get_cluster_config(job_name).current_party


def _get_tls(job_name: str = None):
    return fed_config.get_cluster_config(job_name).tls_config


class FedRemoteFunction:
    def __init__(self, func_or_class) -> None:
        self._node_party = None
        self._func_body = func_or_class
        self._options = {}
        self._fed_call_holder = None

    def party(self, party: str):
        # Mutates and returns self, like the reference (api.py:391-399).
        self._node_party = party
        self._fed_call_holder = FedCallHolder(self._node_party, self._execute_impl, self._options)
        return self

    def options(self, **options):
        self._options = options
        if self._fed_call_holder:
            self._fed_call_holder.options(**options)
        return self

    def remote(self, *args, **kwargs):
        if not self._node_party:
            raise ValueError("You should specify a party name on the fed function.")
        return self._fed_call_holder.internal_remote(*args, **kwargs)

    def _execute_impl(self, args, kwargs):
        num_returns = (self._options or {}).get("num_returns", 1) or 1
        return get_runtime().submit(self._func_body, args, kwargs, num_returns)


class FedRemoteClass:
    def __init__(self, func_or_class) -> None:
        self._party = None
        self._cls = func_or_class
        self._options = {}

    def party(self, party: str):
        self._party = party
        return self

    def options(self, **options):
        self._options = options
        return self

    def remote(self, *cls_args, **cls_kwargs):
        ctx = get_global_context()
        fed_class_task_id = ctx.next_seq_id()  # actor creation burns two ids, like the reference
        job_name = ctx.get_job_name()
        fed_actor_handle = FedActorHandle(
            fed_class_task_id, _get_addresses(job_name), self._cls, _get_party(job_name), self._party,
            self._options,
        )
        fed_call_holder = FedCallHolder(self._party, fed_actor_handle._execute_impl, self._options)
        fed_call_holder.internal_remote(*cls_args, **cls_kwargs)
        return fed_actor_handle


def remote(*args, **kwargs):
    """``@fed.remote`` — make a function or class placeable with ``.party(p)``.

    .. code:: python

        @fed.remote
        def f(a, b): return a + b
        obj = f.party("alice").remote(1, 2)

        @fed.remote
        class Counter:
            def __init__(self, v): self.v = v
            def inc(self, d): self.v += d; return self.v
        c = Counter.party("bob").remote(0)
        fed.get(c.inc.remote(obj))      # -> 3, in every party

    Options (``@fed.remote(num_returns=2)`` or ``.options(num_returns=2)``)
    replace — not merge — any earlier options.
    """

    def _make_fed_remote(function_or_class, **options):
        if inspect.isfunction(function_or_class) or fed_utils.is_cython(function_or_class):
            return FedRemoteFunction(function_or_class).options(**options)
        if inspect.isclass(function_or_class):
            return FedRemoteClass(function_or_class).options(**options)
        raise TypeError("The @fed.remote decorator must be applied to either a function or a class.")

    if len(args) == 1 and len(kwargs) == 0 and callable(args[0]):
        return _make_fed_remote(args[0])
    assert len(args) == 0 and len(kwargs) > 0, "Remote args error."
    return functools.partial(_make_fed_remote, **kwargs)


def get(fed_objects: Union[FedObject, List[FedObject], Any]) -> Any:
    """Fetch the value(s) behind FedObject(s) — a *collective*: every party calls
    it at the same program point, the owner pushes the value to all other
    parties (api.py:564-579), everyone returns it.  Local futures (`Ref`) pass
    straight through."""
    if is_refs(fed_objects):
        return _local_get(fed_objects)

    ctx = get_global_context()
    fake_fed_task_id = ctx.next_seq_id()  # fed.get is a DAG node too
    tl = _tracing._TRACER
    if tl is not None:
        with tl.span("fed.get", seq_id=fake_fed_task_id):
            return _get(ctx, fed_objects, fake_fed_task_id)
    return _get(ctx, fed_objects, fake_fed_task_id)


def _get(ctx, fed_objects, fake_fed_task_id):
    job_name = ctx.get_job_name()
    addresses = _get_addresses(job_name)
    current_party = ctx.get_current_party()
    is_individual_id = isinstance(fed_objects, FedObject)
    if is_individual_id:
        fed_objects = [fed_objects]

    refs = []
    for fed_object in fed_objects:
        if fed_object.get_party() == current_party:
            ref = fed_object.get_ref()
            assert ref is not None
            refs.append(ref)
            todo = [p for p in addresses
                    if p != current_party and not fed_object._was_sending_or_sent_to_party(p)]
            for party_name in todo:
                fed_object._mark_is_sending_to_party(party_name)
            if len(todo) == len(addresses) - 1 and len(todo) >= 1:
                # nobody has it yet: one broadcast (a single multicast push where the
                # transport supports it, N-1 unicast sends otherwise)
                barriers.send_broadcast(todo, ref, fed_object.get_fed_task_id(), fake_fed_task_id)
            else:
                for party_name in todo:
                    send(
                        dest_party=party_name,
                        data=ref,
                        upstream_seq_id=fed_object.get_fed_task_id(),
                        downstream_seq_id=fake_fed_task_id,
                    )
        else:
            ref = fed_object.get_ref()
            if ref is None:
                ref = recv(current_party, fed_object.get_party(), fed_object.get_fed_task_id(), fake_fed_task_id)
                fed_object._cache_ref(ref)
            refs.append(ref)

    try:
        values = _get_with_stall_diagnostic(refs, fake_fed_task_id)
        _raise_on_device_failure()
        return values[0] if is_individual_id else values
    except FedTaskError as e:
        if isinstance(e.cause, FedRemoteError):
            logger.warning(f"Encounter RemoteError happend in other parties, error message: {e.cause.cause}")
            if get_global_context() is not None:
                get_global_context().set_last_recevied_error(e.cause)
        raise


_STALL_WARN_S = float(os.environ.get("RAYFED_B200_STALL_WARN_S", "30"))


def _get_with_stall_diagnostic(refs, fed_task_id):
    """``_local_get`` that never gives up but says *why* it may be stuck: a receiver cannot tell
    a late message from one that will never come, and the reference simply hangs when two
    controllers disagree on the number of fed calls.  Every ``RAYFED_B200_STALL_WARN_S`` seconds
    (0 disables) of blocking we log which keys are awaited and which messages arrived that no
    consumer ever asked for — the latter is the signature of diverged scripts."""
    if _STALL_WARN_S <= 0:
        return _local_get(refs)
    waited = 0.0
    while True:
        try:
            return _local_get(refs, timeout=_STALL_WARN_S)
        except FedGetTimeoutError:
            waited += _STALL_WARN_S
            waiting, unclaimed = barriers.pending_summary()
            msg = (f"fed.get (seq id {fed_task_id}) has been blocked for {waited:.0f}s; awaited message keys "
                   f"(upstream, downstream): {waiting or 'none'}.")
            if unclaimed:
                msg += (f" Messages that arrived but that no task or fed.get of this party asked for: {unclaimed}. "
                        "If these do not get consumed, the parties' driver scripts have most likely diverged "
                        "(a different number of .remote()/fed.get calls was issued in some party).")
            logger.warning(msg)


def _raise_on_device_failure():
    """A device-side wait that timed out leaves its destination untouched and sets a
    status word; surface that as an error instead of handing out garbage."""
    mod = sys.modules.get("rayfed_b200.proxy.nvlink.nvlink_proxy")
    if mod is not None:
        ch = mod.current_channel()
        if ch is not None:
            ch.raise_if_failed()


def kill(actor: FedActorHandle, *, no_restart=True):
    """Forcefully stop an actor — acts only in the owner party (api.py:619-623)."""
    ctx = get_global_context()
    current_party = ctx.get_current_party()
    if actor._node_party == current_party and actor._actor_handle is not None:
        actor._actor_handle.kill()
