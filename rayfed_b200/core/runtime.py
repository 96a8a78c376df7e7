"""PartyRuntime — one per controller process.

What Ray supplied to the reference, and where it lives now:

===========================  =================================================
Ray facility (reference use)  here
===========================  =================================================
``ray.remote(f).remote()``    :meth:`PartyRuntime.submit` → C++ ``Executor`` pool
  (fed/api.py:413-416)          lane; `Ref` args gate the task (no polling)
actors, ordered mailbox,      :meth:`PartyRuntime.create_actor` → C++ ordered lane
  ``max_concurrency``           (+ a dedicated CUDA stream when a device is set)
  (fed/_private/fed_actor.py)
``ray.get`` / ``ray.wait``    :func:`get` / :func:`wait` (GIL-free waits)
``ray.kill``                  :meth:`ActorHandle.kill` → mailbox dropped
named actors                  :meth:`PartyRuntime.register_named` / ``get_named``
===========================  =================================================

CUDA semantics live in the C++ scheduler, not here: a worker
thread of a device party owns a CUDA stream; the executor records a CUDA event
after each task body and attaches it to the output `Ref`s; before a consumer's
body runs — and before the GIL is taken — the executor makes the consumer's
stream wait for the events of its inputs (``cudaStreamWaitEvent``).  The host
never blocks for a device→device dependency, so local tasks overlap in-flight
pushes (``csrc/core/executor.h``, ``csrc/core/device_hooks.h``).
"""
from __future__ import annotations

import atexit
import importlib
import importlib.util
import os
import sys
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence

from .. import _build, _tracing
from ..exceptions import FedActorDiedError, FedGetTimeoutError, FedTaskError

_native = None
_native_lock = threading.Lock()


def native():
    """Import (building on first use) the ``_fedcore`` extension."""
    global _native
    if _native is None:
        with _native_lock:
            if _native is None:
                alt = os.environ.get("RAYFED_B200_CORE_SO")
                if alt:
                    # an instrumented build of the same sources (tools/sanitize_core.sh: TSAN / ASAN)
                    spec = importlib.util.spec_from_file_location("rayfed_b200._fedcore", alt)
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    sys.modules["rayfed_b200._fedcore"] = mod
                    _native = mod
                else:
                    _build.ensure_core()
                    _native = importlib.import_module("rayfed_b200._fedcore")
    return _native


_REF_T = None


def _ref_type():
    global _REF_T
    if _REF_T is None:
        _REF_T = native().Ref
    return _REF_T


class _RefMeta(type):
    def __instancecheck__(cls, obj):
        return isinstance(obj, _ref_type())


class Ref(metaclass=_RefMeta):
    """``isinstance(x, Ref)`` works against the native future type; calling
    ``Ref()`` creates a fresh unresolved native future."""

    def __new__(cls):
        return _ref_type()()


def is_ref(obj) -> bool:
    return type(obj) is (_REF_T or _ref_type())


def is_refs(objs) -> bool:
    """True for a Ref or a non-empty list made only of Refs
    (parity: fed/utils.py:86-96 ``is_ray_object_refs``)."""
    if is_ref(objs):
        return True
    if isinstance(objs, (list, tuple)) and objs:
        return all(is_ref(o) for o in objs)
    return False


def DeviceValue(value, event: int = 0, stream: int = 0, owner=None):
    """``value`` + the CUDA event (raw handle) that orders it after its producer — what a
    device-aware transport's decoder returns and what the sender hand-off receives.  A `Ref`
    completed with one keeps the plain value and adopts the event (csrc/core/ref.h)."""
    return native().DeviceValue(value, int(event), int(stream), owner)


def is_device_value(obj) -> bool:
    return type(obj) is native().DeviceValue


def _wrap_task_error(exc: BaseException):
    return FedTaskError.from_exception(exc)


class ActorHandle:
    """Local actor: an instance pinned to one ordered lane."""

    def __init__(self, runtime: "PartyRuntime", lane_id: int, name: str):
        self._runtime = runtime
        self._lane_id = lane_id
        self._name = name
        self._instance = None
        self._dead = False
        self._ready_ref = None
        self._callers: Dict[str, Callable] = {}

    @property
    def name(self) -> str:
        return self._name

    def submit_method(self, method_name: str, args: tuple, kwargs: dict, num_returns: int = 1):
        call = self._callers.get(method_name)
        if call is None:  # one trampoline per method, not per call
            def call(*a, **kw):
                return getattr(self._instance, method_name)(*a, **kw)

            call.__name__ = f"{self._name}.{method_name}"
            self._callers[method_name] = call
        return self._runtime._submit(self._lane_id, call, args, kwargs, num_returns)

    def submit_call(self, fn: Callable, args: tuple, kwargs: dict, num_returns: int = 1):
        """Run ``fn(instance, *args)`` in mailbox order (used by proxy hosts)."""

        def call(*a, **kw):
            return fn(self._instance, *a, **kw)

        return self._runtime._submit(self._lane_id, call, args, kwargs, num_returns)

    def kill(self):
        if not self._dead:
            self._dead = True
            self._runtime._executor.kill_lane(
                self._lane_id, FedActorDiedError(f"The actor {self._name} was killed by fed.kill().")
            )

    @property
    def instance(self):
        return self._instance

    def backlog(self) -> int:
        return self._runtime._executor.lane_backlog(self._lane_id)


class PartyRuntime:
    def __init__(self, pool_threads: Optional[int] = None, device: Optional[str] = None):
        nat = native()
        if pool_threads is None:
            pool_threads = int(os.environ.get("RAYFED_B200_POOL_THREADS", "4"))
        self._executor = nat.Executor(pool_threads, _wrap_task_error)
        self._device = device
        self._named: Dict[str, Any] = {}
        self._named_lock = threading.Lock()
        self._lane_streams = []
        self._actor_count = 0
        self._closed = False

    # -- configuration ---------------------------------------------------------
    @property
    def device(self) -> Optional[str]:
        return self._device

    def set_device(self, device: Optional[str]):
        """Bind this party to a CUDA device.  From here on every worker thread owns a CUDA
        stream, every task result carries the event recorded after its body, and the C++
        executor orders consumer streams after producer events (``cudaStreamWaitEvent``,
        issued before the GIL is taken) — the host never blocks for a device->device edge."""
        self._device = device
        if device is not None and str(device).startswith("cuda"):
            from ..ops.kernels import cuda_ext

            native().set_device_hooks(cuda_ext().device_hooks())
            self._executor.set_lane_thread_init(0, self._thread_init())
        else:
            self._executor.set_lane_thread_init(0, None)

    def _thread_init(self):
        """Runs once on each worker thread: device + a stream of its own as the thread's current
        stream; returns the raw stream handle to the executor."""
        if self._device is None or not str(self._device).startswith("cuda"):
            return None
        device = self._device

        def init():
            import torch

            dev = torch.device(device)
            torch.cuda.set_device(dev)
            stream = torch.cuda.Stream(device=dev)
            torch.cuda.set_stream(stream)
            with self._named_lock:
                self._lane_streams.append(stream)
            return stream.cuda_stream

        return init

    def device_streams(self) -> List[int]:
        """Raw handles of every stream the runtime's workers run tasks on."""
        with self._named_lock:
            return [s.cuda_stream for s in self._lane_streams]

    # -- tasks -----------------------------------------------------------------
    def _submit(self, lane_id, fn, args, kwargs, num_returns):
        if _tracing._TRACER is not None:
            fn = _tracing.traced_task(fn, getattr(fn, "__name__", "task"))
        refs = self._executor.submit(lane_id, fn, tuple(args), kwargs or None, num_returns)
        return refs[0] if num_returns <= 1 else list(refs)

    def submit(self, fn: Callable, args: Sequence = (), kwargs: Optional[dict] = None, num_returns: int = 1):
        """Stateless task: runs on the shared pool once its `Ref` args resolve."""
        return self._submit(0, fn, args, kwargs, num_returns)

    # -- actors ----------------------------------------------------------------
    def create_actor(self, cls, args: Sequence = (), kwargs: Optional[dict] = None, *, name: str = "",
                     max_concurrency: int = 1) -> ActorHandle:
        self._actor_count += 1
        if not name:
            name = f"{getattr(cls, '__name__', 'actor')}#{self._actor_count}"
        lane_id = self._executor.create_lane(name, max_concurrency or 1, self._thread_init())
        handle = ActorHandle(self, lane_id, name)

        def construct(*a, **kw):
            handle._instance = cls(*a, **kw)
            return True

        construct.__name__ = f"{name}.__init__"
        handle._ready_ref = self._submit(lane_id, construct, args, kwargs, 1)
        return handle

    # -- named registry (ray.get_actor analogue) -------------------------------
    def register_named(self, name: str, obj) -> None:
        with self._named_lock:
            if name in self._named:
                raise ValueError(f"An actor named '{name}' already exists in this party runtime.")
            self._named[name] = obj

    def get_named(self, name: str):
        with self._named_lock:
            if name not in self._named:
                raise ValueError(f"Failed to look up actor with name '{name}'.")
            return self._named[name]

    def has_named(self, name: str) -> bool:
        with self._named_lock:
            return name in self._named

    def unregister_named(self, name: str):
        with self._named_lock:
            return self._named.pop(name, None)

    def named(self) -> List[str]:
        with self._named_lock:
            return list(self._named)

    # -- lifecycle -------------------------------------------------------------
    def stats(self) -> dict:
        return dict(self._executor.stats())

    def shutdown(self, wait: bool = True, timeout: float = 5.0):
        if self._closed:
            return
        self._closed = True
        with self._named_lock:
            named = list(self._named.values())
            self._named.clear()
        for obj in named:
            stop = getattr(obj, "stop", None)
            if callable(stop):
                try:
                    stop()
                except Exception:
                    pass
        self._executor.shutdown(wait, timeout)


# ---------------------------------------------------------------------------
_runtime: Optional[PartyRuntime] = None
_runtime_lock = threading.Lock()


def get_runtime(create: bool = True) -> Optional[PartyRuntime]:
    global _runtime
    if _runtime is None and create:
        with _runtime_lock:
            if _runtime is None:
                _runtime = PartyRuntime()
    return _runtime


def shutdown_runtime(wait: bool = True):
    global _runtime
    with _runtime_lock:
        rt, _runtime = _runtime, None
    if rt is not None:
        rt.shutdown(wait)


def _atexit_shutdown():
    # Join worker/transport threads while the interpreter is still fully alive:
    # a native thread that wakes up during finalisation cannot take the GIL.
    try:
        shutdown_runtime(True)
    except Exception:  # noqa: BLE001
        pass


atexit.register(_atexit_shutdown)


def _host_value(ref, t: float):
    """Host-side read of a task result: its device work must be complete."""
    v = ref.result(t)
    if ref.cuda_event:
        ref.device_sync()  # GIL released
    return v


def get(refs, timeout: Optional[float] = None):
    """``ray.get`` analogue for a Ref or a list of Refs."""
    t = -1.0 if timeout is None else float(timeout)
    try:
        if is_ref(refs):
            return _host_value(refs, t)
        return [_host_value(r, t) for r in refs]
    except TimeoutError as e:
        raise FedGetTimeoutError(str(e)) from None


def wait(refs: List, num_returns: int = 1, timeout: Optional[float] = None):
    """``ray.wait`` analogue → ``(ready, not_ready)``."""
    t = -1.0 if timeout is None else float(timeout)
    ready, rest = native().wait_refs(list(refs), num_returns, t)
    return list(ready), list(rest)
