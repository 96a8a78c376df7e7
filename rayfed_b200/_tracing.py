"""Tracing / profiling hooks (absent in the reference).

* :func:`nvtx_range` — NVTX range per fed op (visible in Nsight) when CUDA is
  present, no-op otherwise.
* :class:`Timeline` — lightweight host timeline of (name, seq-id, t0, t1)
  records, dumpable as Chrome trace JSON (``chrome://tracing`` / Perfetto).
* :class:`DeviceTimer` — CUDA-event timing on the launching stream.
"""
from __future__ import annotations

import contextlib
import json
import threading
import time
from typing import List, Optional


_NVTX = None  # None: unknown, False: unavailable, else torch.cuda.nvtx


def _nvtx():
    global _NVTX
    if _NVTX is None:
        try:
            import torch

            _NVTX = torch.cuda.nvtx if torch.cuda.is_available() else False
        except Exception:  # noqa: BLE001
            _NVTX = False
    return _NVTX


@contextlib.contextmanager
def nvtx_range(name: str):
    nv = _nvtx()
    if nv:
        nv.range_push(name)
    try:
        yield
    finally:
        if nv:
            nv.range_pop()


class Timeline:
    def __init__(self, party: str = ""):
        self.party = party
        self._events: List[dict] = []
        self._device: list = []
        self._lock = threading.Lock()
        self._t0 = time.perf_counter()

    @contextlib.contextmanager
    def span(self, name: str, seq_id=None, **args):
        start = time.perf_counter()
        with nvtx_range(f"{name}#{seq_id}" if seq_id is not None else name):
            yield
        end = time.perf_counter()
        with self._lock:
            self._events.append({
                "name": name, "ph": "X", "pid": self.party or "party", "tid": threading.current_thread().name,
                "ts": (start - self._t0) * 1e6, "dur": (end - start) * 1e6,
                "args": dict(args, seq_id=seq_id),
            })

    def complete(self, name: str, t0: float, t1: float, seq_id=None, **args):
        """Record a finished span given ``time.perf_counter()`` endpoints."""
        with self._lock:
            self._events.append({
                "name": name, "ph": "X", "pid": self.party or "party", "tid": threading.current_thread().name,
                "ts": (t0 - self._t0) * 1e6, "dur": (t1 - t0) * 1e6, "args": dict(args, seq_id=seq_id),
            })

    def device_span(self, name: str, start_event, end_event, **args):
        """A span measured by two CUDA events (resolved lazily at dump time, never synchronises
        the hot path)."""
        with self._lock:
            self._device.append((name, time.perf_counter(), start_event, end_event, args))

    def _resolve_device(self):
        with self._lock:
            pending, self._device = self._device, []
        for name, host_t, a, b, args in pending:
            try:
                b.synchronize()
                dur_us = a.elapsed_time(b) * 1e3
            except Exception:  # noqa: BLE001
                continue
            with self._lock:
                self._events.append({
                    "name": name, "ph": "X", "pid": self.party or "party", "tid": "device",
                    "ts": (host_t - self._t0) * 1e6, "dur": dur_us, "args": dict(args, clock="cuda-events"),
                })

    def events(self) -> List[dict]:
        self._resolve_device()
        with self._lock:
            return list(self._events)

    def dump(self, path: str):
        with open(path, "w") as f:
            json.dump({"traceEvents": self.events()}, f)


# ---------------------------------------------------------------------------
# Process-wide tracer used by the runtime hooks (api.get, barriers.send/recv, task
# execution, device pushes).  Disabled by default: every hook is one global load
# and a ``None`` test.  Enable with ``fed.init(config={"trace": "/path/prefix"})``,
# ``RAYFED_B200_TRACE=/path/prefix`` or :func:`enable`; ``fed.shutdown`` writes
# ``{prefix}.{party}.json`` (Chrome trace / Perfetto).
_TRACER: Optional["Timeline"] = None
_TRACE_PATH: Optional[str] = None


def enable(party: str = "", path: Optional[str] = None) -> "Timeline":
    global _TRACER, _TRACE_PATH
    _TRACER = Timeline(party)
    _TRACE_PATH = path
    return _TRACER


def disable(dump: bool = True) -> Optional[str]:
    """Stop tracing; returns the file written (if a path prefix was configured)."""
    global _TRACER, _TRACE_PATH
    tl, path = _TRACER, _TRACE_PATH
    _TRACER = None
    _TRACE_PATH = None
    if tl is None or not dump or not path:
        return None
    out = f"{path}.{tl.party or 'party'}.json"
    tl.dump(out)
    return out


def tracer() -> Optional["Timeline"]:
    return _TRACER


def traced_task(fn, name: str):
    """Wrap a task body so its execution shows up as a span (+ NVTX range) on the worker thread."""
    tl = _TRACER
    if tl is None:
        return fn

    def run(*args, **kwargs):
        with tl.span(name, cat="task"):
            return fn(*args, **kwargs)

    run.__name__ = getattr(fn, "__name__", name)
    return run


def trace_ref(name: str, ref, seq_id=None, **args) -> None:
    """Span from now until ``ref`` (a native Ref) completes — used for sends / receives."""
    tl = _TRACER
    if tl is None:
        return
    t0 = time.perf_counter()

    def done(r):
        try:
            ok = r.exception() is None
        except BaseException:  # noqa: BLE001
            ok = False
        tl.complete(name, t0, time.perf_counter(), seq_id=seq_id, ok=ok, **args)

    ref.add_done_callback(done)


class DeviceTimer:
    """``with DeviceTimer() as t: ...`` → ``t.ms`` (CUDA events on the current stream)."""

    def __init__(self, stream=None):
        import torch

        self._stream = stream or torch.cuda.current_stream()
        self._a = torch.cuda.Event(enable_timing=True)
        self._b = torch.cuda.Event(enable_timing=True)
        self.ms: Optional[float] = None

    def __enter__(self):
        self._a.record(self._stream)
        return self

    def __exit__(self, *exc):
        self._b.record(self._stream)
        self._b.synchronize()
        self.ms = self._a.elapsed_time(self._b)
        return False
