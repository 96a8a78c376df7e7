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


@contextlib.contextmanager
def nvtx_range(name: str):
    pushed = False
    try:
        import torch

        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(name)
            pushed = True
    except Exception:  # noqa: BLE001
        pushed = False
    try:
        yield
    finally:
        if pushed:
            import torch

            torch.cuda.nvtx.range_pop()


class Timeline:
    def __init__(self, party: str = ""):
        self.party = party
        self._events: List[dict] = []
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

    def events(self) -> List[dict]:
        with self._lock:
            return list(self._events)

    def dump(self, path: str):
        with open(path, "w") as f:
            json.dump({"traceEvents": self.events()}, f)


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
