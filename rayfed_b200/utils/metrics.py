"""Metrics export (absent in the reference — SURVEY §5.5: two op counters on the proxy actors,
fed/proxy/barriers.py:113-183, read with `_get_stats`).

`snapshot()` gathers every counter the party keeps — scheduler, host transport, send bookkeeping,
device channel, kernel launches, spin policy — into one flat ``{name: number}`` dict.
`start_exporter(port)` serves the same numbers in the Prometheus text format (a collector that
reads them at scrape time: nothing is sampled or copied on the hot paths) when
``prometheus_client`` is importable; `fed.init(config={"metrics_port": 9464})` or
``RAYFED_B200_METRICS_PORT`` turn it on, `fed.shutdown()` turns it off.
"""
from __future__ import annotations

import logging
import threading
from typing import Dict, Optional

logger = logging.getLogger(__name__)

_PREFIX = "rayfed_b200_"
_exporter = None
_lock = threading.Lock()


def _flat(out: Dict[str, float], prefix: str, stats) -> None:
    if not stats:
        return
    for k, v in dict(stats).items():
        if isinstance(v, bool):
            out[f"{prefix}{k}"] = float(v)
        elif isinstance(v, (int, float)):
            out[f"{prefix}{k}"] = float(v)


def snapshot() -> Dict[str, float]:
    """All counters of this party as ``{name: value}`` (missing subsystems are simply absent)."""
    out: Dict[str, float] = {}
    try:
        from ..core import get_runtime, native

        _flat(out, "executor_", get_runtime().stats())
        info = native().spin_info()
        _flat(out, "spin_", info)
    except Exception:  # noqa: BLE001
        logger.debug("metrics: runtime stats unavailable", exc_info=True)
    try:
        from ..proxy import barriers

        for role in ("sender", "receiver"):
            host = barriers._HOSTS.get(role)
            if host is not None:
                _flat(out, f"{role}_", host._get_stats())
    except Exception:  # noqa: BLE001
        logger.debug("metrics: proxy stats unavailable", exc_info=True)
    try:
        from .._private.global_context import get_global_context

        ctx = get_global_context()
        if ctx is not None:
            _flat(out, "sends_", ctx.get_cleanup_manager().stats())
            out["seq_id"] = float(ctx.current_seq_id())
    except Exception:  # noqa: BLE001
        logger.debug("metrics: cleanup stats unavailable", exc_info=True)
    try:
        from ..proxy.nvlink import nvlink_proxy

        ch = nvlink_proxy.current_channel()
        if ch is not None:
            _flat(out, "device_channel_", ch.stats)
    except Exception:  # noqa: BLE001
        pass
    try:
        import sys

        kernels = sys.modules.get("rayfed_b200.ops.kernels")
        if kernels is not None:
            for name, n in dict(kernels.launch_counter.by_kernel).items():
                out[f"kernel_launches_{name}"] = float(n)
    except Exception:  # noqa: BLE001
        pass
    return out


class _Collector:
    """prometheus_client custom collector: one untyped sample per counter, read at scrape time."""

    def __init__(self, party: str, job: str):
        self._labels = {"party": party, "job_name": job}

    def collect(self):
        from prometheus_client.core import GaugeMetricFamily

        for name, value in sorted(snapshot().items()):
            safe = "".join(c if c.isalnum() or c == "_" else "_" for c in name)
            g = GaugeMetricFamily(_PREFIX + safe, f"rayfed_b200 counter {name}", labels=list(self._labels))
            g.add_metric(list(self._labels.values()), value)
            yield g


def start_exporter(port: int, party: str = "", job: str = "", addr: str = "127.0.0.1") -> Optional[int]:
    """Serve `/metrics` on ``addr:port`` (port 0 = pick one).  Returns the bound port, or None when
    ``prometheus_client`` is missing.  One exporter per process; a second call replaces the first."""
    global _exporter
    try:
        from prometheus_client import CollectorRegistry, start_http_server
    except ImportError:
        logger.warning("metrics_port is set but prometheus_client is not installed: no exporter")
        return None
    with _lock:
        stop_exporter()
        registry = CollectorRegistry()
        registry.register(_Collector(party, job))
        server, thread = start_http_server(int(port), addr=addr, registry=registry)
        _exporter = (server, thread)
        bound = server.server_address[1]
    logger.info(f"metrics exporter listening on {addr}:{bound}")
    return bound


def stop_exporter() -> None:
    global _exporter
    exp, _exporter = _exporter, None
    if exp is not None:
        server, thread = exp
        try:
            server.shutdown()
            server.server_close()
            thread.join(2)
        except Exception:  # noqa: BLE001
            logger.debug("metrics exporter did not stop cleanly", exc_info=True)
