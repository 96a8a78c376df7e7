"""Public face of the tracer (implementation in :mod:`rayfed_b200._tracing`, which has no
package-level imports so the core runtime can hook it without an import cycle)."""
from .._tracing import (  # noqa: F401
    DeviceTimer,
    Timeline,
    disable,
    enable,
    nvtx_range,
    trace_ref,
    traced_task,
    tracer,
)
