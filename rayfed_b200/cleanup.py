"""CleanupManager — watches every cross-party send to completion.

Parity: fed/cleanup.py:29-203.  Two drainable queues:

* *data-sending*: each element is the future of one ``send``.  If it failed
  because the **producer task** failed (a `FedTaskError` flowing along the
  edge — the reference tests ``isinstance(e, RayError)``, cleanup.py:160) a
  ``FedRemoteError`` is pushed to the peer **under the same
  (upstream, downstream) key** so the peer's receive completes with an error
  instead of hanging (cleanup.py:160-172).  With ``exit_on_sending_failure`` a
  failed send raises SIGINT in this process after taking the once-only shutdown
  flag (cleanup.py:112-128,176-183).
* *error-sending*: futures of those replacement pushes; failures are only logged.
"""
from __future__ import annotations

import logging
import os
import signal
import time
import weakref

from ._private.message_queue import MessageQueueManager
from .exceptions import FedRemoteError, FedTaskError

logger = logging.getLogger(__name__)


class CleanupManager:
    def __init__(self, current_party, acquire_shutdown_flag) -> None:
        self._sending_data_q = MessageQueueManager(
            self._process_data_sending_task_return, thread_name="DataSendingQueueThread"
        )
        self._sending_error_q = MessageQueueManager(
            self._process_error_sending_task_return, thread_name="ErrorSendingQueueThread"
        )
        self._current_party = current_party
        self._acquire_shutdown_flag = acquire_shutdown_flag
        self._last_sending_error = None
        self._exit_on_sending_failure = False
        self._expose_error_trace = False
        # Sends whose outcome is not known yet are counted by a native tracker: an acknowledged
        # send is retired on the transport thread without the GIL (csrc/core/send_tracker.h); only
        # failures — and outcomes that are Python objects — come back through `_on_settled`.
        from .core import native

        # the tracker keeps its handler alive from C++ (invisible to the GC): hand it a weak method
        weak = weakref.WeakMethod(self._on_settled)

        def handler(r, pack, is_error):
            m = weak()
            if m is not None:
                m(r, pack, is_error)

        self._tracker = native().SendTracker(handler)

    def start(self, exit_on_sending_failure=False, expose_error_trace=False):
        self._exit_on_sending_failure = exit_on_sending_failure
        self._expose_error_trace = expose_error_trace
        self._sending_data_q.start()
        self._sending_error_q.start()
        logger.debug("Started the send-checking threads.")

    def stop(self, wait_for_sending=False, drain_timeout_s: float = 600.0, error_drain_timeout_s: float = 5.0):
        """Stop watching sends.

        ``wait_for_sending`` waits for every *unresolved* send first.  Failures
        that were already detected are always processed, and the replacement
        errors they trigger are given a bounded time to reach the peer —
        otherwise a party that shuts down right after a failure would leave its
        peers waiting for data that will never come.
        """
        tracker = self._tracker
        tracker.set_draining(True)
        # A send future may have resolved an instant ago on another thread while its settle step
        # (which queues a failure) is still running: let it finish.
        deadline = time.monotonic() + 2.0
        while time.monotonic() < deadline and tracker.settling():
            time.sleep(0.0005)
        if wait_for_sending:
            deadline = time.monotonic() + drain_timeout_s
            while self._sending_data_q.is_started() and not tracker.wait_idle(0.05) and time.monotonic() < deadline:
                pass
        # Data queue first: draining it may enqueue replacement errors.
        self._sending_data_q.stop(wait_for_sending=True)
        tracker.wait_idle(error_drain_timeout_s, True)
        self._sending_error_q.stop(wait_for_sending=wait_for_sending)
        tracker.bury()

    def push_to_sending(self, obj_ref, dest_party: str = None, upstream_seq_id=-1, downstream_seq_id=-1,
                        is_error: bool = False):
        """Watch the future of a send.  The seq ids are only needed if a
        replacement error has to be shipped under the same key.

        Unlike the reference (one ``ray.get`` per send on a polling thread,
        fed/cleanup.py:150-152) an acknowledged send is retired natively on the thread that
        received the acknowledgement; only failures go through the queue thread."""
        self._tracker.watch(obj_ref, (obj_ref, dest_party, upstream_seq_id, downstream_seq_id), bool(is_error))

    def _on_settled(self, r, pack, is_error):
        """Slow path of the tracker (GIL held): the send did not resolve to a plain acknowledgement."""
        ok = False
        try:
            ok = r.exception() is None and bool(r.result())
        except BaseException:  # noqa: BLE001
            ok = False
        q = self._sending_error_q if is_error else self._sending_data_q
        if ok or not q.is_started():
            self._tracker.retire(is_error)
        else:
            q.append(pack)

    def _retire(self, is_error: bool = False):
        self._tracker.retire(is_error)

    def pending_sends(self) -> int:
        return self._tracker.pending()

    def stats(self) -> dict:
        """Sends being watched and how the finished ones were settled: ``acked_natively`` never
        touched the interpreter, ``settled_in_python`` are failures and Python-object outcomes."""
        t = self._tracker
        return {"pending": t.pending(), "pending_errors": t.pending_errors(),
                "acked_natively": t.fast_retired(), "settled_in_python": t.slow_settled()}

    def get_last_sending_error(self):
        return self._last_sending_error

    def _signal_exit(self):
        # Take the flag first: the SIGINT handler must not need the lock.
        if self._acquire_shutdown_flag():
            logger.warning("Signal SIGINT to exit.")
            os.kill(os.getpid(), signal.SIGINT)

    def _process_data_sending_task_return(self, message):
        try:
            return self._handle_failed_data_send(message)
        finally:
            self._retire()

    def _handle_failed_data_send(self, message):
        obj_ref, dest_party, upstream_seq_id, downstream_seq_id = message
        try:
            res = obj_ref.result()
        except Exception as e:  # noqa: BLE001 - any failure of the send future
            logger.warning(
                f"Failed to send {obj_ref} to {dest_party}, error: {e},"
                f"upstream_seq_id: {upstream_seq_id}, downstream_seq_id: {downstream_seq_id}."
            )
            self._last_sending_error = e
            if isinstance(e, FedTaskError):
                # The producer failed: tell the consumer under the same key.
                from .proxy.barriers import send

                logger.info(f"Sending error {e.cause} to {dest_party}.")
                error_trace = e.cause if self._expose_error_trace else None
                if error_trace is not None and not _is_picklable(error_trace):
                    error_trace = RuntimeError(f"{type(error_trace).__name__}: {error_trace}")
                send(
                    dest_party,
                    FedRemoteError(self._current_party, error_trace),
                    upstream_seq_id,
                    downstream_seq_id,
                    True,
                )
            res = False

        if not res and self._exit_on_sending_failure:
            self._signal_exit()
            return False  # stop polling; the main thread will tear us down
        return True

    def _process_error_sending_task_return(self, error_msg):
        try:
            return self._handle_failed_error_send(error_msg)
        finally:
            self._retire(True)

    def _handle_failed_error_send(self, error_msg):
        error_ref, dest_party, upstream_seq_id, downstream_seq_id = error_msg
        try:
            res = error_ref.result()
            logger.debug(f"Sending error got response: {res}.")
        except Exception:  # noqa: BLE001
            res = False
        if not res:
            logger.warning(
                f"Failed to send error {error_ref} to {dest_party}, "
                f"upstream_seq_id: {upstream_seq_id} downstream_seq_id: {downstream_seq_id}. "
                "In this case, other parties won't sense this error and may cause unknown behaviour."
            )
        return True  # keep going so remaining errors are still delivered


def _is_picklable(obj) -> bool:
    try:
        from ._private.serialization import dumps

        dumps(obj)
        return True
    except Exception:  # noqa: BLE001
        return False
