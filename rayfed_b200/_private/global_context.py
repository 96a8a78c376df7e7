"""Per-process job context (parity: fed/_private/global_context.py:22-120).

Holds the job name, the deterministic sequence counter (native atomic — the
reference's is an unlocked Python int, global_context.py:45-47), failure policy,
the CleanupManager, the last error received from a peer, and the once-only
shutdown flag that makes ``fed.shutdown`` / the SIGINT path idempotent
(global_context.py:70-87).
"""
from __future__ import annotations

import threading
from typing import Callable, Optional

from ..exceptions import FedRemoteError


class GlobalContext:
    def __init__(
        self,
        job_name: str,
        current_party: str,
        sending_failure_handler: Optional[Callable[[Exception], None]],
        exit_on_sending_failure: bool = False,
        continue_waiting_for_data_sending_on_error: bool = False,
    ) -> None:
        from ..cleanup import CleanupManager
        from ..core import native

        self._job_name = job_name
        self._current_party = current_party
        self._seq = native().SeqCounter()
        # the mint is called once per fed call in every party: bind the native method directly
        # (no Python frame in between)
        self.next_seq_id = self._seq.next
        self._sending_failure_handler = sending_failure_handler
        self._exit_on_sending_failure = exit_on_sending_failure
        self._atomic_shutdown_flag_lock = threading.Lock()
        self._atomic_shutdown_flag = True
        self._cleanup_manager = CleanupManager(current_party, self.acquire_shutdown_flag)
        self._last_received_error: Optional[FedRemoteError] = None
        self._continue_waiting_for_data_sending_on_error = continue_waiting_for_data_sending_on_error
        # cached per-job facts used on the per-call hot path
        self._addresses = None
        self._extras = {}

    def next_seq_id(self) -> int:
        return self._seq.next()

    def current_seq_id(self) -> int:
        return self._seq.value()

    def get_cleanup_manager(self):
        return self._cleanup_manager

    def get_job_name(self) -> str:
        return self._job_name

    def get_current_party(self) -> str:
        return self._current_party

    def get_sending_failure_handler(self):
        return self._sending_failure_handler

    def get_exit_on_sending_failure(self) -> bool:
        return self._exit_on_sending_failure

    def get_last_recevied_error(self) -> Optional[FedRemoteError]:
        return self._last_received_error

    def set_last_recevied_error(self, err) -> None:
        self._last_received_error = err

    def get_continue_waiting_for_data_sending_on_error(self) -> bool:
        return self._continue_waiting_for_data_sending_on_error

    def acquire_shutdown_flag(self) -> bool:
        """True exactly once per context. Taken *before* SIGINT is raised so the
        handler never needs the lock (deadlock analysis:
        docs/enhancements/2023-09-01-cross-silo-error.md:30-37 in the reference)."""
        with self._atomic_shutdown_flag_lock:
            if self._atomic_shutdown_flag:
                self._atomic_shutdown_flag = False
                return True
            return False


_global_context: Optional[GlobalContext] = None


def init_global_context(
    current_party: str,
    job_name: str,
    exit_on_sending_failure: bool,
    continue_waiting_for_data_sending_on_error: bool,
    sending_failure_handler: Callable[[Exception], None] = None,
) -> None:
    global _global_context
    if _global_context is None:
        _global_context = GlobalContext(
            job_name,
            current_party,
            exit_on_sending_failure=exit_on_sending_failure,
            continue_waiting_for_data_sending_on_error=continue_waiting_for_data_sending_on_error,
            sending_failure_handler=sending_failure_handler,
        )


def get_global_context() -> Optional[GlobalContext]:
    return _global_context


def clear_global_context(wait_for_sending: bool = False) -> None:
    global _global_context
    if _global_context is not None:
        _global_context.get_cleanup_manager().stop(wait_for_sending=wait_for_sending)
        _global_context = None
