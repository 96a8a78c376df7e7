"""Exception types (parity: fed/exceptions.py:16-25).

``FedRemoteError`` is what a *peer* party sees when a task failed elsewhere.
``FedTaskError`` plays the role Ray's ``RayTaskError`` played in the reference:
the local wrapper whose ``.cause`` is the original exception (owner party) or
the ``FedRemoteError`` (consumer party) — see the contract exercised by
fed/tests/test_cross_silo_error.py:74-82.
"""
from __future__ import annotations

import traceback


class FedError(Exception):
    """Base class of every error raised by the runtime itself."""


class FedRemoteError(FedError):
    def __init__(self, src_party: str, cause=None) -> None:
        super().__init__(src_party, cause)
        self._src_party = src_party
        self._cause = cause

    @property
    def src_party(self):
        return self._src_party

    @property
    def cause(self):
        return self._cause

    def __str__(self) -> str:
        msg = f"FedRemoteError occurred at {self._src_party}"
        if self._cause is not None:
            msg += f" caused by {self._cause}"
        return msg

    def __reduce__(self):
        return (FedRemoteError, (self._src_party, self._cause))


class FedTaskError(FedError):
    """A task body raised; ``cause`` is the original exception instance."""

    def __init__(self, cause=None, traceback_str: str = "", task_name: str = "") -> None:
        super().__init__(cause, traceback_str, task_name)
        self.cause = cause
        self.traceback_str = traceback_str
        self.task_name = task_name

    @classmethod
    def from_exception(cls, exc: BaseException, task_name: str = "") -> "FedTaskError":
        if isinstance(exc, FedTaskError):
            return exc
        tb = "".join(traceback.format_exception(type(exc), exc, exc.__traceback__))
        return cls(exc, tb, task_name)

    def __str__(self) -> str:
        head = f"fed task failed: {type(self.cause).__name__}: {self.cause}"
        if self.traceback_str:
            head += "\n" + self.traceback_str
        return head

    def __reduce__(self):
        return (FedTaskError, (self.cause, self.traceback_str, self.task_name))


class FedActorDiedError(FedError):
    """Raised for calls on an actor whose mailbox was dropped by ``fed.kill``."""


class FedGetTimeoutError(FedError, TimeoutError):
    """A blocking get/ready-check exceeded its deadline."""


class FedDivergenceError(FedError):
    """Controllers disagreed about the (upstream, downstream) key of a slot."""
