"""A drainable work queue served by one daemon thread.

Parity: fed/_private/message_queue.py:28-105 — `append`, a STOP sentinel that is
appended (graceful drain) or pushed to the front (abort), and a handler whose
False return ends the loop.  Unlike the reference (which polls a deque and
sleeps 100 ms when idle, message_queue.py:46) the consumer parks on a
``threading.Condition``, so an idle party costs nothing and a new message is
picked up immediately.
"""
from __future__ import annotations

import logging
import threading
from collections import deque
from typing import Callable

logger = logging.getLogger(__name__)

STOP_SYMBOL = False


class MessageQueueManager:
    def __init__(self, msg_handler: Callable, failure_handler: Callable = None, thread_name: str = ""):
        assert callable(msg_handler), "msg_handler must be a callable function"
        self._queue = deque()
        self._cond = threading.Condition()
        self._msg_handler = msg_handler
        self._failure_handler = failure_handler
        self._thread = None
        if not thread_name:
            thread_name = "MessageQueueManager"
        self._thread_name = thread_name

    # -- lifecycle -------------------------------------------------------------
    def start(self) -> None:
        if self._thread is not None and self._thread.is_alive():
            return

        def _loop():
            while True:
                with self._cond:
                    while not self._queue:
                        self._cond.wait()
                    message = self._queue.popleft()
                if message is STOP_SYMBOL or (isinstance(message, bool) and message == STOP_SYMBOL):
                    break
                try:
                    keep_going = self._msg_handler(message)
                except Exception:  # a handler bug must not wedge shutdown
                    logger.exception("Unhandled error in %s handler.", self._thread_name)
                    keep_going = True
                if not keep_going:
                    break

        self._thread = threading.Thread(target=_loop, name=self._thread_name, daemon=True)
        self._thread.start()

    def append(self, message) -> None:
        with self._cond:
            self._queue.append(message)
            self._cond.notify()

    def appendleft(self, message) -> None:
        with self._cond:
            self._queue.appendleft(message)
            self._cond.notify()

    def _notify_to_exit(self, immediately: bool = False) -> None:
        logger.debug("Notify message polling thread[%s] to exit.", self._thread_name)
        if immediately:
            self.appendleft(STOP_SYMBOL)
        else:
            self.append(STOP_SYMBOL)

    def stop(self, wait_for_sending: bool = True) -> None:
        """Stop the consumer.

        ``wait_for_sending=True`` drains everything queued before the sentinel;
        ``False`` jumps the queue so pending messages are abandoned.  Safe to
        call from the consumer thread itself (it will not join itself).
        """
        if self._thread is None:
            return
        if threading.current_thread() is self._thread:
            logger.error("Can't stop the message queue in the message polling thread[%s].", self._thread_name)
            return
        if self.is_started():
            self._notify_to_exit(immediately=not wait_for_sending)
            self._thread.join()
            logger.info("The message polling thread[%s] was exited.", self._thread_name)

    def is_started(self) -> bool:
        return self._thread is not None and self._thread.is_alive()

    def size(self) -> int:
        return len(self._queue)
