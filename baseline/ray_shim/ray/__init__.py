"""Minimal in-process stand-in for the slice of the Ray API that ray-project/rayfed uses.

See ../README.md.  Pure standard library (+ cloudpickle for parity of `ray.cloudpickle`);
contains no rayfed_b200 code.  Actors are threads (sync actors: a FIFO worker pool of
`max_concurrency` threads; actors with `async def` methods: one asyncio loop thread), tasks run
on a shared thread pool, an ObjectRef passed as a top-level argument delays the call until it
resolves and propagates its failure — the three behaviours the reference relies on
(fed/proxy/barriers.py:478-483, fed/cleanup.py:150-172).
"""
from __future__ import annotations

import asyncio
import concurrent.futures
import inspect
import queue
import threading
import time

from . import exceptions
from .exceptions import GetTimeoutError, RayActorError, RayError, RayTaskError  # noqa: F401

__version__ = "2.9.3"  # what the reference's version shims are compared against
__shim__ = True

_state_lock = threading.RLock()
_initialized = False
_named_actors = {}
_all_actors = []
_pool = None


def _task_pool():
    global _pool
    with _state_lock:
        if _pool is None:
            _pool = concurrent.futures.ThreadPoolExecutor(max_workers=32, thread_name_prefix="ray-shim-task")
        return _pool


def is_initialized():
    return _initialized


def init(*args, **kwargs):
    global _initialized
    with _state_lock:
        _initialized = True
    from ._private import worker

    worker._global_node = worker._Node()
    return {"address": worker._global_node.gcs_address}


def shutdown():
    global _initialized, _pool
    with _state_lock:
        actors = list(_all_actors)
        _all_actors.clear()
        _named_actors.clear()
        _initialized = False
        pool, _pool = _pool, None
    for a in actors:
        a._stop("ray.shutdown")
    if pool is not None:
        pool.shutdown(wait=False)
    from .experimental import internal_kv

    internal_kv._internal_kv_reset()


class ObjectRef:
    """One-shot future."""

    __slots__ = ("_ev", "_value", "_error", "_callbacks", "_lock", "__weakref__")

    def __init__(self):
        self._ev = threading.Event()
        self._value = None
        self._error = None
        self._callbacks = []
        self._lock = threading.Lock()

    def _set(self, value=None, error=None):
        with self._lock:
            if self._ev.is_set():
                return
            self._value, self._error = value, error
            self._ev.set()
            cbs, self._callbacks = self._callbacks, []
        for cb in cbs:
            cb(self)

    def _on_done(self, cb):
        with self._lock:
            if not self._ev.is_set():
                self._callbacks.append(cb)
                return
        cb(self)

    def done(self):
        return self._ev.is_set()

    def __repr__(self):
        return f"ObjectRef(shim:{id(self):x})"


def put(value):
    r = ObjectRef()
    r._set(value)
    return r


def _raise_for(ref):
    raise ref._error


def get(refs, *, timeout=None):
    single = isinstance(refs, ObjectRef)
    lst = [refs] if single else list(refs)
    deadline = None if timeout is None else time.monotonic() + timeout
    out = []
    for r in lst:
        if not isinstance(r, ObjectRef):
            raise ValueError(f"ray.get expects ObjectRef(s), got {type(r)}")
        left = None if deadline is None else max(0.0, deadline - time.monotonic())
        if not r._ev.wait(left):
            raise GetTimeoutError("Get timed out: some object(s) not ready.")
        if r._error is not None:
            _raise_for(r)
        out.append(r._value)
    return out[0] if single else out


def wait(refs, *, num_returns=1, timeout=None, fetch_local=True):
    refs = list(refs)
    deadline = None if timeout is None else time.monotonic() + timeout
    cond = threading.Condition()

    def poke(_):
        with cond:
            cond.notify_all()

    for r in refs:
        r._on_done(poke)
    with cond:
        while True:
            ready = [r for r in refs if r.done()]
            if len(ready) >= num_returns:
                break
            left = None if deadline is None else deadline - time.monotonic()
            if left is not None and left <= 0:
                break
            cond.wait(left)
    ready = [r for r in refs if r.done()][:num_returns]
    rest = [r for r in refs if r not in ready]
    return ready, rest


def _wrap_error(exc, what):
    if isinstance(exc, RayError):
        return exc
    return exceptions.make_task_error(exc, what)


def _resolve_then(args, kwargs, run, fail):
    """Call ``run(args, kwargs)`` once every top-level ObjectRef argument has resolved; if one of
    them failed, ``fail(error)`` instead (Ray: the dependent task fails with the same error)."""
    deps = [a for a in args if isinstance(a, ObjectRef)] + [v for v in kwargs.values() if isinstance(v, ObjectRef)]
    if not deps:
        return run(args, kwargs)
    pending = [len(deps)]
    lock = threading.Lock()

    def one_done(_):
        with lock:
            pending[0] -= 1
            if pending[0]:
                return
        for d in deps:
            if d._error is not None:
                return fail(d._error)
        a2 = tuple(a._value if isinstance(a, ObjectRef) else a for a in args)
        k2 = {k: (v._value if isinstance(v, ObjectRef) else v) for k, v in kwargs.items()}
        run(a2, k2)

    for d in deps:
        d._on_done(one_done)


def _deliver(refs, num_returns, result):
    if num_returns == 1:
        refs[0]._set(result)
        return
    vals = list(result)
    if len(vals) != num_returns:
        err = _wrap_error(ValueError(f"expected {num_returns} return values, got {len(vals)}"), "task")
        for r in refs:
            r._set(error=err)
        return
    for r, v in zip(refs, vals):
        r._set(v)


class RemoteFunction:
    def __init__(self, fn, options=None):
        self._fn = fn
        self._options = dict(options or {})

    def options(self, **options):
        return RemoteFunction(self._fn, options)

    def remote(self, *args, **kwargs):
        n = int(self._options.get("num_returns", 1) or 1)
        refs = [ObjectRef() for _ in range(n)]
        name = getattr(self._fn, "__name__", "task")

        def fail(err):
            for r in refs:
                r._set(error=err)

        def run(a, k):
            def body():
                try:
                    _deliver(refs, n, self._fn(*a, **k))
                except BaseException as e:  # noqa: BLE001
                    fail(_wrap_error(e, name))

            _task_pool().submit(body)

        _resolve_then(args, kwargs, run, fail)
        return refs[0] if n == 1 else refs

    def __call__(self, *a, **k):
        raise TypeError("Remote functions cannot be called directly. Use .remote().")


class ActorMethod:
    def __init__(self, handle, name, options=None):
        self._handle = handle
        self._name = name
        self._options = dict(options or {})

    def options(self, **options):
        return ActorMethod(self._handle, self._name, options)

    def remote(self, *args, **kwargs):
        return self._handle._submit(self._name, args, kwargs, int(self._options.get("num_returns", 1) or 1))


class ActorHandle:
    def __init__(self, cls, options, args, kwargs):
        self._cls = cls
        self._options = dict(options or {})
        self._name = self._options.get("name") or ""
        self._dead = None
        self._instance = None
        self._is_async = any(inspect.iscoroutinefunction(m) for _, m in inspect.getmembers(cls, inspect.isfunction))
        conc = self._options.get("max_concurrency")
        self._concurrency = int(conc) if conc else (1000 if self._is_async else 1)
        self._q = queue.Queue()
        self._loop = None
        self._threads = []
        for name, member in inspect.getmembers(cls, callable):
            if not name.startswith("__"):
                object.__setattr__(self, name, ActorMethod(self, name))
        self._ready = ObjectRef()
        if self._is_async:
            t = threading.Thread(target=self._async_main, name=f"ray-shim-actor-{cls.__name__}", daemon=True)
            t.start()
            self._threads.append(t)
        else:
            for i in range(min(self._concurrency, 16)):
                t = threading.Thread(target=self._sync_main, name=f"ray-shim-actor-{cls.__name__}-{i}", daemon=True)
                t.start()
                self._threads.append(t)

        def finish_ctor(a, k):
            try:
                self._instance = cls(*a, **k)
                self._ready._set(True)
            except BaseException as e:  # noqa: BLE001
                self._dead = RayActorError(f"The actor died because of an error raised in its creation task: {e!r}")
                self._ready._set(error=self._dead)

        def fail(err):
            self._dead = RayActorError(f"actor creation failed: {err!r}")
            self._ready._set(error=self._dead)

        if not self._is_async and self._concurrency == 1:
            # sequential actor: the constructor is the first entry of the mailbox; it waits there
            # for its own arguments, and every method call queues up behind it
            def ctor_body():
                deps = [x for x in args if isinstance(x, ObjectRef)] + \
                       [v for v in kwargs.values() if isinstance(v, ObjectRef)]
                for d in deps:
                    d._ev.wait()
                    if d._error is not None:
                        return fail(d._error)
                finish_ctor(tuple(x._value if isinstance(x, ObjectRef) else x for x in args),
                            {key: (v._value if isinstance(v, ObjectRef) else v) for key, v in kwargs.items()})

            self._q.put(ctor_body)
        else:
            _resolve_then(args, kwargs, lambda a, k: self._enqueue(lambda: finish_ctor(a, k), is_ctor=True), fail)

    # -- execution -------------------------------------------------------------------
    def _enqueue(self, fn, is_ctor=False):
        if self._is_async:
            # the loop may not be up yet
            while self._loop is None and self._dead is None:
                time.sleep(0.0005)
            if is_ctor:
                self._loop.call_soon_threadsafe(fn)
            else:
                self._loop.call_soon_threadsafe(fn)
        else:
            self._q.put(fn)

    def _sync_main(self):
        while True:
            fn = self._q.get()
            if fn is None:
                return
            fn()

    def _async_main(self):
        loop = asyncio.new_event_loop()
        asyncio.set_event_loop(loop)
        self._sem = asyncio.Semaphore(self._concurrency)
        self._loop = loop
        try:
            loop.run_forever()
        finally:
            try:
                loop.close()
            except Exception:  # noqa: BLE001
                pass

    def _submit(self, method, args, kwargs, num_returns):
        refs = [ObjectRef() for _ in range(num_returns)]
        what = f"{self._cls.__name__}.{method}"

        def fail(err):
            for r in refs:
                r._set(error=err)

        if not self._is_async and self._concurrency == 1:
            # A plain (sequential) actor executes its tasks in SUBMISSION order, like Ray does per
            # caller: a task whose arguments are still pending holds the ones behind it back.
            def body():
                self._ready._ev.wait()
                if self._ready._error is not None:
                    return fail(self._ready._error)
                deps = [a for a in args if isinstance(a, ObjectRef)] + \
                       [v for v in kwargs.values() if isinstance(v, ObjectRef)]
                for d in deps:
                    d._ev.wait()
                    if d._error is not None:
                        return fail(d._error)
                if self._dead is not None:
                    return fail(self._dead)
                a = tuple(x._value if isinstance(x, ObjectRef) else x for x in args)
                k = {key: (v._value if isinstance(v, ObjectRef) else v) for key, v in kwargs.items()}
                try:
                    _deliver(refs, num_returns, getattr(self._instance, method)(*a, **k))
                except BaseException as e:  # noqa: BLE001
                    fail(_wrap_error(e, what))

            if self._dead is not None:
                fail(self._dead)
            else:
                self._q.put(body)
            return refs[0] if num_returns == 1 else refs

        def run(a, k):
            # async / threaded actors: a call starts once its arguments and the constructor are done
            def after_ready(_):
                if self._ready._error is not None:
                    return fail(self._ready._error)
                if self._dead is not None:
                    return fail(self._dead)
                if self._is_async:
                    async def go():
                        try:
                            async with self._sem:
                                res = getattr(self._instance, method)(*a, **k)
                                if inspect.isawaitable(res):
                                    res = await res
                            _deliver(refs, num_returns, res)
                        except BaseException as e:  # noqa: BLE001
                            fail(_wrap_error(e, what))

                    self._loop.call_soon_threadsafe(lambda: asyncio.ensure_future(go()))
                else:
                    def body():
                        if self._dead is not None:
                            return fail(self._dead)
                        try:
                            _deliver(refs, num_returns, getattr(self._instance, method)(*a, **k))
                        except BaseException as e:  # noqa: BLE001
                            fail(_wrap_error(e, what))

                    self._q.put(body)

            self._ready._on_done(after_ready)

        _resolve_then(args, kwargs, run, fail)
        return refs[0] if num_returns == 1 else refs

    def _stop(self, why):
        if self._dead is None:
            self._dead = RayActorError(f"The actor is dead: {why}")
        if self._is_async:
            loop = self._loop
            if loop is not None and loop.is_running():
                # A real actor is a process: when it dies the kernel closes its sockets.  Emulate
                # that for what lives on this actor's event loop — gRPC aio servers are shut down,
                # pending tasks cancelled — then stop the loop.
                async def die():
                    import gc

                    for obj in gc.get_objects():
                        try:
                            if type(obj).__module__.startswith("grpc.aio") and type(obj).__name__ == "Server" \
                                    and getattr(obj, "_loop", None) is loop:
                                await obj.stop(0)
                        except Exception:  # noqa: BLE001
                            pass
                    me = asyncio.current_task()
                    for t in asyncio.all_tasks(loop):
                        if t is not me:
                            t.cancel()
                    await asyncio.sleep(0)
                    done.set_result(True)
                    loop.stop()

                done = concurrent.futures.Future()
                loop.call_soon_threadsafe(lambda: asyncio.ensure_future(die()))
                try:
                    done.result(timeout=5)
                except Exception:  # noqa: BLE001
                    try:
                        loop.call_soon_threadsafe(loop.stop)
                    except RuntimeError:
                        pass
            self._instance = None
        else:
            # fail what is queued, then release the workers
            try:
                while True:
                    self._q.get_nowait()
            except queue.Empty:
                pass
            for _ in self._threads:
                self._q.put(None)
            self._instance = None

    def __repr__(self):
        return f"Actor({self._cls.__name__}, name={self._name!r})"


class _UnschedulableActor(ActorHandle):
    def __init__(self, cls):  # noqa: super().__init__ deliberately not called
        self._cls, self._name, self._dead, self._threads, self._is_async = cls, "", None, [], False
        for name, member in inspect.getmembers(cls, callable):
            if not name.startswith("__"):
                object.__setattr__(self, name, ActorMethod(self, name))

    def _submit(self, method, args, kwargs, num_returns):
        refs = [ObjectRef() for _ in range(num_returns)]
        return refs[0] if num_returns == 1 else refs

    def _stop(self, why):
        self._dead = RayActorError(why)


class ActorClass:
    def __init__(self, cls, options=None):
        self._cls = cls
        self._options = dict(options or {})

    def options(self, **options):
        merged = dict(self._options)
        merged.update(options)
        return ActorClass(self._cls, merged)

    def remote(self, *args, **kwargs):
        name = self._options.get("name")
        with _state_lock:
            if name and name in _named_actors and _named_actors[name]._dead is None:
                raise ValueError(f"The name {name} is already taken. Please use a different name or get "
                                 "the existing actor using ray.get_actor()")
        if self._options.get("resources"):
            # a custom resource label this single-node stand-in does not provide: the actor is
            # never scheduled and its calls never resolve (the reference's test
            # fed/tests/test_setup_proxy_actor.py relies on the resulting GetTimeoutError)
            handle = _UnschedulableActor(self._cls)
        else:
            handle = ActorHandle(self._cls, self._options, args, kwargs)
        with _state_lock:
            _all_actors.append(handle)
            if name:
                _named_actors[name] = handle
        return handle

    def __call__(self, *a, **k):
        raise TypeError("Actors cannot be instantiated directly. Use .remote().")


def remote(*args, **kwargs):
    def make(obj, options):
        if inspect.isclass(obj):
            return ActorClass(obj, options)
        if callable(obj):
            return RemoteFunction(obj, options)
        raise TypeError("The @ray.remote decorator must be applied to either a function or a class.")

    if len(args) == 1 and not kwargs and callable(args[0]):
        return make(args[0], {})
    assert not args, "ray.remote(**options) takes keyword arguments only"
    return lambda obj: make(obj, kwargs)


def get_actor(name, namespace=None):
    with _state_lock:
        h = _named_actors.get(name)
    if h is None or h._dead is not None:
        raise ValueError(f"Failed to look up actor with name '{name}'. This could because 1. You are trying to look "
                         "up a named actor you didn't create. 2. The named actor died.")
    return h


def kill(actor, *, no_restart=True):
    if not isinstance(actor, ActorHandle):
        raise ValueError("ray.kill() only supported for actors.")
    actor._stop("ray.kill")
    with _state_lock:
        if actor._name and _named_actors.get(actor._name) is actor:
            del _named_actors[actor._name]


from . import actor as actor  # noqa: E402,F401
from . import util as util  # noqa: E402,F401
from . import worker as worker  # noqa: E402,F401
