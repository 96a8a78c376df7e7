class RayError(Exception):
    pass


class RayTaskError(RayError):
    def __init__(self, function_name="", traceback_str="", cause=None):
        # not super(): the dual-inheritance subclasses made by make_task_error() put the cause's own
        # type (arbitrary __init__ signature) next in the MRO
        Exception.__init__(self, f"{function_name}: {cause!r}")
        self.function_name = function_name
        self.traceback_str = traceback_str
        self.cause = cause

    def as_instanceof_cause(self):
        return self


class RayActorError(RayError):
    pass


class GetTimeoutError(RayError, TimeoutError):
    pass


class ObjectLostError(RayError):
    pass


_derived = {}


def make_task_error(cause, function_name):
    """RayTaskError that is also an instance of the cause's type (Ray: as_instanceof_cause), so
    ``except FedRemoteError`` around ``ray.get`` works as in the reference (fed/api.py:601)."""
    import traceback

    tb = "".join(traceback.format_exception(type(cause), cause, cause.__traceback__))
    ctype = type(cause)
    cls = _derived.get(ctype)
    if cls is None:
        try:
            cls = type(f"RayTaskError({ctype.__name__})", (RayTaskError, ctype), {
                "__init__": lambda self, fn, tbs, c: RayTaskError.__init__(self, fn, tbs, c),
                "__str__": lambda self: f"{self.function_name} failed:\n{self.traceback_str}",
            })
        except TypeError:
            cls = RayTaskError
        _derived[ctype] = cls
    err = cls(function_name, tb, cause)
    for k, v in getattr(cause, "__dict__", {}).items():
        if not hasattr(err, k):
            try:
                setattr(err, k, v)
            except Exception:  # noqa: BLE001
                pass
    return err
