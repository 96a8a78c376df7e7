import threading

_lock = threading.Lock()
_store = {}
_initialized = False


def _initialize_internal_kv(gcs_client):
    global _initialized
    _initialized = True
    return True


def _internal_kv_initialized():
    return _initialized


def _b(x):
    return x.encode() if isinstance(x, str) else x


def _internal_kv_put(key, value, overwrite=True, namespace=None):
    key, value = _b(key), _b(value)
    with _lock:
        existed = key in _store
        if overwrite or not existed:
            _store[key] = value
        return existed


def _internal_kv_get(key, namespace=None):
    key = _b(key)
    with _lock:
        return _store.get(key)


def _internal_kv_del(key, namespace=None):
    key = _b(key)
    with _lock:
        return 1 if _store.pop(key, None) is not None else 0


def _internal_kv_exists(key, namespace=None):
    key = _b(key)
    with _lock:
        return key in _store


def _internal_kv_reset():
    global _initialized
    with _lock:
        _store.clear()
    _initialized = False
