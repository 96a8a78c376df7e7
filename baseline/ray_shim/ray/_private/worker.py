class _Node:
    gcs_address = "shim:0"


class _Worker:
    mode = 0


_global_node = _Node()
global_worker = _Worker()
