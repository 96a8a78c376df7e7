def get_node_ip_address():
    return "127.0.0.1"


from . import client  # noqa: E402,F401
