class GcsClient:
    def __init__(self, address=None, nums_reconnect_retry=0, **kw):
        self.address = address
