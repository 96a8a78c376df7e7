class _ClientApi:
    def is_connected(self):
        return False


ray = _ClientApi()
