is_client_mode_enabled = False
