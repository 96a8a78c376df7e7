"""Cross-party transport layer: plug-in contracts, hosts, and transports."""
