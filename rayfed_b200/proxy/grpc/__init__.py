"""gRPC/TLS transport — wire-compatible with the reference; the off-box fallback."""
