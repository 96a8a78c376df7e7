"""Migration helpers for code written against ray-project/rayfed."""
