"""Sphinx configuration (MyST markdown)."""
import os
import sys

sys.path.insert(0, os.path.abspath("../.."))

project = "rayfed_b200"
author = "rayfed_b200 developers"
release = "0.1.0"
extensions = ["sphinx.ext.autodoc", "sphinx.ext.napoleon", "sphinx.ext.viewcode"]
try:  # markdown sources
    import myst_parser  # noqa: F401

    extensions.append("myst_parser")
except ImportError:  # plain Sphinx still builds the API pages
    pass
source_suffix = {".rst": "restructuredtext", ".md": "markdown"}
master_doc = "index"
exclude_patterns = ["_build"]
html_theme = "alabaster"
autodoc_mock_imports = ["torch", "grpc", "google", "rayfed_b200._fedcore", "rayfed_b200._fedcuda"]
