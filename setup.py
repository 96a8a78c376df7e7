"""Packaging for rayfed_b200 (parity: the reference's setup.py — wheel build with an
optional nightly mode selected by an environment variable, setup.py:23-30).

The native parts are built in-tree by :mod:`rayfed_b200._build` (g++ for the C++
core, nvcc -gencode arch=compute_100a,code=sm_100a for the kernels) and shipped as
package data; ``python setup.py build_ext --inplace`` / ``pip install .`` both go
through the same code path so there is exactly one set of compiler flags.
"""
import datetime
import os
import sys

from setuptools import Command, find_packages, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _version() -> str:
    ns = {}
    with open(os.path.join(ROOT, "rayfed_b200", "__init__.py")) as f:
        for line in f:
            if line.startswith("__version__"):
                exec(line, ns)
    version = ns.get("__version__", "0.0.0")
    if os.environ.get("RAYFED_B200_BUILD_MODE", "") == "nightly":
        version += ".dev" + datetime.datetime.now(datetime.timezone.utc).strftime("%Y%m%d")
    return version


def _build_native(require_cuda: bool = False) -> None:
    from rayfed_b200 import _build

    _build.build_core()
    try:
        _build.build_cuda()
    except Exception:  # no nvcc on this machine: CPU-only wheel (gRPC / socket transports only)
        if require_cuda:
            raise
        print("warning: nvcc not available, building without the sm_100a kernels", file=sys.stderr)


class BuildNative(Command):
    """``python setup.py build_ext [--inplace]``: compile _fedcore and _fedcuda in-tree."""

    description = "build the C++ runtime and the sm_100a CUDA kernels"
    user_options = [("inplace", "i", "kept for familiarity; the build is always in-tree"),
                    ("require-cuda", None, "fail if nvcc is missing")]
    boolean_options = ["inplace", "require-cuda"]

    def initialize_options(self):
        self.inplace = True
        self.require_cuda = False

    def finalize_options(self):
        pass

    def run(self):
        _build_native(bool(self.require_cuda))


class BuildPyWithNative(build_py):
    def run(self):
        _build_native()
        super().run()


setup(
    name="rayfed-b200",
    version=_version(),
    description="Blackwell-native cross-party federated execution engine with the rayfed API",
    long_description=open(os.path.join(ROOT, "README.md")).read(),
    long_description_content_type="text/markdown",
    packages=find_packages(include=["rayfed_b200", "rayfed_b200.*", "fed"]),
    package_data={"rayfed_b200": ["*.so", "csrc/core/*", "csrc/cuda/*"]},
    python_requires=">=3.10",
    install_requires=["cloudpickle", "pybind11", "numpy"],
    extras_require={
        "gpu": ["torch>=2.6"],
        "grpc": ["grpcio>=1.60", "protobuf>=4"],
        "tls-tools": ["cryptography"],
        "dev": ["pytest", "pytest-timeout", "black", "flake8"],
    },
    cmdclass={"build_ext": BuildNative, "build_py": BuildPyWithNative},
    zip_safe=False,
)
