"""In-tree native build for rayfed_b200.

Two shared objects live next to this file after a build:

* ``_fedcore``  — C++17 party runtime (scheduler, futures, mailbox, socket
  transport).  No CUDA dependency; loads on CPU-only machines.
* ``_fedcuda``  — sm_100a kernels (push / wait / barrier / aggregate / LoRA
  merge) + the CUDA-VMM symmetric heap.  Compiled with
  ``-gencode arch=compute_100a,code=sm_100a -lineinfo``; links cudart statically
  and resolves driver entry points at run time, so importing it never needs
  ``libcuda`` until a kernel is launched.

The ``.so`` files are git-ignored but travel with the working tree to the GPU
box, which is why the build is in-tree rather than a JIT cache.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX") or ".so"

CUDA_ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _pybind_includes():
    import pybind11

    return ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]


def _sources_digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(str(p.name).encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def _needs_build(target: Path, sources, extra: str = "") -> bool:
    stamp = target.with_suffix(target.suffix + ".stamp")
    digest = _sources_digest(sources) + extra
    if target.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return False
    return True


def _write_stamp(target: Path, sources, extra: str = ""):
    target.with_suffix(target.suffix + ".stamp").write_text(_sources_digest(sources) + extra)


def _run(cmd, verbose):
    if verbose:
        print("[rayfed_b200 build]", " ".join(map(str, cmd)), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(
            "native build failed:\n" + " ".join(map(str, cmd)) + "\n" + res.stdout + "\n" + res.stderr
        )
    if verbose and (res.stdout.strip() or res.stderr.strip()):
        print(res.stdout + res.stderr, flush=True)
    return res


def core_target() -> Path:
    return PKG_DIR / ("_fedcore" + EXT_SUFFIX)


def cuda_target() -> Path:
    return PKG_DIR / ("_fedcuda" + EXT_SUFFIX)


def build_core(force: bool = False, verbose: bool = False) -> Path:
    srcdir = CSRC / "core"
    sources = list(srcdir.glob("*.cpp")) + list(srcdir.glob("*.h"))
    target = core_target()
    if not force and not _needs_build(target, sources):
        return target
    cxx = os.environ.get("CXX", "g++")
    tmp = target.with_suffix(".tmp%d" % os.getpid())
    cmd = [
        cxx, "-O2", "-g0", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-pthread",
        *_pybind_includes(), "-I" + str(srcdir),
        *[str(p) for p in srcdir.glob("*.cpp")],
        "-o", str(tmp),
    ]
    _run(cmd, verbose)
    os.replace(tmp, target)
    _write_stamp(target, sources)
    return target


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcdir = CSRC / "cuda"
    cu = sorted(srcdir.glob("*.cu"))
    sources = cu + list(srcdir.glob("*.cuh")) + list(srcdir.glob("*.h")) + list(srcdir.glob("*.cpp"))
    target = cuda_target()
    if not force and not _needs_build(target, sources):
        return target
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found; cannot build the sm_100a kernels")
    objdir = PKG_DIR / "build" / "cuda"
    objdir.mkdir(parents=True, exist_ok=True)
    common = [
        "-O3", "-std=c++17", "-lineinfo", *CUDA_ARCH_FLAGS, "-Xcompiler", "-fPIC,-fvisibility=hidden",
        "--expt-relaxed-constexpr", "-I" + str(srcdir), *_pybind_includes(),
    ]
    objs = []
    procs = []
    for src in cu + sorted(srcdir.glob("*.cpp")):
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        stamp_extra = ""
        if not force and not _needs_build(obj, [src] + list(srcdir.glob("*.cuh")) + list(srcdir.glob("*.h")), stamp_extra):
            continue
        cmd = [nvcc, *common, "-c", str(src), "-o", str(obj)]
        if src.suffix == ".cpp":
            cmd.insert(1, "-x")
            cmd.insert(2, "cu")
        if verbose:
            print("[rayfed_b200 build]", " ".join(map(str, cmd)), flush=True)
        procs.append((src, obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, " ".join(map(str, cmd)), out))
        if verbose and out.strip():
            print(out, flush=True)
        _write_stamp(obj, [src] + list(srcdir.glob("*.cuh")) + list(srcdir.glob("*.h")))
    tmp = target.with_suffix(".tmp%d" % os.getpid())
    link = [nvcc, "-shared", *CUDA_ARCH_FLAGS, "-cudart", "static", *[str(o) for o in objs], "-o", str(tmp),
            "-Xlinker", "--no-undefined" if False else "-O1", "-lpthread", "-ldl", "-lrt"]
    _run(link, verbose)
    os.replace(tmp, target)
    _write_stamp(target, sources)
    return target


def ensure_core(verbose: bool = False) -> Path:
    return build_core(force=False, verbose=verbose)


def ensure_cuda(verbose: bool = False) -> Path:
    return build_cuda(force=False, verbose=verbose)


def build_all(force: bool = False, verbose: bool = True):
    a = build_core(force=force, verbose=verbose)
    b = build_cuda(force=force, verbose=verbose)
    return a, b


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_all(force=force))
