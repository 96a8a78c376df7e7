"""Symmetric heap: the same buffer in every party, addressable by every party.

This replaces "a TCP connection per peer" (SURVEY §5.8): after set-up a party
holds a *device pointer* to each peer's copy of the buffer and kernels move data
with plain loads/stores over NVLink.

Backends
--------
``ipc``   cudaMalloc + cudaIpcGetMemHandle; the 64-byte handle travels over any
          PartyGroup exchange.  Works wherever CUDA P2P works.  (default)
``vmm``   cuMemCreate with a POSIX-fd handle, fds passed over a Unix socket
          (SCM_RIGHTS), cuMemMap on every peer, and — when the device supports
          it — a cuMulticast object bound to all copies, giving one *multicast*
          address for ``multimem.ld_reduce`` / ``multimem.st`` (NVLS).
"""
from __future__ import annotations

import logging
import os
import struct
import socket
import time
from typing import List, Optional

from ..ops.kernels import cuda_ext
from .group import PartyGroup

logger = logging.getLogger(__name__)


class _CudaArrayView:
    """``__cuda_array_interface__`` shim so torch can wrap raw device memory."""

    def __init__(self, ptr: int, nbytes: int, owner=None):
        self._owner = owner
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 3,
            "strides": None,
        }


class SymmetricBuffer:
    def __init__(self, name: str, nbytes: int, local_ptr: int, peer_ptrs: List[int], rank: int,
                 device_index: int, mc_ptr: int = 0, backend: str = "ipc"):
        self.name = name
        self.nbytes = nbytes
        self.local_ptr = local_ptr
        self.peer_ptrs = peer_ptrs  # party order; peer_ptrs[rank] == local_ptr
        self.rank = rank
        self.device_index = device_index
        self.mc_ptr = mc_ptr
        self.backend = backend
        self._bytes_tensor = None

    def bytes_tensor(self):
        """uint8 torch view of the local copy."""
        if self._bytes_tensor is None:
            import torch

            self._bytes_tensor = torch.as_tensor(
                _CudaArrayView(self.local_ptr, self.nbytes, self), device=f"cuda:{self.device_index}"
            )
        return self._bytes_tensor

    def tensor(self, dtype, numel: int, byte_offset: int = 0):
        """Typed view ``[byte_offset, byte_offset + numel*itemsize)`` of the local copy."""
        import torch

        itemsize = torch.empty((), dtype=dtype).element_size()
        assert byte_offset % itemsize == 0 and byte_offset + numel * itemsize <= self.nbytes
        return self.bytes_tensor()[byte_offset: byte_offset + numel * itemsize].view(dtype)

    def ptrs(self, byte_offset: int = 0) -> List[int]:
        return [p + byte_offset for p in self.peer_ptrs]


class SymmetricHeap:
    def __init__(self, group: PartyGroup, backend: Optional[str] = None):
        self.group = group
        self.backend = backend or os.environ.get("RAYFED_B200_SYMM_BACKEND", "ipc")
        self._buffers = {}
        self._ext = cuda_ext()
        self._ext.set_device(group.device_index)
        self.multicast_ok = False
        self._vmm_info = None

    # -- public -----------------------------------------------------------------
    def alloc(self, name: str, nbytes: int, multicast: bool = False) -> SymmetricBuffer:
        """Collective: every party calls with the same ``name`` and ``nbytes``."""
        nbytes = (nbytes + 255) // 256 * 256
        self._ext.set_device(self.group.device_index)  # cudaSetDevice is per thread
        if self.backend == "vmm" or multicast:
            buf = self._alloc_vmm(name, nbytes, multicast)  # agrees internally; None => fall back
            if buf is not None:
                self._buffers[name] = buf
                return buf
            if self.backend == "vmm":
                raise RuntimeError("CUDA VMM symmetric allocation is not available on every party")
        buf = self._alloc_ipc(name, nbytes)
        self._buffers[name] = buf
        return buf

    def get(self, name: str) -> SymmetricBuffer:
        return self._buffers[name]

    # -- cudaIpc ------------------------------------------------------------------
    def _alloc_ipc(self, name: str, nbytes: int) -> SymmetricBuffer:
        g = self.group
        ptr, handle = self._ext.ipc_alloc(nbytes, True)
        self._ext.device_synchronize()
        # fixed layout (never pickle: the blob comes from a peer): u64 nbytes | u64 pid | 64-byte handle
        blobs = g.all_gather(g.next_tag(f"symm-ipc/{name}"), struct.pack("<QQ", nbytes, os.getpid()) + bytes(handle))
        peers = []
        for r, blob in enumerate(blobs):
            if r == g.rank:
                peers.append(ptr)
                continue
            if len(blob) != 16 + 64:
                raise RuntimeError(f"malformed cudaIpc handle record from party #{r} ({len(blob)} bytes)")
            pn, ppid = struct.unpack("<QQ", blob[:16])
            ph = blob[16:]
            assert pn == nbytes, f"symmetric alloc size mismatch for {name}: {pn} vs {nbytes}"
            if ppid == os.getpid():
                raise RuntimeError("cudaIpc handles cannot be opened in the exporting process")
            peers.append(self._ext.ipc_open(ph))
        g.barrier(g.next_tag(f"symm-ipc-open/{name}"))
        return SymmetricBuffer(name, nbytes, ptr, peers, g.rank, g.device_index, 0, "ipc")

    # -- VMM + multicast ------------------------------------------------------------
    def vmm_info(self) -> dict:
        if self._vmm_info is None:
            self._vmm_info = dict(self._ext.vmm_info(self.group.device_index))
        return self._vmm_info

    def _sock_path(self, party: str) -> str:
        base = os.environ.get("RAYFED_B200_SOCK_DIR", "/tmp")
        return os.path.join(base, f"rayfed_b200_fd_{os.getuid()}_{party}.sock")

    def _exchange_fds(self, tag: str, fd: int) -> List[int]:
        """All-gather one fd per party over Unix sockets (SCM_RIGHTS)."""
        g = self.group
        path = self._sock_path(g.me)
        try:
            os.unlink(path)
        except FileNotFoundError:
            pass
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(path)
        srv.listen(g.size)
        g.barrier(g.next_tag(f"fdsock/{tag}"))
        # send my fd to every peer
        for p in g.parties:
            if p == g.me:
                continue
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            deadline = time.time() + 30
            while True:
                try:
                    c.connect(self._sock_path(p))
                    break
                except (FileNotFoundError, ConnectionRefusedError):
                    if time.time() > deadline:
                        raise
                    time.sleep(0.01)
            socket.send_fds(c, [g.me.encode()], [fd])
            c.close()
        got = {}
        while len(got) < g.size - 1:
            conn, _ = srv.accept()
            msg, fds, _, _ = socket.recv_fds(conn, 256, 1)
            got[msg.decode()] = fds[0]
            conn.close()
        srv.close()
        try:
            os.unlink(path)
        except FileNotFoundError:
            pass
        return [fd if p == g.me else got[p] for p in g.parties]

    def _agree(self, name: str, ok: bool) -> bool:
        """Collective AND: every party takes the same branch after a step that can fail locally."""
        votes = self.group.all_gather(self.group.next_tag(f"agree/{name}"), b"1" if ok else b"0")
        return all(v == b"1" for v in votes)

    def _alloc_vmm(self, name: str, nbytes: int, multicast: bool) -> Optional[SymmetricBuffer]:
        """VMM-backed symmetric buffer, with an NVLS multicast alias when possible.

        Every step that can fail on one party only is followed by an agreement, so
        the parties never diverge in the exchange sequence: the result is the same
        on all of them — (buffer with mc) / (buffer without mc) / None (→ cudaIpc).
        """
        g = self.group
        info = {}
        ok = True
        try:
            info = self.vmm_info()
            ok = bool(info.get("vmm")) and bool(info.get("posix_fd"))
        except Exception:  # noqa: BLE001
            ok = False
        if not self._agree(f"{name}/vmm-cap", ok):
            return None
        want_mc = multicast and g.size > 1 and self._agree(f"{name}/mc-cap", bool(info.get("multicast")))
        size = nbytes
        ptr = fd = mapped = mem_handle = None
        try:
            if want_mc:
                gran = self._ext.mc_granularity(nbytes, g.size)
                size = (nbytes + gran - 1) // gran * gran
            ptr, fd, mapped, mem_handle = self._ext.vmm_alloc(size, g.device_index)
            self._ext.device_synchronize()  # the zero-fill must land before any peer can write
            ok = True
        except Exception as e:  # noqa: BLE001
            logger.warning(f"vmm_alloc failed: {e}")
            ok = False
        if not self._agree(f"{name}/alloc", ok):
            if fd is not None:
                os.close(fd)
            return None
        fds = self._exchange_fds(name, fd)
        peers = []
        try:
            for r, pfd in enumerate(fds):
                if r == g.rank:
                    peers.append(ptr)
                else:
                    va, _h = self._ext.vmm_import(pfd, mapped, g.device_index)
                    peers.append(va)
            ok = True
        except Exception as e:  # noqa: BLE001
            logger.warning(f"vmm_import failed: {e}")
            ok = False
        for r, pfd in enumerate(fds):
            if r != g.rank:
                try:
                    os.close(pfd)
                except OSError:
                    pass
        if not self._agree(f"{name}/import", ok):
            os.close(fd)
            return None

        mc_ptr = 0
        if want_mc:
            mc_handle, mc_fd = 0, -1
            try:
                if g.rank == 0:
                    mc_handle, mc_fd = self._ext.mc_create(mapped, g.size)
                ok = True
            except Exception as e:  # noqa: BLE001
                logger.warning(f"cuMulticastCreate failed: {e}")
                ok = False
            if self._agree(f"{name}/mc-create", ok):
                mc_fds = self._exchange_fds(name + "/mc", mc_fd if g.rank == 0 else fd)
                try:
                    if g.rank != 0:
                        mc_handle = self._ext.mc_import(mc_fds[0])
                    self._ext.mc_add_device(mc_handle, g.device_index)
                    ok = True
                except Exception as e:  # noqa: BLE001
                    logger.warning(f"multicast import/add_device failed: {e}")
                    ok = False
                for r, pfd in enumerate(mc_fds):
                    if r != g.rank:
                        try:
                            os.close(pfd)
                        except OSError:
                            pass
                if g.rank == 0 and mc_fd >= 0:
                    os.close(mc_fd)
                # all devices must be added before anyone binds
                if self._agree(f"{name}/mc-added", ok):
                    try:
                        mc_ptr = self._ext.mc_bind_and_map(mc_handle, mem_handle, mapped, g.device_index)
                        ok = True
                    except Exception as e:  # noqa: BLE001
                        logger.warning(f"cuMulticastBindMem/map failed: {e}")
                        ok = False
                    if not self._agree(f"{name}/mc-bound", ok):
                        mc_ptr = 0
            if mc_ptr:
                self.multicast_ok = True
        os.close(fd)
        return SymmetricBuffer(name, mapped, ptr, peers, g.rank, g.device_index, mc_ptr, "vmm")
