"""Memory-model litmus tests of the payload→flag protocols.

Message passing, many times, against a *poisoned* destination: the producer's kernel stores a
payload and then publishes a flag; the consumer's stream waits for the flag with a stream memory
operation (``cuStreamWaitValue32``) and must then observe the whole payload — a single stale byte
is a failure.  Variants: SM-issued vector stores + ``fence.acq_rel.sys`` + ``st.release.sys``
(``fed_push_kernel``), TMA-engine bulk stores + ``cp.async.bulk.wait_group`` + proxy fence
(``fed_push_bulk_kernel``), the multi-segment kernel, and — across two party processes —
``multimem.st`` through the NVSwitch when a multicast mapping exists (``fed_push_mc_kernel``).

``RAYFED_B200_LITMUS_ITERS`` scales the iteration count (default 20 000 per variant; the
committed run used 100 000, see profiles/sanitizers/README.md).
"""
import os

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

ITERS = int(os.environ.get("RAYFED_B200_LITMUS_ITERS", "20000"))


def _loopback_litmus(variant: str, nbytes: int, iters: int):
    from rayfed_b200.ops import kernels as K

    ext = K.cuda_ext()
    ext.preload_kernels()
    dev = torch.cuda.current_device()
    arena = torch.full((nbytes + 4096,), 0xEE, dtype=torch.uint8, device="cuda")  # payload at +4096
    ctrl = torch.zeros(256, dtype=torch.int32, device="cuda")  # [0] data flag, [16] consumed flag, [32] done ctr
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    srcs = [torch.full((nbytes,), v, dtype=torch.uint8, device="cuda") for v in range(1, 9)]
    torch.cuda.synchronize()
    prod = ext.stream_create(0)
    cons = ext.stream_create(0)
    tcons = torch.cuda.ExternalStream(cons)
    flag = ctrl.data_ptr()
    consumed = ctrl.data_ptr() + 64
    done = ctrl.data_ptr() + 128
    payload = arena[4096:4096 + nbytes]
    dst = arena.data_ptr() + 4096
    for i in range(1, iters + 1):
        src = srcs[i % 8]
        # producer: wait until the consumer has poisoned the arena again, then push + flag
        if i > 1:
            ext.stream_wait_value32(prod, consumed, i - 1)
        if variant == "multi":
            K.push_multi([(src.data_ptr(), 0, nbytes // 2), (src.data_ptr() + nbytes // 2, nbytes // 2, nbytes // 2)],
                         dst, flag_ptr=flag, flag_value=i, done_counter_ptr=done, stream=prod, device_index=dev)
        else:
            K.push(src.data_ptr(), dst, nbytes, flag_ptr=flag, flag_value=i, done_counter_ptr=done, stream=prod,
                   device_index=dev, use_bulk=(variant == "bulk"))
        # consumer: K4 wait, then every byte must already be the new value
        ext.stream_wait_value32(cons, flag, i)
        with torch.cuda.stream(tcons):
            bad += (payload != src[0]).sum()
            payload.fill_(0xEE)  # poison for the next round
        ext.stream_write_value32(cons, consumed, i)
    ext.stream_synchronize(cons)
    ext.stream_synchronize(prod)
    n_bad = int(bad.item())
    ext.stream_destroy(prod)
    ext.stream_destroy(cons)
    return n_bad


@pytest.mark.parametrize("variant,nbytes", [("sm", 64 << 10), ("sm", 4 << 20), ("bulk", 4 << 20), ("multi", 1 << 20)])
def test_payload_is_visible_once_the_flag_is(variant, nbytes):
    iters = ITERS if nbytes <= (1 << 20) else max(2000, ITERS // 10)
    assert _loopback_litmus(variant, nbytes, iters) == 0


def _mp_litmus(party, addresses, iters):
    """Across two party processes, through the channel: unicast pushes and (when the box has an
    NVLS mapping) multicast pushes; every received tensor must be entirely the announced value."""
    import rayfed_b200 as fed
    from test_gpu_fed_nvlink import _init

    fed, dev = _init(party, addresses, device_arena_bytes=64 << 20, device_zero_copy_min_bytes=1 << 30)

    @fed.remote
    class Src:
        def make(self, i, n):
            return torch.full((n,), float(i), device="cuda")

    @fed.remote
    class Dst:
        def __init__(self):
            self.bad = 0

        def check(self, t, i):
            self.bad += int((t != float(i)).sum())
            return self.bad

    src = Src.party("alice").remote()
    dst = Dst.party("bob").remote()
    last = None
    for i in range(iters):
        last = dst.check.remote(src.make.remote(i, 65536 if i % 3 else 1 << 20), i)  # unicast alice -> bob
        if i % 16 == 0:
            got = fed.get(src.make.remote(i, 1 << 18))  # broadcast (multicast push when available)
            assert bool((got == float(i)).all())
    assert fed.get(last) == 0
    fed.shutdown()


def test_payload_is_visible_across_parties():
    from party_utils import make_addresses, run_parties

    addresses = make_addresses(["alice", "bob"])
    run_parties(_mp_litmus, ["alice", "bob"], args=(addresses, max(300, ITERS // 40)), timeout=240)
