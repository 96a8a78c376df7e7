"""One tiny FedAvg round on a single GPU: exercises fwd/bwd, the fused local SGD
kernel, the fused aggregate kernel (N=1 degenerate all-reduce), push/unpack and the
tcgen05 LoRA merge."""
from __future__ import annotations

import torch


def run(device: str = "cuda:0"):
    from ..models import resnet18
    from ..ops import kernels as K
    from ..parallel.group import FedExchange, PartyGroup
    from ..parallel.symmetric import SymmetricHeap
    from .bucket import FlatBucket
    from .fedavg import FedAggregator, LocalSGD

    dev = torch.device(device)
    torch.cuda.set_device(dev)
    K.cuda_ext()  # fail loudly if the extension is missing
    group = PartyGroup(["solo"], "solo", FedExchange(), device_index=dev.index or 0)
    heap = SymmetricHeap(group)
    torch.manual_seed(0)
    model = resnet18(num_classes=10, width=16).to(dev).to(memory_format=torch.channels_last).to(torch.bfloat16)
    bucket = FlatBucket(model, heap, name="smoke", n_parties=1)
    opt = LocalSGD(bucket, lr=0.05, momentum=0.9)
    agg = FedAggregator(bucket, group, None, kind="fedavgm", lr=1.0, momentum=0.0)
    x = torch.randn(8, 3, 64, 64, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,), device=dev)
    w0 = bucket.flat_w.clone()
    loss = torch.nn.functional.cross_entropy(model(x).float(), y)
    loss.backward()
    opt.step()
    agg.aggregate()
    # push/unpack loopback through the heap
    buf = heap.alloc("smoke-arena", 1 << 20)
    ctrl = torch.zeros(64, dtype=torch.int32, device=dev)
    src = torch.arange(65536, dtype=torch.int32, device=dev)
    K.push(src.data_ptr(), buf.local_ptr, src.numel() * 4, flag_ptr=ctrl.data_ptr(), flag_value=1,
           done_counter_ptr=ctrl.data_ptr() + 8, device_index=dev.index or 0)
    dst = torch.empty_like(src)
    K.unpack(buf.local_ptr, dst.data_ptr(), src.numel() * 4, flag_ptr=ctrl.data_ptr(), flag_value=1,
             done_counter_ptr=ctrl.data_ptr() + 12, status_ptr=ctrl.data_ptr() + 16, timeout_ns=int(5e9),
             device_index=dev.index or 0)
    # the host-free receive path, loopback: multi-tensor push with a device descriptor, the
    # consumer stream ordered by cuStreamWaitValue32, descriptor check, zero-copy view, credit kernel
    from ..proxy.nvlink import channel as C

    ext = K.cuda_ext()
    ext.preload_kernels()
    status = torch.zeros(4, dtype=torch.int32).pin_memory()
    t1 = torch.arange(4096, dtype=torch.float32, device=dev)
    t2 = torch.arange(1024, dtype=torch.float32, device=dev) * 2
    desc = C.pack_desc(seq=1, job_hash=C.fnv1a32("smoke"), key=(1, 0, 2), offset=0, nbytes=C.HEADER + 16384 + 4096,
                       dtype="float32", ndim=1, n_tensors=2, shape=(4096,))
    block = buf.local_ptr + 65536 * 4
    recv = ext.stream_create(0)
    credit_words = torch.zeros(8, dtype=torch.int64, device=dev)
    ext.stream_wait_value32(recv, ctrl.data_ptr() + 32, 1)  # parked BEFORE the push is launched
    ext.desc_check(block, desc, status.data_ptr(), recv)
    ev = ext.event_create(False)
    ext.event_record(ev, recv)
    K.push_multi([(t1.data_ptr(), C.HEADER, 16384), (t2.data_ptr(), C.HEADER + 16384, 4096)], block,
                 flag_ptr=ctrl.data_ptr() + 32, flag_value=1, done_counter_ptr=ctrl.data_ptr() + 8,
                 device_index=dev.index or 0, desc_dst=block, desc=desc)
    ext.stream_wait_event(torch.cuda.current_stream().cuda_stream, ev)
    view = buf.tensor(torch.float32, 4096, 65536 * 4 + C.HEADER)
    got1 = view.clone()
    ext.credit(credit_words.data_ptr(), 7, credit_words.data_ptr() + 8, 1, torch.cuda.current_stream().cuda_stream)
    # exact LoRA merge on the tensor cores (tcgen05/TMEM, TMA-fed persistent pipeline)
    m = n = 256
    r = 64
    B = (torch.randn(m, r, device=dev) * 0.25).to(torch.bfloat16)
    At = (torch.randn(n, r, device=dev) * 0.25).to(torch.bfloat16)
    W = torch.randn(m, n, device=dev).to(torch.bfloat16)
    Wm = W.clone()
    K.lora_merge(b_ptrs=[B.data_ptr()], at_ptrs=[At.data_ptr()], out_ptr=Wm.data_ptr(), m=m, n=n, r=r, scale=0.5,
                 w_base_ptr=Wm.data_ptr())
    torch.cuda.synchronize()
    ref = (W.float() + 0.5 * (B.float() @ At.float().t())).to(torch.bfloat16)
    assert (Wm.float() - ref.float()).abs().max().item() <= 2 ** -5 * max(1.0, ref.float().abs().max().item()), \
        "tcgen05 LoRA merge mismatch"
    assert torch.equal(src, dst), "push/unpack loopback mismatch"
    assert torch.equal(got1, t1) and int(status[0]) == 0, "descriptor / stream-wait receive path mismatch"
    assert credit_words.tolist()[:2] == [7, 1], "credit kernel mismatch"
    ext.event_destroy(ev)
    ext.stream_destroy(recv)
    assert torch.isfinite(loss).item()
    assert not torch.equal(w0, bucket.flat_w), "weights did not change"
    print(f"smoke ok: loss={loss.item():.4f} launches={K.launch_counter.by_kernel}")
