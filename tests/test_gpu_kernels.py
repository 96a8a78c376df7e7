"""GPU numerics: every hand-written kernel vs a plain PyTorch fp32 reference.

Single-GPU tests use loopback (the "peer" pointer is local memory) — the kernels
do not care which GPU a pointer maps to.  Multi-party tests spawn one process
per party and need >= 2 GPUs.
"""
import json
import os
import tempfile

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from rayfed_b200.ops import kernels as K  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ctrl():
    t = torch.zeros(4096, dtype=torch.int32, device=DEV)
    return t


def _flags(ctrl):
    base = ctrl.data_ptr()
    return dict(flag=base, done=base + 64, undone=base + 68, status=base + 72, sig=base + 1024)


def test_extension_is_native_and_loaded():
    ext = K.cuda_ext()
    assert ext.__file__.endswith(".so")
    assert K.sm_count(0) == 148 or K.sm_count(0) > 0


@pytest.mark.parametrize("nbytes", [16, 1000, 4096 + 7, 1 << 20, (1 << 22) + 16])
@pytest.mark.parametrize("bulk", [False, True])
def test_push_unpack_bit_exact(ctrl, nbytes, bulk):
    if bulk and nbytes % 16:
        pytest.skip("bulk path needs 16-byte multiples")
    f = _flags(ctrl)
    ctrl.zero_()
    src = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=DEV)
    arena = torch.zeros(nbytes + 64, dtype=torch.uint8, device=DEV)
    dst = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
    K.push(src.data_ptr(), arena.data_ptr(), nbytes, flag_ptr=f["flag"], flag_value=7, done_counter_ptr=f["done"],
           use_bulk=bulk)
    K.unpack(arena.data_ptr(), dst.data_ptr(), nbytes, flag_ptr=f["flag"], flag_value=7,
             done_counter_ptr=f["undone"], status_ptr=f["status"], timeout_ns=int(5e9))
    torch.cuda.synchronize()
    assert torch.equal(src, dst)
    assert ctrl[0].item() == 7  # flag published
    assert ctrl[18].item() == 0  # no timeout
    assert arena[nbytes:].sum().item() == 0  # no overrun


def test_unpack_waits_for_flag_on_device(ctrl):
    """K4: the unpack kernel parks on the flag and only then copies.  The flag lives in
    pinned host memory and is flipped by a CPU store, so no second GPU stream is needed
    (a later-submitted kernel may share a front-end queue with the spinning one — the
    hazard described in profiles/README.md §8)."""
    f = _flags(ctrl)
    ctrl.zero_()
    n = 1 << 16
    src = torch.arange(n, dtype=torch.int32, device=DEV)
    arena = src.clone()  # the payload is already in the arena; only the flag is missing
    dst = torch.zeros(n, dtype=torch.int32, device=DEV)
    flag = torch.zeros(4, dtype=torch.int32).pin_memory()
    s_recv = torch.cuda.Stream()
    torch.cuda.synchronize()
    K.unpack(arena.data_ptr(), dst.data_ptr(), n * 4, flag_ptr=flag.data_ptr(), flag_value=1,
             done_counter_ptr=f["undone"], status_ptr=f["status"], timeout_ns=int(20e9), stream=s_recv)
    import time

    time.sleep(0.05)
    assert not s_recv.query()  # still waiting on the device
    flag[0] = 1  # host store, observed by ld.acquire.sys over PCIe
    s_recv.synchronize()
    assert torch.equal(src, dst)
    assert ctrl[18].item() == 0


def test_unpack_timeout_sets_status(ctrl):
    f = _flags(ctrl)
    ctrl.zero_()
    a = torch.zeros(64, dtype=torch.uint8, device=DEV)
    K.unpack(a.data_ptr(), a.data_ptr(), 16, flag_ptr=f["flag"], flag_value=5, done_counter_ptr=f["undone"],
             status_ptr=f["status"], timeout_ns=int(2e6))
    torch.cuda.synchronize()
    assert ctrl[18].item() == 1


def test_push_cast_f32_to_bf16_and_back(ctrl):
    f = _flags(ctrl)
    ctrl.zero_()
    n = 100_003
    x = torch.randn(n, device=DEV)
    wire = torch.zeros(n, dtype=torch.bfloat16, device=DEV)
    K.push(x.data_ptr(), wire.data_ptr(), n, mode=K.MODE_F32_TO_BF16, scale=0.5, flag_ptr=f["flag"], flag_value=1,
           done_counter_ptr=f["done"])
    y = torch.zeros(n, device=DEV)
    K.unpack(wire.data_ptr(), y.data_ptr(), n, mode=K.MODE_BF16_TO_F32, flag_ptr=f["flag"], flag_value=1,
             done_counter_ptr=f["undone"], status_ptr=f["status"], timeout_ns=int(5e9))
    torch.cuda.synchronize()
    ref = (x * 0.5).to(torch.bfloat16)
    assert torch.equal(wire, ref)
    assert torch.equal(y, ref.float())


def test_push_bf16_scale(ctrl):
    f = _flags(ctrl)
    ctrl.zero_()
    n = 8 * 1000 + 3
    x = torch.randn(n, device=DEV).to(torch.bfloat16)
    out = torch.zeros_like(x)
    K.push(x.data_ptr(), out.data_ptr(), n, mode=K.MODE_BF16_SCALE, scale=0.25, flag_ptr=f["flag"], flag_value=1,
           done_counter_ptr=f["done"])
    torch.cuda.synchronize()
    assert torch.equal(out, (x.float() * 0.25).to(torch.bfloat16))


@pytest.mark.parametrize("n", [32, 32 * 7, 1 << 16, (1 << 18) + 32 * 5])
def test_fp8_block_scaled_wire_format_matches_torch(ctrl, n):
    f = _flags(ctrl)
    ctrl.zero_()
    x = (torch.randn(n, device=DEV) * torch.logspace(-3, 3, n, device=DEV)).to(torch.bfloat16)
    payload = torch.zeros(n, dtype=torch.uint8, device=DEV)
    scales = torch.zeros(n // 32, dtype=torch.uint8, device=DEV)
    K.push(x.data_ptr(), payload.data_ptr(), n, mode=K.MODE_BF16_TO_FP8, dst_scales_ptr=scales.data_ptr(),
           flag_ptr=f["flag"], flag_value=1, done_counter_ptr=f["done"])
    y = torch.zeros(n, dtype=torch.bfloat16, device=DEV)
    K.unpack(payload.data_ptr(), y.data_ptr(), n, mode=K.MODE_FP8_TO_BF16, scales_ptr=scales.data_ptr(),
             flag_ptr=f["flag"], flag_value=1, done_counter_ptr=f["undone"], status_ptr=f["status"],
             timeout_ns=int(5e9))
    torch.cuda.synchronize()
    xf = x.float().view(-1, 32)
    amax = xf.abs().amax(dim=1, keepdim=True)
    e = torch.ceil(torch.log2(amax.clamp_min(2.0 ** -120) / 448.0)).clamp(-126, 126)
    sc = torch.exp2(e)
    ref_q = (xf / sc).to(torch.float8_e4m3fn)
    assert torch.equal(scales.view(-1, 1).int() - 127, e.int())
    assert torch.equal(payload.view(torch.float8_e4m3fn).view(-1, 32).float(), ref_q.float())
    assert torch.equal(y.float().view(-1, 32), (ref_q.float() * sc).to(torch.bfloat16).float())


def test_local_sgd_matches_torch():
    n = 8 * 12345
    w = torch.randn(n, device=DEV).to(torch.bfloat16)
    g = torch.randn(n, device=DEV).to(torch.bfloat16)
    m = torch.randn(n, device=DEV)
    w0, g0, m0 = w.float().clone(), g.float().clone(), m.clone()
    K.local_sgd(w.data_ptr(), g.data_ptr(), m.data_ptr(), n, lr=0.1, momentum=0.9, weight_decay=0.01,
                grad_scale=0.5, zero_grad=True)
    torch.cuda.synchronize()
    gg = g0 * 0.5 + 0.01 * w0
    m_ref = 0.9 * m0 + gg
    w_ref = (w0 - 0.1 * m_ref).to(torch.bfloat16)
    assert torch.allclose(m, m_ref, rtol=1e-5, atol=1e-6)
    # the kernel contracts to FMAs: allow a 1-ulp bf16 rounding flip on a tiny fraction
    diff = (w.float() - w_ref.float()).abs()
    assert (diff > 0).float().mean().item() < 1e-3
    assert (diff <= 2.0 ** -7 * w_ref.float().abs().clamp_min(1e-3)).all()
    assert g.abs().sum().item() == 0


def _agg_single(kind, n=32 * 1000, **kw):
    """N=1 degenerate aggregate: exercises the optimizer math and the store path."""
    pad = torch.zeros(16384, dtype=torch.int32, device=DEV)
    x = torch.randn(n, device=DEV).to(torch.bfloat16)
    buf = x.clone()
    K.aggregate(inputs=[buf.data_ptr()], outputs=[buf.data_ptr()], my_rank=0, shard_start=0, shard_len=n, kind=kind,
                sig_ptr=pad.data_ptr(), sig_peers=[pad.data_ptr()], epoch=0, status_ptr=pad.data_ptr() + 60000,
                timeout_ns=int(5e9), **kw)
    torch.cuda.synchronize()
    return x, buf


def test_aggregate_avg_identity_n1():
    x, out = _agg_single(K.AGG_AVG)
    assert torch.equal(x, out)


def test_aggregate_fedavgm_n1():
    n = 32 * 999
    master = torch.randn(n, device=DEV)
    mom = torch.randn(n, device=DEV)
    m0, q0 = master.clone(), mom.clone()
    x, out = _agg_single(K.AGG_FEDAVGM, n=n, master_ptr=master.data_ptr(), mom_ptr=mom.data_ptr(), lr=0.7,
                         momentum=0.5)
    d = m0 - x.float()
    q = 0.5 * q0 + d
    ref = m0 - 0.7 * q
    assert torch.allclose(master, ref, rtol=1e-6, atol=1e-6)
    assert torch.equal(out, ref.to(torch.bfloat16))


def test_aggregate_adam_n1_matches_torch_adamw():
    n = 32 * 500
    w0 = torch.randn(n, device=DEV)
    master = w0.clone()
    mom = torch.zeros(n, device=DEV)
    var = torch.zeros(n, device=DEV)
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.AdamW([p], lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1)
    pad = torch.zeros(16384, dtype=torch.int32, device=DEV)
    out = torch.zeros(n, dtype=torch.bfloat16, device=DEV)
    for step in range(1, 4):
        g = torch.randn(n, device=DEV).to(torch.bfloat16)
        p.grad = g.float()
        opt.step()
        K.aggregate(inputs=[g.data_ptr()], outputs=[out.data_ptr()], my_rank=0, shard_start=0, shard_len=n,
                    kind=K.AGG_ADAM, master_ptr=master.data_ptr(), mom_ptr=mom.data_ptr(), var_ptr=var.data_ptr(),
                    lr=1e-2, beta1=0.9, beta2=0.99, eps=1e-8, weight_decay=0.1, step=step, sig_ptr=pad.data_ptr(),
                    sig_peers=[pad.data_ptr()], epoch=2 * step, status_ptr=pad.data_ptr() + 60000,
                    timeout_ns=int(5e9))
        torch.cuda.synchronize()
        assert torch.allclose(master, p.detach(), rtol=2e-5, atol=2e-6), step
    assert torch.equal(out, master.to(torch.bfloat16))


def test_aggregate_fp8_out_n1():
    n = 32 * 64
    pad = torch.zeros(16384, dtype=torch.int32, device=DEV)
    x = (torch.randn(n, device=DEV) * 2).to(torch.bfloat16)
    out = torch.zeros(n, dtype=torch.uint8, device=DEV)
    scales = torch.zeros(n // 32, dtype=torch.uint8, device=DEV)
    K.aggregate(inputs=[x.data_ptr()], outputs=[out.data_ptr()], out_scales=[scales.data_ptr()], out_fp8=True,
                my_rank=0, shard_start=0, shard_len=n, kind=K.AGG_AVG, sig_ptr=pad.data_ptr(),
                sig_peers=[pad.data_ptr()], epoch=0, status_ptr=pad.data_ptr() + 60000, timeout_ns=int(5e9))
    torch.cuda.synchronize()
    xf = x.float().view(-1, 32)
    e = torch.ceil(torch.log2(xf.abs().amax(1, keepdim=True).clamp_min(2.0 ** -120) / 448.0))
    ref = (xf / torch.exp2(e)).to(torch.float8_e4m3fn)
    assert torch.equal(scales.int() - 127, e.view(-1).int())
    assert torch.equal(out.view(torch.float8_e4m3fn).view(-1, 32).float(), ref.float())


def test_two_virtual_parties_on_one_gpu_rendezvous_in_kernel():
    """Two 'parties' = two streams on one GPU: the in-kernel signal-pad barrier and
    the cross-party reduce are exercised without a second device."""
    n = 32 * 2 * 4096
    shard = n // 2
    xs = [torch.randn(n, device=DEV).to(torch.bfloat16) for _ in range(2)]
    bufs = [x.clone() for x in xs]
    pads = [torch.zeros(16384, dtype=torch.int32, device=DEV) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for r in range(2):
        K.aggregate(inputs=[b.data_ptr() for b in bufs], outputs=[b.data_ptr() for b in bufs], my_rank=r,
                    shard_start=r * shard, shard_len=shard, kind=K.AGG_AVG, sig_ptr=pads[r].data_ptr(),
                    sig_peers=[p.data_ptr() for p in pads], epoch=0, status_ptr=pads[r].data_ptr() + 60000,
                    timeout_ns=int(3e9), stream=streams[r], blocks=32)
    torch.cuda.synchronize()
    if pads[0][15000].item() or pads[1][15000].item():
        # the second launch was queued behind the first one (both streams mapped to the same
        # GPU front-end queue), so the two kernels could not rendezvous: not a kernel bug.
        pytest.skip("streams shared a hardware queue; the two virtual parties could not co-run")
    ref = ((xs[0].float() + xs[1].float()) * 0.5).to(torch.bfloat16)
    assert torch.equal(bufs[0], ref) and torch.equal(bufs[1], ref)


@pytest.mark.parametrize("nv", [4, 8])
def test_virtual_parties_on_one_gpu_pipelined_fedavg(nv):
    """N 'parties' = N streams on one GPU, each launching the pipelined FedAvg kernel of its rank
    (`fed_allreduce_avg_kernel<false, 4, 2>` / `<false, 8, 1>`): weighted mean, in place, identical
    result in every party's buffer."""
    per = 32 * 512  # elements per shard
    n = per * nv
    xs = [torch.randn(n, device=DEV).to(torch.bfloat16) for _ in range(nv)]
    bufs = [x.clone() for x in xs]
    pads = [torch.zeros(16384, dtype=torch.int32, device=DEV) for _ in range(nv)]
    ext = K.cuda_ext()
    ext.preload_kernels()
    streams = [ext.stream_create(0) for _ in range(nv)]
    weights = [(i + 1.0) for i in range(nv)]
    tot = sum(weights)
    weights = [w / tot for w in weights]
    torch.cuda.synchronize()
    for r in range(nv):
        K.aggregate(inputs=[b.data_ptr() for b in bufs], outputs=[b.data_ptr() for b in bufs], my_rank=r,
                    shard_start=r * per, shard_len=per, kind=K.AGG_AVG, weights=weights, sig_ptr=pads[r].data_ptr(),
                    sig_peers=[p.data_ptr() for p in pads], epoch=0, status_ptr=pads[r].data_ptr() + 60000,
                    timeout_ns=int(3e9), stream=streams[r], blocks=4)
    torch.cuda.synchronize()
    for h in streams:
        ext.stream_destroy(h)
    if any(p[15000].item() for p in pads):
        pytest.skip("the virtual parties' streams could not co-run on this GPU")
    ref = torch.zeros(n, device=DEV)
    for shard in range(nv):
        sl = slice(shard * per, (shard + 1) * per)
        a = torch.zeros(per, device=DEV)
        for j in range(nv):
            k = (shard + j) % nv  # the owner of a shard sums the parties in rotation order from itself
            a += weights[k] * xs[k][sl].float()
        ref[sl] = a
    for b in bufs[1:]:
        assert torch.equal(b, bufs[0])  # every party holds the same bits
    # fp32 accumulation in the same order, but FFMA vs mul+add: allow the last bf16 bit
    err = (bufs[0].float() - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -7 + 1e-6).all()), float(err.max())


# -- K3: tcgen05/TMEM LoRA merge -----------------------------------------------------------------
@pytest.mark.parametrize("impl", ["staged3", "staged", "tma", "direct"])
@pytest.mark.parametrize("parties,r,m,n", [(2, 16, 256, 256), (8, 16, 512, 384), (4, 32, 128, 1024), (1, 64, 200, 96),
                                           (8, 16, 200, 192)])
def test_lora_merge_tcgen05_matches_torch(parties, r, m, n, impl):
    torch.manual_seed(parties * 1000 + r)
    Bs = [(torch.randn(m, r, device=DEV) * 0.5).to(torch.bfloat16) for _ in range(parties)]
    Ats = [(torch.randn(n, r, device=DEV) * 0.5).to(torch.bfloat16) for _ in range(parties)]
    base = torch.randn(m, n, device=DEV).to(torch.bfloat16)
    out = torch.zeros(m, n, dtype=torch.bfloat16, device=DEV)
    weights = [1.0] * parties
    weights[0] = 0.5
    K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats], out_ptr=out.data_ptr(),
                 m=m, n=n, r=r, scale=1.0 / parties, weights=weights, w_base_ptr=base.data_ptr(), impl=impl)
    torch.cuda.synchronize()
    dw = torch.zeros(m, n, device=DEV)
    for i, (b, at) in enumerate(zip(Bs, Ats)):
        bw = (b.float() * weights[i]).to(torch.bfloat16).float() if weights[i] != 1.0 else b.float()
        dw += bw @ at.float().t()
    ref = (base.float() + dw / parties).to(torch.bfloat16)
    err = (out.float() - ref.float()).abs().max().item()
    assert err <= 2 ** -6 * max(1.0, ref.float().abs().max().item()), err
    assert torch.equal(out, ref) or (out.float() - ref.float()).abs().mean().item() < 1e-3


@pytest.mark.parametrize("has_base", [True, False])
def test_lora_merge_staged_persistent_ring_wraps_in_place(has_base):
    """512 tiles on <=148 persistent CTAs: every smem/TMEM stage is reused (both mbarrier
    parities), and the update runs in place (out aliases W_base) as FedLoRAExactMerge does."""
    parties, r, m, n = 8, 16, 4096, 2048
    torch.manual_seed(7)
    Bs = [(torch.randn(m, r, device=DEV) * 0.5).to(torch.bfloat16) for _ in range(parties)]
    Ats = [(torch.randn(n, r, device=DEV) * 0.5).to(torch.bfloat16) for _ in range(parties)]
    w = torch.randn(m, n, device=DEV).to(torch.bfloat16)
    dw = torch.cat(Bs, 1).float() @ torch.cat(Ats, 1).float().t()
    ref = ((w.float() if has_base else 0) + dw / parties).to(torch.bfloat16)
    for _ in range(2):  # second call reuses the cached workspace
        out = w.clone()
        K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats], out_ptr=out.data_ptr(),
                     m=m, n=n, r=r, w_base_ptr=out.data_ptr() if has_base else 0, impl="staged")
        torch.cuda.synchronize()
        err = (out.float() - ref.float()).abs().max().item()
        assert err <= 2 ** -6 * max(1.0, ref.float().abs().max().item()), err
        assert (out.float() - ref.float()).abs().mean().item() < 1e-3


def test_lora_merge_fp8_base_and_out():
    parties, r, m, n = 4, 16, 256, 256
    torch.manual_seed(7)
    Bs = [(torch.randn(m, r, device=DEV) * 0.3).to(torch.bfloat16) for _ in range(parties)]
    Ats = [(torch.randn(n, r, device=DEV) * 0.3).to(torch.bfloat16) for _ in range(parties)]
    w = torch.randn(m, n, device=DEV)
    e = torch.ceil(torch.log2(w.view(m, -1, 32).abs().amax(-1, keepdim=True) / 448.0))
    wq = (w.view(m, -1, 32) / torch.exp2(e)).to(torch.float8_e4m3fn)
    base_q = wq.view(m, n).view(torch.uint8).contiguous()
    base_s = (e.view(m, n // 32) + 127).to(torch.uint8).contiguous()
    out_q = torch.zeros(m, n, dtype=torch.uint8, device=DEV)
    out_s = torch.zeros(m, n // 32, dtype=torch.uint8, device=DEV)
    K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats], out_ptr=out_q.data_ptr(),
                 m=m, n=n, r=r, w_base_ptr=base_q.data_ptr(), w_base_scales_ptr=base_s.data_ptr(),
                 out_scales_ptr=out_s.data_ptr(), base_fp8=True, out_fp8=True)
    torch.cuda.synchronize()
    base_deq = (wq.float() * torch.exp2(e)).view(m, n)
    dw = sum(b.float() @ at.float().t() for b, at in zip(Bs, Ats)) / parties
    ref = base_deq + dw
    got = (out_q.view(torch.float8_e4m3fn).float().view(m, -1, 32)
           * torch.exp2(out_s.float().view(m, -1, 1) - 127)).view(m, n)
    rel = ((got - ref).abs().mean() / ref.abs().mean()).item()
    assert rel < 0.04, rel  # e4m3 has 3 mantissa bits
    e_ref = torch.ceil(torch.log2(ref.view(m, -1, 32).abs().amax(-1) / 448.0))
    assert (out_s.float() - 127 - e_ref).abs().max().item() <= 1


def test_timed_out_wait_leaves_destination_untouched_and_flags_pinned_status():
    """A device wait that times out must not hand out garbage: no copy, credit still
    returned, and the host sees the failure through a pinned status word (no CUDA call)."""
    status = torch.zeros(16, dtype=torch.int32).pin_memory()
    ctrlw = torch.zeros(64, dtype=torch.int32, device=DEV)
    credit = torch.zeros(2, dtype=torch.int64, device=DEV)
    arena = torch.full((4096,), 7, dtype=torch.uint8, device=DEV)
    dst = torch.full((4096,), 42, dtype=torch.uint8, device=DEV)
    K.unpack(arena.data_ptr(), dst.data_ptr(), 4096, flag_ptr=ctrlw.data_ptr(), flag_value=3,
             done_counter_ptr=ctrlw.data_ptr() + 8, credit_ptr=credit.data_ptr(), credit_value=1234,
             status_ptr=status.data_ptr(), timeout_ns=int(3e6))
    torch.cuda.synchronize()
    assert int(status[0]) == 1  # visible to the host without a device read
    assert torch.all(dst == 42)  # destination untouched
    assert credit[0].item() == 1234  # flow control still advances
