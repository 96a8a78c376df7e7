"""Every hand-written kernel once, on small loopback problems, with results checked — the target of
``compute-sanitizer --tool memcheck|racecheck|synccheck`` (tools/sanitize_kernels.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rayfed_b200.ops import kernels as K  # noqa: E402
from rayfed_b200.proxy.nvlink import channel as C  # noqa: E402

dev = "cuda:0"
ext = K.cuda_ext()
ctrl = torch.zeros(16384, dtype=torch.int32, device=dev)
base = ctrl.data_ptr()
cur = torch.cuda.current_stream().cuda_stream

# K1 push (all modes) + unpack
n = 1 << 16
src = torch.randn(n, device=dev)
arena = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
K.push(src.data_ptr(), arena.data_ptr() + 256, n * 4, flag_ptr=base, flag_value=1, done_counter_ptr=base + 64)
out = torch.empty_like(src)
K.unpack(arena.data_ptr() + 256, out.data_ptr(), n * 4, flag_ptr=base, flag_value=1, done_counter_ptr=base + 68,
         status_ptr=base + 72, timeout_ns=int(5e9))
torch.cuda.synchronize()
assert torch.equal(src, out)
K.push(src.data_ptr(), arena.data_ptr() + 256, n, mode=K.MODE_F32_TO_BF16, flag_ptr=base, flag_value=2,
       done_counter_ptr=base + 64)
b16 = src.to(torch.bfloat16)
pay = torch.zeros(n, dtype=torch.uint8, device=dev)
sc = torch.zeros(n // 32, dtype=torch.uint8, device=dev)
K.push(b16.data_ptr(), pay.data_ptr(), n, mode=K.MODE_BF16_TO_FP8, dst_scales_ptr=sc.data_ptr(), flag_ptr=base,
       flag_value=3, done_counter_ptr=base + 64)
back = torch.empty(n, dtype=torch.bfloat16, device=dev)
K.unpack(pay.data_ptr(), back.data_ptr(), n, mode=K.MODE_FP8_TO_BF16, scales_ptr=sc.data_ptr(), flag_ptr=0,
         flag_value=0, done_counter_ptr=base + 68, status_ptr=base + 72)
K.push(src.data_ptr(), arena.data_ptr() + 256, n * 4, flag_ptr=base, flag_value=4, done_counter_ptr=base + 64,
       use_bulk=True)
# multi-tensor push with a descriptor + check + credit
desc = C.pack_desc(seq=5, job_hash=7, key=(1, 0, 2), offset=0, nbytes=256 + 2 * 65536)
K.push_multi([(src.data_ptr(), 256, 65536), (src.data_ptr() + 65536, 256 + 65536, 65536)], arena.data_ptr(),
             flag_ptr=base + 4, flag_value=5, done_counter_ptr=base + 64, desc_dst=arena.data_ptr(), desc=desc)
ext.stream_wait_value32(cur, base + 4, 5)
ext.desc_check(arena.data_ptr(), desc, base + 72, cur)
ext.credit(base + 1024, 9, base + 1032, 1, cur)
torch.cuda.synchronize()
assert int(ctrl[18]) == 0, "descriptor check raised the status word"
# aggregate: pipelined FedAvg + generic (fedavgm, adam) at N=1, local SGD
m = 1 << 18
w = torch.randn(m, device=dev).to(torch.bfloat16)
for kind, extra in ((K.AGG_AVG, {}), (K.AGG_FEDAVGM, {"momentum": 0.9}), (K.AGG_ADAM, {})):
    master, mom, var = w.float().clone(), torch.zeros(m, device=dev), torch.zeros(m, device=dev)
    buf = w.clone()
    plan = K.AggregatePlan(inputs=[buf.data_ptr()], outputs=[buf.data_ptr()], my_rank=0, shard_start=0, shard_len=m,
                           kind=kind, master_ptr=master.data_ptr(), mom_ptr=mom.data_ptr(), var_ptr=var.data_ptr(),
                           lr=0.1, sig_ptr=base + 4096, sig_peers=[base + 4096], epoch_dev_ptr=base + 2048,
                           agg_done_ptr=base + 2052, status_ptr=base + 72, timeout_ns=int(5e9), **extra)
    plan.launch(0, 1)
    plan.launch(0, 2)
g = torch.randn(m, device=dev).to(torch.bfloat16)
mo = torch.zeros(m, device=dev)
K.local_sgd(w.data_ptr(), g.data_ptr(), mo.data_ptr(), m, lr=0.01, momentum=0.9)
K.barrier(n_parties=1, my_rank=0, sig_ptr=base + 4096, sig_peers=[base + 4096], epoch=100, status_ptr=base + 72,
          timeout_ns=int(5e9))
# K3: every LoRA merge implementation on a small problem
P, r, M, N = 4, 16, 256, 256
Bs = [(torch.randn(M, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
Ats = [(torch.randn(N, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
W = torch.randn(M, N, device=dev).to(torch.bfloat16)
ref = (W.float() + torch.cat(Bs, 1).float() @ torch.cat(Ats, 1).float().t() / P).to(torch.bfloat16)
for impl in ("staged", "staged3", "tma", "direct"):
    o = torch.empty_like(W)
    K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats], out_ptr=o.data_ptr(), m=M,
                 n=N, r=r, w_base_ptr=W.data_ptr(), impl=impl)
    torch.cuda.synchronize()
    assert (o.float() - ref.float()).abs().max().item() <= 2 ** -5 * max(1.0, ref.float().abs().max().item()), impl
torch.cuda.synchronize()
print("sanitizer target done")
