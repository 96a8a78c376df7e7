"""Launch each hand-written kernel a few times on ONE GPU (loopback peers) — the
target of `ncu --set full -k regex:fed_`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rayfed_b200.ops import kernels as K  # noqa: E402

dev = "cuda:0"
ctrl = torch.zeros(65536, dtype=torch.int32, device=dev)
base = ctrl.data_ptr()
n = 25_557_056
x = torch.randn(n, device=dev).to(torch.bfloat16)
buf = x.clone()
g = torch.randn(n, device=dev).to(torch.bfloat16)
mo = torch.zeros(n, device=dev)
master = x.float().clone()
mom = torch.zeros_like(master)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
big2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pay = torch.empty(n, dtype=torch.uint8, device=dev)
sc = torch.empty(n // 32, dtype=torch.uint8, device=dev)
P, r, M, N = 8, 16, 4096, 4096
Bs = [(torch.randn(M, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
Ats = [(torch.randn(N, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
W = torch.randn(M, N, device=dev).to(torch.bfloat16)
Wout = torch.empty_like(W)
plan = K.AggregatePlan(inputs=[buf.data_ptr()], outputs=[buf.data_ptr()], my_rank=0, shard_start=0, shard_len=n,
                       kind=K.AGG_FEDAVGM, master_ptr=master.data_ptr(), mom_ptr=mom.data_ptr(), lr=1.0,
                       momentum=0.9, sig_ptr=base + 4096, sig_peers=[base + 4096], epoch_dev_ptr=base + 256,
                       agg_done_ptr=base + 260, status_ptr=base + 264, timeout_ns=int(5e9))
from rayfed_b200.proxy.nvlink import channel as C  # noqa: E402

ext = K.cuda_ext()
avg_plan = K.AggregatePlan(inputs=[buf.data_ptr()], outputs=[buf.data_ptr()], my_rank=0, shard_start=0, shard_len=n,
                           kind=K.AGG_AVG, sig_ptr=base + 4096, sig_peers=[base + 4096], epoch_dev_ptr=base + 512,
                           agg_done_ptr=base + 516, status_ptr=base + 264, timeout_ns=int(5e9))
desc = C.pack_desc(seq=1, job_hash=1, key=(1, 0, 2), offset=0, nbytes=(64 << 20) + 256)
segs = [(big.data_ptr() + i * (4 << 20), 256 + i * (4 << 20), 4 << 20) for i in range(16)]
for _ in range(2):
    avg_plan.launch(0, 1)  # fed_allreduce_avg_kernel (pipelined FedAvg, N=1 loopback)
    K.push_multi(segs, big2.data_ptr(), flag_ptr=base + 128, flag_value=1, done_counter_ptr=base + 64,
                 desc_dst=big2.data_ptr(), desc=desc)
    ext.desc_check(big2.data_ptr(), desc, base + 264, torch.cuda.current_stream().cuda_stream)
    ext.credit(base + 1024, 1, base + 1032, 1, torch.cuda.current_stream().cuda_stream)
    K.push(big.data_ptr(), big2.data_ptr(), 256 << 20, flag_ptr=base, flag_value=1, done_counter_ptr=base + 64,
           use_bulk=False)
    K.push(big.data_ptr(), big2.data_ptr(), 256 << 20, flag_ptr=base, flag_value=1, done_counter_ptr=base + 64,
           use_bulk=True)
    K.push(x.data_ptr(), pay.data_ptr(), n, mode=K.MODE_BF16_TO_FP8, dst_scales_ptr=sc.data_ptr(), flag_ptr=base,
           flag_value=1, done_counter_ptr=base + 64)
    plan.launch(0, 1)
    K.local_sgd(buf.data_ptr(), g.data_ptr(), mo.data_ptr(), n, lr=0.01, momentum=0.9)
    for impl in ("staged", "tma", "direct") + (("staged3",) if os.environ.get("PROFILE_STAGED3") == "1" else ()):  # fed_lora_gather + fed_lora_gemm / fed_lora_merge_tma / fed_lora_merge
        K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats],
                     out_ptr=Wout.data_ptr(), m=M, n=N, r=r, w_base_ptr=W.data_ptr(), impl=impl)
torch.cuda.synchronize()
print("profile target done")
