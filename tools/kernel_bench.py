"""Single-GPU kernel microbenchmarks (CUDA events, L2-busting sizes): local HBM
roofline fractions of push / unpack / aggregate(N=1) / local_sgd."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rayfed_b200.ops import kernels as K  # noqa: E402


def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm = peaks.get("hbm_gbs", 6650.0) * 1e9
    dev = "cuda:0"
    out = {"hbm_peak_GBps": hbm / 1e9}
    ctrl = torch.zeros(65536, dtype=torch.int32, device=dev)
    base = ctrl.data_ptr()
    n = 25_557_056  # ResNet-50 bucket
    x = torch.randn(n, device=dev).to(torch.bfloat16)
    buf = x.clone()
    # torch copy reference
    y = torch.empty_like(x)
    t = timed(lambda: y.copy_(x))
    out["torch_copy_51MB"] = {"us": t * 1e6, "GBps": 2 * n * 2 / t / 1e9}
    big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    big2 = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    t = timed(lambda: big2.copy_(big), iters=5)
    out["torch_copy_1GiB"] = {"us": t * 1e6, "GBps": 2 * (1 << 30) / t / 1e9}
    for blocks in (148, 296, 592):
        t = timed(lambda: K.push(big.data_ptr(), big2.data_ptr(), 1 << 30, flag_ptr=base, flag_value=1,
                                 done_counter_ptr=base + 64, blocks=blocks, use_bulk=False), iters=5)
        out[f"push_local_1GiB_b{blocks}"] = {"us": t * 1e6, "GBps": 2 * (1 << 30) / t / 1e9,
                                              "frac_hbm": 2 * (1 << 30) / t / hbm}
    t = timed(lambda: K.push(big.data_ptr(), big2.data_ptr(), 1 << 30, flag_ptr=base, flag_value=1,
                             done_counter_ptr=base + 64, use_bulk=True), iters=5)
    out["push_bulk_local_1GiB"] = {"us": t * 1e6, "GBps": 2 * (1 << 30) / t / 1e9, "frac_hbm": 2 * (1 << 30) / t / hbm}
    for blocks in (148, 296):
        e = [0]

        def agg():
            K.aggregate(inputs=[buf.data_ptr()], outputs=[buf.data_ptr()], my_rank=0, shard_start=0, shard_len=n,
                        kind=K.AGG_AVG, sig_ptr=base + 4096, sig_peers=[base + 4096], epoch=e[0],
                        status_ptr=base + 128, timeout_ns=int(5e9), blocks=blocks)
            e[0] += 2

        t = timed(agg)
        out[f"aggregate_avg_n1_b{blocks}"] = {"us": t * 1e6, "GBps": 2 * n * 2 / t / 1e9, "frac_hbm": 2 * n * 2 / t / hbm}
    master = x.float().clone()
    mom = torch.zeros_like(master)
    e = [1000]

    def agg_sgd():
        K.aggregate(inputs=[buf.data_ptr()], outputs=[buf.data_ptr()], my_rank=0, shard_start=0, shard_len=n,
                    kind=K.AGG_FEDAVGM, master_ptr=master.data_ptr(), mom_ptr=mom.data_ptr(), lr=1.0, momentum=0.9,
                    sig_ptr=base + 4096, sig_peers=[base + 4096], epoch=e[0], status_ptr=base + 128,
                    timeout_ns=int(5e9))
        e[0] += 2

    t = timed(agg_sgd)
    byts = n * (2 + 2 + 16)
    out["aggregate_fedavgm_n1"] = {"us": t * 1e6, "GBps": byts / t / 1e9, "frac_hbm": byts / t / hbm}
    # pre-baked plan (what FedAggregator uses): launch overhead = one pybind call
    for kind, name in ((K.AGG_AVG, "avg"), (K.AGG_FEDAVGM, "fedavgm")):
        plan = K.AggregatePlan(inputs=[buf.data_ptr()], outputs=[buf.data_ptr()], my_rank=0, shard_start=0,
                               shard_len=n, kind=kind, master_ptr=master.data_ptr(), mom_ptr=mom.data_ptr(), lr=1.0,
                               momentum=0.9, sig_ptr=base + 8192, sig_peers=[base + 8192], epoch_dev_ptr=base + 256,
                               agg_done_ptr=base + 260, status_ptr=base + 264, timeout_ns=int(5e9))
        t = timed(lambda: plan.launch(0, 1))
        byts = n * (4 if kind == K.AGG_AVG else 20)
        out[f"aggregate_plan_{name}_n1"] = {"us": t * 1e6, "GBps": byts / t / 1e9, "frac_hbm": byts / t / hbm}
    tiny = torch.zeros(4096, dtype=torch.bfloat16, device=dev)
    plan = K.AggregatePlan(inputs=[tiny.data_ptr()], outputs=[tiny.data_ptr()], my_rank=0, shard_start=0,
                           shard_len=4096, kind=K.AGG_AVG, sig_ptr=base + 8192, sig_peers=[base + 8192],
                           epoch_dev_ptr=base + 256, agg_done_ptr=base + 260, status_ptr=base + 264,
                           timeout_ns=int(5e9))
    t = timed(lambda: plan.launch(0, 1), iters=200)
    out["aggregate_plan_launch_floor_us"] = t * 1e6
    # K3 tcgen05 LoRA merge, Llama-3-8B q_proj shape, 8 parties x r=16 (loopback operands)
    P, r, M, N = 8, 16, 4096, 4096
    Bs = [(torch.randn(M, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
    Ats = [(torch.randn(N, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
    W = torch.randn(M, N, device=dev).to(torch.bfloat16)
    Wout = torch.empty_like(W)
    t = timed(lambda: K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats],
                                   out_ptr=Wout.data_ptr(), m=M, n=N, r=r, w_base_ptr=W.data_ptr()))
    byts = 2 * M * N * 2
    flops = 2.0 * M * N * P * r
    out["lora_merge_staged_4096x4096_k128"] = {"us": t * 1e6, "GBps": byts / t / 1e9, "frac_hbm": byts / t / hbm,
                                               "TFLOPs": flops / t / 1e12}
    t = timed(lambda: K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats],
                                   out_ptr=Wout.data_ptr(), m=M, n=N, r=r, w_base_ptr=W.data_ptr(), impl="tma"))
    out["lora_merge_tma_4096x4096_k128"] = {"us": t * 1e6, "GBps": byts / t / 1e9, "frac_hbm": byts / t / hbm,
                                            "TFLOPs": flops / t / 1e12}
    t = timed(lambda: K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats],
                                   out_ptr=Wout.data_ptr(), m=M, n=N, r=r, w_base_ptr=W.data_ptr(), impl="direct"))
    out["lora_merge_direct_4096x4096_k128"] = {"us": t * 1e6, "GBps": byts / t / 1e9, "frac_hbm": byts / t / hbm,
                                               "TFLOPs": flops / t / 1e12}
    Bc = torch.cat(Bs, 1)
    Ac = torch.cat(Ats, 1)
    t = timed(lambda: torch.addmm(W, Bc, Ac.t(), alpha=1.0 / P, out=Wout))
    out["torch_addmm_same_shape"] = {"us": t * 1e6, "GBps": byts / t / 1e9}
    g = torch.randn(n, device=dev).to(torch.bfloat16)
    mo = torch.zeros(n, device=dev)
    for blocks in (296, 592, 1184):
        t = timed(lambda: K.local_sgd(buf.data_ptr(), g.data_ptr(), mo.data_ptr(), n, lr=0.01, momentum=0.9,
                                      blocks=blocks))
        byts = n * (2 + 2 + 2 + 2 + 4 + 4)
        out[f"local_sgd_b{blocks}"] = {"us": t * 1e6, "GBps": byts / t / 1e9, "frac_hbm": byts / t / hbm}
    # fp8 push (bf16 -> e4m3 + scales), local
    pay = torch.empty(n, dtype=torch.uint8, device=dev)
    sc = torch.empty(n // 32, dtype=torch.uint8, device=dev)
    t = timed(lambda: K.push(x.data_ptr(), pay.data_ptr(), n, mode=K.MODE_BF16_TO_FP8, dst_scales_ptr=sc.data_ptr(),
                             flag_ptr=base, flag_value=1, done_counter_ptr=base + 64))
    byts = n * 3 + n // 32
    out["push_fp8_51MB"] = {"us": t * 1e6, "GBps": byts / t / 1e9, "frac_hbm": byts / t / hbm}
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/kernel_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
