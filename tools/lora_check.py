"""Quick single-GPU check + timing of the three LoRA-merge implementations
(staged persistent / per-tile TMA / direct) against torch.addmm (cuBLAS).
Writes gpurun_out/lora_check.json.  Loopback operands (all parties on one GPU)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rayfed_b200.ops import kernels as K  # noqa: E402


def timed(fn, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def main():
    dev = "cuda:0"
    out = {}
    for (P, r, M, N) in [(8, 16, 4096, 4096), (8, 16, 4096, 14336), (2, 16, 4096, 4096)]:
        Bs = [(torch.randn(M, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
        Ats = [(torch.randn(N, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
        W = torch.randn(M, N, device=dev).to(torch.bfloat16)
        Wout = torch.empty_like(W)
        Bc, Ac = torch.cat(Bs, 1), torch.cat(Ats, 1)
        ref = torch.addmm(W.float(), Bc.float(), Ac.float().t(), alpha=1.0 / P).to(torch.bfloat16)
        res = {}
        for impl in ("staged", "staged3", "tma", "direct"):
            try:
                Wout.zero_()
                call = lambda: K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats],  # noqa: E731
                                            out_ptr=Wout.data_ptr(), m=M, n=N, r=r, w_base_ptr=W.data_ptr(), impl=impl)
                call()
                torch.cuda.synchronize()
                err = (Wout.float() - ref.float()).abs().max().item()
                res[impl] = {"max_err": err, "us": timed(call)}
            except Exception as e:  # noqa: BLE001
                res[impl] = {"error": repr(e)[:300]}
        res["torch_addmm"] = {"us": timed(lambda: torch.addmm(W, Bc, Ac.t(), alpha=1.0 / P, out=Wout))}
        res["torch_copy_2x"] = {"us": timed(lambda: Wout.copy_(W))}
        byts = 2 * M * N * 2
        for v in res.values():
            if "us" in v:
                v["GBps"] = byts / v["us"] / 1e3
        out[f"P{P}_r{r}_{M}x{N}"] = res
        print(f"P{P}_r{r}_{M}x{N}", json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/lora_check.json", "w"), indent=1)


if __name__ == "__main__":
    main()
