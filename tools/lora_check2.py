"""Staged LoRA merge vs K (parties*r) and operand padding; single GPU, loopback."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rayfed_b200.ops import kernels as K  # noqa: E402
from tools.lora_check import timed  # noqa: E402


def main():
    dev = "cuda:0"
    out = {}
    M = N = 4096
    W = torch.randn(M, N, device=dev).to(torch.bfloat16)
    Wout = torch.empty_like(W)
    for (P, r) in [(2, 16), (4, 16), (6, 16), (8, 16)]:
        Bs = [(torch.randn(M, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
        Ats = [(torch.randn(N, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
        ref = torch.addmm(W.float(), torch.cat(Bs, 1).float(), torch.cat(Ats, 1).float().t(), alpha=1.0 / P).to(torch.bfloat16)
        res = {}
        for name, impl, kmin in (("staged", "staged", "0"), ("staged_pad128", "staged", "128"), ("tma", "tma", "0")):
            os.environ["RAYFED_B200_LORA_KPAD_MIN"] = kmin
            call = lambda: K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats],  # noqa: E731
                                        out_ptr=Wout.data_ptr(), m=M, n=N, r=r, w_base_ptr=W.data_ptr(), impl=impl)
            call()
            torch.cuda.synchronize()
            res[name] = {"max_err": (Wout.float() - ref.float()).abs().max().item(), "us": timed(call, iters=20, warmup=3)}
        out[f"K{P * r}"] = res
        print(f"K{P * r}", json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/lora_check2.json", "w"), indent=1)


if __name__ == "__main__":
    main()
