"""CUDA-event timing of the LoRA-merge implementations (1 GPU, loopback operands, 8 parties x r=16)."""
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
    return a.elapsed_time(b) / iters * 1e3


out = {}
dev = "cuda:0"
for (M, N) in ((4096, 4096), (4096, 14336), (4096, 1024)):
    P, r = 8, 16
    Bs = [(torch.randn(M, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
    Ats = [(torch.randn(N, r, device=dev) * 0.1).to(torch.bfloat16) for _ in range(P)]
    # distinct W buffers per iteration set so the 126 MB L2 does not hide the traffic at 4096x14336
    Ws = [torch.randn(M, N, device=dev).to(torch.bfloat16) for _ in range(4)]
    row = {}
    for impl in ("staged", "staged3", "tma"):
        i = [0]

        def f():
            w = Ws[i[0] % len(Ws)]
            i[0] += 1
            K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats], out_ptr=w.data_ptr(),
                         m=M, n=N, r=r, w_base_ptr=w.data_ptr(), impl=impl)

        row[impl] = timed(f)
    Bcat = torch.cat(Bs, 1)
    Acat = torch.cat(Ats, 1)
    j = [0]

    def g():
        w = Ws[j[0] % len(Ws)]
        j[0] += 1
        torch.addmm(w, Bcat, Acat.t(), alpha=1.0 / P, out=w)

    row["cublas_addmm_local_operands"] = timed(g)
    k = [0]

    def c():
        w = Ws[k[0] % len(Ws)]
        k[0] += 1
        Ws[(k[0]) % len(Ws)].copy_(w)

    row["torch_copy_floor"] = timed(c)
    out[f"{M}x{N}"] = row
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/lora_time.json", "w"), indent=1)
