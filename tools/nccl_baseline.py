"""NCCL *library* baseline for the two hot paths (run under torchrun):
  * aggregate: ncclAllReduce(bf16, 25.56M elements) + a separate PyTorch SGD-momentum
    update (what "a path that only calls NCCL" costs), device-timed, max over ranks;
  * push: ncclSend/ncclRecv rank0 -> rank1 bandwidth sweep.
Prints one JSON line on rank 0."""
import json
import os

import torch
import torch.distributed as dist


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
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    n = 25_557_056
    w = torch.randn(n, device="cuda").to(torch.bfloat16)
    master = w.float().clone()
    mom = torch.zeros_like(master)
    out = {"world": world}

    def allreduce_only():
        dist.all_reduce(w)

    def allreduce_plus_update():
        dist.all_reduce(w)
        avg = w.float().div_(world)
        d = master - avg
        mom.mul_(0.9).add_(d)
        master.sub_(mom)
        w.copy_(master)

    dist.barrier()
    t = timed(allreduce_only)
    tt = torch.tensor([t], device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out["nccl_allreduce_51MB_us"] = tt.item() * 1e6
    t = timed(allreduce_plus_update)
    tt = torch.tensor([t], device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out["nccl_allreduce_plus_fedavgm_update_us"] = tt.item() * 1e6
    if world >= 2:
        sweep = {}
        for k in range(10, 31, 4):
            nbytes = 1 << k
            buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")

            def xfer():
                if rank == 0:
                    dist.send(buf, 1)
                elif rank == 1:
                    dist.recv(buf, 0)

            dist.barrier()
            t = timed(xfer, iters=10 if nbytes > (1 << 26) else 30)
            tt = torch.tensor([t], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sweep[str(nbytes)] = {"us": tt.item() * 1e6, "GBps": nbytes / tt.item() / 1e9}
        out["nccl_send_recv"] = sweep
    if rank == 0:
        print(json.dumps(out), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/nccl_baseline_n{world}.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
