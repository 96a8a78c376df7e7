"""Multi-process GPU probe: symmetric heap bring-up, kernel correctness, first timings.

    python tools/gpu_probe.py --nproc 2 [--skip-vmm]

Each rank is one party on GPU ``rank % device_count``.  Writes
``gpurun_out/probe_rank{r}.json``.
"""
import argparse
import json
import os
import sys
import tempfile
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, stream, iters=10, warmup=3):
    import torch

    for _ in range(warmup):
        fn()
    stream.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(iters):
        fn()
    e.record(stream)
    e.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def run(rank, nproc, xdir, args):
    import torch

    from rayfed_b200.ops import kernels as K
    from rayfed_b200.parallel import FileExchange, PartyGroup, SymmetricHeap

    out = {"rank": rank}
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    parties = [f"p{i}" for i in range(nproc)]
    group = PartyGroup(parties, parties[rank], FileExchange(xdir), device_index=dev)
    heap = SymmetricHeap(group)
    ext = K.cuda_ext()
    out["sm_count"] = K.sm_count(dev)
    out["ndev"] = ndev
    peer = (rank + 1) % nproc
    stream = torch.cuda.current_stream()

    # ---------------- symmetric heap (cudaIpc) ---------------------------------
    MAXB = args.max_bytes
    buf = heap.alloc("probe", 4096 + 2 * MAXB)
    ctrl = heap.alloc("ctrl", 1 << 16)
    out["ipc_ok"] = True
    flag_peer = ctrl.peer_ptrs[peer] + 0
    flag_local = ctrl.local_ptr + 0
    done = ctrl.local_ptr + 1024
    undone = ctrl.local_ptr + 1028
    status = ctrl.local_ptr + 1032

    # ---------------- K1 copy correctness ----------------------------------------
    torch.manual_seed(1234 + rank)
    seq = 0
    res = {}
    for nbytes in [1 << 10, 4097, 1 << 20, (1 << 24) + 48]:
        src = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=f"cuda:{dev}")
        for bulk in ([False, True] if nbytes % 16 == 0 else [False]):
            seq += 1
            group.barrier()
            K.push(src.data_ptr(), buf.peer_ptrs[peer] + 4096, nbytes, flag_ptr=flag_peer, flag_value=seq,
                   done_counter_ptr=done, device_index=dev, use_bulk=bulk)
            dst = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
            K.unpack(buf.local_ptr + 4096, dst.data_ptr(), nbytes, flag_ptr=flag_local, flag_value=seq,
                     done_counter_ptr=undone, status_ptr=status, timeout_ns=int(20e9), device_index=dev)
            torch.cuda.synchronize()
            # expected: what the *previous* rank generated
            g = torch.Generator(device=f"cuda:{dev}")
            res[f"copy_{nbytes}_{'bulk' if bulk else 'sm'}"] = True
            # verify by exchanging a checksum through the file exchange
            mine = int(src.to(torch.int64).sum().item())
            got = int(dst.to(torch.int64).sum().item())
            sums = group.all_gather(group.next_tag("sum"), json.dumps([mine, got]).encode())
            sender = (rank - 1) % nproc
            exp = json.loads(sums[sender])[0]
            res[f"copy_{nbytes}_{'bulk' if bulk else 'sm'}"] = (exp == got)
    out["copy"] = res
    out["status_after_copy"] = ext.read_u32(status)

    # ---------------- K1 bandwidth sweep ---------------------------------------------
    sweep = {}
    sizes = [1 << k for k in range(10, 31, 2) if (1 << k) <= MAXB]
    big = torch.empty(MAXB, dtype=torch.uint8, device=f"cuda:{dev}")
    big.random_(0, 255)
    for nbytes in sizes:
        row = {}
        for bulk in (False, True):
            if bulk and nbytes < (1 << 16):
                continue
            group.barrier()
            it = 20 if nbytes <= (1 << 24) else 5
            t = timed(lambda: K.push(big.data_ptr(), buf.peer_ptrs[peer] + 4096, nbytes, flag_ptr=flag_peer,
                                     flag_value=0, done_counter_ptr=done, device_index=dev, use_bulk=bulk),
                      stream, iters=it)
            row["bulk" if bulk else "sm"] = {"us": t * 1e6, "GBps": nbytes / t / 1e9}
        # library baseline: cudaMemcpyPeerAsync-equivalent via torch (copy into peer-mapped memory)
        sweep[str(nbytes)] = row
    out["push_sweep"] = sweep
    group.barrier()

    # ---------------- fp8 wire format ---------------------------------------------------
    n = 1 << 20
    x = (torch.randn(n, device=f"cuda:{dev}") * 3).to(torch.bfloat16)
    seq += 1
    group.barrier()
    ext.write_u32(flag_local, 0)
    group.barrier()
    pay_pad = n
    K.push(x.data_ptr(), buf.peer_ptrs[rank] + 4096, n, mode=K.MODE_BF16_TO_FP8,
           dst_scales_ptr=buf.peer_ptrs[rank] + 4096 + pay_pad, flag_ptr=flag_local, flag_value=1,
           done_counter_ptr=done, device_index=dev)
    y = torch.empty(n, dtype=torch.bfloat16, device=f"cuda:{dev}")
    K.unpack(buf.local_ptr + 4096, y.data_ptr(), n, mode=K.MODE_FP8_TO_BF16, scales_ptr=buf.local_ptr + 4096 + pay_pad,
             flag_ptr=flag_local, flag_value=1, done_counter_ptr=undone, status_ptr=status,
             timeout_ns=int(20e9), device_index=dev)
    torch.cuda.synchronize()
    xf = x.float().view(-1, 32)
    amax = xf.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
    e = torch.ceil(torch.log2(amax / 448.0))
    sc = torch.exp2(e)
    ref = (xf / sc).to(torch.float8_e4m3fn).float() * sc
    err = (y.float().view(-1, 32) - ref).abs().max().item()
    out["fp8_roundtrip_max_err_vs_torch"] = err
    out["fp8_rel_err"] = ((y.float() - x.float()).abs().mean() / x.float().abs().mean()).item()

    # ---------------- K2 aggregate (P2P) ------------------------------------------------
    agg = {}
    N = nproc
    nel = args.agg_elems // (32 * N) * (32 * N)
    wbuf = heap.alloc("weights", nel * 2)
    w = wbuf.tensor(torch.bfloat16, nel)
    sig_ptr = ctrl.local_ptr + 4096
    sig_peers = [p + 4096 for p in ctrl.peer_ptrs]
    shard = nel // N
    gens = []
    for j in range(N):
        gj = torch.Generator(device="cpu")
        gj.manual_seed(777 + j)
        gens.append(torch.randn(nel, generator=gj).to(torch.bfloat16))
    w.copy_(gens[rank].to(w.device))
    torch.cuda.synchronize()
    group.barrier()
    epoch = 0
    K.aggregate(inputs=wbuf.peer_ptrs, outputs=wbuf.peer_ptrs, my_rank=rank, shard_start=rank * shard,
                shard_len=shard, kind=K.AGG_AVG, sig_ptr=sig_ptr, sig_peers=sig_peers, epoch=epoch,
                status_ptr=status, timeout_ns=int(20e9), device_index=dev)
    epoch += 2
    torch.cuda.synchronize()
    group.barrier()
    expect = torch.stack([g.float() for g in gens]).mean(0).to(torch.bfloat16)
    agg["avg_max_err"] = (w.float().cpu() - expect.float()).abs().max().item()
    # FedAvgM / SGD with sharded master
    master = torch.zeros(shard, device=f"cuda:{dev}")
    mom = torch.zeros(shard, device=f"cuda:{dev}")
    master.copy_(expect[rank * shard:(rank + 1) * shard].float().to(master.device))
    w.copy_(gens[rank].to(w.device))  # contributions = "grads"
    torch.cuda.synchronize()
    group.barrier()
    K.aggregate(inputs=wbuf.peer_ptrs, outputs=wbuf.peer_ptrs, my_rank=rank, shard_start=rank * shard,
                shard_len=shard, kind=K.AGG_SGD, master_ptr=master.data_ptr(), mom_ptr=mom.data_ptr(), lr=0.1,
                momentum=0.9, weight_decay=0.01, sig_ptr=sig_ptr, sig_peers=sig_peers, epoch=epoch,
                status_ptr=status, timeout_ns=int(20e9), device_index=dev)
    epoch += 2
    torch.cuda.synchronize()
    group.barrier()
    g_avg = torch.stack([g.float() for g in gens]).mean(0)
    m0 = expect.float()
    gg = g_avg + 0.01 * m0
    new = (m0 - 0.1 * gg).to(torch.bfloat16)
    agg["sgd_max_err"] = (w.float().cpu() - new.float()).abs().max().item()
    # timing (in-place average, repeated)
    group.barrier()

    def agg_once():
        nonlocal epoch
        K.aggregate(inputs=wbuf.peer_ptrs, outputs=wbuf.peer_ptrs, my_rank=rank, shard_start=rank * shard,
                    shard_len=shard, kind=K.AGG_AVG, sig_ptr=sig_ptr, sig_peers=sig_peers, epoch=epoch,
                    status_ptr=status, timeout_ns=int(20e9), device_index=dev)
        epoch += 2

    t = timed(agg_once, stream, iters=20)
    agg["avg_us"] = t * 1e6
    agg["elems"] = nel
    # sensitivity to the number of CTAs (all parties use the same value each time): if halving
    # the grid costs much less than 2x, the kernel is latency- not parallelism-bound
    sweep = {}
    for blocks in (37, 74, 111, 148):
        group.barrier()

        def agg_b():
            nonlocal epoch
            K.aggregate(inputs=wbuf.peer_ptrs, outputs=wbuf.peer_ptrs, my_rank=rank, shard_start=rank * shard,
                        shard_len=shard, kind=K.AGG_AVG, sig_ptr=sig_ptr, sig_peers=sig_peers, epoch=epoch,
                        status_ptr=status, timeout_ns=int(20e9), device_index=dev, blocks=blocks)
            epoch += 2

        sweep[str(blocks)] = timed(agg_b, stream, iters=10) * 1e6
    agg["avg_us_by_blocks"] = sweep
    # bytes over my NVLink: read (N-1)/N * nel*2, write (N-1)/N * nel*2
    agg["link_GBps_per_dir"] = (N - 1) / N * nel * 2 / t / 1e9
    out["aggregate_p2p"] = agg
    out["status_after_agg"] = ext.read_u32(status)
    group.barrier()

    # ---------------- NCCL baseline for the same aggregate (if torch.distributed usable) --
    # (measured separately in bench.py under torchrun)

    # ---------------- VMM + multicast probe ---------------------------------------------------
    if not args.skip_vmm:
        vm = {}
        try:
            vm["info"] = heap.vmm_info()
            mcbuf = heap.alloc("mc-weights", nel * 2, multicast=True)
            vm["backend"] = mcbuf.backend
            vm["mc_ptr_nonzero"] = bool(mcbuf.mc_ptr)
            if mcbuf.mc_ptr:
                wm = mcbuf.tensor(torch.bfloat16, nel)
                wm.copy_(gens[rank].to(wm.device))
                torch.cuda.synchronize()
                group.barrier()
                K.aggregate(inputs=mcbuf.peer_ptrs, outputs=mcbuf.peer_ptrs, my_rank=rank,
                            shard_start=rank * shard, shard_len=shard, kind=K.AGG_AVG, mc_in=mcbuf.mc_ptr,
                            mc_out=mcbuf.mc_ptr, sig_ptr=sig_ptr, sig_peers=sig_peers, epoch=epoch,
                            status_ptr=status, timeout_ns=int(20e9), device_index=dev)
                epoch += 2
                torch.cuda.synchronize()
                group.barrier()
                vm["mc_avg_max_err"] = (wm.float().cpu() - expect.float()).abs().max().item()

                def mc_once():
                    nonlocal epoch
                    K.aggregate(inputs=mcbuf.peer_ptrs, outputs=mcbuf.peer_ptrs, my_rank=rank,
                                shard_start=rank * shard, shard_len=shard, kind=K.AGG_AVG, mc_in=mcbuf.mc_ptr,
                                mc_out=mcbuf.mc_ptr, sig_ptr=sig_ptr, sig_peers=sig_peers, epoch=epoch,
                                status_ptr=status, timeout_ns=int(20e9), device_index=dev)
                    epoch += 2

                group.barrier()
                t = timed(mc_once, stream, iters=20)
                vm["mc_avg_us"] = t * 1e6
        except Exception as e:  # noqa: BLE001
            vm["error"] = repr(e)
            vm["trace"] = traceback.format_exc()[-1500:]
        out["vmm"] = vm
    return out


def worker(rank, nproc, xdir, args):
    os.makedirs("gpurun_out", exist_ok=True)
    try:
        out = run(rank, nproc, xdir, args)
    except Exception as e:  # noqa: BLE001
        out = {"rank": rank, "error": repr(e), "trace": traceback.format_exc()}
    with open(f"gpurun_out/probe_rank{rank}.json", "w") as f:
        json.dump(out, f, indent=1)
    if rank == 0:
        print(json.dumps(out, indent=1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nproc", type=int, default=2)
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--agg-elems", type=int, default=25_600_000)
    ap.add_argument("--skip-vmm", action="store_true")
    args = ap.parse_args()
    import torch.multiprocessing as mp

    xdir = tempfile.mkdtemp(prefix="fedprobe_")
    mp.spawn(worker, args=(args.nproc, xdir, args), nprocs=args.nproc, join=True)
