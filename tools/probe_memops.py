"""Hardware probe for the host-free receive path (run on a 2-GPU box).

    python tools/probe_memops.py --nproc 2

Every step logs to ``gpurun_out/probe_memops_rank{r}.log`` as it goes and never blocks on the
GPU without a deadline (a parked wait is force-released from the host), so a hang is located,
not suffered.  Questions:

1. zero-copy torch views of raw arena memory + deleter callback;
2. does a parked ``cuStreamWaitValue32`` stall work on *other* streams?
3. wait-value on a flag in my own device memory released by a kernel on another stream of the
   same process (cudaMalloc / cudaIpc / VMM backed);
4. the same, released by the *peer GPU's* push kernel (st.release.sys over NVLink), payload
   visibility included; resume latency.
"""
import argparse
import gc
import json
import os
import sys
import tempfile
import time
import traceback

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def chk(res, what):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"{what}: {err}")


def run(rank, nproc, xdir):
    os.makedirs("gpurun_out", exist_ok=True)
    logf = open(f"gpurun_out/probe_memops_rank{rank}.log", "w", buffering=1)
    out = {"rank": rank}

    def log(*a):
        print(f"[{time.time():.3f}]", *a, file=logf, flush=True)

    def save():
        with open(f"gpurun_out/probe_memops_rank{rank}.json", "w") as f:
            json.dump(out, f, indent=1)

    try:
        _run(rank, nproc, xdir, out, log, save)
    except Exception:  # noqa: BLE001
        out["fatal"] = traceback.format_exc()
        log("FATAL", out["fatal"])
    save()
    if rank == 0:
        print(json.dumps(out, indent=1), flush=True)
    logf.close()
    os._exit(0)


def _run(rank, nproc, xdir, out, log, save):
    import torch
    from cuda.bindings import driver as cu

    from rayfed_b200.ops import kernels as K
    from rayfed_b200.parallel import FileExchange, PartyGroup, SymmetricHeap
    from rayfed_b200.parallel.symmetric import _CudaArrayView

    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    torch.zeros(1, device=f"cuda:{dev}")
    ext = K.cuda_ext()
    GEQ = cu.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ
    log("cuda up; MAX_CONNECTIONS", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))

    def wait_event(ev, seconds):
        t0 = time.time()
        while time.time() - t0 < seconds:
            if ev.query():
                return True
            time.sleep(0.0005)
        return bool(ev.query())

    # ---- 1. zero-copy arena views + deleter --------------------------------------------
    try:
        raw = torch.arange(1 << 20, dtype=torch.uint8, device=f"cuda:{dev}")
        freed = []

        class Block:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False),
                                                 "version": 3, "strides": None}

            def __del__(self):
                freed.append(1)

        t = torch.as_tensor(Block(raw.data_ptr() + 256, 4096), device=f"cuda:{dev}")
        v = t.view(torch.bfloat16).view(32, 64)
        same_ptr = v.data_ptr() == raw.data_ptr() + 256
        del t
        gc.collect()
        alive_with_view = len(freed) == 0
        w = v[3:5]
        del v
        gc.collect()
        alive_with_slice = len(freed) == 0
        del w
        gc.collect()
        out["arena_view"] = {"zero_copy": bool(same_ptr), "alive_with_view": alive_with_view,
                             "alive_with_slice": alive_with_slice, "freed_after_last": len(freed) == 1}
    except Exception:  # noqa: BLE001
        out["arena_view"] = "ERR " + traceback.format_exc()
    log("arena_view", out["arena_view"])
    save()

    # ---- 2. does a parked wait-value stall other streams? (flag in pinned host memory) ---
    try:
        flagbuf = torch.zeros(64, dtype=torch.int32).pin_memory()
        x = torch.ones(1024, device=f"cuda:{dev}")
        stall = {}
        own = [ext.stream_create(0) for _ in range(48)]  # raw, non-pool streams
        for nstreams in (2, 4, 8, 16, 24, 31, 40, 48):
            flagbuf.zero_()
            torch.cuda.synchronize()
            chk(cu.cuStreamWaitValue32(own[0], flagbuf.data_ptr(), 1, GEQ), "waitvalue")
            evs = []
            for h in own[1:nstreams]:
                s = torch.cuda.ExternalStream(h, device=f"cuda:{dev}")
                with torch.cuda.stream(s):
                    y = x * 2  # noqa: F841
                    e = torch.cuda.Event()
                    e.record()
                    evs.append(e)
            time.sleep(0.1)
            blocked = [i + 1 for i, e in enumerate(evs) if not e.query()]
            flagbuf[0] = 1  # CPU store releases the wait
            ok = all(wait_event(e, 5) for e in evs)
            stall[str(nstreams)] = {"blocked": blocked, "all_done_after_release": ok}
            log("stall", nstreams, stall[str(nstreams)])
        out["stall_vs_streams_own"] = stall
        # torch pool streams (what lanes use)
        flagbuf.zero_()
        pool = [torch.cuda.Stream() for _ in range(31)]
        chk(cu.cuStreamWaitValue32(own[0], flagbuf.data_ptr(), 1, GEQ), "waitvalue")
        evs = []
        for s in pool:
            with torch.cuda.stream(s):
                y = x * 2  # noqa: F841
                e = torch.cuda.Event()
                e.record()
                evs.append(e)
        y = x * 3  # default stream too
        e = torch.cuda.Event()
        e.record()
        evs.append(e)
        time.sleep(0.1)
        out["stall_pool31_plus_default"] = [i for i, e in enumerate(evs) if not e.query()]
        flagbuf[0] = 1
        out["stall_pool_released"] = all(wait_event(e, 5) for e in evs)
        log("pool", out["stall_pool31_plus_default"])
    except Exception:  # noqa: BLE001
        out["stall_vs_streams_own"] = "ERR " + traceback.format_exc()
        log(out["stall_vs_streams_own"])
    save()

    # ---- 3. local: flag in device memory, released by a kernel on another stream -----------
    wait_s = ext.stream_create(0)
    wait_ts = torch.cuda.ExternalStream(wait_s, device=f"cuda:{dev}")
    push_ts = torch.cuda.Stream()

    def local_case(name, flag_ptr, done_ptr, payload_dst, seq):
        src = torch.randint(0, 255, (1 << 20,), dtype=torch.uint8, device=f"cuda:{dev}")
        torch.cuda.synchronize()
        e0 = torch.cuda.Event()
        with torch.cuda.stream(wait_ts):
            chk(cu.cuStreamWaitValue32(wait_s, flag_ptr, seq, GEQ), "waitvalue")
            e0.record()
        time.sleep(0.02)
        early = e0.query()
        K.push(src.data_ptr(), payload_dst, 1 << 20, flag_ptr=flag_ptr, flag_value=seq, done_counter_ptr=done_ptr,
               device_index=dev, stream=push_ts)
        ok = wait_event(e0, 5)
        if not ok:
            log(name, "HANG: forcing the flag from the host")
            ext.write_u32(flag_ptr, seq)
            wait_event(e0, 5)
        out[f"local_{name}"] = {"released_by_kernel": ok, "released_early": bool(early)}
        log("local", name, out[f"local_{name}"])
        save()

    try:
        plain = torch.zeros((2 << 20) // 4, dtype=torch.int32, device=f"cuda:{dev}")
        local_case("cudamalloc", plain.data_ptr(), plain.data_ptr() + 1024, plain.data_ptr() + 4096, 1)
    except Exception:  # noqa: BLE001
        out["local_cudamalloc"] = "ERR " + traceback.format_exc()
        log(out["local_cudamalloc"])

    parties = [f"p{i}" for i in range(nproc)]
    group = PartyGroup(parties, parties[rank], FileExchange(xdir, timeout_s=60), device_index=dev)
    peer = (rank + 1) % nproc
    for backend in ("ipc", "vmm"):
        try:
            heap = SymmetricHeap(group, backend=backend)
            NB = 64 << 20
            buf = heap.alloc(f"probe-{backend}", 8192 + NB, multicast=(backend == "vmm"))
            log(backend, "alloc ok; mc", bool(buf.mc_ptr))
            local_case(backend, buf.local_ptr + 2048, buf.local_ptr + 3072, buf.local_ptr + 8192, 1)

            # ---- 4. cross-GPU ------------------------------------------------------------
            flag_local = buf.local_ptr
            flag_peer = buf.peer_ptrs[peer]
            done = buf.local_ptr + 1024
            src = torch.randint(0, 255, (NB,), dtype=torch.uint8, device=f"cuda:{dev}")
            res = {}
            seq = 0
            for nbytes in (4096, 1 << 20, NB):
                seq += 1
                dst = torch.zeros(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
                torch.cuda.synchronize()
                ev_ready = torch.cuda.Event()
                ev_copied = torch.cuda.Event()
                with torch.cuda.stream(wait_ts):
                    chk(cu.cuStreamWaitValue32(wait_s, flag_local, seq, GEQ), "waitvalue")
                    ev_ready.record()
                    view = torch.as_tensor(_CudaArrayView(buf.local_ptr + 8192, nbytes), device=f"cuda:{dev}")
                    dst.copy_(view)
                    ev_copied.record()
                group.barrier()
                time.sleep(0.02)  # the wait is parked well before the push is even launched
                early = ev_ready.query()
                K.push(src.data_ptr(), buf.peer_ptrs[peer] + 8192, nbytes, flag_ptr=flag_peer, flag_value=seq,
                       done_counter_ptr=done, device_index=dev, stream=push_ts)
                ok = wait_event(ev_copied, 10)
                if not ok:
                    log(backend, nbytes, "HANG: forcing the flag from the host")
                    ext.write_u32(flag_local, seq)
                    wait_event(ev_copied, 5)
                push_ts.synchronize()
                mine = int(src[:nbytes].to(torch.int64).sum().item())
                got = int(dst.to(torch.int64).sum().item())
                sums = group.all_gather(group.next_tag("sum"), json.dumps([mine, got]).encode())
                exp = json.loads(sums[(rank - 1) % nproc])[0]
                res[str(nbytes)] = {"released_by_peer": ok, "payload_ok": exp == got, "released_early": bool(early)}
                log(backend, nbytes, res[str(nbytes)])
            out[f"cross_{backend}"] = res
            save()

            lat = []
            for it in range(20):
                seq += 1
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                group.barrier()
                t0 = time.time()
                K.push(src.data_ptr(), buf.peer_ptrs[peer] + 8192, 4096, flag_ptr=flag_peer, flag_value=seq,
                       done_counter_ptr=done, device_index=dev, stream=push_ts)
                with torch.cuda.stream(wait_ts):
                    chk(cu.cuStreamWaitValue32(wait_s, flag_local, seq, GEQ), "waitvalue")
                    e1.record()
                if not wait_event(e1, 5):
                    ext.write_u32(flag_local, seq)
                    wait_event(e1, 5)
                    lat.append(-1.0)
                else:
                    lat.append((time.time() - t0) * 1e6)
            out[f"push4k_to_resume_wall_us_{backend}"] = sorted(lat)[len(lat) // 2]
            log(backend, "latency", out[f"push4k_to_resume_wall_us_{backend}"])
            save()
            group.barrier()
        except Exception:  # noqa: BLE001
            out[f"cross_{backend}"] = "ERR " + traceback.format_exc()
            log(out[f"cross_{backend}"])
            save()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nproc", type=int, default=2)
    args = ap.parse_args()
    import torch.multiprocessing as mp

    xdir = tempfile.mkdtemp(prefix="fedprobe")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=run, args=(r, args.nproc, xdir)) for r in range(args.nproc)]
    for p in procs:
        p.start()
    deadline = time.time() + 150
    for p in procs:
        p.join(max(1.0, deadline - time.time()))
    for p in procs:
        if p.is_alive():
            p.kill()
    sys.exit(0)


if __name__ == "__main__":
    main()
