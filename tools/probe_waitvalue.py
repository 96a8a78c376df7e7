"""Single-GPU, step-logged probe of cuStreamWaitValue32 (see tools/probe_memops.py).

    python tools/probe_waitvalue.py          # log: gpurun_out/probe_waitvalue.log

Each step is logged BEFORE it is attempted, and a watchdog thread releases every flag after a
few seconds, so a step that blocks the host is identified by name."""
import os
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    mode = sys.argv[1] if len(sys.argv) > 1 else "lazy"
    logf = open(f"gpurun_out/probe_waitvalue_{mode}.log", "w", buffering=1)

    def log(*a):
        print(f"[{time.time():.3f}]", *a, file=logf, flush=True)
        print(*a, flush=True)

    import torch
    from cuda.bindings import driver as cu

    from rayfed_b200.ops import kernels as K

    torch.cuda.set_device(0)
    x = torch.ones(1024, device="cuda")
    torch.cuda.synchronize()
    ext = K.cuda_ext()
    log("attrs")
    for name in ("CU_DEVICE_ATTRIBUTE_CAN_USE_STREAM_WAIT_VALUE_NOR_V2", "CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS_V2",
                 "CU_DEVICE_ATTRIBUTE_CAN_FLUSH_REMOTE_WRITES", "CU_DEVICE_ATTRIBUTE_CAN_USE_STREAM_MEM_OPS_V1",
                 "CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS", "CU_DEVICE_ATTRIBUTE_CAN_USE_STREAM_WAIT_VALUE_NOR"):
        try:
            a = getattr(cu.CUdevice_attribute, name)
            log(name, cu.cuDeviceGetAttribute(a, 0))
        except Exception as e:  # noqa: BLE001
            log(name, "ERR", e)

    host_flag = torch.zeros(64, dtype=torch.int32).pin_memory()
    dev_flag = torch.zeros(64, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    log("mode", mode, "CUDA_MODULE_LOADING", os.environ.get("CUDA_MODULE_LOADING"))
    if mode == "warm":
        # every kernel used below has been launched once before any wait parks
        wbuf = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
        wsrc = torch.ones(1 << 16, dtype=torch.uint8, device="cuda")
        y = x * 2  # noqa: F841
        dev_flag.fill_(0)
        K.push(wsrc.data_ptr(), wbuf.data_ptr() + 4096, 1 << 16, flag_ptr=dev_flag.data_ptr() + 192, flag_value=1,
               done_counter_ptr=dev_flag.data_ptr() + 128, device_index=0)
        torch.cuda.synchronize()

    released = threading.Event()

    def watchdog():
        # releases everything after 4 s whatever happens
        time.sleep(6)
        host_flag[:] = 1000
        log("watchdog: host flag released")
        try:
            ext.write_u32(dev_flag.data_ptr(), 1000)
            log("watchdog: device flag released (cudaMemcpy)")
        except Exception as e:  # noqa: BLE001
            log("watchdog device release failed", e)
        released.set()

    threading.Thread(target=watchdog, daemon=True).start()

    def hard_exit():
        time.sleep(25)
        log("HARD EXIT (a step never returned)")
        os._exit(3)

    threading.Thread(target=hard_exit, daemon=True).start()

    s_wait = ext.stream_create(0)
    s_other = ext.stream_create(0)
    ts_other = torch.cuda.ExternalStream(s_other)
    log("streams", s_wait, s_other)

    # --- A: host-memory flag ---
    t0 = time.time()
    log("A1 calling stream_wait_value32(host flag >= 1) via own binding")
    ext.stream_wait_value32(s_wait, host_flag.data_ptr(), 1)
    log("A1 returned after", time.time() - t0)
    t0 = time.time()
    log("A2 launching kernel on OTHER stream")
    with torch.cuda.stream(ts_other):
        y = x * 2  # noqa: F841
        e = torch.cuda.Event()
        e.record()
    log("A2 launch returned after", time.time() - t0)
    time.sleep(0.05)
    log("A3 other-stream event done while wait parked:", e.query())
    ew = ext.event_create(False)
    ext.event_record(ew, s_wait)
    log("A4 waiter event done (expect False):", ext.event_query(ew))
    t0 = time.time()
    host_flag[0] = 1
    while not ext.event_query(ew) and time.time() - t0 < 3:
        time.sleep(0.0001)
    log("A5 waiter released by CPU store:", ext.event_query(ew), "after", time.time() - t0)

    # --- A6: cudaMalloc while a wait is parked ---
    host_flag[3] = 0
    ext.stream_wait_value32(s_wait, host_flag.data_ptr() + 12, 1)
    t0 = time.time()
    log("A6 cudaMalloc (256 MiB, fresh) while parked")
    p = ext.dev_malloc(256 << 20, False)
    log("A6 cudaMalloc returned after", time.time() - t0)
    t0 = time.time()
    big = torch.empty(300 << 20, dtype=torch.uint8, device="cuda")  # caching allocator miss
    log("A6 torch.empty(300MiB) returned after", time.time() - t0)
    host_flag[3] = 1
    ext.dev_free(p)

    dev_flag.zero_()
    torch.cuda.synchronize()
    # --- B: device-memory flag released by a kernel on the other stream ---
    t0 = time.time()
    log("B1 calling stream_wait_value32(device flag >= 1)")
    ext.stream_wait_value32(s_wait, dev_flag.data_ptr(), 1)
    log("B1 returned after", time.time() - t0)
    ext.event_record(ew, s_wait)
    log("B2 waiter done (expect False):", ext.event_query(ew))
    t0 = time.time()
    log("B3 releasing with a torch kernel on the other stream (dev_flag.fill_(1))")
    with torch.cuda.stream(ts_other):
        dev_flag.fill_(1)
        e2 = torch.cuda.Event()
        e2.record()
    log("B3 launch returned after", time.time() - t0)
    while not ext.event_query(ew) and time.time() - t0 < 3:
        time.sleep(0.0001)
    log("B4 waiter released by kernel store:", ext.event_query(ew), "after", time.time() - t0, "fill done", e2.query())

    # --- C: released by the push kernel (st.release.sys) ---
    buf = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    src = torch.ones(1 << 16, dtype=torch.uint8, device="cuda")
    cflag = torch.zeros(256, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for variant in ("push_after_wait", "wait_after_push", "credit_kernel", "push_multi"):
        cflag.zero_()
        torch.cuda.synchronize()
        fp = cflag.data_ptr() + 64
        if variant == "wait_after_push":
            K.push(src.data_ptr(), buf.data_ptr() + 4096, 1 << 16, flag_ptr=fp, flag_value=7,
                   done_counter_ptr=cflag.data_ptr() + 512, device_index=0, stream=s_other)
            ext.stream_synchronize(s_other)
            log("C", variant, "flag after push sync =", ext.read_u32(fp), "done ctr", ext.read_u32(cflag.data_ptr() + 512))
        ext.stream_wait_value32(s_wait, fp, 7)
        ext.event_record(ew, s_wait)
        t0 = time.time()
        if variant == "push_after_wait":
            K.push(src.data_ptr(), buf.data_ptr() + 4096, 1 << 16, flag_ptr=fp, flag_value=7,
                   done_counter_ptr=cflag.data_ptr() + 512, device_index=0, stream=s_other)
        elif variant == "credit_kernel":
            ext.credit(0, 0, fp, 7, s_other)  # st.release.sys.u64 of 7 (low word = 7)
        elif variant == "push_multi":
            K.push_multi([(src.data_ptr(), 256, 1 << 16), (src.data_ptr(), 256 + (1 << 16), 4096)], buf.data_ptr(),
                         flag_ptr=fp, flag_value=7, done_counter_ptr=cflag.data_ptr() + 512, device_index=0,
                         stream=s_other)
        while not ext.event_query(ew) and time.time() - t0 < 2:
            time.sleep(0.0001)
        ok = ext.event_query(ew)
        log("C", variant, "released:", ok, "after", time.time() - t0)
        if not ok:
            ext.stream_synchronize(s_other)
            log("C", variant, "kernel done; flag =", ext.read_u32(fp), "done ctr", ext.read_u32(cflag.data_ptr() + 512))
            ext.write_u32(fp, 7)
            time.sleep(0.05)
            log("C", variant, "after host write released:", ext.event_query(ew))
        else:
            ext.stream_synchronize(s_other)
            log("C", variant, "payload ok:", bool((buf[4096:4096 + 16] == 1).all().item()) if variant != "credit_kernel" else "-")

    # --- D: stall map: which other streams does a parked wait delay? ---
    own = [ext.stream_create(0) for _ in range(40)]
    host_flag[1] = 0
    ext.stream_wait_value32(own[0], host_flag.data_ptr() + 4, 1)
    evs = []
    for h in own[1:]:
        ev = ext.event_create(False)
        with torch.cuda.stream(torch.cuda.ExternalStream(h)):
            y = x * 2  # noqa: F841
        ext.event_record(ev, h)
        evs.append(ev)
    time.sleep(0.1)
    log("D blocked other streams (indices):", [i + 1 for i, ev in enumerate(evs) if not ext.event_query(ev)])
    host_flag[1] = 1
    time.sleep(0.1)
    log("D after release, still blocked:", [i + 1 for i, ev in enumerate(evs) if not ext.event_query(ev)])
    # torch pool streams + default stream
    host_flag[2] = 0
    ext.stream_wait_value32(own[0], host_flag.data_ptr() + 8, 1)
    pool = [torch.cuda.Stream() for _ in range(32)]
    evs = []
    for s in pool + [torch.cuda.default_stream()]:
        ev = ext.event_create(False)
        with torch.cuda.stream(s):
            y = x * 2  # noqa: F841
        ext.event_record(ev, s.cuda_stream)
        evs.append(ev)
    time.sleep(0.1)
    log("D2 blocked torch pool/default streams:", [i for i, ev in enumerate(evs) if not ext.event_query(ev)])
    host_flag[2] = 1
    time.sleep(0.1)
    log("done")
    os._exit(0)


if __name__ == "__main__":
    main()
