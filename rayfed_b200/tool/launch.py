"""Start one controller process per party on this machine.

    python -m rayfed_b200.tool.launch --parties alice,bob [--gpus 0,1 | --gpus auto] \\
        [--base-port 11000] [--transport auto] script.py [script args...]

Every party runs the *same* script (multi-controller execution).  The launcher passes the
party's identity through the environment; the script calls :func:`rayfed_b200.init_from_env`::

    import rayfed_b200 as fed
    fed.init_from_env()            # addresses / party / device / transport from the launcher
    ...
    fed.shutdown()

Exit status: 0 iff every party exited 0 (the first failing party's status otherwise); when a
party fails the others get SIGINT after a grace period so the job does not linger.
(The reference has no launcher — parties are started by hand or by the test harness.)
"""
from __future__ import annotations

import argparse
import json
import os
import signal
import socket
import subprocess
import sys
import time
from typing import Dict, List, Optional

ENV_PARTY = "RAYFED_B200_PARTY"
ENV_ADDRESSES = "RAYFED_B200_ADDRESSES"
ENV_DEVICE = "RAYFED_B200_DEVICE"
ENV_TRANSPORT = "RAYFED_B200_TRANSPORT"
ENV_JOB = "RAYFED_B200_JOB"


def _free_ports(n: int) -> List[int]:
    socks, ports = [], []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        socks.append(s)
        ports.append(s.getsockname()[1])
    for s in socks:
        s.close()
    return ports


def party_env(party: str, addresses: Dict[str, str], device: Optional[str], transport: str, job: str) -> Dict[str, str]:
    env = dict(os.environ)
    env[ENV_PARTY] = party
    env[ENV_ADDRESSES] = json.dumps(addresses)
    env[ENV_TRANSPORT] = transport
    env[ENV_JOB] = job
    if device:
        env[ENV_DEVICE] = device
    else:
        env.pop(ENV_DEVICE, None)
    return env


def launch(parties: List[str], argv: List[str], gpus: Optional[List[int]] = None, base_port: int = 0,
           transport: str = "auto", job: Optional[str] = None, grace_s: float = 10.0) -> int:
    ports = [base_port + i for i in range(len(parties))] if base_port else _free_ports(len(parties))
    addresses = {p: f"127.0.0.1:{port}" for p, port in zip(parties, ports)}
    job = job or f"job-{os.getpid()}"
    procs = []
    for i, p in enumerate(parties):
        device = f"cuda:{gpus[i % len(gpus)]}" if gpus else None
        procs.append(subprocess.Popen([sys.executable] + argv, env=party_env(p, addresses, device, transport, job)))
    status = 0
    failed_at = None
    try:
        while any(pr.poll() is None for pr in procs):
            for pr in procs:
                rc = pr.poll()
                if rc not in (None, 0) and status == 0:
                    status, failed_at = rc, time.time()
            if failed_at is not None and time.time() - failed_at > grace_s:
                for pr in procs:
                    if pr.poll() is None:
                        pr.send_signal(signal.SIGINT)
                failed_at = time.time() + 3600  # signal once
            time.sleep(0.05)
    except KeyboardInterrupt:
        for pr in procs:
            if pr.poll() is None:
                pr.send_signal(signal.SIGINT)
        status = 130
    for pr in procs:
        rc = pr.wait()
        if rc != 0 and status == 0:
            status = rc
    return status


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m rayfed_b200.tool.launch", description=__doc__.split("\n\n")[0])
    ap.add_argument("--parties", required=True, help="comma-separated party names, or an integer N for p0..pN-1")
    ap.add_argument("--gpus", default=None, help="comma-separated GPU indices (one per party, reused round-robin), "
                                                 "'auto' = 0..N-1, omit for CPU parties")
    ap.add_argument("--base-port", type=int, default=0, help="first TCP port (default: pick free ports)")
    ap.add_argument("--transport", default="auto", choices=["auto", "native", "nvlink", "grpc"])
    ap.add_argument("--job", default=None)
    ap.add_argument("script", nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)
    if not args.script:
        ap.error("missing script")
    parties = [f"p{i}" for i in range(int(args.parties))] if args.parties.isdigit() else args.parties.split(",")
    gpus = None
    if args.gpus == "auto":
        gpus = list(range(len(parties)))
    elif args.gpus:
        gpus = [int(g) for g in args.gpus.split(",")]
    return launch(parties, args.script, gpus, args.base_port, args.transport, args.job)


if __name__ == "__main__":
    sys.exit(main())
