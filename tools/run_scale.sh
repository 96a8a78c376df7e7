#!/bin/bash
# One multi-GPU session: bench (ours) + NCCL baseline + probe at N GPUs.  Usage: tools/run_scale.sh N [steps]
N=$1; STEPS=${2:-10}; PORT=$((29600 + N))
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
  bench.py --gpus $N --steps $STEPS --warmup 3 --tiny-iters 3000 > gpurun_out/bench_n$N.log 2>&1
tail -1 gpurun_out/bench_n$N.log > gpurun_out/bench_n$N.json
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+50)) \
  tools/nccl_baseline.py > gpurun_out/nccl_n$N.log 2>&1
timeout 150 python tools/gpu_probe.py --nproc $N --max-bytes 268435456 > gpurun_out/probe_n$N.log 2>&1
for r in 0; do cp gpurun_out/probe_rank$r.json gpurun_out/probe_n${N}_rank$r.json; done
python - <<PY
import json
b=json.loads(open("gpurun_out/bench_n$N.json").read())
print("BENCH", {k:b[k] for k in ("value","ms_per_step","rounds_per_sec","gpu_launches")}, b["e2e"]["value"], b["many_tiny_tasks"], b.get("aggregate"))
try: print("NCCL", open("gpurun_out/nccl_baseline_n$N.json").read()[:600])
except Exception as e: print("nccl missing", e)
p=json.load(open("gpurun_out/probe_n${N}_rank0.json"))
print("PROBE agg", p.get("aggregate_p2p"), p.get("vmm"), p.get("error"))
PY
