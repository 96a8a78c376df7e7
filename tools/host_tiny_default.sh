#!/bin/bash
# many_tiny_tasks on the GPU host with the shipped defaults (CPU parties, shm rings), N = 2, 4, 8.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; OUT=gpurun_out/host_tiny_final.txt; : > $OUT
echo "nproc=$(nproc) cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null) load=$(cut -d' ' -f1-3 /proc/loadavg)" | tee -a $OUT
port=25000
for n in 2 4 8 2 8; do port=$((port + 40))
  r=$(timeout 60 python benchmarks/many_tiny_tasks_benchmark.py --parties $n --iters 3000 --base-port $port 2>/dev/null | tail -n 1 | cut -c60-125)
  echo "n=$n defaults $r" | tee -a $OUT
done
