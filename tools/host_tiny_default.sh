#!/bin/bash
# many_tiny_tasks on the GPU host with the shipped defaults (CPU parties, shm rings), N = 8, 2, 4.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; OUT=gpurun_out/host_tiny_final.txt; : > $OUT
echo "nproc=$(nproc) cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null) load=$(cut -d' ' -f1-3 /proc/loadavg)" | tee -a $OUT
port=26000
for n in 8 2 4 8 2; do port=$((port + 40))
  r=$(timeout 30 python benchmarks/many_tiny_tasks_benchmark.py --parties $n --iters 3000 --base-port $port 2>/dev/null | tail -n 1 | cut -c60-125)
  echo "n=$n defaults $r" | tee -a $OUT
done
