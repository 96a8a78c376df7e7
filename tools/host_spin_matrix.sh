#!/bin/bash
# many_tiny_tasks on the host the GPUs sit in, adaptive spinning off / on (CPU parties, shm rings).
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; OUT=gpurun_out/host_spin_matrix.txt; : > $OUT
{ echo "nproc=$(nproc)"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|Hypervisor" ; cat /sys/fs/cgroup/cpu.max 2>/dev/null;
  python -c "import rayfed_b200; from rayfed_b200.core import native; print(native().spin_info())"; } 2>&1 | tee -a $OUT
port=21000
for rep in 1 2; do for n in 8 2 4; do for cap in 0 500; do
  port=$((port + 40))
  r=$(RAYFED_B200_ADAPTIVE_SPIN_US=$cap timeout 100 python benchmarks/many_tiny_tasks_benchmark.py --parties $n --iters 3000 --base-port $port 2>/dev/null | tail -n 1 | cut -c1-120)
  echo "n=$n adaptive_cap=$cap rep=$rep $r" | tee -a $OUT
done; done; done
port=$((port + 40))
r=$(timeout 100 python benchmarks/many_tiny_tasks_benchmark.py --parties 8 --iters 3000 --base-port $port 2>/dev/null | tail -n 1 | cut -c1-120)
echo "n=8 default $r" | tee -a $OUT
echo "== transport send cost, default policy" | tee -a $OUT; timeout 60 python tools/transport_send_cost.py 2>&1 | tail -n 4 | tee -a $OUT
echo "== transport send cost, adaptive off" | tee -a $OUT; RAYFED_B200_ADAPTIVE_SPIN_US=0 timeout 60 python tools/transport_send_cost.py 2>&1 | tail -n 4 | tee -a $OUT
