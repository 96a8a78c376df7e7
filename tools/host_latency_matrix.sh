#!/bin/bash
# Host-side latency knobs, measured where it matters (the B200 host, many cores):
# many_tiny_tasks at N parties for every combination of the shared-memory data path, the wait
# spin budget and the GIL hand-off spin.  Usage: bash tools/host_latency_matrix.sh [N=2] [iters=5000]
N=${1:-2}; ITERS=${2:-5000}
mkdir -p gpurun_out
OUT=gpurun_out/host_latency_n$N.txt; : > $OUT
for shm in 0 1; do for spin in 40 200; do for gil in 0 80; do for adapt in 0 300; do
  r=$(RAYFED_B200_SHM=$shm RAYFED_B200_SPIN_US=$spin RAYFED_B200_GIL_SPIN_US=$gil RAYFED_B200_ADAPTIVE_SPIN_US=$adapt \
      timeout 300 python benchmarks/many_tiny_tasks_benchmark.py --parties $N --iters $ITERS 2>/dev/null | tail -1)
  echo "shm=$shm spin_us=$spin gil_spin_us=$gil adaptive_spin_us=$adapt $r" | tee -a $OUT
done; done; done; done
for shm in 0 1; do RAYFED_B200_SHM=$shm python tools/transport_pingpong.py | tee -a $OUT; done
