#!/bin/bash
# many_tiny_tasks on the host the GPUs sit in: N parties (CPU processes), shared-memory rings on/off.
mkdir -p gpurun_out; OUT=gpurun_out/host_tiny_matrix.txt; : > $OUT
for n in 2 4 8; do for shm in 1 0; do
  r=$(RAYFED_B200_SHM=$shm timeout 120 python benchmarks/many_tiny_tasks_benchmark.py --parties $n --iters 3000 --base-port $((14000 + n * 20 + shm * 10)) 2>/dev/null | tail -n 1)
  echo "n=$n shm=$shm $r" | tee -a $OUT
done; done
