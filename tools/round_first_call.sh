#!/bin/bash
# Everything the first GPU call of a round should collect, in ONE gpurun invocation on ONE GPU
# (≈6-8 min of box time).  Usage:  gpurun --timeout 900 -- 'bash tools/round_first_call.sh'
# Outputs land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
export RAYFED_B200_NEXT_ROUND_TESTS=1   # include the tests that have never run on hardware
echo "== gpu tests"; timeout 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/gpu_tests.txt
echo "== smoke";     timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== bench N=1"; timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1.json
echo "== kernels";   timeout 200 python tools/kernel_bench.py 2>&1 | tail -3 | tee gpurun_out/kernel_bench.txt
echo "== lora";      timeout 100 python tools/lora_check.py 2>&1 | tail -4
echo "== launch list (never a bench value)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_target.py > gpurun_out/launches.log 2>&1
echo "== ncu full: staged LoRA GEMM, multicast-free single-GPU kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fed_lora_gemm|fed_lora_gather" -c 2 \
    -o gpurun_out/ncu_lora_staged -f python tools/profile_target.py > gpurun_out/ncu_lora.log 2>&1
echo "== compute-sanitizer memcheck on the kernel tests (single GPU)"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x \
    -k "push or local_sgd or aggregate_avg or lora_merge_staged" -p no:cacheprovider > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/memcheck.log
tail -3 gpurun_out/memcheck.log
