#!/bin/bash
# The first multi-GPU call of a round (gpurun --gpus N): device-channel tests incl. the ones that
# were never run on hardware, then the scaling bench at N.  Usage: bash tools/round_multi_gpu_call.sh N
N=${1:-2}
mkdir -p gpurun_out
export RAYFED_B200_NEXT_ROUND_TESTS=1   # include the tests that have never run on hardware
timeout 400 python -m pytest tests/test_gpu_fed_nvlink.py tests/test_gpu_zz_next_round.py -m gpu -q -p no:cacheprovider 2>&1 \
    | tail -15 | tee gpurun_out/gpu_tests_multi.txt
bash tools/run_scale.sh "$N" 10
