#!/bin/bash
# One 1-GPU gpurun call: smoke, the new GPU suites, litmus at 100k iterations, BASELINE config #4 at N=1,
# the staged3 LoRA pipeline check, and the ncu capture of every hand-written kernel.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "== device path + litmus"; RAYFED_B200_LITMUS_ITERS=${LITMUS:-100000} timeout 900 python -m pytest tests/test_gpu_device_path.py tests/test_gpu_litmus.py -q -p no:cacheprovider --durations=12 > gpurun_out/t_dev.log 2>&1; tail -n 22 gpurun_out/t_dev.log
echo "== staged3"; timeout 400 python -m pytest tests/test_gpu_lora_staged3.py -q -p no:cacheprovider > gpurun_out/t_staged3.log 2>&1; tail -n 4 gpurun_out/t_staged3.log
echo "== llama-lora N=1"; timeout 400 python bench.py --config llama-lora --gpus 1 --steps 3 --warmup 3 > gpurun_out/llama_n1.log 2>&1; tail -n 3 gpurun_out/llama_n1.log | cut -c1-1500
echo "== ncu"; PROFILE_STAGED3=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fed_ -c 18 -f -o gpurun_out/prof_all python tools/profile_target.py > gpurun_out/ncu.log 2>&1; tail -n 3 gpurun_out/ncu.log; ls -la gpurun_out/prof_all.ncu-rep
