"""Hand-written sm_100a kernels (Python entry points).

===========================  ==================================================
``ops.push`` / ``ops.unpack``  K1/K4 — cross-party tensor push with fused
                               cast / scale / block-fp8 quantisation; device-side
                               flag wait + unpack + credit return
``ops.aggregate``              K2/K3 — fused all-reduce + 1/N + optimizer +
                               all-gather (FedAvg/FedAvgM/SGD/Adam/FedAdam, bf16
                               or block-scaled fp8 out, P2P or NVLS multimem)
``ops.barrier``                K5 — N-party device barrier on the signal pad
``ops.local_sgd``              fused local SGD step on flat buckets
``ops.lora_merge``             tcgen05/TMEM GEMM: exact LoRA FedAvg Σ BᵢAᵢ with
                               operands gathered from peers
===========================  ==================================================

All of them raise if the ``_fedcuda`` extension cannot be loaded — there is no
silent PyTorch fallback on a GPU box.
"""
from .kernels import (  # noqa: F401
    AGG_ADAM,
    AGG_AVG,
    AGG_FEDADAM,
    AGG_FEDAVGM,
    AGG_SGD,
    MODE_BF16_SCALE,
    MODE_BF16_TO_F32,
    MODE_BF16_TO_FP8,
    MODE_COPY,
    MODE_F32_TO_BF16,
    MODE_FP8_TO_BF16,
    AggregatePlan,
    aggregate,
    barrier,
    cuda_ext,
    launch_counter,
    local_sgd,
    lora_merge,
    push,
    push_mc,
    sm_count,
    unpack,
)
