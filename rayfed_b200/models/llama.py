"""Llama-3 style decoder with LoRA adapters — the model of BASELINE.json's
"FedAvg Llama-3-8B LoRA fp8" config.

* ``LlamaConfig.llama3_8b()`` is the real 8B shape (32 layers, d=4096, 32 heads /
  8 KV heads, FFN 14336, vocab 128256); tiny configs exist for tests.
* Base weights are frozen; every attention / MLP projection can carry a LoRA
  adapter.  Adapters use the layout the K3 kernel wants: ``lora_B`` is
  ``[d_out, r]`` and ``lora_At`` is ``A^T = [d_in, r]`` (both K-major for
  tcgen05), so ``y = x W^T + (alpha/r) * (x At) B^T``.
* Attention uses ``F.scaled_dot_product_attention`` (library flash kernel) and
  cuBLAS GEMMs — plain library calls; the federated-specific work (adapter
  buckets in the symmetric heap, fused aggregate, tcgen05 merge) is in
  :mod:`rayfed_b200.fl.lora`.
* ``base_dtype="fp8"`` keeps the frozen base weights block-scaled e4m3
  (UE8M0 per 32 along the input dim) — what the K3 merge epilogue reads and
  writes — and dequantises per layer for the GEMM.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn


@dataclass
class LlamaConfig:
    vocab_size: int = 128256
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 8
    rope_theta: float = 500000.0
    rms_eps: float = 1e-5
    lora_r: int = 16
    lora_alpha: float = 32.0
    lora_targets: Tuple[str, ...] = ("q_proj", "k_proj", "v_proj", "o_proj")
    base_dtype: str = "bf16"  # "bf16" | "fp8"

    @staticmethod
    def llama3_8b(**kw) -> "LlamaConfig":
        return LlamaConfig(**kw)

    @staticmethod
    def tiny(**kw) -> "LlamaConfig":
        d = dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=4,
                 num_kv_heads=2, lora_r=16)
        d.update(kw)
        return LlamaConfig(**d)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads


def quantize_block_fp8(w: torch.Tensor):
    """[out, in] -> (uint8 e4m3 payload [out, in], uint8 UE8M0 scales [out, in/32])."""
    o, i = w.shape
    assert i % 32 == 0
    blocks = w.float().view(o, i // 32, 32)
    amax = blocks.abs().amax(-1, keepdim=True).clamp_min(2.0 ** -120)
    e = torch.ceil(torch.log2(amax / 448.0)).clamp(-126, 126)
    q = (blocks / torch.exp2(e)).to(torch.float8_e4m3fn)
    return q.view(o, i).view(torch.uint8).contiguous(), (e.view(o, i // 32) + 127).to(torch.uint8).contiguous()


def dequantize_block_fp8(payload: torch.Tensor, scales: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    o, i = payload.shape
    q = payload.view(torch.float8_e4m3fn).float().view(o, i // 32, 32)
    s = torch.exp2(scales.float() - 127.0).view(o, i // 32, 1)
    return (q * s).view(o, i).to(dtype)


class LoRALinear(nn.Module):
    """Frozen base projection + optional LoRA adapter in the K3 layout."""

    def __init__(self, d_in: int, d_out: int, r: int = 0, alpha: float = 32.0, base_dtype: str = "bf16"):
        super().__init__()
        self.d_in, self.d_out, self.r = d_in, d_out, r
        self.scaling = alpha / r if r else 0.0
        self.base_dtype = base_dtype
        w = torch.empty(d_out, d_in)
        nn.init.normal_(w, std=0.02)
        if base_dtype == "fp8":
            q, s = quantize_block_fp8(w)
            self.register_buffer("weight_q", q)
            self.register_buffer("weight_s", s)
            self.weight = None
        else:
            self.weight = nn.Parameter(w.to(torch.bfloat16), requires_grad=False)
        if r:
            at = torch.empty(d_in, r)
            nn.init.kaiming_uniform_(at, a=5 ** 0.5)
            self.lora_At = nn.Parameter(at.to(torch.bfloat16))          # A^T  [d_in, r]
            self.lora_B = nn.Parameter(torch.zeros(d_out, r, dtype=torch.bfloat16))  # [d_out, r]

    def base_weight(self) -> torch.Tensor:
        if self.base_dtype == "fp8":
            return dequantize_block_fp8(self.weight_q, self.weight_s)
        return self.weight

    def forward(self, x):
        y = F.linear(x, self.base_weight())
        if self.r:
            y = y + (x @ self.lora_At) @ self.lora_B.t() * self.scaling
        return y


class RMSNorm(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d, dtype=torch.bfloat16), requires_grad=False)
        self.eps = eps

    def forward(self, x):
        v = x.float()
        return (v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + self.eps)).to(x.dtype) * self.weight


def _rope(x, theta):
    # x: [B, H, T, D]
    b, h, t, d = x.shape
    inv = 1.0 / (theta ** (torch.arange(0, d, 2, device=x.device, dtype=torch.float32) / d))
    ang = torch.arange(t, device=x.device, dtype=torch.float32)[:, None] * inv[None, :]
    cos, sin = ang.cos()[None, None], ang.sin()[None, None]
    x1, x2 = x.float()[..., 0::2], x.float()[..., 1::2]
    out = torch.stack((x1 * cos - x2 * sin, x1 * sin + x2 * cos), dim=-1).flatten(-2)
    return out.to(x.dtype)


class Attention(nn.Module):
    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.cfg = cfg
        d, hd = cfg.hidden_size, cfg.head_dim

        def proj(name, d_in, d_out):
            r = cfg.lora_r if name in cfg.lora_targets else 0
            return LoRALinear(d_in, d_out, r, cfg.lora_alpha, cfg.base_dtype)

        self.q_proj = proj("q_proj", d, cfg.num_heads * hd)
        self.k_proj = proj("k_proj", d, cfg.num_kv_heads * hd)
        self.v_proj = proj("v_proj", d, cfg.num_kv_heads * hd)
        self.o_proj = proj("o_proj", cfg.num_heads * hd, d)

    def forward(self, x):
        cfg = self.cfg
        b, t, _ = x.shape
        q = self.q_proj(x).view(b, t, cfg.num_heads, cfg.head_dim).transpose(1, 2)
        k = self.k_proj(x).view(b, t, cfg.num_kv_heads, cfg.head_dim).transpose(1, 2)
        v = self.v_proj(x).view(b, t, cfg.num_kv_heads, cfg.head_dim).transpose(1, 2)
        q, k = _rope(q, cfg.rope_theta), _rope(k, cfg.rope_theta)
        rep = cfg.num_heads // cfg.num_kv_heads
        if rep > 1:
            k = k.repeat_interleave(rep, dim=1)
            v = v.repeat_interleave(rep, dim=1)
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.o_proj(o.transpose(1, 2).reshape(b, t, -1))


class MLPBlock(nn.Module):
    def __init__(self, cfg: LlamaConfig):
        super().__init__()

        def proj(name, d_in, d_out):
            r = cfg.lora_r if name in cfg.lora_targets else 0
            return LoRALinear(d_in, d_out, r, cfg.lora_alpha, cfg.base_dtype)

        self.gate_proj = proj("gate_proj", cfg.hidden_size, cfg.intermediate_size)
        self.up_proj = proj("up_proj", cfg.hidden_size, cfg.intermediate_size)
        self.down_proj = proj("down_proj", cfg.intermediate_size, cfg.hidden_size)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class Block(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.attn_norm = RMSNorm(cfg.hidden_size, cfg.rms_eps)
        self.attn = Attention(cfg)
        self.mlp_norm = RMSNorm(cfg.hidden_size, cfg.rms_eps)
        self.mlp = MLPBlock(cfg)

    def forward(self, x):
        x = x + self.attn(self.attn_norm(x))
        return x + self.mlp(self.mlp_norm(x))


class LlamaLoRA(nn.Module):
    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.cfg = cfg
        self.embed = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.embed.weight.requires_grad_(False)
        self.layers = nn.ModuleList([Block(cfg) for _ in range(cfg.num_layers)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_eps)
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        self.lm_head.weight.requires_grad_(False)

    def forward(self, tokens):
        x = self.embed(tokens).to(torch.bfloat16)
        for layer in self.layers:
            x = layer(x)
        return self.lm_head(self.norm(x).to(self.lm_head.weight.dtype))

    def lora_modules(self) -> List[Tuple[str, LoRALinear]]:
        return [(n, m) for n, m in self.named_modules() if isinstance(m, LoRALinear) and m.r]

    def num_lora_params(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


def llama_lora(cfg: Optional[LlamaConfig] = None, **kw) -> LlamaLoRA:
    return LlamaLoRA(cfg or LlamaConfig.tiny(**kw))
