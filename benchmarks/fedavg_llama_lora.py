"""FedAvg of Llama-3-8B LoRA adapters (BASELINE.json config 4), one party per GPU.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/fedavg_llama_lora.py \
        [--layers 32] [--seq 512] [--batch 1] [--rounds 5] [--base fp8|bf16] [--rule adam|merge]

Per round every party runs ``local_steps`` fwd/bwd on synthetic tokens (frozen base,
trainable adapters), then ONE collective:
  rule=adam   fused all-reduce(grads) + AdamW on the sharded master + all-gather  (K3)
  rule=merge  exact  W <- W + (alpha/r)/N * sum_i B_i A_i  on tcgen05, operands over NVLink
Prints rounds/s (device-timed, max over parties) and the aggregate-only time.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--local-steps", type=int, default=1)
    ap.add_argument("--base", default="bf16", choices=["bf16", "fp8"])
    ap.add_argument("--rule", default="adam", choices=["adam", "merge"])
    args = ap.parse_args()
    out = run(args)
    if out is not None:
        print(json.dumps(out), flush=True)


def bench_main(bargs):
    """``python bench.py --config llama-lora --gpus N --steps K --warmup W``: BASELINE config #4 —
    Llama-3-8B shapes (all 32 layers, q/k/v/o adapters r=16), block-scaled fp8 base, one party per
    GPU, fused all-reduce + AdamW aggregate (K3); prints the driver-contract JSON line."""
    ns = argparse.Namespace(layers=int(os.environ.get("LLAMA_LAYERS", "32")), seq=512, batch=1,
                            rounds=bargs.steps, warmup=max(bargs.warmup, 3), local_steps=1,
                            base=os.environ.get("LLAMA_BASE", "fp8"), rule="adam")
    r = run(ns)
    if r is None:
        return
    n = r["parties"]
    out = {"metric": "fedavg_llama3_8b_lora_fp8_rounds_per_sec", "value": r["rounds_per_sec"], "unit": "rounds/s",
           "n_gpus": n, "steps": ns.rounds, "warmup": ns.warmup, "ms_per_step": r["ms_per_round"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 compute, fp8 base weights",
           "data": "synthetic tokens (pinned host memory, H2D each step), random-init Llama-3-8B shapes",
           "impl": "ours", "tokens_per_sec": r["tokens_per_sec"],
           "config": {"model": f"Llama-3-8B shapes, {ns.layers} layers, LoRA r=16 on q/k/v/o "
                               f"({r['lora_params'] / 1e6:.1f}M adapter params)", "global_batch": n * ns.batch,
                      "seq_len": ns.seq, "parallelism": f"fedavg-dp{n} (one party per GPU)", "base": ns.base,
                      "rule": "fused all-reduce(grads) + AdamW on the sharded fp32 master + all-gather (K3)",
                      "l2_policy": "inputs larger than L2 (8-16 GB of base weights stream per step)"},
           "aggregate": {"us": r["aggregate_only_us"]}, "e2e": {"value": 1e3 / r["wall_ms_per_round"], "unit": "rounds/s",
                                                                "h2d_bytes_per_step": ns.batch * ns.seq * 8,
                                                                "d2h_bytes_per_step": 4 * n},
           "lora_merge": r.get("lora_merge"), "mem_GB": r["mem_GB"], "losses": r["losses"]}
    print(json.dumps(out), flush=True)


def run(args):

    import rayfed_b200 as fed
    from rayfed_b200.fl import FedLoRAAdam, FedLoRAExactMerge, device_context
    from rayfed_b200.models import LlamaConfig, LlamaLoRA

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    parties = [f"p{i}" for i in range(world)]
    base_port = int(os.environ.get("MASTER_PORT", 29500)) + 300
    fed.init(addresses={p: f"127.0.0.1:{base_port + i}" for i, p in enumerate(parties)}, party=parties[rank],
             logging_level="warning",
             config={"device": f"cuda:{local}", "barrier_on_initializing": True,
                     "cross_silo_comm": {"transport": "nvlink", "retry_max_attempts": 200,
                                         "retry_initial_backoff_s": 0.05, "retry_max_backoff_s": 0.5}})

    class Worker:
        def __init__(self, r):
            torch.cuda.set_device(local)
            self.group, self.heap, self.channel = device_context(parties, parties[r])
            cfg = LlamaConfig.llama3_8b(num_layers=args.layers, base_dtype=args.base)
            torch.manual_seed(0)
            with torch.device("cuda"):
                self.model = LlamaLoRA(cfg)
            self.model.embed.to(torch.bfloat16)
            self.model.lm_head.to(torch.bfloat16)
            if args.rule == "adam":
                self.fl = FedLoRAAdam(self.model, self.group, self.heap, self.channel, lr=1e-4, on_grads=True)
            else:
                self.fl = FedLoRAExactMerge(self.model, self.group, self.heap, self.channel)
                self.opt = torch.optim.SGD([p for p in self.model.parameters() if p.requires_grad], lr=1e-3)
            g = torch.Generator().manual_seed(r)
            self.tokens = torch.randint(0, cfg.vocab_size, (args.batch, args.seq), generator=g).pin_memory()
            self.ev = {}

        def info(self):
            return {"lora_params": self.model.num_lora_params(),
                    "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}

        def round(self):
            loss = None
            for _ in range(args.local_steps):
                tok = self.tokens.cuda(non_blocking=True)
                out = self.model(tok)
                loss = torch.nn.functional.cross_entropy(out.float().view(-1, out.shape[-1]), tok.view(-1))
                loss.backward()
                if args.rule == "merge":
                    self.opt.step()
                    self.opt.zero_grad(set_to_none=False)
            if args.rule == "adam":
                self.fl.step()
            else:
                self.fl.merge()
            return loss.detach()

        def mark(self, k):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.ev[k] = e
            return True

        def ms(self, a, b):
            self.ev[b].synchronize()
            return self.ev[a].elapsed_time(self.ev[b])

        def aggregate_only_us(self, iters=10):
            step = self.fl.step if args.rule == "adam" else self.fl.merge
            for _ in range(2):
                step()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                step()
            b.record()
            b.synchronize()
            return a.elapsed_time(b) / iters * 1e3

    W = fed.remote(Worker)
    ws = {p: W.party(p).remote(i) for i, p in enumerate(parties)}
    for _ in range(getattr(args, "warmup", 2)):
        fed.get([ws[p].round.remote() for p in parties])
    for p in parties:
        ws[p].mark.remote("a")
    t0 = time.perf_counter()
    for _ in range(args.rounds):
        losses = [ws[p].round.remote() for p in parties]
    vals = [float(x) for x in fed.get(losses)]
    wall = time.perf_counter() - t0
    for p in parties:
        ws[p].mark.remote("b")
    ms = max(fed.get([ws[p].ms.remote("a", "b") for p in parties])) / args.rounds
    agg = max(fed.get([ws[p].aggregate_only_us.remote() for p in parties]))
    info = fed.get(ws[parties[0]].info.remote())
    out = {
        "bench": "fedavg_llama3_8b_lora", "parties": world, "layers": args.layers, "seq": args.seq,
        "batch": args.batch, "base": args.base, "rule": args.rule, "rounds_per_sec": 1e3 / ms,
        "ms_per_round": ms, "wall_ms_per_round": wall / args.rounds * 1e3, "aggregate_only_us": agg,
        "tokens_per_sec": world * args.batch * args.seq * args.local_steps * 1e3 / ms, "losses": vals, **info,
    }
    fed.shutdown()
    return out if rank == 0 else None


if __name__ == "__main__":
    main()
