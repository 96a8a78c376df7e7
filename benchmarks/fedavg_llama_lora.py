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
    ap.add_argument("--local-steps", type=int, default=1)
    ap.add_argument("--base", default="bf16", choices=["bf16", "fp8"])
    ap.add_argument("--rule", default="adam", choices=["adam", "merge"])
    args = ap.parse_args()

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
    for _ in range(2):
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
    if rank == 0:
        print(json.dumps({
            "bench": "fedavg_llama3_8b_lora", "parties": world, "layers": args.layers, "seq": args.seq,
            "batch": args.batch, "base": args.base, "rule": args.rule, "rounds_per_sec": 1e3 / ms,
            "ms_per_round": ms, "wall_ms_per_round": wall / args.rounds * 1e3, "aggregate_only_us": agg,
            "tokens_per_sec": world * args.batch * args.seq * args.local_steps * 1e3 / ms, "losses": vals, **info,
        }), flush=True)
    fed.shutdown()


if __name__ == "__main__":
    main()
