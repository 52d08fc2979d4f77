"""BASELINE.json configs[4]: Top-P threshold sweep (p = 0.5 .. 0.95) with skewed router logits stressing load
imbalance, expert-parallel over all visible ranks (launch with torch.distributed.run; also runs on 1 GPU).
The skew is a fixed bias linspace(+2, -2) on the 9 dynamic gate logits; logits are computed once
with torch and fed to the router so every implementation sees identical values.

    python -m torch.distributed.run --nproc-per-node 8 tools/bench_topp_sweep.py [--batch 64 --seq 4096]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE  # noqa: E402
from unimoe_audio_b200.ep import ExpertParallelDCMoE  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--skew", type=float, default=2.0)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dt = torch.bfloat16
    B = a.batch // world
    T = B * a.seq
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, a.seq, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt)
    for p in (0.5, 0.6, 0.7, 0.8, 0.9, 0.95):
        cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=p,
                   mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
                   shared_intermediate_size=1376, router_jitter_noise=0.01)
        with torch.device("meta"):
            m = DCMoE(cfg)
        m = m.to(dt).to_empty(device=dev).eval()
        g0 = torch.Generator(device=dev).manual_seed(0)
        with torch.no_grad():
            for _, prm in sorted(m.named_parameters(), key=lambda kv: kv[0]):
                prm.copy_((torch.randn(prm.shape, generator=g0, device=dev, dtype=torch.float32) * 0.02).to(dt))
        logits = torch.nn.functional.linear(x.view(T, 2048), m.gate.weight).float()
        logits[:, :9] += torch.linspace(a.skew, -a.skew, 9, device=dev)
        logits = logits.to(dt).contiguous()
        layer = ExpertParallelDCMoE(m, dist.group.WORLD) if world > 1 else m
        for _ in range(3):
            out = layer(x, None, None, router_logits=logits)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.steps):
            out = layer(x, None, None, router_logits=logits)
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / a.steps], device=dev, dtype=torch.float64)
        sent = out[3][:, :8].sum(0).to(torch.float64)            # rows this rank routes to each expert
        parity = None
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            dist.all_reduce(sent)
            # parity gate: the ranks' outputs, gathered, against ONE single-GPU forward of the concatenated batch
            x_all = torch.empty((world * B, a.seq, 2048), dtype=dt, device=dev)
            dist.all_gather_into_tensor(x_all, x)
            lg_all = torch.empty((world * T, 11), dtype=dt, device=dev)
            dist.all_gather_into_tensor(lg_all, logits)
            got = []
            for i in (0, 2, 3, 4):
                t = out[i].contiguous()
                g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                dist.all_gather_into_tensor(g, t)
                got.append(g)
            if rank == 0:
                ref = m(x_all, None, None, router_logits=lg_all)
                torch.cuda.synchronize()
                parity = all(bool(torch.equal(g.reshape(ref[i].shape), ref[i])) for g, i in zip(got, (0, 2, 3, 4)))
                del ref
            del x_all, lg_all, got
            torch.cuda.empty_cache()
        else:
            allms = [ms]
        if rank == 0:
            tms = [t.item() for t in allms]
            load = sent.cpu()
            print(json.dumps({"top_p": p, "n_gpus": world, "tokens_global": T * world, "ms_per_step_max": max(tms),
                              "ms_per_step_min": min(tms), "rank_ms_per_step": tms,
                              "rank_time_spread": (max(tms) - min(tms)) / max(tms),
                              "ep_path": getattr(layer, "last_path", None), "ep_equals_single_gpu": parity,
                              "tokens_per_s": T * world / (max(tms) * 1e-3),
                              "mean_routed_experts": load.sum().item() / (T * world),
                              "expert_load_max_over_mean": (load.max() / load.mean()).item(),
                              "rows_per_expert": [int(v) for v in load.tolist()]}), flush=True)
        del layer, m
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
