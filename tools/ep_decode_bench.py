"""Expert-parallel layer latency at decode sizes (launch with torch.distributed.run, one rank per GPU):
T tokens per rank, eager calls, device time per call.  python -m torch.distributed.run ... tools/ep_decode_bench.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE  # noqa: E402
from unimoe_audio_b200.ep import ExpertParallelDCMoE  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dt = torch.bfloat16
    cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
               mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
               shared_intermediate_size=1376, router_jitter_noise=0.01)
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev).eval()
    gen = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]):
            p.copy_((torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt))
    ep = ExpertParallelDCMoE(m, dist.group.WORLD)
    cases = [(pol, T) for pol in ("replicate", "exchange") for T in (2, 64 // world)] + [("large-T paths", 16)]
    for pol, T in cases:
        ep.decode_policy = pol if pol != "large-T paths" else "off"
        ep.decode_mode = ep.decode_policy != "off"
        x = torch.randn(1, T, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt)
        for _ in range(10):
            ep(x, None, None)
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 100
        s.record()
        for _ in range(n):
            ep(x, None, None)
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / n * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the same call replayed from a CUDA graph (NCCL collectives are capturable): no host work per call
        g_us = float("nan")
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    ep(x, None, None)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = ep(x, None, None)
            for _ in range(5):
                graph.replay()
            torch.cuda.synchronize()
            dist.barrier()
            s.record()
            for _ in range(n):
                graph.replay()
            e.record()
            torch.cuda.synchronize()
            tg = torch.tensor([s.elapsed_time(e) / n * 1e3], device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            g_us = tg.item()
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                print("graph capture failed:", str(exc)[:200], flush=True)
        if rank == 0:
            print(f"EP{world} T/rank={T:4d} [{pol}, path {ep.last_path}]: eager {t.item():8.1f} us per layer call, CUDA-graph replay {g_us:8.1f} us", flush=True)
    # captured graphs hold NCCL work: skip the orderly teardown (it was seen to hang) and leave at once
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
