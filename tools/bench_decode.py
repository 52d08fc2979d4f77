"""Decode-sized latency of one DCMoE forward: eager (6 launches + Python) vs CUDA-graph replay.
    python tools/bench_decode.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE  # noqa: E402
from unimoe_audio_b200.host import GraphedDCMoE  # noqa: E402


def main():
    dev = torch.device("cuda:0")
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
    print("weights per layer: 304.4 MB bf16 -> %.1f us at 6.53 TB/s (floor when every expert is hit)" % (304.4e6 / 6.5297e12 * 1e6))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    for T in (2, 8, 32, 128, 512):
        x = torch.randn(T, 1, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt)
        for _ in range(5):
            m(x, None, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            m(x, None, None)
        torch.cuda.synchronize()
        eager = (time.perf_counter() - t0) / n * 1e6
        g = GraphedDCMoE(m, T, 1, dt, device=dev)
        for _ in range(5):
            g(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            g(x)
        torch.cuda.synchronize()
        graph = (time.perf_counter() - t0) / n * 1e6
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            g.graph.replay()
        e.record()
        torch.cuda.synchronize()
        dev_us = s.elapsed_time(e) / n * 1e3
        print(f"T={T:4d}  eager {eager:8.1f} us/call   graph replay {graph:8.1f} us/call   (device time per replay {dev_us:7.1f} us)")


if __name__ == "__main__":
    main()
