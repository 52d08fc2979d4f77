"""Per-kernel timings in isolation (CUDA events, rotating inputs).  python tools/bench_stages.py [T]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE, ops  # noqa: E402


def timeit(fn, n=20, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
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
    xs = [torch.randn(T, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt) for _ in range(4)]
    m(xs[0].view(1, T, 2048), None, None)
    ws = m.last_workspace
    wg = m.gate.weight.detach()
    lg, tk, mk, gw = ops.router(xs[0], wg, ws)
    ops.plan(ws)
    torch.cuda.synchronize()
    A = int(ws.counts.sum().item())
    print(f"T={T} routed rows A={A} (r={A / T:.2f})")
    lgs = [lg.clone() for _ in range(4)]
    out = torch.empty_like(xs[0])
    res = {}
    res["router (gate + routing)"] = timeit(lambda i: ops.router(xs[i % 4], wg, ws))
    res["router (logits_in, routing only)"] = timeit(lambda i: ops.router(None, None, ws, logits_in=lgs[i % 4]))
    ops.router(xs[0], wg, ws)
    res["plan"] = timeit(lambda i: ops.plan(ws))
    res["permute"] = timeit(lambda i: ops.permute(xs[0], mk, gw, ws))
    res["ffn gemm1"] = timeit(lambda i: ops.grouped_ffn(xs[0], m._w13, m._w2, ws, 0, 1))
    res["ffn gemm2"] = timeit(lambda i: ops.grouped_ffn(xs[0], m._w13, m._w2, ws, 0, 2))
    res["combine"] = timeit(lambda i: ops.combine(ws, out))
    res["layer"] = timeit(lambda i: m(xs[i % 4].view(1, T, 2048), None, None))
    for k, v in res.items():
        print(f"{k:36s} {v:9.1f} us")
    hbm = {"router (gate + routing)": T * 4192, "permute": (T + A) * 4096, "combine": (2 * T + A) * 4096}
    for k, b in hbm.items():
        print(f"  {k}: {b / res[k] / 1e3:.0f} GB/s")
    print(f"  gemm1 {(T + A) * 4 * 2048 * 2752 / res['ffn gemm1'] / 1e6:.0f} TFLOP/s  gemm2 {(T + A) * 2 * 2048 * 2752 / res['ffn gemm2'] / 1e6:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
